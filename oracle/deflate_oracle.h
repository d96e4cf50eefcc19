/*
 * deflate_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the DEFLATE encode path of image-rs/deflate-rs
 * (crate `deflate` 1.0.0).  Every function in deflate_oracle.c cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py may load this library; the product library
 * (deflate-rs_b200/libdeflate_b200.so) never links, loads or calls it.
 *
 * Parity status: PINNED against the reference's own known-answer tests and fixtures
 * (tests/test_oracle_kat.py lists each vector with its reference file:line).  The
 * reference itself cannot be compiled here (no rustc/cargo in the image), so there is
 * no oracle/_ref build; see DESIGN.md.
 */
#ifndef DEFLATE_ORACLE_H
#define DEFLATE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* compression_options.rs:78-120 */
typedef struct {
    uint16_t max_hash_checks;
    uint16_t lazy_if_less_than;
    uint8_t matching_type; /* 0 = Greedy, 1 = Lazy (lz77.rs:26-37) */
    uint8_t special;       /* 0 = Normal (compression_options.rs:52-59) */
} dfo_options;

enum { DFO_RAW = 0, DFO_ZLIB = 1, DFO_GZIP = 2 };
enum { DFO_FLUSH_NONE = 0, DFO_FLUSH_SYNC = 1, DFO_FLUSH_FINISH = 2 }; /* compress.rs:17-30 */

/* A token as produced by the LZ77 stage: dist == 0 -> literal `litlen`,
 * else match of length litlen+3 at distance dist (lzvalue.rs:42-76). */
typedef struct {
    uint16_t dist;
    uint8_t litlen;
    uint8_t pad;
} dfo_token;

/* One-shot API: lib.rs:137 (raw), :182 (zlib), :242 (gzip, default GzBuilder header).
 * *out is malloc'd; release with dfo_free. Returns 0 on success. */
int dfo_compress(const uint8_t *in, size_t n, const dfo_options *opt, int wrap,
                 uint8_t **out, size_t *out_len);
void dfo_free(void *p);

/* Streaming API mirroring write::{DeflateEncoder,ZlibEncoder,GzEncoder} over a Vec sink
 * (writer.rs:89-290). */
typedef struct dfo_stream dfo_stream;
dfo_stream *dfo_stream_new(const dfo_options *opt, int wrap);
/* write_all(buf) -- loops Write::write until everything is consumed. */
int dfo_stream_write(dfo_stream *s, const uint8_t *buf, size_t n);
int dfo_stream_flush(dfo_stream *s);  /* Write::flush == Z_SYNC_FLUSH (writer.rs:134) */
int dfo_stream_finish(dfo_stream *s); /* finish(): output_all + trailer */
int dfo_stream_reset(dfo_stream *s);  /* reset(W): finish current, start a new stream;
                                         the old sink's bytes stay available below */
uint32_t dfo_stream_checksum(const dfo_stream *s); /* ZlibEncoder::checksum (writer.rs:248) */
const uint8_t *dfo_stream_output(const dfo_stream *s, size_t *len);
void dfo_stream_clear_output(dfo_stream *s);
void dfo_stream_free(dfo_stream *s);

/* ---- hooks for known-answer tests (each restates one reference function) ---- */

/* lz77.rs:869-905 lz77_compress_conf: all tokens of the stream, blocks concatenated.
 * block_ends (optional, malloc'd) receives the cumulative token count at each block end. */
int dfo_lz77_tokens(const uint8_t *in, size_t n, const dfo_options *opt, dfo_token **toks,
                    size_t *ntoks, size_t **block_ends, size_t *nblocks);
/* matching.rs:13-73 */
size_t dfo_get_match_length(const uint8_t *data, size_t len, size_t cur, size_t pos_to_check);
/* matching.rs:87-166 over a table filled like chained_hash_table.rs:222-229
 * (filled_hash_table on data[..fill_len], then add_hash_value semantics). */
int dfo_longest_match_filled(const uint8_t *data, size_t len, size_t fill_len, size_t position,
                             size_t prev_length, uint16_t max_hash_checks, size_t *out_len,
                             size_t *out_dist);
/* length_encode.rs:347-415 */
void dfo_huffman_lengths(const uint16_t *freqs, size_t n, size_t max_len, uint8_t *lens);
/* length_encode.rs:82-155. out_sym[i]: 0..15 literal length, 16/17/18 repeat codes;
 * out_arg[i]: repeat count (0 for literals). Returns number of symbols. */
size_t dfo_encode_lengths(const uint8_t *lens, size_t n, uint8_t *out_sym, uint8_t *out_arg,
                          uint16_t freqs19[19]);
/* huffman_table.rs:253-278 */
void dfo_create_codes(const uint8_t *lens, size_t n, uint16_t *codes);
/* huffman_table.rs:143-182 */
unsigned dfo_length_code(uint16_t length, unsigned *extra_bits, unsigned *extra_val);
unsigned dfo_distance_code(uint16_t distance, unsigned *extra_bits, unsigned *extra_val);
/* bitstream.rs:76-106: write (v[i], nbits[i]) pairs then flush_raw. Returns byte count. */
size_t dfo_bitwriter_kat(const uint16_t *v, const uint8_t *nbits, size_t n, uint8_t *out,
                         size_t out_cap);
/* huffman_lengths.rs:113-124 */
uint64_t dfo_stored_padding(uint8_t pending_bits);
/* bit_reverse.rs:3-10 */
uint16_t dfo_reverse_bits(uint16_t v, uint8_t nbits);
/* compress.rs:43-57 compress_data_fixed: one fixed block, high() options. */
int dfo_compress_fixed(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len);
/* checksum.rs:33-57 / adler32 crate 1.2.0 (RFC 1950 Adler-32). */
uint32_t dfo_adler32(uint32_t adler, const uint8_t *buf, size_t n);
/* gzip-header 1.0 Crc (RFC 1952 CRC-32). */
uint32_t dfo_crc32(uint32_t crc, const uint8_t *buf, size_t n);
/* chained_hash_table.rs: snapshot of head/prev after filled_hash_table(data). */
void dfo_hash_table_filled(const uint8_t *data, size_t len, uint16_t *head, uint16_t *prev);

#ifdef __cplusplus
}
#endif
#endif
