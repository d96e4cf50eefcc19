/*
 * deflate_b200.h -- C ABI of the B200-native DEFLATE encoder (libdeflate_b200.so).
 *
 * This is the drop-in boundary for the encode path of image-rs/deflate-rs (crate `deflate`
 * 1.0.0).  Each entry point names the reference interface it replaces (file:line under the
 * reference's src/).  The reference has no FFI of its own (pure Rust, forbid(unsafe_code));
 * the seam is the single internal function every public encoder funnels into,
 *     compress_data_dynamic_n(input, &mut DeflateState<W>, Flush) -> io::Result<usize>
 * (compress.rs:80-84), reached from deflate_bytes* via compress_until_done (writer.rs:15-58,
 * lib.rs:110-122) and from write::{DeflateEncoder,ZlibEncoder,GzEncoder} (writer.rs:124-136,
 * 254-276).  INTEGRATION.md shows the Rust shim (extern "C" block + safe wrappers with the
 * crate's public names) that a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; 0 = success, > 0 = retryable, < 0 = fatal; nothing
 * throws or aborts across this boundary.  All compute runs in hand-written sm_100a CUDA kernels;
 * there is no CPU fallback -- without a CUDA device every compute entry point returns
 * DFL_E_NODEVICE.
 */
#ifndef DEFLATE_B200_H
#define DEFLATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFL_VERSION 100 /* 0.1.0 */

/* ---- CompressionOptions (compression_options.rs:78-120); layout crosses the ABI --------- */
typedef struct dfl_options {
    uint16_t max_hash_checks;   /* compression_options.rs:85  */
    uint16_t lazy_if_less_than; /* compression_options.rs:100 */
    uint8_t matching_type;      /* lz77.rs:26-37: 0 = Greedy, 1 = Lazy */
    uint8_t special;            /* compression_options.rs:52-59: 0 = Normal (others are
                                   unimplemented placeholders in the reference too) */
} dfl_options;

/* Compression::{Fast,Default,Best} (compression_options.rs:31-42) and the named presets
 * (compression_options.rs:126-178). */
enum dfl_preset {
    DFL_PRESET_FAST = 0,        /* 1 check, greedy              (:141-148) */
    DFL_PRESET_DEFAULT = 1,     /* 128 checks, lazy < 32        (:19-20,67-72) */
    DFL_PRESET_BEST = 2,        /* == high(): 1768, lazy < 128  (:14-15,126-133) */
    DFL_PRESET_HUFFMAN_ONLY = 3,/* 0 checks, greedy             (:155-162) */
    DFL_PRESET_RLE = 4          /* 0 checks, lazy => RLE        (:171-178) */
};
int dfl_options_preset(int preset, dfl_options *out);

/* Container written around the raw DEFLATE stream. */
enum dfl_wrap {
    DFL_RAW = 0,  /* deflate_bytes[_conf]       lib.rs:137,163 / DeflateEncoder writer.rs:89  */
    DFL_ZLIB = 1, /* deflate_bytes_zlib[_conf]  lib.rs:182,216 / ZlibEncoder    writer.rs:183 */
    DFL_GZIP = 2  /* deflate_bytes_gzip[_conf]  lib.rs:242,284 / GzEncoder      writer.rs:331 */
};

/* Flush (compress.rs:17-30). */
enum dfl_flush {
    DFL_FLUSH_SYNC = 1,  /* Flush::Sync: close the block, append 00 00 FF FF (compress.rs:258-261) */
    DFL_FLUSH_FINISH = 2 /* Flush::Finish */
};

enum dfl_status {
    DFL_OK = 0,
    DFL_AGAIN = 1,          /* io::ErrorKind::Interrupted, "internal buffer full" (compress.rs:117-120) */
    DFL_E_ARG = -1,
    DFL_E_NOMEM = -2,
    DFL_E_CUDA = -3,
    DFL_E_NODEVICE = -4,    /* no CUDA device / extension unusable: there is no CPU fallback */
    DFL_E_OVERFLOW = -5,    /* out_cap too small; *out_len receives a capacity that suffices (see dfl_bound) */
    DFL_E_STATE = -6,       /* e.g. write after finish */
    DFL_E_UNSUPPORTED = -7,
    DFL_E_INTERNAL = -8,
    DFL_E_NCCL = -10        /* NCCL missing or a collective failed (dfl_comm_last_error has the text) */
};
const char *dfl_strerror(int status);
/* Text of the last CUDA error seen by the calling thread ("" if none). */
const char *dfl_last_cuda_error(void);
int dfl_version(void);
/* Number of usable CUDA devices (0 on a CPU-only box; never an error). */
int dfl_device_count(void);

/* Upper bound of the output size for n input bytes (any options, any wrapper with its default header; a caller
 * that passes its own gzip header adds gz_hdr_len).  Device output buffers must be 16-byte aligned and the kernels
 * write whole 16-byte units: a capacity of dfl_bound() is always enough; when a smaller one turns out too small the
 * call returns DFL_E_OVERFLOW and *out_len holds a capacity that is enough for this input. */
size_t dfl_bound(size_t n, int wrap);

/* ---- one-shot, host buffers: deflate_bytes_conf / _zlib_conf / _gzip_conf -------------------
 * (lib.rs:137, 182, 242).  gz_hdr/gz_hdr_len: complete RFC 1952 member header produced by the
 * caller (the crate's GzBuilder::into_header(), lib.rs:251); NULL/0 = the default header.
 * Copies in -> device, encodes, copies the stream back.  Blocking. */
int dfl_compress(const uint8_t *in, size_t n, const dfl_options *opt, int wrap,
                 const uint8_t *gz_hdr, size_t gz_hdr_len, uint8_t *out, size_t out_cap,
                 size_t *out_len);

/* ---- one-shot, device buffers (the measured path) -----------------------------------------
 * d_in/d_out are device pointers on the current CUDA device; `stream` is a cudaStream_t (NULL =
 * the context's own stream).  The complete container (header, stream, trailer) is left in d_out;
 * its size is returned in *out_len after the stream has been synchronised. */
int dfl_compress_device(const void *d_in, size_t n, const dfl_options *opt, int wrap,
                        const uint8_t *gz_hdr, size_t gz_hdr_len, void *d_out, size_t out_cap,
                        size_t *out_len, void *stream);

/* ---- one piece of a longer stream, device buffers ------------------------------------------
 * compress_data_dynamic_n(piece, &mut state, Flush::{Sync,Finish}) (compress.rs:80-84) for a state
 * whose window already holds the plaintext in front of the piece: d_in points at `dict_len` bytes
 * of that plaintext (the last 32768 are what matters, matching.rs:102-106) followed by the piece,
 * n_total bytes in all.  Produces the piece's raw DEFLATE blocks only (no container bytes):
 * DFL_FLUSH_SYNC ends them with the empty stored block 00 00 FF FF (compress.rs:258-261), so the
 * next piece starts byte aligned and pieces encoded independently -- e.g. one per GPU -- concatenate
 * into the stream the reference's writer produces with flush() called at the same offsets;
 * DFL_FLUSH_FINISH sets BFINAL on the last block. */
int dfl_compress_device_piece(const void *d_in, size_t n_total, size_t dict_len, const dfl_options *opt,
                              int flush_mode, void *d_out, size_t out_cap, size_t *out_len, void *stream);

/* ---- many independent streams, device buffers ----------------------------------------------
 * `count` separate calls of deflate_bytes*_conf (lib.rs:137,182,242) with the same options, e.g. the
 * IDAT chunks of a batch of PNGs: stream i is d_in[i][0..n[i]) -> d_out[i], size in out_len[i].
 * Several pipelines are kept in flight on different CUDA streams so that small inputs still fill the
 * GPU.  Blocking.  status (may be NULL) receives one code per stream; the return value is the first
 * non-zero one.  gzip streams get the default header. */
int dfl_compress_device_batch(size_t count, const void *const *d_in, const size_t *n, const dfl_options *opt,
                              int wrap, void *const *d_out, const size_t *out_cap, size_t *out_len, int *status);

/* ---- many independent streams, host buffers ---------------------------------------------------
 * The same `count` calls of deflate_bytes*_conf (lib.rs:137,182,242) for callers whose data lives in
 * host memory (the image encoders this crate serves: one IDAT stream per picture).  The members rotate
 * over a pool of pipelines, each with its own stream: a member's input copy, kernels and output copy
 * overlap those of the others.  Pinned buffers copy asynchronously; pageable ones work, staged by the
 * driver (members are small).  Byte-for-byte what dfl_compress returns for each member.  Blocking; status as above. */
int dfl_compress_batch(size_t count, const uint8_t *const *in, const size_t *n, const dfl_options *opt, int wrap,
                       uint8_t *const *out, const size_t *out_cap, size_t *out_len, int *status);

/* Per-stage device timings of the most recent dfl_compress_device call on this thread, in
 * milliseconds (CUDA events on the launching stream): names[i] points to a static string.
 * Returns the number of stages (0 if timing was not enabled with dfl_set_profiling(1)). */
int dfl_set_profiling(int enabled);
int dfl_last_stage_times(const char **names, float *ms, int cap);
/* Counters of the most recent call: [0] tokens, [1] deflate blocks, [2] parse segments,
 * [3] segments re-parsed in parallel rounds, [4] segments re-parsed sequentially,
 * [5] kernels launched, [6] stored blocks, [7] fixed blocks. */
int dfl_last_counters(uint64_t *out, int cap);

/* Releases the scratch the library keeps between calls: the calling thread's one-shot context (about 29 bytes of
 * device memory per input byte of its largest call so far) and batch pools, and the process-wide pool of parked
 * handle resources.  Nothing in the reference corresponds to it (a Vec is dropped when deflate_bytes returns,
 * lib.rs:141-146); here allocations are kept because cudaMalloc of a context costs more than encoding with it.
 * Live dfl_encoder handles are not touched.  Always DFL_OK. */
int dfl_trim(void);

/* ---- streaming: write::{DeflateEncoder,ZlibEncoder,GzEncoder} (writer.rs:89-467) ----------
 * The generic sink W cannot cross an FFI, so output is pulled: the shim forwards the bytes lent
 * by dfl_encoder_take_output to its `inner.write(..)` and reports progress with
 * dfl_encoder_advance_output, which preserves the reference's partial-write bookkeeping
 * (compress.rs:96-111,286-299). */
typedef struct dfl_encoder dfl_encoder;

dfl_encoder *dfl_encoder_new(const dfl_options *opt, int wrap, const uint8_t *gz_hdr,
                             size_t gz_hdr_len);                     /* ::new / from_builder */
/* Write::write (writer.rs:124-127,254-267): consumes up to n bytes, *consumed <= n.  `buf` may be reused
 * when the call returns: its bytes are on their way to the device (pageable memory goes through the
 * library's pinned staging slots, pinned memory is copied directly and waited for). */
int dfl_encoder_write(dfl_encoder *e, const uint8_t *buf, size_t n, size_t *consumed);
/* Write::flush = DFL_FLUSH_SYNC (writer.rs:134-136); finish()/Drop = DFL_FLUSH_FINISH
 * (writer.rs:103-108,139-152).  After FINISH the trailer is part of the output. */
int dfl_encoder_flush(dfl_encoder *e, int mode);
/* Written input is encoded at every flush and, on its own, whenever `bytes` of it have accumulated
 * (default 256 MiB; 4096 <= bytes <= 2 GiB), so memory stays bounded and a stream can be longer than
 * 4 GiB.  The output does not depend on this value: without a flush the reference's stream has no seam
 * (lib.rs:408-433), and neither has this one.  Such a piece runs while later writes arrive; its bytes
 * become available at a later write, flush or take_output call, always in stream order. */
int dfl_encoder_set_piece_bytes(dfl_encoder *e, size_t bytes);
/* Lends the bytes produced so far and not yet advanced over; the pointer is valid until the next call on
 * the handle.  After a flush everything up to the flush point is included. */
int dfl_encoder_take_output(dfl_encoder *e, const uint8_t **p, size_t *len);
void dfl_encoder_advance_output(dfl_encoder *e, size_t n);
/* ZlibEncoder::checksum (writer.rs:248) / GzEncoder::checksum (writer.rs:429): checksum of the
 * bytes consumed so far (Adler-32 or CRC-32; 1 for DFL_RAW like NoChecksum, checksum.rs:26-28). */
uint32_t dfl_encoder_checksum(dfl_encoder *e);
/* reset(W) (writer.rs:112-115,218-223,394-401): finishes the current stream (its bytes stay
 * available through take_output) and starts a new one with the same options. */
int dfl_encoder_reset(dfl_encoder *e, const uint8_t *gz_hdr, size_t gz_hdr_len);
void dfl_encoder_free(dfl_encoder *e);

/* ---- building blocks exposed for tests and for callers that keep data on the device --------
 * Adler-32 (checksum.rs:33-57) of a device buffer, computed on the device. */
int dfl_adler32_device(const void *d_in, size_t n, uint32_t *adler, void *stream);
/* CRC-32 (gzip-header 1.0 `Crc`, the reference's gzip trailer: lib.rs:257-265, writer.rs:408-426)
 * of a device buffer, computed on the device. */
int dfl_crc32_device(const void *d_in, size_t n, uint32_t *crc, void *stream);
/* Runs only the entropy stage (block cut at 31744 tokens, code construction, block-type choice,
 * bit packing: huffman_lengths.rs:167-369, encoder_state.rs:58-105, compress.rs:187-247) on a
 * caller-supplied token stream (host memory; token = literal byte, or len | dist << 9), so the
 * stage can be compared bit-for-bit with the reference fed the same tokens. `in` is the
 * uncompressed data the tokens describe (needed for stored blocks). */
int dfl_encode_tokens(const uint8_t *in, size_t n, const uint32_t *tokens, size_t n_tokens,
                      uint8_t *out, size_t out_cap, size_t *out_len);
/* Runs only the LZ77 stage and returns the token stream (host memory, same encoding). */
int dfl_lz77_tokens(const uint8_t *in, size_t n, const dfl_options *opt, uint32_t *tokens,
                    size_t tokens_cap, size_t *n_tokens);

/* ---- multi-GPU: bringing the compressed streams of all ranks to one rank ----------
 * The reference is single-process; this is what replaces "append to the caller's Vec" (lib.rs:110-122)
 * when independent inputs -- or the pieces of one stream, dfl_compress_device_piece -- are encoded on several
 * GPUs, one process per GPU.  Encoding needs no collective; the one exchange step is a gather-v of byte
 * streams over NCCL (NVLink / NVSwitch): ncclAllGather of the 8-byte sizes, then grouped ncclSend / ncclRecv
 * at the prefix offsets, all on the caller's stream.  NCCL is loaded at run time; without it these calls
 * return DFL_E_NCCL and everything else keeps working. */
typedef struct dfl_comm dfl_comm;
#define DFL_COMM_ID_BYTES 128
/* Rank 0 creates the id and hands it to the other ranks out of band (MPI, torch.distributed, a file ...). */
int dfl_comm_unique_id(uint8_t *id /* DFL_COMM_ID_BYTES */);
/* Collective over all `world` ranks; binds the communicator to the current CUDA device. */
int dfl_comm_init(dfl_comm **comm, int world, int rank, const uint8_t *id);
void dfl_comm_free(dfl_comm *comm);
int dfl_comm_world(const dfl_comm *comm);
int dfl_comm_rank(const dfl_comm *comm);
const char *dfl_comm_last_error(void);
/* Collective.  Every rank contributes the first n bytes at d_src (device); on `root` they arrive back to back in
 * rank order at d_dst (device, dst_cap bytes; ignored elsewhere).  sizes (host, world entries, may be NULL)
 * receives every rank's n on every rank, so rank r's bytes are d_dst[sum(sizes[0..r)) ...).  The size exchange
 * costs one short host wait; the payload is queued on `stream` (a cudaStream_t) and the call returns without
 * waiting for it, so the transfer overlaps whatever the caller encodes next on another stream. */
int dfl_gather_device(dfl_comm *comm, const void *d_src, size_t n, void *d_dst, size_t dst_cap, size_t *sizes,
                      int root, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DEFLATE_B200_H */
