/*
 * deflate_oracle.c -- TEST INFRASTRUCTURE ONLY (see deflate_oracle.h).
 *
 * Plain-C restatement of the encode path of image-rs/deflate-rs @ 3262c25
 * (crate `deflate` 1.0.0).  It is written from the behaviour of the reference, one
 * reference function per C function, each citing `file:line` under /root/reference/src.
 * It is deliberately sequential and single-threaded, like the reference.
 *
 * Parity: pinned by tests/test_oracle_kat.py against the reference's known-answer tests.
 */
#include "deflate_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ constants */
#define WINDOW_SIZE 32768u /* chained_hash_table.rs:1 */
#define WINDOW_MASK (WINDOW_SIZE - 1u)
#define HASH_SHIFT 5u    /* chained_hash_table.rs:5 */
#define HASH_MASK 0x7fffu /* chained_hash_table.rs:6 */
#define MIN_MATCH 3u     /* huffman_table.rs:20 */
#define MAX_MATCH 258u   /* huffman_table.rs:21 */
#define BUFFER_SIZE (WINDOW_SIZE * 2u + MAX_MATCH) /* input_buffer.rs:8 */
#define MAX_BUFFER_LENGTH (1024u * 31u)            /* output_writer.rs:19 */
#define NUM_LL 286u                                /* huffman_table.rs:13 */
#define NUM_DIST 30u                               /* huffman_table.rs:9 */
#define END_OF_BLOCK 256u                          /* huffman_table.rs:28 */
#define MAX_CODE_LENGTH 15u                        /* huffman_table.rs:17 */
#define MAX_STORED_BLOCK_LENGTH 32767u             /* stored_block.rs:11 */
#define LARGEST_OUTPUT_BUF_SIZE (1024u * 32u)      /* compress.rs:12 */
#define LAZY_CLAMP 32768u                          /* deflate_state.rs:105 (MAX_HASH_CHECKS) */

static void die(const char *msg) {
    fprintf(stderr, "deflate_oracle: reference would panic: %s\n", msg);
    abort();
}

/* ------------------------------------------------------------------ byte vector */
typedef struct {
    uint8_t *p;
    size_t len, cap;
} bytevec;

static void bv_reserve(bytevec *v, size_t extra) {
    if (v->len + extra > v->cap) {
        size_t nc = v->cap ? v->cap * 2 : 4096;
        while (nc < v->len + extra) nc *= 2;
        v->p = (uint8_t *)realloc(v->p, nc);
        if (!v->p) die("out of memory");
        v->cap = nc;
    }
}
static void bv_push(bytevec *v, uint8_t b) {
    bv_reserve(v, 1);
    v->p[v->len++] = b;
}
static void bv_extend(bytevec *v, const uint8_t *d, size_t n) {
    bv_reserve(v, n);
    if (n) memcpy(v->p + v->len, d, n);
    v->len += n;
}

/* ------------------------------------------------------------------ bit_reverse.rs:3-10 */
uint16_t dfo_reverse_bits(uint16_t v, uint8_t nbits) {
    uint16_t r = 0;
    for (int i = 0; i < 16; i++)
        if (v & (1u << i)) r |= (uint16_t)(1u << (15 - i));
    return (uint16_t)(r >> (16 - nbits)); /* callers never pass nbits == 0 */
}

/* ------------------------------------------------------------------ bitstream.rs:54-106 */
typedef struct {
    bytevec w;
    uint8_t bits;
    uint64_t acc;
} lsb_writer;

/* bitstream.rs:76-86 (64-bit arch_dep: FLUSH_AT = 48, six bytes per push) */
static void lsb_write_bits(lsb_writer *s, uint16_t v, uint8_t n) {
    s->acc |= ((uint64_t)v) << s->bits;
    s->bits = (uint8_t)(s->bits + n);
    while (s->bits >= 48) {
        uint8_t six[6];
        for (int i = 0; i < 6; i++) six[i] = (uint8_t)(s->acc >> (8 * i));
        bv_extend(&s->w, six, 6);
        s->acc >>= 48;
        s->bits = (uint8_t)(s->bits - 48);
    }
}
/* bitstream.rs:88-97 */
static void lsb_write_bits_finish(lsb_writer *s, uint16_t v, uint8_t n) {
    s->acc |= ((uint64_t)v) << s->bits;
    s->bits = (uint8_t)(s->bits + n % 8);
    while (s->bits >= 8) {
        bv_push(&s->w, (uint8_t)s->acc);
        s->acc >>= 8;
        s->bits = (uint8_t)(s->bits - 8);
    }
}
/* bitstream.rs:99-106 */
static void lsb_flush_raw(lsb_writer *s) {
    uint8_t missing = (uint8_t)(48 - s->bits);
    if (missing > 0 && s->bits > 0) lsb_write_bits_finish(s, 0, missing);
}
/* bitstream.rs:109-119 (impl Write) */
static void lsb_write_bytes(lsb_writer *s, const uint8_t *buf, size_t n) {
    if (s->acc == 0) {
        bv_extend(&s->w, buf, n);
    } else {
        for (size_t i = 0; i < n; i++) lsb_write_bits(s, buf[i], 8);
    }
}

size_t dfo_bitwriter_kat(const uint16_t *v, const uint8_t *nbits, size_t n, uint8_t *out,
                         size_t out_cap) {
    lsb_writer w;
    memset(&w, 0, sizeof w);
    for (size_t i = 0; i < n; i++) lsb_write_bits(&w, v[i], nbits[i]);
    lsb_flush_raw(&w);
    size_t len = w.w.len;
    if (len <= out_cap && len) memcpy(out, w.w.p, len);
    free(w.w.p);
    return len;
}

/* ------------------------------------------------------------------ huffman_table.rs tables */
/* huffman_table.rs:32-42 */
static uint8_t FIXED_CODE_LENGTHS[288];
/* huffman_table.rs:45-47 */
static const uint8_t LENGTH_EXTRA_BITS_LENGTH[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                                     2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
/* huffman_table.rs:65-68 */
static const uint8_t BASE_LENGTH[29] = {0,  1,  2,  3,  4,  5,  6,  7,  8,   10,  12,  14,  16,  20, 24,
                                        28, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 255};
/* huffman_table.rs:108-111 */
static const uint16_t DISTANCE_BASE[30] = {0,   1,   2,   3,   4,    6,    8,    12,   16,   24,
                                           32,  48,  64,  96,  128,  192,  256,  384,  512,  768,
                                           1024, 1536, 2048, 3072, 4096, 6144, 8192, 12288, 16384, 24576};
/* huffman_table.rs:50-62 LENGTH_CODE and :77-99 DISTANCE_CODES are regular; they are
 * generated here from the base tables instead of being typed in (tests pin the values). */
static uint8_t LENGTH_CODE[256];
static uint8_t DISTANCE_CODES[512];
static int tables_ready = 0;

static void init_tables(void) {
    if (tables_ready) return;
    for (int i = 0; i < 288; i++)
        FIXED_CODE_LENGTHS[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
    /* LENGTH_CODE[stored_len] = greatest code whose base <= stored_len; 255 -> 28 */
    for (int l = 0; l < 256; l++) {
        int c = 0;
        for (int k = 0; k < 29; k++)
            if (BASE_LENGTH[k] <= l) c = k;
        LENGTH_CODE[l] = (uint8_t)c;
    }
    /* DISTANCE_CODES[0..256): code of distance d = index+1.
     * DISTANCE_CODES[256..512): code of distance ((index-256) << 7) + 1 (upper-bit lookup). */
    for (int i = 0; i < 256; i++) {
        int d0 = i; /* distance - 1 */
        int c = 0;
        for (int k = 0; k < 30; k++)
            if (DISTANCE_BASE[k] <= d0) c = k;
        DISTANCE_CODES[i] = (uint8_t)c;
    }
    for (int i = 0; i < 256; i++) {
        int d0 = i << 7;
        int c = 0;
        for (int k = 0; k < 30; k++)
            if (DISTANCE_BASE[k] <= d0) c = k;
        /* entries 256,257 correspond to d0 = 0,128 which the reference never looks up via the
         * upper table (distance <= 256 uses the lower table); the reference stores 0 there. */
        DISTANCE_CODES[256 + i] = (uint8_t)(i < 2 ? 0 : c);
    }
    tables_ready = 1;
}

/* huffman_table.rs:113-115 */
static uint8_t num_extra_bits_for_length_code(uint8_t code) { return LENGTH_EXTRA_BITS_LENGTH[code]; }
/* huffman_table.rs:120-126 */
static uint8_t num_extra_bits_for_distance_code(uint8_t code) {
    uint8_t c = (uint8_t)(code >> 1);
    c = (uint8_t)(c - (c != 0));
    return c;
}
/* huffman_table.rs:143-147 */
static unsigned get_length_code(uint16_t length) {
    return (unsigned)LENGTH_CODE[(uint8_t)(length - MIN_MATCH)] + 257u;
}
/* huffman_table.rs:170-182 */
static uint8_t get_distance_code(uint16_t distance) {
    unsigned d = distance;
    if (d >= 1 && d <= 256) return DISTANCE_CODES[d - 1];
    if (d >= 257 && d <= 32768) return DISTANCE_CODES[256 + ((d - 1) >> 7)];
    return 0;
}

unsigned dfo_length_code(uint16_t length, unsigned *extra_bits, unsigned *extra_val) {
    init_tables();
    /* huffman_table.rs:150-165 get_length_code_and_extra_bits */
    uint8_t stored = (uint8_t)(length - MIN_MATCH);
    uint8_t n = LENGTH_CODE[stored];
    if (extra_bits) *extra_bits = num_extra_bits_for_length_code(n);
    if (extra_val) *extra_val = (unsigned)(stored - BASE_LENGTH[n]);
    return (unsigned)n + 257u;
}
unsigned dfo_distance_code(uint16_t distance, unsigned *extra_bits, unsigned *extra_val) {
    init_tables();
    /* huffman_table.rs:184-195 get_distance_code_and_extra_bits */
    uint8_t c = get_distance_code(distance);
    if (extra_bits) *extra_bits = num_extra_bits_for_distance_code(c);
    if (extra_val) *extra_val = (unsigned)(uint16_t)(distance - (DISTANCE_BASE[c] + 1));
    return c;
}

/* huffman_table.rs:253-278 create_codes_in_place (with build_length_count_table :229-249) */
static void create_codes_in_place(uint16_t *code_table, const uint8_t *length_table, size_t n) {
    uint16_t len_counts[16];
    memset(len_counts, 0, sizeof len_counts);
    if (n == 0) die("BUG! Empty lengths!");
    unsigned max_length = 0;
    size_t max_length_pos = 0;
    for (size_t i = 0; i < n; i++)
        if (length_table[i] > max_length) max_length = length_table[i];
    if (max_length > MAX_CODE_LENGTH) die("code length > 15");
    for (size_t i = 0; i < n; i++) {
        if (length_table[i] > 0) {
            len_counts[length_table[i]]++;
            max_length_pos = i;
        }
    }
    uint16_t next_code[17];
    uint16_t code = 0;
    next_code[0] = code;
    for (unsigned bits = 1; bits <= max_length; bits++) {
        code = (uint16_t)((code + len_counts[bits - 1]) << 1);
        next_code[bits] = code;
    }
    for (size_t i = 0; i <= max_length_pos; i++) {
        unsigned length = length_table[i];
        if (length != 0) {
            code_table[i] = dfo_reverse_bits(next_code[length], (uint8_t)length);
            next_code[length] = (uint16_t)(next_code[length] + 1);
        }
    }
}
void dfo_create_codes(const uint8_t *lens, size_t n, uint16_t *codes) {
    create_codes_in_place(codes, lens, n);
}

/* huffman_table.rs:281-288 */
typedef struct {
    uint16_t codes[288];
    uint8_t code_lengths[288];
    uint16_t distance_codes[32];
    uint8_t distance_code_lengths[32];
} huffman_table;

/* huffman_table.rs:331-337 */
static void ht_update_from_lengths(huffman_table *t) {
    create_codes_in_place(t->codes, t->code_lengths, 288);
    create_codes_in_place(t->distance_codes, t->distance_code_lengths, 32);
}
/* huffman_table.rs:339-343 */
static void ht_set_to_fixed(huffman_table *t) {
    memcpy(t->code_lengths, FIXED_CODE_LENGTHS, 288);
    memset(t->distance_code_lengths, 5, 32);
    ht_update_from_lengths(t);
}

/* ------------------------------------------------------------------ encoder_state.rs */
typedef struct {
    huffman_table huffman_table;
    lsb_writer writer;
} encoder_state;

/* encoder_state.rs:85-99 */
static void es_write_start_of_block(encoder_state *es, int fixed, int final_block) {
    uint16_t v;
    if (final_block)
        v = fixed ? 3 /*0b011*/ : 5 /*0b101*/;
    else
        v = fixed ? 2 /*0b010*/ : 4 /*0b100*/;
    lsb_write_bits(&es->writer, v, 3);
}
/* encoder_state.rs:58-82 */
static void es_write_lzvalue(encoder_state *es, dfo_token t) {
    huffman_table *h = &es->huffman_table;
    if (t.dist == 0) {
        lsb_write_bits(&es->writer, h->codes[t.litlen], h->code_lengths[t.litlen]);
    } else {
        /* get_length_huffman (huffman_table.rs:373-386) */
        uint8_t n = LENGTH_CODE[t.litlen];
        unsigned code_number = (unsigned)n + 257u;
        lsb_write_bits(&es->writer, h->codes[code_number], h->code_lengths[code_number]);
        lsb_write_bits(&es->writer, (uint16_t)(t.litlen - BASE_LENGTH[n]),
                       num_extra_bits_for_length_code(n));
        /* get_distance_huffman (huffman_table.rs:392-410) */
        uint8_t dc = get_distance_code(t.dist);
        lsb_write_bits(&es->writer, h->distance_codes[dc], h->distance_code_lengths[dc]);
        lsb_write_bits(&es->writer, (uint16_t)(t.dist - (DISTANCE_BASE[dc] + 1)),
                       num_extra_bits_for_distance_code(dc));
    }
}
/* encoder_state.rs:102-105 */
static void es_write_end_of_block(encoder_state *es) {
    lsb_write_bits(&es->writer, es->huffman_table.codes[END_OF_BLOCK],
                   es->huffman_table.code_lengths[END_OF_BLOCK]);
}

/* ------------------------------------------------------------------ stored_block.rs */
/* stored_block.rs:13-23 */
static void write_stored_header(lsb_writer *w, int final_block) {
    lsb_write_bits(w, final_block ? 1 : 0, 3);
    lsb_flush_raw(w);
}
/* stored_block.rs:26-40 */
static void compress_block_stored(const uint8_t *input, size_t n, lsb_writer *w) {
    if (n > 65535) die("Stored block too long!");
    uint8_t hdr[4];
    uint16_t len = (uint16_t)n, nlen = (uint16_t)~n;
    hdr[0] = (uint8_t)len;
    hdr[1] = (uint8_t)(len >> 8);
    hdr[2] = (uint8_t)nlen;
    hdr[3] = (uint8_t)(nlen >> 8);
    lsb_write_bytes(w, hdr, 2);
    lsb_write_bytes(w, hdr + 2, 2);
    lsb_write_bytes(w, input, n);
}
/* compress.rs:59-77 */
static void write_stored_block(const uint8_t *input, size_t n, lsb_writer *w, int final_block) {
    if (n != 0) {
        size_t off = 0;
        while (off < n) {
            size_t chunk = n - off < MAX_STORED_BLOCK_LENGTH ? n - off : MAX_STORED_BLOCK_LENGTH;
            int last_chunk = (off + chunk >= n);
            write_stored_header(w, final_block && last_chunk);
            compress_block_stored(input + off, chunk, w);
            off += chunk;
        }
    } else {
        write_stored_header(w, final_block);
        compress_block_stored(NULL, 0, w);
    }
}

/* ------------------------------------------------------------------ length_encode.rs */
enum { EL_LENGTH = 0, EL_COPY_PREVIOUS = 16, EL_REPEAT_ZERO_3 = 17, EL_REPEAT_ZERO_7 = 18 };
typedef struct {
    uint8_t sym; /* 0..15 literal length value, or 16/17/18 */
    uint8_t arg; /* repeat count for 16/17/18 */
} encoded_length;

typedef struct {
    encoded_length *p;
    size_t len, cap;
} elvec;

/* length_encode.rs:42-57 update_out_and_freq */
static void el_push(elvec *out, uint16_t freqs[19], uint8_t sym, uint8_t arg) {
    if (out->len == out->cap) {
        out->cap = out->cap ? out->cap * 2 : 512;
        out->p = (encoded_length *)realloc(out->p, out->cap * sizeof(encoded_length));
        if (!out->p) die("out of memory");
    }
    freqs[sym]++;
    out->p[out->len].sym = sym;
    out->p[out->len].arg = arg;
    out->len++;
}
/* length_encode.rs:19-32 from_prev_and_repeat */
static void el_push_repeat(elvec *out, uint16_t freqs[19], uint8_t prev, uint8_t repeat) {
    if (prev == 0) {
        if (repeat <= 10)
            el_push(out, freqs, EL_REPEAT_ZERO_3, repeat);
        else
            el_push(out, freqs, EL_REPEAT_ZERO_7, repeat);
    } else if (prev <= 15) {
        el_push(out, freqs, EL_COPY_PREVIOUS, repeat);
    } else {
        die("from_prev_and_repeat: prev > 15");
    }
}
/* length_encode.rs:60-62 */
static int not_max_repetitions(uint8_t length_value, uint8_t repeats) {
    return (length_value == 0 && repeats < 138) || repeats < 6;
}
/* length_encode.rs:82-155 encode_lengths_m */
static void encode_lengths_m(const uint8_t *lengths, size_t n, elvec *out, uint16_t freqs[19]) {
    out->len = 0;
    if (n == 0) die("No length values!");
    uint8_t repeat = 0;
    uint8_t prev = (uint8_t)~lengths[0];
    for (size_t i = 0; i < n; i++) {
        uint8_t l = lengths[i];
        int at_end = (i + 1 == n); /* iter.peek().is_none() */
        if (l == prev && not_max_repetitions(l, repeat)) repeat++;
        if (l != prev || at_end || !not_max_repetitions(l, repeat)) {
            if (repeat >= 3) {
                el_push_repeat(out, freqs, prev, repeat);
                repeat = 0;
                if (l != prev) {
                    if (l != 0 || at_end) {
                        el_push(out, freqs, l, 0);
                        repeat = 0;
                    } else {
                        repeat = 1;
                    }
                }
            } else {
                size_t extra_skip = (at_end && l == prev) ? 1 : 0;
                size_t skip = i + extra_skip - repeat;
                size_t extra = (l != 0 || at_end) ? 1 : 0;
                size_t take = (size_t)repeat + extra;
                for (size_t k = skip; k < n && k < skip + take; k++) el_push(out, freqs, lengths[k], 0);
                repeat = (uint8_t)(1 - extra);
            }
        }
        prev = l;
    }
}

size_t dfo_encode_lengths(const uint8_t *lens, size_t n, uint8_t *out_sym, uint8_t *out_arg,
                          uint16_t freqs19[19]) {
    elvec v = {0, 0, 0};
    memset(freqs19, 0, 19 * sizeof(uint16_t));
    encode_lengths_m(lens, n, &v, freqs19);
    for (size_t i = 0; i < v.len; i++) {
        out_sym[i] = v.p[i].sym;
        out_arg[i] = v.p[i].arg;
    }
    size_t r = v.len;
    free(v.p);
    return r;
}

/* length_encode.rs:211-215 Node */
typedef struct {
    uint32_t value;
    uint16_t symbol;
} leaf_node;

/* length_encode.rs:218-246 step_1 */
static void mk_step_1(leaf_node *leaves, size_t n) {
    size_t root = 0, leaf = 2;
    leaves[0].value += leaves[1].value;
    for (size_t next = 1; next + 1 < n; next++) {
        if (leaf >= n || leaves[root].value < leaves[leaf].value) {
            leaves[next].value = leaves[root].value;
            leaves[root].value = (uint32_t)next;
            root++;
        } else {
            leaves[next].value = leaves[leaf].value;
            leaf++;
        }
        if (leaf >= n || (root < next && leaves[root].value < leaves[leaf].value)) {
            leaves[next].value += leaves[root].value;
            leaves[root].value = (uint32_t)next;
            root++;
        } else {
            leaves[next].value += leaves[leaf].value;
            leaf++;
        }
    }
}
/* length_encode.rs:248-278 step_2 */
static void mk_step_2(leaf_node *leaves, size_t n) {
    leaves[n - 2].value = 0;
    for (size_t t = n - 2; t-- > 0;) leaves[t].value = leaves[leaves[t].value].value + 1;
    size_t available = 1, used = 0;
    uint32_t depth = 0;
    long root = (long)n - 2, next = (long)n - 1;
    while (available > 0) {
        while (root >= 0 && leaves[root].value == depth) {
            used++;
            root--;
        }
        while (available > used) {
            leaves[next].value = depth;
            next--;
            available--;
        }
        available = 2 * used;
        depth++;
        used = 0;
    }
}
/* length_encode.rs:290-327 enforce_max_code_lengths */
static void enforce_max_code_lengths(uint16_t num_codes[33], size_t num_used, size_t max_len) {
    if (num_used > 1) {
        uint16_t num_above_max = 0;
        for (size_t i = max_len + 1; i < 33; i++) num_above_max = (uint16_t)(num_above_max + num_codes[i]);
        num_codes[max_len] = (uint16_t)(num_codes[max_len] + num_above_max);
        uint32_t total = 0;
        for (size_t i = max_len; i >= 1; i--) total += ((uint32_t)num_codes[i]) << (max_len - i);
        while (total != (1u << max_len)) {
            num_codes[max_len]--;
            for (size_t i = max_len - 1; i >= 1; i--) {
                if (num_codes[i] != 0) {
                    num_codes[i]--;
                    num_codes[i + 1] = (uint16_t)(num_codes[i + 1] + 2);
                    break;
                }
            }
            total--;
        }
    }
}
/* stable merge sort by value (Rust's slice::sort_by is stable; length_encode.rs:384-386) */
static void stable_sort_leaves(leaf_node *a, size_t n) {
    if (n < 2) return;
    leaf_node *tmp = (leaf_node *)malloc(n * sizeof(leaf_node));
    for (size_t w = 1; w < n; w *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (a[j].value < a[i].value) ? a[j++] : a[i++];
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, n * sizeof(leaf_node));
    }
    free(tmp);
}
/* length_encode.rs:347-415 in_place_lengths; `lengths` has `nl` entries, all zeroed first */
static void in_place_lengths(const uint16_t *freqs, size_t nf, size_t max_len, uint8_t *lengths,
                             size_t nl) {
    leaf_node leaves[288];
    size_t n = 0;
    memset(lengths, 0, nl);
    for (size_t i = 0; i < nf; i++) {
        if (freqs[i] > 0) {
            leaves[n].value = freqs[i];
            leaves[n].symbol = (uint16_t)i;
            n++;
        }
    }
    if (n == 1) {
        lengths[leaves[0].symbol] = 1;
        return;
    } else if (n == 0) {
        return;
    }
    stable_sort_leaves(leaves, n);
    mk_step_1(leaves, n);
    mk_step_2(leaves, n);
    uint16_t num_codes[33];
    memset(num_codes, 0, sizeof num_codes);
    for (size_t i = 0; i < n; i++) {
        if (leaves[i].value >= 33) die("code depth >= 33");
        num_codes[leaves[i].value]++;
    }
    enforce_max_code_lengths(num_codes, n, max_len);
    size_t it = n; /* leaves.iter().rev() */
    for (size_t len = 1; len <= max_len; len++) {
        for (uint16_t c = 0; c < num_codes[len]; c++) {
            if (it == 0) die("leaf iterator exhausted");
            it--;
            lengths[leaves[it].symbol] = (uint8_t)len;
        }
    }
}
void dfo_huffman_lengths(const uint16_t *freqs, size_t n, size_t max_len, uint8_t *lens) {
    in_place_lengths(freqs, n, max_len, lens, n);
}

/* ------------------------------------------------------------------ huffman_lengths.rs */
static const uint8_t HUFFMAN_LENGTH_ORDER[19] = {16, 17, 18, 0, 8,  7, 9,  6, 10, 5,
                                                 11, 4,  12, 3, 13, 2, 14, 1, 15}; /* :27-29 */

/* huffman_lengths.rs:44-47 remove_trailing_zeroes (u16 and u8 variants) */
static size_t trimmed_len_u16(const uint16_t *a, size_t n, size_t min_length) {
    size_t z = 0;
    while (z < n && a[n - 1 - z] == 0) z++;
    return (n - z) > min_length ? (n - z) : min_length;
}
static size_t trimmed_len_u8(const uint8_t *a, size_t n, size_t min_length) {
    size_t z = 0;
    while (z < n && a[n - 1 - z] == 0) z++;
    return (n - z) > min_length ? (n - z) : min_length;
}
/* huffman_lengths.rs:50-56: NB 16 and 17 both count 3 (the reference's own cost model) */
static uint8_t extra_bits_for_huffman_length_code(uint8_t code) {
    if (code == 16 || code == 17) return 3;
    if (code == 18) return 7;
    return 0;
}
/* huffman_lengths.rs:113-124 */
uint64_t dfo_stored_padding(uint8_t pending_bits) {
    if (pending_bits > 8) die("stored_padding: pending_bits > 8");
    uint8_t free_space = (uint8_t)(8 - pending_bits);
    if (free_space >= 3) return (uint64_t)(free_space - 3);
    return (uint64_t)(8 - (3 - free_space));
}
/* huffman_lengths.rs:132-143 */
static uint64_t stored_length(uint64_t input_bytes) {
    if (input_bytes == 0) die("Underflow calculating stored block length!");
    uint64_t num_blocks = (input_bytes - 1) / MAX_STORED_BLOCK_LENGTH + 1;
    return (input_bytes + 4 * num_blocks + (num_blocks - 1)) * 8;
}

typedef struct {
    uint8_t huffman_table_lengths[19];
    size_t used_hclens;
} dynamic_block_header;
enum { BT_STORED = 0, BT_FIXED = 1, BT_DYNAMIC = 2 };

/* huffman_lengths.rs:167-287 gen_huffman_lengths */
static int gen_huffman_lengths(const uint16_t *l_freqs_full, const uint16_t *d_freqs_full,
                               uint64_t num_input_bytes, uint8_t pending_bits, uint8_t *l_lengths,
                               uint8_t *d_lengths, elvec *length_buf, dynamic_block_header *hdr) {
    if (num_input_bytes <= 4) return BT_FIXED;

    size_t nl = trimmed_len_u16(l_freqs_full, NUM_LL, 257);
    size_t nd = trimmed_len_u16(d_freqs_full, NUM_DIST, 1);

    in_place_lengths(l_freqs_full, nl, MAX_CODE_LENGTH, l_lengths, 288);
    in_place_lengths(d_freqs_full, nd, MAX_CODE_LENGTH, d_lengths, 32);

    uint16_t freqs[19];
    memset(freqs, 0, sizeof freqs);
    uint8_t chained[288 + 32];
    memcpy(chained, l_lengths, nl);
    memcpy(chained + nl, d_lengths, nd);
    encode_lengths_m(chained, nl + nd, length_buf, freqs);

    memset(hdr->huffman_table_lengths, 0, 19);
    in_place_lengths(freqs, 19, 7, hdr->huffman_table_lengths, 19);

    size_t trailing = 0;
    while (trailing < 19 && hdr->huffman_table_lengths[HUFFMAN_LENGTH_ORDER[18 - trailing]] == 0)
        trailing++;
    hdr->used_hclens = 19 - trailing;

    /* calculate_block_length (:75-107): NB the distance pass also zips with the
     * literal/length FIXED_CODE_LENGTHS table, exactly as the reference does. */
    uint64_t d_ll = 0, s_ll = 0, d_dist = 0, s_dist = 0;
    for (size_t c = 0; c < nl; c++) {
        uint64_t f = l_freqs_full[c];
        uint64_t eb = num_extra_bits_for_length_code((uint8_t)(c >= 257 ? c - 257 : 0));
        d_ll += f * ((uint64_t)l_lengths[c] + eb);
        s_ll += f * ((uint64_t)FIXED_CODE_LENGTHS[c] + eb);
    }
    for (size_t c = 0; c < nd; c++) {
        uint64_t f = d_freqs_full[c];
        uint64_t eb = num_extra_bits_for_distance_code((uint8_t)c);
        d_dist += f * ((uint64_t)d_lengths[c] + eb);
        s_dist += f * ((uint64_t)FIXED_CODE_LENGTHS[c] + eb);
    }
    /* calculate_huffman_length (:59-68) */
    uint64_t huff_table_length = 0;
    for (size_t i = 0; i < 19; i++)
        huff_table_length += (uint64_t)freqs[i] * ((uint64_t)hdr->huffman_table_lengths[i] +
                                                   extra_bits_for_huffman_length_code((uint8_t)i));

    uint64_t dynamic_length = d_ll + d_dist + huff_table_length + (uint64_t)hdr->used_hclens * 3 + 5 + 5 + 4;
    uint64_t static_length = s_ll + s_dist;
    uint64_t stored_len = stored_length(num_input_bytes) + dfo_stored_padding((uint8_t)(pending_bits % 8));

    uint64_t used = dynamic_length < static_length ? dynamic_length : static_length;
    if (stored_len < used) used = stored_len;

    if (used == static_length) return BT_FIXED;
    if (used == stored_len) return BT_STORED;
    return BT_DYNAMIC;
}

/* huffman_lengths.rs:290-369 write_huffman_lengths */
static void write_huffman_lengths(const dynamic_block_header *hdr, const huffman_table *ht,
                                  const elvec *encoded, lsb_writer *w) {
    size_t nl = trimmed_len_u8(ht->code_lengths, 288, 257);
    size_t nd = trimmed_len_u8(ht->distance_code_lengths, 32, 1);
    if (nl > NUM_LL || nd > NUM_DIST) die("write_huffman_lengths: too many lengths");
    lsb_write_bits(w, (uint16_t)(nl - 257), 5);
    lsb_write_bits(w, (uint16_t)(nd - 1), 5);
    size_t hclen = hdr->used_hclens >= 4 ? hdr->used_hclens - 4 : 0;
    lsb_write_bits(w, (uint16_t)hclen, 4);
    for (size_t i = 0; i < hdr->used_hclens; i++)
        lsb_write_bits(w, hdr->huffman_table_lengths[HUFFMAN_LENGTH_ORDER[i]], 3);
    uint16_t codes[19];
    memset(codes, 0, sizeof codes);
    create_codes_in_place(codes, hdr->huffman_table_lengths, 19);
    for (size_t i = 0; i < encoded->len; i++) {
        encoded_length e = encoded->p[i];
        if (e.sym <= 15) {
            lsb_write_bits(w, codes[e.sym], hdr->huffman_table_lengths[e.sym]);
        } else if (e.sym == EL_COPY_PREVIOUS) {
            lsb_write_bits(w, codes[16], hdr->huffman_table_lengths[16]);
            lsb_write_bits(w, (uint16_t)(e.arg - 3), 2);
        } else if (e.sym == EL_REPEAT_ZERO_3) {
            lsb_write_bits(w, codes[17], hdr->huffman_table_lengths[17]);
            lsb_write_bits(w, (uint16_t)(e.arg - 3), 3);
        } else {
            lsb_write_bits(w, codes[18], hdr->huffman_table_lengths[18]);
            lsb_write_bits(w, (uint16_t)(e.arg - 11), 7);
        }
    }
}

/* ------------------------------------------------------------------ chained_hash_table.rs */
typedef struct {
    uint16_t current_hash;
    uint16_t head[WINDOW_SIZE];
    uint16_t prev[WINDOW_SIZE];
} hash_table;

/* chained_hash_table.rs:55-62 */
static uint16_t update_hash(uint16_t h, uint8_t b) {
    return (uint16_t)(((uint16_t)(h << HASH_SHIFT) ^ (uint16_t)b) & HASH_MASK);
}
/* chained_hash_table.rs:34-51 / :92-103 */
static void ht_reset(hash_table *t) {
    t->current_hash = 0;
    for (unsigned n = 0; n < WINDOW_SIZE; n++) t->head[n] = t->prev[n] = (uint16_t)n;
}
/* chained_hash_table.rs:148-158 */
static void ht_add_with_hash(hash_table *t, size_t position, uint16_t hash) {
    t->prev[position & WINDOW_MASK] = t->head[hash];
    t->head[hash] = (uint16_t)position;
}
/* chained_hash_table.rs:118-139 */
static void ht_add_hash_value(hash_table *t, size_t position, uint8_t value) {
    uint16_t nh = update_hash(t->current_hash, value);
    ht_add_with_hash(t, position, nh);
    t->current_hash = nh;
}
/* chained_hash_table.rs:197-219 */
static void ht_slide(hash_table *t, size_t bytes) {
    uint16_t b = (uint16_t)bytes;
    for (unsigned n = 0; n < WINDOW_SIZE; n++) t->head[n] = t->head[n] >= b ? (uint16_t)(t->head[n] - b) : (uint16_t)n;
    for (unsigned n = 0; n < WINDOW_SIZE; n++) t->prev[n] = t->prev[n] >= b ? (uint16_t)(t->prev[n] - b) : (uint16_t)n;
}

void dfo_hash_table_filled(const uint8_t *data, size_t len, uint16_t *head, uint16_t *prev) {
    /* chained_hash_table.rs:222-229 filled_hash_table */
    hash_table *t = (hash_table *)malloc(sizeof(hash_table));
    ht_reset(t);
    if (len >= 2) {
        t->current_hash = update_hash(t->current_hash, data[0]);
        t->current_hash = update_hash(t->current_hash, data[1]);
        for (size_t n = 0; n + 2 < len; n++) ht_add_hash_value(t, n, data[n + 2]);
    }
    memcpy(head, t->head, sizeof t->head);
    memcpy(prev, t->prev, sizeof t->prev);
    free(t);
}

/* ------------------------------------------------------------------ matching.rs */
/* matching.rs:13-73 get_match_length (naive variant, the one compiled) */
size_t dfo_get_match_length(const uint8_t *data, size_t len, size_t cur, size_t chk) {
    size_t n = 0;
    while (n < MAX_MATCH && cur + n < len && chk + n < len && data[cur + n] == data[chk + n]) n++;
    return n;
}
/* matching.rs:87-166 longest_match */
static void longest_match(const uint8_t *data, size_t len, const hash_table *t, size_t position,
                          size_t prev_length, uint16_t max_hash_checks, size_t *out_len,
                          size_t *out_dist) {
    *out_len = 0;
    *out_dist = 0;
    if (prev_length >= MAX_MATCH || position + prev_length >= len) return;
    size_t limit = position > WINDOW_SIZE ? position - WINDOW_SIZE : 0;
    if (prev_length < 1) prev_length = 1;
    size_t max_length = len - position < MAX_MATCH ? len - position : MAX_MATCH;
    size_t current_head = position;
    size_t best_length = prev_length, best_distance = 0;
    for (uint16_t i = 0; i < max_hash_checks; i++) {
        size_t prev_head = current_head;
        current_head = t->prev[current_head & WINDOW_MASK];
        if (current_head >= prev_head || current_head < limit) break;
        if (data[position + best_length - 1] == data[current_head + best_length - 1] &&
            data[position + best_length] == data[current_head + best_length]) {
            size_t length = dfo_get_match_length(data, len, position, current_head);
            if (length > best_length) {
                best_length = length;
                best_distance = position - current_head;
                if (length == max_length) break;
            }
        }
    }
    if (best_length > prev_length) {
        *out_len = best_length;
        *out_dist = best_distance;
    }
}

int dfo_longest_match_filled(const uint8_t *data, size_t len, size_t fill_len, size_t position,
                             size_t prev_length, uint16_t max_hash_checks, size_t *out_len,
                             size_t *out_dist) {
    hash_table *t = (hash_table *)malloc(sizeof(hash_table));
    ht_reset(t);
    if (fill_len >= 2) {
        t->current_hash = update_hash(t->current_hash, data[0]);
        t->current_hash = update_hash(t->current_hash, data[1]);
        for (size_t n = 0; n + 2 < fill_len; n++) ht_add_hash_value(t, n, data[n + 2]);
    }
    longest_match(data, len, t, position, prev_length, max_hash_checks, out_len, out_dist);
    free(t);
    return 0;
}

/* ------------------------------------------------------------------ output_writer.rs */
typedef struct {
    dfo_token *buffer;
    size_t len;
    uint16_t frequencies[NUM_LL];
    uint16_t distance_frequencies[NUM_DIST];
} dynamic_writer;

/* output_writer.rs:100-117 */
static void dw_clear(dynamic_writer *w) {
    memset(w->frequencies, 0, sizeof w->frequencies);
    memset(w->distance_frequencies, 0, sizeof w->distance_frequencies);
    w->frequencies[END_OF_BLOCK] = 1;
    w->len = 0;
}
/* output_writer.rs:47-52; returns 1 when BufferStatus::Full */
static int dw_write_literal(dynamic_writer *w, uint8_t lit) {
    dfo_token t = {0, lit, 0};
    w->buffer[w->len++] = t;
    w->frequencies[lit]++;
    return w->len >= MAX_BUFFER_LENGTH;
}
/* output_writer.rs:55-65 */
static int dw_write_length_distance(dynamic_writer *w, uint16_t length, uint16_t distance) {
    dfo_token t = {distance, (uint8_t)(length - MIN_MATCH), 0};
    w->buffer[w->len++] = t;
    w->frequencies[get_length_code(length)]++;
    w->distance_frequencies[get_distance_code(distance)]++;
    return w->len >= MAX_BUFFER_LENGTH;
}
/* output_writer.rs:90-98 */
static int dw_write_length_rle(dynamic_writer *w, uint16_t length) {
    dfo_token t = {1, (uint8_t)(length - MIN_MATCH), 0};
    w->buffer[w->len++] = t;
    w->frequencies[get_length_code(length)]++;
    w->distance_frequencies[0]++;
    return w->len >= MAX_BUFFER_LENGTH;
}

/* ------------------------------------------------------------------ lz77.rs */
typedef struct { /* lz77.rs:162-173 ChunkState */
    uint16_t current_length, current_distance;
    uint8_t prev_byte, cur_byte;
    int add;
} chunk_state;

typedef struct { /* lz77.rs:49-74 LZ77State */
    hash_table hash_table;
    int is_first_window, is_last_block;
    size_t overlap;
    uint64_t current_block_input_bytes;
    uint16_t max_hash_checks, lazy_if_less_than;
    int matching_type;
    chunk_state match_state;
    size_t bytes_to_hash;
    int was_synced;
} lz77_state;

/* lz77.rs:98-106 reset (is also the state after new(), :76-96) */
static void lz77_reset(lz77_state *s) {
    ht_reset(&s->hash_table);
    s->is_first_window = 1;
    s->is_last_block = 0;
    s->overlap = 0;
    s->current_block_input_bytes = 0;
    memset(&s->match_state, 0, sizeof s->match_state);
    s->bytes_to_hash = 0;
}

typedef struct {
    size_t overlap;
    int full;        /* ProcessStatus::BufferFull */
    size_t full_pos; /* its argument */
} process_result;

/* the two iterators of lz77.rs:281-303 create_iterators, as indices */
typedef struct {
    size_t ipos, end; /* insert_it: next position, exclusive end */
    size_t hpos, hend; /* hash_it: next hash byte index, data.len() */
} chunk_iters;

static chunk_iters create_iterators(size_t data_len, size_t start, size_t range_end) {
    chunk_iters it;
    it.end = data_len < range_end ? data_len : range_end;
    if (start > it.end) die("create_iterators: start > end");
    it.ipos = start;
    it.hpos = (data_len - start > 2) ? start + 2 : data_len;
    it.hend = data_len;
    return it;
}
/* lz77.rs:236-256 add_to_hash_table */
static void add_to_hash_table(size_t bytes_to_add, chunk_iters *it, const uint8_t *data, hash_table *t) {
    uint16_t hash = t->current_hash;
    size_t hash_taken = 0;
    for (size_t k = 0; k < bytes_to_add; k++) {
        if (it->ipos >= it->end) break; /* taker exhausted */
        size_t ipos = it->ipos++;
        if (hash_taken < bytes_to_add && it->hpos < it->hend) {
            uint8_t hb = data[it->hpos++];
            hash_taken++;
            hash = update_hash(hash, hb);
            ht_add_with_hash(t, ipos, hash);
        }
    }
    t->current_hash = hash;
}
/* lz77.rs:274-278 */
static int match_too_far(size_t match_len, size_t match_dist) {
    return match_len == MIN_MATCH && match_dist > 8 * 1024;
}

/* lz77.rs:305-486 process_chunk_lazy */
static process_result process_chunk_lazy(const uint8_t *data, size_t data_len, size_t start,
                                         size_t range_end, chunk_state *state, hash_table *t,
                                         dynamic_writer *writer, uint16_t max_hash_checks,
                                         size_t lazy_if_less_than) {
    process_result r = {0, 0, 0};
    chunk_iters it = create_iterators(data_len, start, range_end);
    size_t end = it.end;
    uint16_t prev_length = state->current_length;
    uint16_t prev_distance = state->current_distance;
    state->current_length = 0;
    state->current_distance = 0;
    size_t overlap = 0;
    int ignore_next = (size_t)prev_length >= lazy_if_less_than;
    state->prev_byte = state->cur_byte;

    while (it.ipos < it.end) {
        size_t position = it.ipos++;
        uint8_t b = data[position];
        state->cur_byte = b;
        if (it.hpos < it.hend) {
            uint8_t hash_byte = data[it.hpos++];
            ht_add_hash_value(t, position, hash_byte);
            if (!ignore_next) {
                uint16_t checks = prev_length >= 32 ? (uint16_t)(max_hash_checks >> 2) : max_hash_checks;
                size_t match_len, match_dist;
                longest_match(data, data_len, t, position, prev_length, checks, &match_len, &match_dist);
                if (match_too_far(match_len, match_dist)) match_len = 0;
                if (match_len >= lazy_if_less_than) ignore_next = 1;
                state->current_length = (uint16_t)match_len;
                state->current_distance = (uint16_t)match_dist;
            } else {
                state->current_length = 0;
                state->current_distance = 0;
                ignore_next = 0;
            }
            if (prev_length >= state->current_length && prev_length >= MIN_MATCH) {
                int full = dw_write_length_distance(writer, prev_length, prev_distance);
                size_t bytes_to_add = (size_t)prev_length - 2;
                add_to_hash_table(bytes_to_add, &it, data, t);
                if (position + prev_length > end) overlap = position + prev_length - end - 1;
                state->add = 0;
                state->current_length = 0;
                state->current_distance = 0;
                if (full) {
                    r.overlap = overlap;
                    r.full = 1;
                    r.full_pos = position + prev_length - 1;
                    return r;
                }
                ignore_next = 0;
            } else if (state->add) {
                if (dw_write_literal(writer, state->prev_byte)) {
                    r.overlap = 0;
                    r.full = 1;
                    r.full_pos = position + 1;
                    return r;
                }
            } else {
                state->add = 1;
            }
            prev_length = state->current_length;
            prev_distance = state->current_distance;
            state->prev_byte = b;
        } else {
            if (prev_length >= MIN_MATCH) {
                int full = dw_write_length_distance(writer, prev_length, prev_distance);
                state->current_length = 0;
                state->current_distance = 0;
                state->add = 0;
                size_t o = position + prev_length;
                o = o > end ? o - end : 0;
                o = o > 1 ? o - 1 : 0;
                overlap = o;
                r.overlap = overlap;
                if (full) {
                    r.full = 1;
                    r.full_pos = end;
                }
                return r;
            }
            if (state->add) {
                state->add = 0;
                if (dw_write_literal(writer, state->prev_byte)) {
                    r.overlap = 0;
                    r.full = 1;
                    r.full_pos = position;
                    return r;
                }
            }
            if (dw_write_literal(writer, b)) {
                r.overlap = 0;
                r.full = 1;
                r.full_pos = position + 1;
                return r;
            }
        }
    }
    r.overlap = overlap;
    return r;
}

/* lz77.rs:488-547 process_chunk_greedy */
static process_result process_chunk_greedy(const uint8_t *data, size_t data_len, size_t start,
                                           size_t range_end, hash_table *t, dynamic_writer *writer,
                                           uint16_t max_hash_checks) {
    process_result r = {0, 0, 0};
    chunk_iters it = create_iterators(data_len, start, range_end);
    size_t end = it.end;
    size_t overlap = 0;
    while (it.ipos < it.end) {
        size_t position = it.ipos++;
        uint8_t b = data[position];
        if (it.hpos < it.hend) {
            uint8_t hash_byte = data[it.hpos++];
            ht_add_hash_value(t, position, hash_byte);
            size_t match_len, match_dist;
            longest_match(data, data_len, t, position, 0, max_hash_checks, &match_len, &match_dist);
            if (match_len >= MIN_MATCH && !match_too_far(match_len, match_dist)) {
                int full = dw_write_length_distance(writer, (uint16_t)match_len, (uint16_t)match_dist);
                add_to_hash_table(match_len - 1, &it, data, t);
                if (position + match_len > end) overlap = position + match_len - end;
                if (full) {
                    r.overlap = overlap;
                    r.full = 1;
                    r.full_pos = position + match_len;
                    return r;
                }
            } else {
                if (dw_write_literal(writer, b)) {
                    r.overlap = 0;
                    r.full = 1;
                    r.full_pos = position + 1;
                    return r;
                }
            }
        } else {
            if (dw_write_literal(writer, b)) {
                r.overlap = 0;
                r.full = 1;
                r.full_pos = position + 1;
                return r;
            }
        }
    }
    r.overlap = overlap;
    return r;
}

/* rle.rs:23-71 process_chunk_greedy_rle */
static process_result process_chunk_greedy_rle(const uint8_t *data, size_t data_len, size_t range_start,
                                               size_t range_end, dynamic_writer *writer) {
    process_result r = {0, 0, 0};
    if (data_len == 0) return r;
    size_t end = data_len < range_end ? data_len : range_end;
    size_t start = range_start > 1 ? range_start : 1;
    uint8_t prev = data[start - 1];
    size_t cstart = start < end ? start : end;
    size_t overlap = 0;
    if (range_start == 0) {
        if (dw_write_literal(writer, data[0])) {
            r.full = 1;
            r.full_pos = 1;
            return r;
        }
    }
    size_t n = 0, cn = end - cstart; /* enumerate() over current_chunk */
    while (n < cn) {
        uint8_t b = data[cstart + n];
        size_t position = n + start;
        n++;
        size_t match_len = 0;
        if (prev == b) {
            /* rle.rs:13-18 get_match_length_rle(&data[position..], prev) */
            while (match_len < MAX_MATCH && position + match_len < data_len && data[position + match_len] == prev)
                match_len++;
        }
        if (match_len >= MIN_MATCH) {
            if (position + match_len > end) overlap = position + match_len - end;
            int full = dw_write_length_rle(writer, (uint16_t)match_len);
            if (full) {
                r.overlap = overlap;
                r.full = 1;
                r.full_pos = position + match_len;
                return r;
            }
            n += match_len - 1; /* insert_it.nth(match_len - 2) consumes match_len-1 items */
        } else {
            if (dw_write_literal(writer, b)) {
                r.overlap = 0;
                r.full = 1;
                r.full_pos = position + 1;
                return r;
            }
        }
        prev = b;
    }
    r.overlap = overlap;
    return r;
}

/* lz77.rs:192-232 process_chunk */
static process_result process_chunk(const uint8_t *data, size_t data_len, size_t start, size_t range_end,
                                    chunk_state *ms, hash_table *t, dynamic_writer *writer,
                                    uint16_t max_hash_checks, size_t lazy_if_less_than, int matching_type) {
    if (matching_type == 0) return process_chunk_greedy(data, data_len, start, range_end, t, writer, max_hash_checks);
    if (max_hash_checks > 0)
        return process_chunk_lazy(data, data_len, start, range_end, ms, t, writer, max_hash_checks, lazy_if_less_than);
    return process_chunk_greedy_rle(data, data_len, start, range_end, writer);
}

/* ------------------------------------------------------------------ input_buffer.rs */
typedef struct {
    uint8_t buffer[BUFFER_SIZE];
    size_t len;
} input_buffer;

typedef struct { /* Option<&[u8]> */
    int some;
    const uint8_t *p;
    size_t n;
} opt_slice;

/* input_buffer.rs:31-46 add_data */
static opt_slice ib_add_data(input_buffer *b, const uint8_t *data, size_t n) {
    opt_slice r = {0, NULL, 0};
    if (b->len + n > BUFFER_SIZE) {
        size_t space_left = BUFFER_SIZE - b->len;
        memcpy(b->buffer + b->len, data, space_left);
        b->len += space_left;
        r.some = 1;
        r.p = data + space_left;
        r.n = n - space_left;
    } else {
        if (n) memcpy(b->buffer + b->len, data, n);
        b->len += n;
    }
    return r;
}
/* input_buffer.rs:56-91 slide */
static opt_slice ib_slide(input_buffer *b, const uint8_t *data, size_t n) {
    opt_slice r = {0, NULL, 0};
    if (!(b->len > WINDOW_SIZE * 2)) die("InputBuffer::slide on a non-full buffer");
    size_t upper_total = b->len - WINDOW_SIZE; /* upper window + lookahead */
    memcpy(b->buffer, b->buffer + WINDOW_SIZE, WINDOW_SIZE);
    size_t lookahead_len = upper_total - WINDOW_SIZE;
    memmove(b->buffer + WINDOW_SIZE, b->buffer + 2 * WINDOW_SIZE, lookahead_len);
    size_t upper_len = upper_total - lookahead_len; /* == WINDOW_SIZE */
    size_t end = n < upper_len ? n : upper_len;
    if (end) memcpy(b->buffer + WINDOW_SIZE + lookahead_len, data, end);
    b->len = WINDOW_SIZE + lookahead_len + end;
    if (n > upper_len) {
        r.some = 1;
        r.p = data + end;
        r.n = n - end;
    }
    return r;
}

/* ------------------------------------------------------------------ lz77.rs:581-770 */
enum { LZ_NEED_INPUT = 0, LZ_END_BLOCK = 1, LZ_FINISHED = 2 };

static void lz77_compress_block(const uint8_t *data, size_t data_n, lz77_state *state, input_buffer *buffer,
                                dynamic_writer *writer, int flush, size_t *consumed, int *status_out,
                                size_t *position_out) {
    const size_t window_size = WINDOW_SIZE;
    int finish = (flush == DFO_FLUSH_FINISH || flush == DFO_FLUSH_SYNC);
    int sync = (flush == DFO_FLUSH_SYNC);
    size_t current_position = 0;
    int status = LZ_END_BLOCK;
    int add_initial = 1;

    if (state->was_synced) {
        if (buffer->len > 2) {
            size_t pos_add = buffer->len - 2;
            for (size_t n = 0; n < 2 && n < data_n; n++)
                ht_add_hash_value(&state->hash_table, n + pos_add, data[n]);
            add_initial = 0;
        }
        state->was_synced = 0;
    }

    opt_slice remaining = ib_add_data(buffer, data, data_n);

    for (;;) {
        size_t pending_previous = state->match_state.add ? 1 : 0;
        if (!(writer->len <= window_size * 2)) die("writer.buffer_length() > 2 * window");
        if (buffer->len >= window_size * 2 + MAX_MATCH || finish) {
            if (state->is_first_window) {
                if (buffer->len >= 2 && add_initial && state->current_block_input_bytes == 0) {
                    state->hash_table.current_hash = update_hash(state->hash_table.current_hash, buffer->buffer[0]);
                    state->hash_table.current_hash = update_hash(state->hash_table.current_hash, buffer->buffer[1]);
                    add_initial = 0;
                }
            } else if (buffer->len >= window_size + 2) {
                size_t avail = buffer->len - (window_size + 2);
                for (size_t n = 0; n < avail && n < state->bytes_to_hash; n++)
                    ht_add_hash_value(&state->hash_table, window_size + n, buffer->buffer[window_size + 2 + n]);
                state->bytes_to_hash = 0;
            }

            size_t window_start = state->is_first_window ? 0 : window_size;
            size_t start = state->overlap + window_start;
            size_t end = window_size + window_start < buffer->len ? window_size + window_start : buffer->len;

            process_result pr = process_chunk(buffer->buffer, buffer->len, start, end, &state->match_state,
                                              &state->hash_table, writer, state->max_hash_checks,
                                              (size_t)state->lazy_if_less_than, state->matching_type);
            size_t overlap = pr.overlap;
            state->bytes_to_hash = overlap;

            if (pr.full) {
                size_t written = pr.full_pos;
                size_t pend_now = state->match_state.add ? 1 : 0;
                state->current_block_input_bytes += (uint64_t)(written - start + pending_previous - pend_now);
                if (overlap > 0) {
                    if (!state->is_first_window) {
                        if (state->max_hash_checks > 0) ht_slide(&state->hash_table, window_size);
                        remaining = ib_slide(buffer, remaining.some ? remaining.p : NULL, remaining.some ? remaining.n : 0);
                    } else {
                        state->is_first_window = 0;
                    }
                    state->overlap = overlap;
                } else {
                    state->overlap = written - window_start;
                }
                current_position = written - pend_now;
                break;
            }

            {
                size_t pend_now = state->match_state.add ? 1 : 0;
                state->current_block_input_bytes += (uint64_t)(end - start + overlap + pending_previous - pend_now);
            }
            state->overlap = overlap;

            if ((state->is_first_window || !remaining.some) && finish && end >= buffer->len) {
                if (state->is_first_window) {
                    current_position = end - (state->match_state.add ? 1 : 0);
                } else {
                    current_position = buffer->len;
                }
                if (!sync) {
                    state->is_last_block = 1;
                    state->is_first_window = 0;
                } else {
                    state->overlap = state->is_first_window ? end : buffer->len - window_size;
                    state->was_synced = 1;
                }
                status = LZ_FINISHED;
                break;
            } else if (state->is_first_window) {
                state->is_first_window = 0;
            } else {
                if (state->max_hash_checks > 0) ht_slide(&state->hash_table, window_size);
                remaining = ib_slide(buffer, remaining.some ? remaining.p : NULL, remaining.some ? remaining.n : 0);
            }
        } else {
            status = LZ_NEED_INPUT;
            break;
        }
    }
    *consumed = data_n - (remaining.some ? remaining.n : 0);
    *status_out = status;
    *position_out = current_position;
}

/* ------------------------------------------------------------------ deflate_state.rs */
typedef struct {
    input_buffer input_buffer;
    lz77_state lz77_state;
    encoder_state encoder_state;
    dynamic_writer lz77_writer;
    elvec length_buf; /* LengthBuffers::length_buf */
    uint64_t bytes_written;
    bytevec inner; /* the wrapped writer W, here always a Vec<u8> */
    size_t output_buf_pos;
    int flush_mode;
    int needs_flush;
} deflate_state;

/* deflate_state.rs:100-124 new */
static deflate_state *ds_new(const dfo_options *opt) {
    init_tables();
    deflate_state *s = (deflate_state *)calloc(1, sizeof(deflate_state));
    if (!s) die("out of memory");
    s->lz77_state.max_hash_checks = opt->max_hash_checks;
    s->lz77_state.lazy_if_less_than =
        (uint16_t)(opt->lazy_if_less_than < LAZY_CLAMP ? opt->lazy_if_less_than : LAZY_CLAMP);
    s->lz77_state.matching_type = opt->matching_type;
    s->lz77_state.was_synced = 0;
    lz77_reset(&s->lz77_state);
    s->lz77_writer.buffer = (dfo_token *)malloc((MAX_BUFFER_LENGTH + 8) * sizeof(dfo_token));
    dw_clear(&s->lz77_writer);
    s->flush_mode = DFO_FLUSH_NONE;
    return s;
}
static void ds_free(deflate_state *s) {
    if (!s) return;
    free(s->lz77_writer.buffer);
    free(s->length_buf.p);
    free(s->encoder_state.writer.w.p);
    free(s->inner.p);
    free(s);
}
/* the sink is a Vec: Write::write takes everything */
static size_t sink_write(deflate_state *s, const uint8_t *p, size_t n) {
    bv_extend(&s->inner, p, n);
    return n;
}

/* compress.rs:80-302 compress_data_dynamic_n. Returns bytes consumed, or -1 for
 * Err(ErrorKind::Interrupted) ("Internal buffer full"). */
static long compress_data_dynamic_n(const uint8_t *input, size_t input_n, deflate_state *ds, int flush) {
    size_t bytes_written = 0;
    const uint8_t *slice = input;
    size_t slice_n = input_n;
    bytevec *outbuf = &ds->encoder_state.writer.w;

    while (!ds->needs_flush) {
        size_t output_buf_len = outbuf->len;
        size_t output_buf_pos = ds->output_buf_pos;
        if (output_buf_len > LARGEST_OUTPUT_BUF_SIZE) {
            size_t written = sink_write(ds, outbuf->p + output_buf_pos, output_buf_len - output_buf_pos);
            if (written < output_buf_len - output_buf_pos) {
                ds->output_buf_pos += written;
            } else {
                ds->needs_flush = 0;
                ds->output_buf_pos = 0;
                outbuf->len = 0;
            }
            if (bytes_written == 0) return -1;
            return (long)bytes_written;
        }
        if (ds->lz77_state.is_last_block) break;

        size_t written;
        int status;
        size_t position;
        lz77_compress_block(slice, slice_n, &ds->lz77_state, &ds->input_buffer, &ds->lz77_writer, flush,
                            &written, &status, &position);
        bytes_written += written;
        ds->bytes_written += written;
        if (status == LZ_NEED_INPUT) return (long)bytes_written;
        slice += written;
        slice_n -= written;

        int last_block = ds->lz77_state.is_last_block;
        uint64_t current_block_input_bytes = ds->lz77_state.current_block_input_bytes;
        uint8_t partial_bits = ds->encoder_state.writer.bits;

        dynamic_block_header header;
        int res = gen_huffman_lengths(ds->lz77_writer.frequencies, ds->lz77_writer.distance_frequencies,
                                      current_block_input_bytes, partial_bits,
                                      ds->encoder_state.huffman_table.code_lengths,
                                      ds->encoder_state.huffman_table.distance_code_lengths, &ds->length_buf,
                                      &header);
        if (res == BT_DYNAMIC) {
            es_write_start_of_block(&ds->encoder_state, 0, last_block);
            write_huffman_lengths(&header, &ds->encoder_state.huffman_table, &ds->length_buf,
                                  &ds->encoder_state.writer);
            ht_update_from_lengths(&ds->encoder_state.huffman_table);
            for (size_t i = 0; i < ds->lz77_writer.len; i++) es_write_lzvalue(&ds->encoder_state, ds->lz77_writer.buffer[i]);
            es_write_end_of_block(&ds->encoder_state);
        } else if (res == BT_FIXED) {
            es_write_start_of_block(&ds->encoder_state, 1, last_block);
            ht_set_to_fixed(&ds->encoder_state.huffman_table);
            for (size_t i = 0; i < ds->lz77_writer.len; i++) es_write_lzvalue(&ds->encoder_state, ds->lz77_writer.buffer[i]);
            es_write_end_of_block(&ds->encoder_state);
        } else {
            if (position < current_block_input_bytes)
                die("Error! Trying to output a stored block with forgotten data!");
            size_t start_pos = position - (size_t)current_block_input_bytes;
            write_stored_block(ds->input_buffer.buffer + start_pos, position - start_pos,
                               &ds->encoder_state.writer, flush == DFO_FLUSH_FINISH && last_block);
        }
        dw_clear(&ds->lz77_writer);
        ds->lz77_state.current_block_input_bytes = 0;

        if (status == LZ_FINISHED) {
            if (flush == DFO_FLUSH_SYNC) {
                write_stored_block(NULL, 0, &ds->encoder_state.writer, 0);
                ds->needs_flush = 1;
            } else if (!ds->lz77_state.is_last_block) {
                ht_set_to_fixed(&ds->encoder_state.huffman_table);
                es_write_start_of_block(&ds->encoder_state, 1, 1);
                es_write_end_of_block(&ds->encoder_state);
            }
            break;
        }
    }

    lsb_flush_raw(&ds->encoder_state.writer);
    size_t output_buf_pos = ds->output_buf_pos;
    size_t written_to_writer = sink_write(ds, outbuf->p + output_buf_pos, outbuf->len - output_buf_pos);
    if (written_to_writer < outbuf->len - output_buf_pos) {
        ds->output_buf_pos += written_to_writer;
    } else {
        ds->output_buf_pos = 0;
        outbuf->len = 0;
        ds->needs_flush = 0;
    }
    return (long)bytes_written;
}

/* writer.rs:15-58 compress_until_done */
static void compress_until_done(const uint8_t *input, size_t n, deflate_state *ds, int flush_mode) {
    if (flush_mode == DFO_FLUSH_NONE) die("compress_until_done with Flush::None");
    for (;;) {
        long r = compress_data_dynamic_n(input, n, ds, flush_mode);
        if (r == 0) {
            if (ds->encoder_state.writer.w.len == 0) break;
            n = 0;
        } else if (r > 0) {
            if ((size_t)r < n) {
                input += r;
                n -= (size_t)r;
            } else {
                n = 0;
            }
        } /* r < 0: Interrupted -> retry */
    }
}

/* deflate_state.rs:133-152 reset (sink kept: the caller swaps it) */
static void ds_reset(deflate_state *ds) {
    lsb_flush_raw(&ds->encoder_state.writer);
    sink_write(ds, ds->encoder_state.writer.w.p, ds->encoder_state.writer.w.len);
    ds->encoder_state.writer.w.len = 0;
    ds->input_buffer.len = 0;
    dw_clear(&ds->lz77_writer);
    lz77_reset(&ds->lz77_state);
    ds->bytes_written = 0;
    ds->output_buf_pos = 0;
    ds->flush_mode = DFO_FLUSH_NONE;
    ds->needs_flush = 0;
}

/* ------------------------------------------------------------------ checksums */
/* adler32 crate 1.2.0 == RFC 1950 section 8.2 */
uint32_t dfo_adler32(uint32_t adler, const uint8_t *buf, size_t n) {
    uint32_t a = adler & 0xffff, b = (adler >> 16) & 0xffff;
    while (n > 0) {
        size_t k = n < 5552 ? n : 5552;
        for (size_t i = 0; i < k; i++) {
            a += buf[i];
            b += a;
        }
        a %= 65521;
        b %= 65521;
        buf += k;
        n -= k;
    }
    return (b << 16) | a;
}
/* gzip-header 1.0 Crc == RFC 1952 section 8 */
uint32_t dfo_crc32(uint32_t crc, const uint8_t *buf, size_t n) {
    static uint32_t table[256];
    static int ready = 0;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = 1;
    }
    uint32_t c = crc ^ 0xffffffffu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ buf[i]) & 0xff] ^ (c >> 8);
    return c ^ 0xffffffffu;
}

/* zlib.rs:40-62: CMF 0x78, FLG from CompressionLevel::Default (2 << 6) + FCHECK */
static void zlib_header(uint8_t out[2]) {
    uint8_t cmf = 8 | (7 << 4);
    uint8_t flg = 2 << 6;
    unsigned rem = ((unsigned)cmf * 256 + flg) % 31;
    flg = (uint8_t)((flg & 0xe0) + (31 - rem));
    out[0] = cmf;
    out[1] = flg;
}
/* gzip-header 1.0 GzBuilder::new().into_header(): ID1 ID2 CM FLG MTIME(4) XFL OS.
 * Layout is parity-unpinned by the reference; any RFC 1952 header is valid. */
static const uint8_t GZIP_DEFAULT_HEADER[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};

/* ------------------------------------------------------------------ one-shot API */
/* lib.rs:110-122 compress_data_dynamic + :137 / :182 / :242 */
int dfo_compress(const uint8_t *in, size_t n, const dfo_options *opt, int wrap, uint8_t **out, size_t *out_len) {
    deflate_state *ds = ds_new(opt);
    uint8_t hdr[2];
    if (wrap == DFO_ZLIB) {
        zlib_header(hdr);
        bv_extend(&ds->inner, hdr, 2);
    } else if (wrap == DFO_GZIP) {
        bv_extend(&ds->inner, GZIP_DEFAULT_HEADER, 10);
    }
    compress_until_done(in, n, ds, DFO_FLUSH_FINISH);
    if (wrap == DFO_ZLIB) {
        uint32_t h = dfo_adler32(1, in, n);
        uint8_t t[4] = {(uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h};
        bv_extend(&ds->inner, t, 4);
    } else if (wrap == DFO_GZIP) {
        uint32_t c = dfo_crc32(0, in, n);
        uint32_t amt = (uint32_t)n;
        uint8_t t[8] = {(uint8_t)c,   (uint8_t)(c >> 8),   (uint8_t)(c >> 16),   (uint8_t)(c >> 24),
                        (uint8_t)amt, (uint8_t)(amt >> 8), (uint8_t)(amt >> 16), (uint8_t)(amt >> 24)};
        bv_extend(&ds->inner, t, 8);
    }
    *out = ds->inner.p;
    *out_len = ds->inner.len;
    ds->inner.p = NULL;
    ds_free(ds);
    return 0;
}
void dfo_free(void *p) { free(p); }

/* ------------------------------------------------------------------ streaming API */
struct dfo_stream {
    deflate_state *ds;
    int wrap;
    int header_written;
    uint32_t adler;
    uint32_t crc;
    uint32_t amt;
};

dfo_stream *dfo_stream_new(const dfo_options *opt, int wrap) {
    dfo_stream *s = (dfo_stream *)calloc(1, sizeof *s);
    s->ds = ds_new(opt);
    s->wrap = wrap;
    s->adler = 1;
    return s;
}
/* writer.rs:226-232 / gzip check_write_header: the header goes into the output buffer */
static void stream_check_write_header(dfo_stream *s) {
    if (s->header_written) return;
    if (s->wrap == DFO_ZLIB) {
        uint8_t h[2];
        zlib_header(h);
        bv_extend(&s->ds->encoder_state.writer.w, h, 2);
    } else if (s->wrap == DFO_GZIP) {
        bv_extend(&s->ds->encoder_state.writer.w, GZIP_DEFAULT_HEADER, 10);
    }
    s->header_written = 1;
}
/* one Write::write call (writer.rs:124-127, 254-267); returns consumed or -1 */
static long stream_write_once(dfo_stream *s, const uint8_t *buf, size_t n) {
    stream_check_write_header(s);
    long r = compress_data_dynamic_n(buf, n, s->ds, s->ds->flush_mode);
    if (r >= 0) {
        size_t upd = (r == 0) ? n : (size_t)r;
        if (s->wrap == DFO_ZLIB) s->adler = dfo_adler32(s->adler, buf, upd);
        if (s->wrap == DFO_GZIP) {
            s->crc = dfo_crc32(s->crc, buf, upd);
            s->amt += (uint32_t)upd;
        }
    }
    return r;
}
int dfo_stream_write(dfo_stream *s, const uint8_t *buf, size_t n) {
    /* io::Write::write_all: Interrupted is retried, Ok(0) on a non-empty buffer is WriteZero */
    while (n > 0) {
        long r = stream_write_once(s, buf, n);
        if (r < 0) continue;
        if (r == 0) return -1;
        buf += r;
        n -= (size_t)r;
    }
    return 0;
}
int dfo_stream_flush(dfo_stream *s) {
    compress_until_done(NULL, 0, s->ds, DFO_FLUSH_SYNC);
    return 0;
}
/* writer.rs:201-207 output_all */
static void stream_output_all(dfo_stream *s) {
    stream_check_write_header(s);
    compress_until_done(NULL, 0, s->ds, DFO_FLUSH_FINISH);
    if (s->wrap == DFO_ZLIB) {
        uint32_t h = s->adler;
        uint8_t t[4] = {(uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h};
        bv_extend(&s->ds->inner, t, 4);
    } else if (s->wrap == DFO_GZIP) {
        uint32_t c = s->crc, amt = s->amt;
        uint8_t t[8] = {(uint8_t)c,   (uint8_t)(c >> 8),   (uint8_t)(c >> 16),   (uint8_t)(c >> 24),
                        (uint8_t)amt, (uint8_t)(amt >> 8), (uint8_t)(amt >> 16), (uint8_t)(amt >> 24)};
        bv_extend(&s->ds->inner, t, 8);
    }
}
int dfo_stream_finish(dfo_stream *s) {
    stream_output_all(s);
    return 0;
}
int dfo_stream_reset(dfo_stream *s) {
    stream_output_all(s);
    s->header_written = 0;
    s->adler = 1;
    s->crc = 0;
    s->amt = 0;
    ds_reset(s->ds);
    return 0;
}
uint32_t dfo_stream_checksum(const dfo_stream *s) {
    if (s->wrap == DFO_ZLIB) return s->adler;
    if (s->wrap == DFO_GZIP) return s->crc;
    return 1; /* NoChecksum::current_hash, checksum.rs:26-28 */
}
const uint8_t *dfo_stream_output(const dfo_stream *s, size_t *len) {
    *len = s->ds->inner.len;
    return s->ds->inner.p;
}
void dfo_stream_clear_output(dfo_stream *s) { s->ds->inner.len = 0; }
void dfo_stream_free(dfo_stream *s) {
    if (!s) return;
    ds_free(s->ds);
    free(s);
}

/* ------------------------------------------------------------------ KAT hooks */
int dfo_lz77_tokens(const uint8_t *in, size_t n, const dfo_options *opt, dfo_token **toks, size_t *ntoks,
                    size_t **block_ends, size_t *nblocks) {
    /* lz77.rs:869-905 lz77_compress_conf */
    deflate_state *ds = ds_new(opt);
    size_t cap = n + 16, len = 0;
    dfo_token *out = (dfo_token *)malloc(cap * sizeof(dfo_token));
    size_t bcap = 16, bn = 0;
    size_t *be = (size_t *)malloc(bcap * sizeof(size_t));
    const uint8_t *slice = in;
    size_t slice_n = n;
    while (!ds->lz77_state.is_last_block) {
        size_t consumed, position;
        int status;
        lz77_compress_block(slice, slice_n, &ds->lz77_state, &ds->input_buffer, &ds->lz77_writer,
                            DFO_FLUSH_FINISH, &consumed, &status, &position);
        slice += consumed;
        slice_n -= consumed;
        if (len + ds->lz77_writer.len > cap) {
            cap = (len + ds->lz77_writer.len) * 2;
            out = (dfo_token *)realloc(out, cap * sizeof(dfo_token));
        }
        memcpy(out + len, ds->lz77_writer.buffer, ds->lz77_writer.len * sizeof(dfo_token));
        len += ds->lz77_writer.len;
        if (bn == bcap) {
            bcap *= 2;
            be = (size_t *)realloc(be, bcap * sizeof(size_t));
        }
        be[bn++] = len;
        dw_clear(&ds->lz77_writer);
        ds->lz77_state.current_block_input_bytes = 0;
    }
    ds_free(ds);
    *toks = out;
    *ntoks = len;
    if (block_ends) {
        *block_ends = be;
        *nblocks = bn;
    } else {
        free(be);
    }
    return 0;
}

int dfo_compress_fixed(const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
    /* compress.rs:43-57 compress_data_fixed, with lz77_compress defaults (lz77.rs:852-859:
     * HIGH_MAX_HASH_CHECKS 1768, HIGH_LAZY_IF_LESS_THAN 128, Lazy) */
    init_tables();
    dfo_options opt = {1768, 128, 1, 0};
    dfo_token *toks;
    size_t nt;
    dfo_lz77_tokens(in, n, &opt, &toks, &nt, NULL, NULL);
    encoder_state es;
    memset(&es, 0, sizeof es);
    ht_set_to_fixed(&es.huffman_table);
    es_write_start_of_block(&es, 1, 1);
    for (size_t i = 0; i < nt; i++) es_write_lzvalue(&es, toks[i]);
    es_write_end_of_block(&es);
    lsb_flush_raw(&es.writer);
    free(toks);
    *out = es.writer.w.p;
    *out_len = es.writer.w.len;
    return 0;
}
