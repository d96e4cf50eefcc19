// dfl_internal.h -- private interface between the C-ABI layer (dfl_api.cu) and the kernels
// (dfl_kernels.cu).  Not installed; include/deflate_b200.h is the public boundary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dfl_core.h"

namespace dfl {

// Tunables of the parse stage (see DESIGN.md "parse").
#ifndef DFL_PARSE_SEG
#define DFL_PARSE_SEG 8192
#endif
#ifndef DFL_PARSE_WARM
#define DFL_PARSE_WARM 256
#endif
constexpr uint32_t kParseSeg = DFL_PARSE_SEG;     // positions owned by one parse thread (large inputs)
constexpr uint32_t kParseWarm = DFL_PARSE_WARM;   // speculative warm-up before the segment start
#ifndef DFL_PARSE_WARM_GREEDY
#define DFL_PARSE_WARM_GREEDY 256
#endif
constexpr uint32_t kParseSegGreedy = kParseSeg, kParseWarmGreedy = DFL_PARSE_WARM_GREEDY;   // the greedy parser (large inputs)
static_assert(kParseSeg + kParseWarm + 300u + 264u < 16384u && kParseSegGreedy + kParseWarmGreedy + 564u < 16384u,
              "deferred literal tokens carry a 14-bit position relative to the segment (kTokRelMask)");
// One parse thread walks its segment sequentially, so for small inputs the segment length *is* the
// latency of the stage: shorter segments there (any length gives the same tokens, hand-offs are
// verified and repaired).  seg + warm + 264 tokens of buffer per segment.
// Large inputs, 1 GiB at Default on B200 (tools/tune_variants.sh): 8 KiB + 256 B parses in 74 ms, 8 KiB + 512 B
// in 78, 4 KiB + 512 B in 80, 8 KiB + 1 KiB in 82, 6 KiB or 10 KiB + 512 B in 96 / 89 (segments that are not a
// power of two fall across the strided segment order of a warp).  The warm-up is parsed twice; 256 bytes
// resynchronise all but a few segments per GiB, which the repair rounds re-parse.
struct ParseGeom { uint32_t seg, warm; };
// About one segment per parse lane (148 SMs x 7 CTAs x 128 threads = 132 608): a power of two between 256 bytes and
// the large-input segment, such that the input has between 98 304 and 196 608 of them.  A parse lane walks its
// segment (and the warm-up in front of it) sequentially, resolving long records on the way, so on inputs that do not
// fill the chip the segment length is the latency of the whole stage -- tools/latency.py, one Default call on
// Silesia-mix text, 1 KiB + 512 B segments against these: 64 KiB 4.9 -> 1.6 ms, 1 MiB 10.4 -> 2.2 ms, 4 MiB 7.9 ->
// 3.1 ms, 16 MiB 12.2 -> 7.2 ms; 256 x 4 MiB PNG-like chunks through the batch call 5 242 -> 6 994 MiB/s.
//   256 MiB in 8 KiB segments would leave three quarters of the lanes without work (parse 69 ms against 30 ms with
// 2 KiB segments, CompressionOptions::high()).  The warm-up is parsed twice, so it is short: 64 bytes resynchronise
// all but a few segments (the repair rounds re-parse those).
inline ParseGeom parse_geom(size_t payload, int mode) {
    if (payload <= (2u << 20)) return {128u, 64u};
    uint32_t seg = 256u;
    while (seg < kParseSeg && payload / seg > 196608u) seg <<= 1;
    if (seg <= 512u) return {seg, 64u};
    if (seg <= 1024u) return {seg, 128u};
    if (seg < kParseSeg) return {seg, 256u};
    return {kParseSeg, mode == kGreedy ? kParseWarmGreedy : kParseWarm};
}
inline uint32_t parse_tok_cap(ParseGeom g) { return g.seg + g.warm + 264u; }
inline size_t parse_n_seg(size_t payload, ParseGeom g) { return (payload + g.seg - 1) / g.seg; }
// u32 words of segment token buffers / number of segments needed for a payload of at most `cap` bytes, whatever its
// geometry: the largest payload of every geometry (the sizes at which parse_geom changes) is what counts
template <class F>
inline size_t parse_max_over_geometries(size_t cap, F f) {
    size_t best = 0;
    const size_t edges[8] = {(size_t)2u << 20, (size_t)196608u * 256u, (size_t)196608u * 512u, (size_t)196608u * 1024u,
                             (size_t)196608u * 2048u, (size_t)196608u * 4096u, (size_t)196608u * 8192u, cap};
    for (size_t e0 : edges) {
        const size_t e = e0 < cap ? e0 : cap;
        for (int mode : {(int)kGreedy, (int)kLazy}) {
            const size_t v = f(e, parse_geom(e, mode));
            if (v > best) best = v;
        }
    }
    return best;
}
inline size_t parse_buffer_words(size_t cap) {
    return parse_max_over_geometries(cap, [](size_t e, ParseGeom g) { return (parse_n_seg(e, g) + 1) * (size_t)parse_tok_cap(g); });
}
inline size_t parse_max_segments(size_t cap) {
    return parse_max_over_geometries(cap, [](size_t e, ParseGeom g) { return parse_n_seg(e, g) + 1; });
}
// Parallel repair rounds before the sequential fallback; every second one predicts the phase of chains of
// maximum-length matches (k_chain_predict).  A round is three small launches; short inputs, where launch latency
// is what counts, get fewer.
inline uint32_t repair_rounds(size_t payload) { return payload <= (4u << 20) ? 4u : 12u; }

// Device-resident bookkeeping of one encode call (one instance per context).
struct DevMeta {
    unsigned long long n_tokens;
    unsigned long long stream_bits;   // bits of the raw deflate stream (incl. sync marker)
    unsigned long long stream_bytes;  // ceil(stream_bits / 8)
    unsigned long long out_bytes;     // container size: header + stream + trailer
    uint32_t n_blocks;
    uint32_t adler;
    uint32_t crc;                     // CRC-32 of the payload (gzip)
    uint32_t n_bad;                   // segments whose hand-off check failed in the latest verify
    uint32_t n_repaired_par;
    uint32_t n_repaired_seq;
    uint32_t n_stored;
    uint32_t n_fixed;
    uint32_t err;                     // nonzero = internal invariant violated on the device
    // open pieces of a stream (streaming handle): where the parse stopped and what is left for the next piece
    uint32_t end_pos, end_key;        // parse state at the first iteration at or after EncodeJob::parse_end
    unsigned long long in_coded_end;  // input offset (in d_in) of the first token that was not coded
};

// Per-block cost summary (SoA-friendly, read by the bit-offset scan).
struct BlockCost {
    unsigned long long dynamic_cost, static_cost, stored_cost;   // reference's estimates
    unsigned long long dynamic_bits, fixed_bits;                  // bits actually emitted (body)
    unsigned long long input_bytes;
    uint32_t tiny;
    uint32_t hdr_bits;   // dynamic header bits after the 3-bit block marker
};

// Per-block code tables and header symbol stream (written by k_block_codes, read by k_pack).
struct BlockTables {
    uint16_t ll_code[288];
    uint16_t d_code[32];
    uint16_t cl_code[19];
    uint16_t hdr_sym[322];
    uint8_t ll_len[288];
    uint8_t d_len[32];
    uint8_t cl_len[19];
    uint8_t pad_;
    uint32_t n_hdr_sym, hlit, hdist, used_hclens;
};

struct Buffers {   // device scratch of one context, grown on demand
    uint2* K = nullptr;             // candidate entries in bucket order (dfl_core.h Entry), n_windows * 32768
    uint2* K2 = nullptr;            // bytes 8..15 of every entry of K, same order (the parse stage settles common prefixes below 16 from it)
    uint16_t* off = nullptr;        // bucket start offsets, n_windows * 32768
    uint32_t* Mf = nullptr;         // per-position match (full chain budget)
    uint32_t* Mq = nullptr;         // per-position match (quarter budget), only if needed
    uint32_t* segtok = nullptr;     // per parse segment token buffers, n_pseg * parse_tok_cap
    uint32_t* seg_e_pos = nullptr;  // hand-off records (SoA), n_pseg each
    uint32_t* seg_e_key = nullptr;
    uint32_t* seg_e_tok = nullptr;
    uint32_t* seg_x_pos = nullptr;
    uint32_t* seg_x_key = nullptr;
    uint32_t* seg_x_tok = nullptr;
    uint32_t* seg_start_pos = nullptr;   // repair start states
    uint32_t* seg_start_key = nullptr;
    uint8_t* seg_bad = nullptr;
    uint32_t* seg_cnt = nullptr;    // valid tokens per segment
    unsigned long long* seg_off = nullptr;   // exclusive prefix of seg_cnt
    uint32_t* seq_tab = nullptr;    // head / prev chains of the sequential lazy < 3 path (3 x 32768 words), on first use
    uint32_t* tok = nullptr;        // compacted token stream
    uint32_t* hist = nullptr;       // per block 320 counters (286 ll + 30 dist + pad)
    BlockCost* cost = nullptr;
    BlockTables* tables = nullptr;
    int* blk_type = nullptr;
    unsigned long long* blk_bit = nullptr;   // bit offset of each block in the stream
    unsigned long long* blk_in = nullptr;    // input offset of each block
    unsigned long long* adler_part = nullptr;// per chunk (A, B) sums
    DevMeta* meta = nullptr;
    size_t cap_n = 0;               // input size the buffers were sized for
    bool cap_quarter = false;
};

// MatchingType::Lazy with lazy_if_less_than < 3 on a one-shot stream takes the sequential kernel (k_lz77_seq):
// only there can the reference's length-2 results and its per-call ignore_next be observed.
inline bool use_seq_lz77(const Params& p, uint32_t begin, int open_piece, uint32_t init_key, uint32_t n_carry_tok) {
    return p.mode == kLazy && p.lazy < 3u && begin == 0 && !open_piece && init_key == 0 && n_carry_tok == 0;
}

// max_hash_checks == 1 (Compression::Fast): the one candidate of a position is its predecessor in the bucket; the
// sort kernel settles the matches itself (k_window_sort<true>, k_match_first).  Not with a quarter-budget record
// (lazy_if_less_than > 32: that budget is 0).
inline bool one_candidate(const Params& p) { return p.checks == 1u && p.mode != kRle && !p.need_quarter; }

struct EncodeJob {
    const uint8_t* d_in;     // device input (history + payload)
    uint32_t n;              // bytes available at d_in
    uint32_t begin;          // first byte to encode (bytes before it are dictionary only)
    Params prm;
    int final_block;         // set BFINAL on the last block (Flush::Finish)
    int sync_marker;         // append an empty stored block (Flush::Sync)
    uint8_t* d_out;          // device output, zero-filled by the pipeline
    size_t out_cap;
    uint32_t hdr_bytes;      // container header already accounted for at the start of d_out
    uint32_t isize;          // gzip ISIZE: total input length modulo 2^32
    const uint32_t* d_tokens_override;   // test hook: skip the LZ77 stage, use these tokens
    unsigned long long n_tokens_override;
    int stop_after_tokens;   // test hook: run only the LZ77 stage
    uint32_t peers;          // pipelines running beside this one (dfl_compress_device_batch): they fill the GPU together
    // ---- continuation of a stream in pieces without flushes (streaming handle; all zero for a fresh stream)
    uint32_t init_key;       // parse state at `begin` (parse_state_key), 0 = blank
    uint32_t parse_end;      // positions >= parse_end are left to the next piece (n for a closed piece)
    int open_piece;          // code complete 31744-token blocks only; the rest is carried
    uint32_t n_carry_tok;    // tokens carried in front of this piece's (already at the start of Buffers::tok)
    uint32_t carry_in_pos;   // offset in d_in of the first carried token's input byte
    uint32_t carry_bits_n;   // bits (0..7) of the last, incomplete byte of the previous piece ...
    uint32_t carry_bits_v;   // ... and their value
};

constexpr uint32_t kAdlerChunk = 1u << 16;

// Stage launchers (dfl_kernels.cu).  All asynchronous on `st`.
uint32_t n_windows(const EncodeJob& j);
uint32_t first_match_window(const EncodeJob& j);
uint32_t first_sort_window(const EncodeJob& j);
cudaError_t launch_window_sort(const EncodeJob& j, Buffers& b, cudaStream_t st, uint32_t w_lo, uint32_t w_hi);
cudaError_t launch_match(const EncodeJob& j, Buffers& b, cudaStream_t st, uint32_t w_lo, uint32_t w_hi);
cudaError_t launch_parse(const EncodeJob& j, Buffers& b, cudaStream_t st);
cudaError_t launch_lz77_seq(const EncodeJob& j, Buffers& b, cudaStream_t st);
cudaError_t launch_token_layout(const EncodeJob& j, Buffers& b, cudaStream_t st);
cudaError_t launch_block_stats(const EncodeJob& j, Buffers& b, cudaStream_t st);
cudaError_t launch_block_codes(const EncodeJob& j, Buffers& b, cudaStream_t st);
cudaError_t launch_block_scan(const EncodeJob& j, Buffers& b, cudaStream_t st);
cudaError_t launch_pack(const EncodeJob& j, Buffers& b, cudaStream_t st);
cudaError_t launch_adler32(const uint8_t* d_in, size_t n, Buffers& b, cudaStream_t st);
cudaError_t launch_crc32(const uint8_t* d_in, size_t n, Buffers& b, cudaStream_t st);
cudaError_t launch_finalize(const EncodeJob& j, Buffers& b, int wrap, cudaStream_t st);
uint32_t max_blocks_for(uint32_t n_payload);
extern thread_local int g_launch_count;   // kernels launched by the calling thread since its last reset

}  // namespace dfl
