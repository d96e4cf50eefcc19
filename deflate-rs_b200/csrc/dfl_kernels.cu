// dfl_kernels.cu -- sm_100a kernels of the DEFLATE encode pipeline.
//
// Stage order (one encode call; all launches on one stream, no host round trips):
//   k_window_sort   per 32 KiB window: stable counting sort of positions by the reference's 15-bit
//                   3-byte hash -> contiguous most-recent-first candidate lists (replaces the
//                   head/prev chains of chained_hash_table.rs)
//   k_match         per sorted entry: longest match among the first `max_hash_checks` candidates,
//                   nearest wins ties (matching.rs:87-166), too-far rule (lz77.rs:274-278)
//   k_parse*        greedy / lazy / RLE token selection (lz77.rs:305-547, rle.rs:23-71) run
//                   speculatively per 8 KiB segment, hand-offs verified, mismatches re-parsed
//   k_seg_scan, k_compact   token stream layout
//   k_block_stats   286+30 bin histograms per 31744-token block (output_writer.rs:19,47-65)
//   k_block_codes   code lengths, canonical codes, header RLE, costs (huffman_lengths.rs:167-287)
//   k_block_scan    block type per bit alignment + exclusive scan of block bit lengths
//   k_pack          Huffman-code the tokens and scatter the bits (encoder_state.rs:58-105,
//                   bitstream.rs:76-86, stored_block.rs:13-40)
//   k_adler32*      Adler-32 of the input (checksum.rs:33-57), k_finalize container bytes
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "dfl_internal.h"

// tuning knobs (tools/tune_variants.sh)
#ifndef DFL_WALK_UNROLL
#define DFL_WALK_UNROLL 4
#endif
#define DFL_PRAGMA_(x) _Pragma(#x)
#define DFL_PRAGMA(x) DFL_PRAGMA_(x)

namespace dfl {

thread_local int g_launch_count = 0;   // kernels launched by the calling thread since it last reset the counter

#define DFL_LAUNCH_CHECK()                         \
    do {                                           \
        g_launch_count++;                          \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return e__;        \
    } while (0)

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if ((int)lane_id() >= d) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread across the CTA (blockDim.x multiple of 32, <= 1024).
// `total` receives the CTA-wide sum.  `ws` is 33 words of shared scratch.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* ws, uint32_t& total) {
    uint32_t incl = warp_incl_scan(v);
    uint32_t nw = blockDim.x >> 5;
    if (lane_id() == 31) ws[warp_id()] = incl;
    __syncthreads();
    if (warp_id() == 0) {
        uint32_t x = lane_id() < nw ? ws[lane_id()] : 0;
        uint32_t xi = warp_incl_scan(x);
        ws[lane_id()] = xi - x;
        if (lane_id() == 31) ws[32] = xi;
    }
    __syncthreads();
    uint32_t r = ws[warp_id()] + incl - v;
    total = ws[32];
    __syncthreads();
    return r;
}

// Unaligned little-endian 32-bit read from a shared byte array viewed as words.
__device__ __forceinline__ uint32_t lds32(const uint32_t* w, uint32_t byte_idx) {
    uint32_t a = byte_idx >> 2;
    uint32_t lo = w[a], hi = w[a + 1];
    return __funnelshift_r(lo, hi, (byte_idx & 3u) * 8u);
}

// Copy in[src_lo .. src_lo+count) into shared bytes, zero-filling positions outside [0, n).
__device__ __forceinline__ void stage_bytes(uint8_t* dst, const uint8_t* __restrict__ in, long long src_lo,
                                            uint32_t count, uint32_t n) {
    // count and dst are multiples of 16; src_lo is a multiple of 16 (windows are 32 KiB aligned).
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) & 15u) == 0);
    for (uint32_t i = threadIdx.x * 16u; i < count; i += blockDim.x * 16u) {
        long long a = src_lo + (long long)i;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (a >= 0 && a + 16 <= (long long)n && aligned) {
            v = __ldg(reinterpret_cast<const uint4*>(in + a));
        } else if (a + 16 > 0 && a < (long long)n) {
            uint8_t tmp[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                long long x = a + k;
                tmp[k] = (x >= 0 && x < (long long)n) ? in[x] : (uint8_t)0;
            }
            v.x = tmp[0] | (tmp[1] << 8) | (tmp[2] << 16) | ((uint32_t)tmp[3] << 24);
            v.y = tmp[4] | (tmp[5] << 8) | (tmp[6] << 16) | ((uint32_t)tmp[7] << 24);
            v.z = tmp[8] | (tmp[9] << 8) | (tmp[10] << 16) | ((uint32_t)tmp[11] << 24);
            v.w = tmp[12] | (tmp[13] << 8) | (tmp[14] << 16) | ((uint32_t)tmp[15] << 24);
        }
        *reinterpret_cast<uint4*>(dst + i) = v;
    }
}

__host__ __device__ inline uint32_t window_count(uint32_t n, uint32_t w) {
    // number of hashable positions (p + 2 < n) in window w
    uint32_t hashable = n >= 2u ? n - 2u : 0u;
    uint32_t base = w * kWindow;
    if (hashable <= base) return 0u;
    uint32_t c = hashable - base;
    return c < kWindow ? c : kWindow;
}

// =====================================================================================
// k_window_sort: one CTA (1024 threads) per 32 KiB window.
//   A stable LSD radix sort (3 passes x 5 bits) of the window's positions by the reference's 15-bit
//   hash; every thread owns 32 consecutive positions and private digit counters, so no atomics and
//   no warp votes are needed and the order inside a bucket is position order by construction.
//   Output: the window's candidate lists as 64-bit entries (dfl_core.h Entry) in bucket order, and
//   the bucket start offsets.
//   shared: 64 KiB of packed u16 counters (also the byte staging area) + 132 KiB of padded items
// =====================================================================================
constexpr uint32_t kSortThreads = 1024;
constexpr uint32_t kSortItems = kWindow / kSortThreads;          // 32 items per thread
constexpr uint32_t kSortCntWords = 32 * (kSortThreads / 2);      // 32 digits x 512 packed counter pairs
constexpr uint32_t kSortBufWords = kWindow + kWindow / 32;       // +1 pad word per 32 (conflict-free blocked reads)
constexpr uint32_t kSortSmem = (kSortCntWords + kSortBufWords) * 4 + 256;
static_assert(kSortItems == 32, "one item per bit of the digit/prefix packing below");
static_assert(kWindow + 32 <= kSortCntWords * 4, "byte staging must fit in the counter area");

// word w of a counter row is stored at slot_of(w): its 32-word group rotated by the group number
__device__ __forceinline__ uint32_t sort_slot_of(uint32_t w) { return (w & ~31u) | ((w + (w >> 5)) & 31u); }

//   ONE (max_hash_checks == 1, e.g. Compression::Fast): the only candidate of a position is its predecessor in the
//   bucket (chained_hash_table.rs:148-158: the head of the chain when the position is inserted), and that is the
//   neighbouring item of the sorted order.  So the match is settled right here, on the bytes that are staged
//   anyway: the kernel writes the per-position match records instead of candidate lists, and `off` receives the
//   last position of every bucket (0xffff: none) for k_match_first, which settles the positions whose predecessor
//   lies in the previous window.  No lists, no k_match.
constexpr uint32_t kRecFirst = 0xffffffffu;   // ONE: provisional record of the first position of a bucket in its window
template <bool ONE>
__global__ void __launch_bounds__(kSortThreads, 1) k_window_sort(const uint8_t* __restrict__ in, uint32_t n,
                                                                 uint32_t w_first, uint2* __restrict__ K,
                                                                 uint2* __restrict__ K2, uint16_t* __restrict__ off,
                                                                 uint32_t* __restrict__ Mf) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* cntw = reinterpret_cast<uint32_t*>(smem);                        // kSortCntWords
    uint16_t* cnt16 = reinterpret_cast<uint16_t*>(smem);
    uint32_t* buf = reinterpret_cast<uint32_t*>(smem) + kSortCntWords;         // kSortBufWords
    uint32_t* xw = buf + kSortBufWords;                                        // 64 words of scratch
    const uint32_t t = threadIdx.x;
    const uint32_t w = w_first + blockIdx.x;
    const uint32_t base = w * kWindow;
    const uint32_t cnt = window_count(n, w);

    // ---- stage the window's bytes (+16) in the counter area and form the items (hash << 15 | pos)
    stage_bytes(smem, in, (long long)base, kWindow + 16, n);
    __syncthreads();
    {
        const uint4* s4 = reinterpret_cast<const uint4*>(smem) + t * 2u;
        uint4 a = s4[0], b = s4[1];
        uint32_t c = reinterpret_cast<const uint32_t*>(smem)[t * 8u + 8u];
        uint32_t wv[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c};
#pragma unroll
        for (uint32_t i = 0; i < kSortItems; i++) {
            uint32_t lo = wv[i >> 2], hi = wv[(i >> 2) + 1];
            uint32_t v = __funnelshift_r(lo, hi, (i & 3u) * 8u);   // bytes i, i+1, i+2 of this thread's run
            uint32_t pos = t * kSortItems + i;
            uint32_t h = pos < cnt ? hash3(v & 0xffu, (v >> 8) & 0xffu, (v >> 16) & 0xffu) : 0x7fffu;
            buf[t * 33u + i] = (h << 15) | pos;   // positions past the end sort behind every real entry
        }
    }
    __syncthreads();

    // Items live in `buf` in blocked order (thread t owns buf[33 t .. 33 t + 31]) between passes.
    // u16 index of this thread's counter inside a digit row.  Word w of a row holds the counters of threads w and
    // 512 + w; the 32-word groups of a row are rotated by their group number (word w lives at slot_of(w)), which
    // keeps the per-thread accesses of the counting and scattering loops conflict free *and* lets the row scan
    // below read 16 consecutive words per lane without bank conflicts (unrotated, the lanes of a warp would meet
    // in two bank groups: ncu showed 4x the ideal number of shared-memory wavefronts there).
    const uint32_t cw = sort_slot_of(t & 511u) * 2u + (t >> 9);
#pragma unroll 1
    for (uint32_t pass = 0; pass < 3; pass++) {
        const uint32_t shift = 15u + 5u * pass;
        for (uint32_t i = t; i < kSortCntWords; i += kSortThreads) cntw[i] = 0;
        __syncthreads();
        // thread-private digit counts; pre = number of earlier items of this thread with the same digit
        uint32_t pre[kSortItems / 4];                     // 8 bits each
#pragma unroll
        for (uint32_t i = 0; i < kSortItems; i++) {
            uint32_t d = (buf[t * 33u + i] >> shift) & 31u;
            uint32_t idx = d * 1024u + cw;
            uint32_t c = cnt16[idx];
            cnt16[idx] = (uint16_t)(c + 1u);
            if ((i & 3u) == 0) pre[i >> 2] = c; else pre[i >> 2] |= c << (8u * (i & 3u));
        }
        __syncthreads();
        // exclusive scan of all counters in (digit, thread) order.  Row d lives in words [512 d, 512 d + 512):
        // low halves are threads 0..511, high halves threads 512..1023.  Warp d scans row d.
        {
            const uint32_t d = t >> 5, l = t & 31u;
            uint32_t* row = cntw + d * 512u;
            uint32_t sum = 0;                              // packed (hi << 16 | lo); totals <= 32768 each
#pragma unroll
            for (uint32_t k = 0; k < 16; k++) sum += row[sort_slot_of(l * 16u + k)];
            uint32_t incl = sum;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, incl, dd);
                if ((int)l >= dd) incl += o;
            }
            uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t lo_tot = tot & 0xffffu, hi_tot = tot >> 16;
            if (l == 0) xw[d] = lo_tot + hi_tot;
            __syncthreads();
            uint32_t dbase = 0;
            {
                uint32_t x = xw[l];
                uint32_t xi = x;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    uint32_t o = __shfl_up_sync(0xffffffffu, xi, dd);
                    if ((int)l >= dd) xi += o;
                }
                dbase = __shfl_sync(0xffffffffu, xi - x, d);
            }
            uint32_t excl = incl - sum;
            uint32_t run_lo = dbase + (excl & 0xffffu);
            uint32_t run_hi = dbase + lo_tot + (excl >> 16);
#pragma unroll
            for (uint32_t k = 0; k < 16; k++) {
                uint32_t* slot = row + sort_slot_of(l * 16u + k);
                uint32_t vk = *slot;
                *slot = run_lo | (run_hi << 16);           // both < 65536 (at most 32768 items)
                run_lo += vk & 0xffffu;
                run_hi += vk >> 16;
            }
        }
        // every thread takes its items out of `buf` before anybody scatters into it
        uint32_t item[kSortItems];
#pragma unroll
        for (uint32_t i = 0; i < kSortItems; i++) item[i] = buf[t * 33u + i];
        __syncthreads();
#pragma unroll
        for (uint32_t i = 0; i < kSortItems; i++) {
            uint32_t d = (item[i] >> shift) & 31u;
            uint32_t r = cnt16[d * 1024u + cw] + ((pre[i >> 2] >> (8u * (i & 3u))) & 0xffu);
            buf[r + (r >> 5)] = item[i];
        }
        __syncthreads();
    }

    if (ONE) {
        // ---- match records: every item against its predecessor in the sorted order, if that is in the same bucket
        stage_bytes(smem, in, (long long)base, kWindow + 272, n);   // counters are dead; bytes again, with the look-ahead
        __syncthreads();
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(smem);
        for (uint32_t r = t; r < cnt; r += kSortThreads) {
            const uint32_t it = buf[r + (r >> 5)];
            const uint32_t pos = it & 0x7fffu, h = it >> 15;
            uint32_t rec = kRecFirst;
            if (r > 0) {
                const uint32_t itp = buf[(r - 1) + ((r - 1) >> 5)];
                if ((itp >> 15) == h) {
                    const uint32_t q = itp & 0x7fffu;                 // q < pos: the sort is stable
                    const uint32_t left = n - (base + pos), maxl = left < kMaxMatch ? left : kMaxMatch;
                    uint32_t l = 0;
                    while (l < maxl) {
                        const uint32_t x = lds32(sw, pos + l) ^ lds32(sw, q + l);
                        if (x != 0u) { l += (uint32_t)(__ffs((int)x) - 1) >> 3; break; }
                        l += 4u;
                    }
                    rec = finalize_match(l < maxl ? l : maxl, pos - q);
                }
            }
            Mf[base + pos] = rec;
        }
        __syncthreads();
        // ---- last position of every bucket
        uint16_t* st = reinterpret_cast<uint16_t*>(smem);
        for (uint32_t i = t; i < kWindow / 2; i += kSortThreads) reinterpret_cast<uint32_t*>(st)[i] = 0xffffffffu;
        __syncthreads();
        for (uint32_t r = t; r < cnt; r += kSortThreads) {
            const uint32_t it = buf[r + (r >> 5)];
            const uint32_t hn = r + 1u < cnt ? (buf[(r + 1) + ((r + 1) >> 5)] >> 15) : 0xffffffffu;
            if ((it >> 15) != hn) st[it >> 15] = (uint16_t)(it & 0x7fffu);
        }
        __syncthreads();
        uint4* o4 = reinterpret_cast<uint4*>(off + (size_t)w * kWindow);
        const uint4* s4 = reinterpret_cast<const uint4*>(smem);
        for (uint32_t i = t; i < kWindow * 2u / 16u; i += kSortThreads) o4[i] = s4[i];
        return;
    }
    // ---- output: entries in bucket order
    {
    stage_bytes(smem, in, (long long)base, kWindow + 32, n);   // counters are dead; bytes again
    __syncthreads();
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(smem);
    uint2* Kw = K + (size_t)w * kWindow;
    uint2* K2w = K2 + (size_t)w * kWindow;
    for (uint32_t r = t; r < kWindow; r += kSortThreads) {
        uint2 e = make_uint2(0u, 0xfffe0000u);               // filler beyond cnt; never read as a candidate
        if (r < cnt) {
            uint32_t it = buf[r + (r >> 5)];
            uint32_t pos = it & 0x7fffu;
            uint32_t a = pos >> 2, sh = (pos & 3u) * 8u;
            uint32_t w0 = sw[a], w1 = sw[a + 1], w2 = sw[a + 2];
            uint32_t v0 = __funnelshift_r(w0, w1, sh);       // bytes 0..3
            uint32_t v1 = __funnelshift_r(w1, w2, sh);       // bytes 4..7
            e.x = (v0 >> 24) | (v1 << 8);                    // bytes 3..6
            e.y = (v1 >> 24) | (tag9(v0 & 0xffu, (v0 >> 8) & 0xffu) << 8) | (pos << 17);
            // bytes 8..15, for the parse stage: what decides between candidates that share the 8 entry bytes
            const uint32_t w3 = sw[a + 3], w4 = sw[a + 4];
            K2w[r] = make_uint2(__funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
        Kw[r] = e;
    }
    }
    __syncthreads();
    // ---- bucket start offsets: off[h] = number of entries with hash < h
    uint16_t* st = reinterpret_cast<uint16_t*>(smem);        // 32768 u16, reuses the byte staging area
    for (uint32_t i = t; i < kWindow / 2; i += kSortThreads) reinterpret_cast<uint32_t*>(st)[i] = 0xffffffffu;
    __syncthreads();
    for (uint32_t r = t; r < cnt; r += kSortThreads) {
        uint32_t h = buf[r + (r >> 5)] >> 15;
        uint32_t hp = r > 0 ? (buf[(r - 1) + ((r - 1) >> 5)] >> 15) : 0xffffffffu;
        if (h != hp) st[h] = (uint16_t)r;
    }
    __syncthreads();
    {
        // backward fill: an empty bucket starts where the next non-empty one starts (or at cnt)
        uint32_t v[32];
        const uint4* s4 = reinterpret_cast<const uint4*>(st) + t * 4u;
        uint32_t first = 0xffffu;                             // first non-empty start in this thread's 32 buckets
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            uint4 q = s4[k];
            uint32_t ww[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) { v[k * 8 + j * 2] = ww[j] & 0xffffu; v[k * 8 + j * 2 + 1] = ww[j] >> 16; }
        }
#pragma unroll
        for (int k = 31; k >= 0; k--) if (v[k] != 0xffffu) first = v[k];
        // suffix "first valid" across threads: starts increase with the bucket index, so it is a suffix minimum
        uint32_t m = first;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            uint32_t o = __shfl_down_sync(0xffffffffu, m, dd);
            if ((int)(t & 31u) + dd < 32) m = m < o ? m : o;
        }
        if ((t & 31u) == 0) xw[t >> 5] = m;                  // warp minimum
        __syncthreads();
        uint32_t carry = 0xffffu;                             // minimum over later warps
        for (uint32_t k = (t >> 5) + 1; k < 32; k++) { uint32_t o = xw[k]; carry = carry < o ? carry : o; }
        uint32_t nxt = __shfl_down_sync(0xffffffffu, m, 1);   // suffix minimum of the lanes after this one
        if ((t & 31u) == 31u) nxt = 0xffffu;
        uint32_t fill = nxt < carry ? nxt : carry;
        if (fill == 0xffffu) fill = cnt;
#pragma unroll
        for (int k = 31; k >= 0; k--) { if (v[k] == 0xffffu) v[k] = fill; else fill = v[k]; }
        uint4* o4 = reinterpret_cast<uint4*>(off + (size_t)w * kWindow) + t * 4u;
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            uint4 q;
            q.x = v[k * 8 + 0] | (v[k * 8 + 1] << 16); q.y = v[k * 8 + 2] | (v[k * 8 + 3] << 16);
            q.z = v[k * 8 + 4] | (v[k * 8 + 5] << 16); q.w = v[k * 8 + 6] | (v[k * 8 + 7] << 16);
            o4[k] = q;
        }
    }
}

// =====================================================================================
// k_match: one CTA per window; a thread owns one sorted entry (the target) and visits its candidates
// most recent first: the entries before it in its own bucket, then the tail of the same bucket of
// the previous window (only positions at distance <= 32768), at most max_hash_checks in total.
//   Entries only: a visit is a coalesced 8-byte load and three logic ops against a pair of masks that
//   tighten as the running best grows (dfl_core.h EntryWalk).  Results shorter than 8 bytes are final;
//   a target whose best candidate shares all 8 entry bytes gets a "long" record (rank + visit index of
//   the nearest such candidate) and is resolved on the data by the parse stage -- if the parser ever
//   searches there.  No data is touched here beyond the three bytes that give a target its bucket.
//   shared: the window's bytes (+16)
// =====================================================================================
// 8 bytes at in[idx ..] (idx < n), little endian; bytes past the last word of the input read as 0 (callers clamp
// the lengths they derive to the bytes that exist).  Works for any alignment of `in`.
__device__ __forceinline__ unsigned long long ld8(const uint8_t* __restrict__ in, const uint32_t* last_word, uint32_t idx) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(in + idx);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3u) * 8u;
    const uint32_t x0 = __ldg(w);
    const uint32_t x1 = (w + 1 <= last_word) ? __ldg(w + 1) : 0u;
    const uint32_t x2 = (w + 2 <= last_word) ? __ldg(w + 2) : 0u;
    const uint32_t lo = __funnelshift_r(x0, x1, sh), hi = __funnelshift_r(x1, x2, sh);
    return ((unsigned long long)hi << 32) | lo;
}

constexpr uint32_t kMatchThreads = 512;
constexpr uint32_t kMatchStage = kWindow + 16;        // multiple of 16
constexpr uint32_t kMatchSmem = kMatchStage + 16;

// Per-lane state of one entry walk, kept in registers (the device-side form of dfl_core.h EntryWalk).
struct Walk {
    uint32_t me_lo, me_hi;   // the target's entry; me_hi is poisoned once the walk has stopped so that nothing passes
    uint32_t mlo, mhi;       // entry_mask(best_len)
    uint32_t best_len;       // 1 = nothing yet
    uint32_t best_pos, best_k;
    uint32_t stop;
    uint32_t q_len, q_pos, q_k;   // state after checks_quarter visits (NEEDQ)
};

// entry_mask() by best length 3..8 (index 0..2 unused): what a candidate must share to be longer
__constant__ uint2 c_walk_masks[9] = {
    {0u, 0u}, {0u, 0u}, {0u, 0u},
    {0x000000ffu, kEntryTagHi}, {0x0000ffffu, kEntryTagHi}, {0x00ffffffu, kEntryTagHi}, {0xffffffffu, kEntryTagHi},
    {0xffffffffu, kEntryKeyHi}, {0xffffffffu, kEntryKeyHi}};

// The visit itself: ((v ^ me) & mask) == 0 over both words, one LOP3 per word (written as PTX so that the
// compiler does not split the xor off for the rare path below and pay for it on every visit).
__device__ __forceinline__ bool walk_test(const Walk& wk, uint2 v) {
    uint32_t t1, t2;
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(t1) : "r"(v.x), "r"(wk.me_lo), "r"(wk.mlo));
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(t2) : "r"(v.y), "r"(wk.me_hi), "r"(wk.mhi));
    return (t1 | t2) == 0u;
}

// A candidate passed the masks: it shares more bytes with the target than the running best (unless the end
// of the input clamps it).  kg = visit index.
template <bool NEEDQ>
__device__ __forceinline__ void walk_improve(Walk& wk, uint2 v, uint32_t kg, uint32_t maxl, uint32_t qbudget) {
    if (wk.stop) return;
    uint32_t l = entry_lcp(v.x ^ wk.me_lo, v.y ^ wk.me_hi);
    l = l < maxl ? l : maxl;
    if (l > wk.best_len) {                      // strictly longer: the nearest candidate wins ties
        wk.best_len = l;
        wk.best_pos = v.y >> 17;
        wk.best_k = kg;
        if (NEEDQ) {
            if (kg < qbudget) { wk.q_len = l; wk.q_pos = wk.best_pos; wk.q_k = kg; }
        }
        const uint2 m = c_walk_masks[l];
        wk.mlo = m.x; wk.mhi = m.y;
        if (l >= kEntryBytes || l == maxl) {    // matching.rs:152-156, or nothing longer can be proven from entries
            wk.stop = 1;
            wk.mlo = 0xffffffffu; wk.mhi = kEntryKeyHi;
            wk.me_hi ^= 0x100u;                 // a flipped tag bit: (almost) nothing passes any more; the guard above catches the rest
        }
    }
}

__device__ __forceinline__ uint32_t warp_max(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ uint32_t warp_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }

#ifndef DFL_MATCH_CTAS
#define DFL_MATCH_CTAS 3      // CTAs of 512 threads per SM k_match is compiled for: 40 registers (at 4, 32 registers, the walk is 7 % slower)
#endif
template <bool NEEDQ>
__global__ void __launch_bounds__(kMatchThreads, DFL_MATCH_CTAS)
k_match(const uint8_t* __restrict__ in, uint32_t n, uint32_t begin, uint32_t w_first, Params prm,
        const uint2* __restrict__ K, const uint16_t* __restrict__ off, uint32_t* __restrict__ Mf,
        uint32_t* __restrict__ Mq) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(smem);
    const uint32_t w = w_first + blockIdx.x;
    const uint32_t base = w * kWindow;
    const uint32_t cnt = window_count(n, w);
    stage_bytes(smem, in, (long long)base, kMatchStage, n);
    __syncthreads();

    const uint2* Kw = K + (size_t)w * kWindow;
    const uint16_t* ow = off + (size_t)w * kWindow;
    const uint2* Kp = w > 0 ? K + (size_t)(w - 1) * kWindow : Kw;
    const uint16_t* op = w > 0 ? off + (size_t)(w - 1) * kWindow : ow;
    const uint32_t cnt_prev = w > 0 ? window_count(n, w - 1) : 0u;
    const uint32_t budget = prm.checks;
    const uint32_t qbudget = prm.checks_quarter;
    const uint32_t* last_word = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(in + (n ? n - 1 : 0)) & ~(uintptr_t)3);

    // gridDim.y CTAs share one window when the input is small (the stage's latency is one CTA's run through
    // 32768 entries otherwise): each takes a contiguous, 32-aligned slice of the sorted entries
    const uint32_t slice = ((cnt + gridDim.y - 1) / gridDim.y + 31u) & ~31u;
    const uint32_t i_lo = blockIdx.y * slice, i_hi = i_lo + slice < cnt ? i_lo + slice : cnt;
    for (uint32_t i0 = i_lo + (threadIdx.x & ~31u); i0 < i_hi; i0 += blockDim.x) {   // warp-uniform loop
        const uint32_t i = i0 + lane_id();
        bool act = i < i_hi;
        Walk wk;
        wk.me_lo = 0; wk.me_hi = 0;
        if (act) { uint2 v = __ldg(Kw + i); wk.me_lo = v.x; wk.me_hi = v.y; }
        const uint32_t pl = entry_pos(wk.me_hi);
        const uint32_t p = base + pl;
        if (p < begin) act = false;
        uint32_t n_own = 0, n_tot = 0, pe = 0;
        uint32_t maxl = 0;
        if (act) {
            const uint32_t w0 = lds32(sw, pl);
            const uint32_t h = hash3(w0 & 0xffu, (w0 >> 8) & 0xffu, (w0 >> 16) & 0xffu);
            maxl = (n - p) < kMaxMatch ? (n - p) : kMaxMatch;
            const uint32_t s0 = ow[h];
            n_own = i - s0 < budget ? i - s0 : budget;
            n_tot = n_own;
            if (w > 0 && n_own < budget) {
                // previous window: entries [ps, pe) of the same bucket whose position is >= pl
                // (distance <= 32768, matching.rs:102-106,127); they are position sorted.
                const uint32_t ps = op[h];
                pe = (h + 1u < kWindow) ? op[h + 1u] : cnt_prev;
                const uint32_t rem = budget - n_own;
                uint32_t lo = (pe - ps > rem) ? pe - rem : ps, hi = pe;
                // first j in [lo, pe) with pos >= pl; the oldest admissible entry is probed first
                // because usually the whole tail qualifies
                if (lo < hi) { if (entry_pos(__ldg(&Kp[lo].y)) >= pl) hi = lo; else lo++; }
                while (lo < hi) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (entry_pos(__ldg(&Kp[mid].y)) >= pl) hi = mid; else lo = mid + 1;
                }
                n_tot += pe - lo;
            }
        }
        wk.best_len = 1; wk.best_pos = 0; wk.best_k = 0; wk.stop = 0; wk.mlo = 0; wk.mhi = kEntryTagHi;
        wk.q_len = 1; wk.q_pos = 0; wk.q_k = 0;
        // Visit k is Kw[i - 1 - k] while k < n_own and Kp[pe - 1 - (k - n_own)] after that: one index space, so
        // that a warp whose lanes sit on both sides of a bucket boundary runs max(n_tot) steps, not the sum of
        // the two maxima.  Steps below the warp minimum of n_own need no per-lane case distinction.
        const uint32_t tA = warp_min(n_own), tB = warp_max(n_own), tmax = warp_max(n_tot);
        uint32_t k = 0;
        {   // every lane is inside its own window's list
            const uint2* ptr = Kw + i - 1;
            for (; k + 8u <= tA; k += 8u, ptr -= 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint2 v = __ldg(ptr - u);
                    if (walk_test(wk, v)) walk_improve<NEEDQ>(wk, v, k + u, maxl, qbudget);
                }
                if (__all_sync(0xffffffffu, wk.stop != 0u)) { k = tmax; break; }
            }
        }
        if (k < tB) {   // some lanes have crossed into the previous window's list, some have not
            const uint32_t oa = i - 1u + kWindow, ob = pe - 1u + n_own;    // entry index relative to Kw - kWindow, minus k
            const uint2* Kb = Kw - kWindow;
#pragma unroll 2
            for (; k < tB; k++) {
                if (k < n_tot) {
                    const uint2 v = __ldg(Kb + ((k < n_own ? oa : ob) - k));
                    if (walk_test(wk, v)) walk_improve<NEEDQ>(wk, v, k, maxl, qbudget);
                }
                if ((k & 15u) == 15u && __all_sync(0xffffffffu, wk.stop != 0u || k >= n_tot)) { k = tmax; break; }
            }
        }
        if (k < tmax && !__all_sync(0xffffffffu, wk.stop != 0u || k >= n_tot)) {   // only the previous window's list is left
            const uint2* ptr = Kp + pe - 1 + n_own - k;
#pragma unroll 2
            for (; k < tmax; k++, ptr--) {
                if (k < n_tot) {
                    const uint2 v = __ldg(ptr);
                    if (walk_test(wk, v)) walk_improve<NEEDQ>(wk, v, k, maxl, qbudget);
                }
                if ((k & 15u) == 15u && __all_sync(0xffffffffu, wk.stop != 0u || k >= n_tot)) break;
            }
        }
        if (act) {
            // distance of visit k: own window pl - pos, previous window pl + 32768 - pos
            const uint32_t d = (wk.best_k < n_own ? pl : pl + kWindow) - wk.best_pos;
            uint32_t rec = (wk.best_len >= kEntryBytes && maxl > kEntryBytes) ? rec_long(i, wk.best_k) : finalize_match(wk.best_len, d);
            if (rec_is_long(rec) && n_tot == 1u) {
                // a single candidate (always so at Compression::Fast, max_hash_checks = 1): nothing to choose from, so the
                // length is settled here instead of by the parser (the bytes are in L2: this CTA has just staged them)
                uint32_t l = kEntryBytes;
                while (l < maxl) {
                    const unsigned long long x = ld8(in, last_word, p + l) ^ ld8(in, last_word, p - d + l);
                    if (x != 0ull) { l += (uint32_t)(__ffsll((long long)x) - 1) >> 3; break; }
                    l += 8u;
                }
                rec = finalize_match(l < maxl ? l : maxl, d);
            }
            Mf[p] = rec;
            if (NEEDQ) {
                const uint32_t dq = (wk.q_k < n_own ? pl : pl + kWindow) - wk.q_pos;
                uint32_t rq = (wk.q_len >= kEntryBytes && maxl > kEntryBytes) ? rec_long(i, wk.q_k) : finalize_match(wk.q_len, dq);
                if (rec_is_long(rq) && n_tot == 1u) rq = rec;      // the one candidate there is: settled above
                Mq[p] = rq;
            }
        }
    }
}

// k_match_first (max_hash_checks == 1): the positions k_window_sort<true> left open -- the first of their bucket in
// their window -- have their one candidate in the previous window: the last position of the same bucket there, if
// it is within the window (matching.rs:102-106,127).
__global__ void __launch_bounds__(256) k_match_first(const uint8_t* __restrict__ in, uint32_t n, uint32_t w_first,
                                                     const uint16_t* __restrict__ off, uint32_t* __restrict__ Mf) {
    const uint32_t w = w_first + blockIdx.x / 4u, part = blockIdx.x % 4u;   // four CTAs per window
    const uint32_t base = w * kWindow, cnt = window_count(n, w);
    const uint32_t* last_word = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(in + (n ? n - 1 : 0)) & ~(uintptr_t)3);
    // four records per thread and load: almost all of them are settled already, so this is a streaming pass
    for (uint32_t g = part * (kWindow / 4u) + threadIdx.x * 4u; g < (part + 1u) * (kWindow / 4u) && g < cnt; g += blockDim.x * 4u) {
        const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(Mf + base + g));
        const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (uint32_t k = 0; k < 4u; k++) {
            const uint32_t pl = g + k;
            if (rr[k] != kRecFirst || pl >= cnt) continue;
            const uint32_t p = base + pl;
            uint32_t rec = 0u;
            if (w > 0u) {
                const uint32_t h = hash3(in[p], in[p + 1u], in[p + 2u]);
                const uint32_t ql = off[(size_t)(w - 1u) * kWindow + h];
                if (ql != 0xffffu && ql >= pl) {                            // distance <= 32768
                    const uint32_t d = pl + kWindow - ql;
                    const uint32_t maxl = (n - p) < kMaxMatch ? (n - p) : kMaxMatch;
                    uint32_t l = 0;
                    while (l < maxl) {
                        const unsigned long long x = ld8(in, last_word, p + l) ^ ld8(in, last_word, p - d + l);
                        if (x != 0ull) { l += (uint32_t)(__ffsll((long long)x) - 1) >> 3; break; }
                        l += 8u;
                    }
                    rec = finalize_match(l < maxl ? l : maxl, d);
                }
            }
            Mf[p] = rec;
        }
    }
}

// =====================================================================================
// parse: the reference's token selection (lz77.rs:305-547, rle.rs:23-71) over the match records, one lane per
// segment, with the long records resolved on the data by the whole warp.
//   A lane runs its parser until the position it examines has a long record; it then parks with a request
//   (position, candidate range, floor = prev_length).  When few lanes are left running the warp serves the
//   parked lanes one after the other: 32 candidates per round, lane-private entry test (all 8 entry bytes
//   equal), the reference's quick reject on the byte that would extend the best match (matching.rs:141-143),
//   a lock-step byte comparison 8 bytes at a time, and a warp max that keeps the nearest of the longest
//   (matching.rs:148-157).  Only positions the reference's parser searches are ever resolved: everything
//   inside an emitted match is skipped (lz77.rs:398-405), which is 70-85 % of all positions on text.
// =====================================================================================
__device__ __forceinline__ ParseState state_from_key(uint32_t pos, uint32_t key) { return parse_state_from_key(pos, key); }

struct ParseArgs {
    const uint8_t* in; uint32_t n; uint32_t begin; Params prm;
    const uint32_t* Mf; const uint32_t* Mq;
    const uint2* K; const uint16_t* off;   // sorted entries and bucket offsets per window (long records refer to them)
    const uint2* K2;                       // bytes 8..15 of every sorted entry
    uint32_t* segtok;
    uint32_t *e_pos, *e_key, *e_tok, *x_pos, *x_key, *x_tok;
    uint32_t n_seg;
    uint32_t seg, warm, tok_cap;   // parse_geom of this call
    uint32_t end;                  // parse bound (EncodeJob::parse_end)
    uint32_t init_key;             // state at `begin`
};

// What a parked lane asks for.  Filled by the lane itself; all parked lanes of a warp prepare at the same time.
struct Resolve {
    uint32_t p;        // position
    uint32_t rank;     // index of the position in its window's sorted list
    uint32_t k0;       // first visit to look at (rec_k8)
    uint32_t start;    // max(prev_length, 7): only a longer result is of use (matching.rs:161-165)
    uint32_t maxl;     // min(258, bytes left)
    uint32_t n_own;    // visits k < n_own are Kw[rank - 1 - k]
    uint32_t n_vis;    // visits allowed in total (chain budget; the previous window's share may end earlier)
    uint32_t pe;       // visit k >= n_own is Kp[pe - 1 - (k - n_own)]
    uint32_t me_lo, me_hi;
    uint32_t b8_lo, b8_hi;   // bytes 8..15 of the target
};

__device__ __forceinline__ void resolve_prepare(const ParseArgs& A, Resolve& r, uint32_t budget, bool quarter) {
    const uint32_t p = r.p, w = p >> 15;
    const uint32_t rec = quarter ? A.Mq[p] : A.Mf[p];   // a long record: rank of p in its window's list, first visit to look at
    r.rank = rec_rank(rec); r.k0 = rec_k8(rec);
    const uint2 me = __ldg(A.K + (size_t)w * kWindow + r.rank);
    const uint2 b8 = __ldg(A.K2 + (size_t)w * kWindow + r.rank);
    r.me_lo = me.x; r.me_hi = me.y; r.b8_lo = b8.x; r.b8_hi = b8.y;
    const uint32_t h = hash3(A.in[p], A.in[p + 1], A.in[p + 2]);
    const uint16_t* ow = A.off + (size_t)w * kWindow;
    const uint32_t full = A.prm.checks;
    const uint32_t s0 = __ldg(ow + h);
    r.n_own = r.rank - s0 < full ? r.rank - s0 : full;
    uint32_t n_tot = r.n_own;
    r.pe = 0;
    if (w > 0 && r.n_own < full) {
        const uint16_t* op = ow - kWindow;
        const uint32_t ps = __ldg(op + h);
        r.pe = (h + 1u < kWindow) ? (uint32_t)__ldg(op + h + 1u) : window_count(A.n, w - 1);
        const uint32_t rem = full - r.n_own;
        n_tot += (r.pe - ps > rem) ? rem : r.pe - ps;   // entries with a position below pl are cut off during the scan
    }
    r.n_vis = n_tot < budget ? n_tot : budget;
    if (r.start >= r.maxl) r.n_vis = 0;                 // nothing can be longer than the floor (matching.rs:99-101)
}

// Entry of visit k of a request (warp-uniform request, lane-private k).
__device__ __forceinline__ const uint2* resolve_entry(const ParseArgs& A, uint32_t w, uint32_t rank, uint32_t n_own, uint32_t pe,
                                                      uint32_t k) {
    const uint2* Kw = A.K + (size_t)w * kWindow;
    return k < n_own ? Kw + (rank - 1u - k) : Kw - kWindow + (pe - 1u - (k - n_own));
}

#ifndef DFL_PARSE_CTAS
#define DFL_PARSE_CTAS 7      // CTAs of 128 threads per SM the parse kernels are compiled for (72 registers)
#endif
#ifndef DFL_PARSE_KEEP
#define DFL_PARSE_KEEP 6      // keep parsing while at least this many lanes of a warp are running
#endif
constexpr uint32_t kParseThreads = 128;
constexpr uint32_t kParseWarps = kParseThreads / 32;
constexpr uint32_t kCandCap = 128;       // candidates a warp collects before it compares them
#ifndef DFL_LC_BYTES
#define DFL_LC_BYTES 128
#endif
constexpr uint32_t kLcBytes = DFL_LC_BYTES;   // length codes (dfl_core.h rec_len_code, one byte per position) a lane keeps in shared memory
constexpr uint32_t kLcWords = kLcBytes / 4;
constexpr uint32_t kLcHalf = kLcBytes / 2;
struct ParseShared {
    uint32_t cand_q[kCandCap];           // absolute position of a candidate that shares the target's 8 entry bytes
    uint32_t cand_meta[kCandCap];        // owner lane | visit index << 5
    uint2 cand_b8[kCandCap];             // the candidate's bytes 8..15
    uint32_t res_key[32];                // per owner lane: best length << 16 | (0xffff - visit index)
    uint32_t lc[32 * (kLcWords + 1)];    // per lane kLcWords words of length codes (+1: rows fall on different banks)
};

// Tokens as the parser writes them into its segment buffer.  What the parser does not need in order to decide --
// the byte of a literal, the distance of a match whose record is final -- is not loaded by it (a dependent load
// per step is what bounds a sequential parser on a GPU): such tokens name a position instead and k_compact,
// fully parallel, fills them in.  Positions are relative to (segment start - origin_back).
constexpr uint32_t kTokLitAt = 0x80000000u;       // | relative position
constexpr uint32_t kTokMatchAt = 0x40000000u;     // | quarter-budget record << 29 | length << 14 | relative position
constexpr uint32_t kTokQuarter = 0x20000000u;
constexpr uint32_t kTokRelMask = 0x3fffu;
__host__ __device__ inline uint32_t parse_origin_back(uint32_t warm) { return warm + 300u; }
struct DeferSink {
    uint32_t* tk; uint32_t nt, cap; long long origin;
    __device__ __forceinline__ void put(uint32_t t) { if (nt < cap) tk[nt] = t; nt++; }
    __device__ __forceinline__ void literal(uint32_t pos) { put(kTokLitAt | (uint32_t)((long long)pos - origin)); }
    __device__ __forceinline__ void match(uint32_t len, uint32_t ref, uint32_t kind) {
        if (kind == kRefDist) put(tok_match(len, ref));
        else put(kTokMatchAt | (kind == kRefQuarter ? kTokQuarter : 0u) | (len << 14) | (uint32_t)((long long)ref - origin));
    }
};

// Compares the collected candidates of all owners, one candidate per lane and round.  Bytes 8..15 of target and
// candidate are at hand (K2), so a common prefix below 16 bytes is settled without touching the input; only a
// candidate that shares all 16 goes on: the byte that would extend the owner's running best
// (matching.rs:141-143), then a lock-step comparison 8 bytes at a time.  The longest candidate wins, the nearest
// among equals (matching.rs:148-157): a max over length << 16 | ~visit.
__device__ __forceinline__ void resolve_compare(const ParseArgs& A, ParseShared& S, uint32_t cnt, const Resolve& rq,
                                                const uint32_t* last_word) {
    const uint32_t lane = lane_id();
    for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t c = base + lane;
        const bool have = c < cnt;
        const uint32_t meta = have ? S.cand_meta[c] : 0u;
        const uint32_t q = have ? S.cand_q[c] : 0u;
        const uint2 cb8 = have ? S.cand_b8[c] : make_uint2(0u, 0u);
        const uint32_t owner = meta & 31u, k = meta >> 5;
        const uint32_t po = __shfl_sync(0xffffffffu, rq.p, owner);
        const uint32_t st = __shfl_sync(0xffffffffu, rq.start, owner);
        const uint32_t ml = __shfl_sync(0xffffffffu, rq.maxl, owner);
        const uint32_t tlo = __shfl_sync(0xffffffffu, rq.b8_lo, owner), thi = __shfl_sync(0xffffffffu, rq.b8_hi, owner);
        const uint32_t cur = S.res_key[owner] >> 16;          // the owner's best from earlier rounds: all of them nearer
        const uint32_t s0 = st > cur ? st : cur;
        uint32_t mine = 0;
        bool alive = false;
        if (have && s0 < ml) {
            const uint32_t xlo = cb8.x ^ tlo, xhi = cb8.y ^ thi;
            if (xlo) mine = kEntryBytes + ((uint32_t)(__ffs((int)xlo) - 1) >> 3);
            else if (xhi) mine = kEntryBytes + 4u + ((uint32_t)(__ffs((int)xhi) - 1) >> 3);
            else if (ml <= 2u * kEntryBytes) mine = ml;
            else alive = (s0 < 2u * kEntryBytes) || (A.in[q + s0] == A.in[po + s0]);
        }
        uint32_t l = 2u * kEntryBytes;
        while (__any_sync(0xffffffffu, alive)) {              // lock step: l is the same in every lane that is alive
            if (alive) {
                const unsigned long long x = ld8(A.in, last_word, po + l) ^ ld8(A.in, last_word, q + l);
                if (x != 0ull) { mine = l + ((uint32_t)(__ffsll((long long)x) - 1) >> 3); alive = false; }
                else if (l + 8u >= ml) { mine = ml; alive = false; }
            }
            l += 8u;
        }
        mine = mine < ml ? mine : ml;
        if (mine > s0) atomicMax(&S.res_key[owner], (mine << 16) | (0xffffu - k));
        __syncwarp();
    }
}

// One owner's request, broadcast to the warp.
struct Owner { uint32_t p, rank, k0, n_own, n_vis, pe, lo, hi; };
__device__ __forceinline__ Owner resolve_owner(const Resolve& rq, int j) {
    Owner o;
    o.p = __shfl_sync(0xffffffffu, rq.p, j); o.rank = __shfl_sync(0xffffffffu, rq.rank, j);
    o.k0 = __shfl_sync(0xffffffffu, rq.k0, j); o.n_own = __shfl_sync(0xffffffffu, rq.n_own, j);
    o.n_vis = __shfl_sync(0xffffffffu, rq.n_vis, j); o.pe = __shfl_sync(0xffffffffu, rq.pe, j);
    o.lo = __shfl_sync(0xffffffffu, rq.me_lo, j); o.hi = __shfl_sync(0xffffffffu, rq.me_hi, j);
    return o;
}
// Entries of visits kb + 32 t + lane, t = 0..3 (up to 128 entries of one owner in flight); slots past the owner's
// last visit are not touched.
__device__ __forceinline__ void resolve_load4(const ParseArgs& A, const Owner& o, uint32_t kb, uint2 e[4]) {
    const uint2* Kw = A.K + (size_t)(o.p >> 15) * kWindow;
    const uint2* pa = Kw + (o.rank - 1u);                 // visit k of the own window: pa - k
    const uint2* pb = Kw - kWindow + (o.pe - 1u + o.n_own);   // visit k of the previous window: pb - k
#pragma unroll
    for (uint32_t t = 0; t < 4; t++) {
        const uint32_t k = kb + t * 32u + lane_id();
        e[t] = make_uint2(0u, 0u);
        if (k < o.n_vis) e[t] = __ldg((k < o.n_own ? pa : pb) - k);
    }
}

// Every lane of the calling warp enters (work == false: the lane only helps with resolutions).  A lane with work
// runs the reference's token selection from `st` until the first iteration position >= b.
__device__ void parse_lanes(const ParseArgs& A, ParseShared& S, bool work, uint32_t s, ParseState st, uint32_t a, uint32_t b) {
    DeferSink out;
    out.tk = A.segtok + (size_t)s * A.tok_cap; out.nt = 0; out.cap = A.tok_cap;
    out.origin = (long long)A.begin + (long long)s * A.seg - (long long)parse_origin_back(A.warm);
    bool have_e = false;
    uint32_t epos = 0, ekey = 0, etok = 0;
    const uint32_t n = A.n;
    const int mode = A.prm.mode;
    const bool has_m = (mode != kRle) && (A.prm.checks > 0);
    const uint32_t lane = lane_id();
    const uint32_t* last_word = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(A.in + (n ? n - 1 : 0)) & ~(uintptr_t)3);
    uint32_t* lc = S.lc + lane * (kLcWords + 1u);
    uint32_t cb = 0xffffff00u;                       // position of the first cached length code (a multiple of kLcHalf); nothing yet
    bool running = work, parked = false, have_m = false, rq_quarter = false;
    uint32_t m_ready = 0, rq_budget = 0;
    Resolve rq;
    rq.p = rq.rank = rq.k0 = rq.start = rq.maxl = rq.n_own = rq.n_vis = rq.pe = rq.me_lo = rq.me_hi = rq.b8_lo = rq.b8_hi = 0;
    // the hand-off key carries the pending match's distance: looked up here, once per segment boundary
    auto state_key = [&](const ParseState& x) -> uint32_t {
        uint32_t d = 0;
        if (x.prev_len) d = x.prev_kind == kRefDist ? x.prev_ref : match_dist((x.prev_kind == kRefQuarter ? A.Mq : A.Mf)[x.prev_ref]);
        return parse_state_key(x, d);
    };
    for (;;) {
        // ---- parse: every running lane takes one step per iteration
        for (;;) {
            bool miss = false;
            if (running && !parked) {
                bool stop_here = st.pos >= n;
                if (!stop_here) {
                    if (!have_e && st.pos >= a) { epos = st.pos; ekey = state_key(st); etok = out.nt; have_e = true; }
                    stop_here = st.pos >= b;
                }
                if (stop_here) {
                    if (!have_e) { epos = st.pos; ekey = state_key(st); etok = out.nt; }
                    A.e_pos[s] = epos; A.e_key[s] = ekey; A.e_tok[s] = etok;
                    A.x_pos[s] = st.pos; A.x_key[s] = state_key(st); A.x_tok[s] = out.nt;
                    running = false;
                } else {
                    const uint32_t p = st.pos;
                    uint32_t m_len = 0, m_ref = 0, m_kind = kRefDist;
                    if (have_m) {                                            // the answer to this lane's request
                        m_len = match_len(m_ready); m_ref = m_len ? match_dist(m_ready) : 0u; have_m = false;
                    } else if (has_m && p + 2u < n && (mode != kLazy || !st.ign)) {   // the reference searches here (lz77.rs:347,505)
                        const bool quarter = mode == kLazy && st.prev_len >= 32u;        // lz77.rs:351-355
                        if (!quarter || A.prm.need_quarter) {
                            uint32_t code;
                            if (quarter) code = rec_len_code(A.Mq[p]);
                            else if (p - cb < kLcBytes) code = (lc[(p - cb) >> 2] >> (8u * (p & 3u))) & 0xffu;   // cb is a multiple of 4
                            else { code = 0; miss = true; }
                            if (code == kLenLong) {
                                rq.p = p;
                                const uint32_t floor = mode == kLazy ? st.prev_len : 0u;
                                rq.start = floor > kEntryBytes - 1u ? floor : kEntryBytes - 1u;
                                rq.maxl = (n - p) < kMaxMatch ? (n - p) : kMaxMatch;
                                rq_budget = quarter ? A.prm.checks_quarter : A.prm.checks;
                                rq_quarter = quarter;
                                parked = true;
                            } else {
                                m_len = code == 0u ? 0u : (code == kLenSeeRecord ? match_len(quarter ? A.Mq[p] : A.Mf[p]) : code + 2u);
                                m_ref = p; m_kind = quarter ? kRefQuarter : kRefFull;
                            }
                        }
                    }
                    if (!parked && !miss) {
                        if (mode == kLazy) lazy_step(st, n, m_len, m_ref, m_kind, A.prm.lazy, out);
                        else if (mode == kGreedy) greedy_step(st, n, m_len, m_ref, m_kind, out);
                        else rle_step(st, n, A.in, out);       // rle.rs runs over the buffer from its first byte
                    }
                }
            }
            // Length codes: the match records of the next kLcBytes positions of a lane, one byte each, fetched by the
            // whole warp (coalesced 16-byte loads, four records per lane and load) when the lane runs out; lanes that
            // are past the middle of what they hold are refreshed in the same round, so that round trips are shared.
            uint32_t want = __ballot_sync(0xffffffffu, miss);
            if (want) {
                want |= __ballot_sync(0xffffffffu, running && !parked && has_m && st.pos - cb >= kLcHalf);
                const uint32_t mine = st.pos & ~(kLcHalf - 1u);
                for (uint32_t left = want; left;) {
                    const int j = __ffs((int)left) - 1;
                    left &= left - 1u;
                    const uint32_t base = __shfl_sync(0xffffffffu, mine, j);
#pragma unroll
                    for (uint32_t u = 0; u < kLcWords / 32u; u++) {
                        const uint32_t q = base + 4u * (u * 32u + lane);      // this lane converts records q .. q + 3
                        uint32_t codes = 0;
                        if (q + 4u <= n && ((reinterpret_cast<uintptr_t>(A.Mf + q) & 15u) == 0)) {
                            const uint4 r = __ldg(reinterpret_cast<const uint4*>(A.Mf + q));
                            codes = rec_len_code(r.x) | (rec_len_code(r.y) << 8) | (rec_len_code(r.z) << 16) | (rec_len_code(r.w) << 24);
                        } else {
#pragma unroll
                            for (uint32_t t = 0; t < 4; t++) if (q + t < n) codes |= rec_len_code(A.Mf[q + t]) << (8u * t);
                        }
                        S.lc[(uint32_t)j * (kLcWords + 1u) + u * 32u + lane] = codes;
                    }
                }
                if ((want >> lane) & 1u) cb = mine;
                __syncwarp();
            }
            const uint32_t going = __ballot_sync(0xffffffffu, running && !parked);
            if (going == 0u) break;
            if ((uint32_t)__popc(going) < DFL_PARSE_KEEP && __any_sync(0xffffffffu, parked)) break;
        }
        // ---- resolve the parked lanes' long records together
        const uint32_t waiting = __ballot_sync(0xffffffffu, parked);
        if (waiting == 0u) {
            if (!__any_sync(0xffffffffu, running)) break;
            continue;
        }
        if (parked) resolve_prepare(A, rq, rq_budget, rq_quarter);   // every parked lane at once: their loads overlap
        S.res_key[lane] = 0u;
        __syncwarp();
        // scan: per owner, every candidate from visit k0 on whose 8 entry bytes equal the target's goes on the list;
        // the entries of the next owner are requested before the current one's are looked at
        uint32_t cnt = 0;
        uint32_t left = waiting;
        int j = __ffs((int)left) - 1;
        left &= left - 1u;
        Owner o = resolve_owner(rq, j);
        uint2 e[4];
        resolve_load4(A, o, o.k0, e);
        for (;;) {
            const int jn = left ? __ffs((int)left) - 1 : -1;
            Owner on = o;
            uint2 en[4];
            if (jn >= 0) {
                left &= left - 1u;
                on = resolve_owner(rq, jn);
                resolve_load4(A, on, on.k0, en);
            }
            const uint32_t w = o.p >> 15, pl = o.p & kWindowMask;
            bool over = false;
            for (uint32_t kb = o.k0; kb < o.n_vis && !over; kb += 128u) {
                if (kb != o.k0) resolve_load4(A, o, kb, e);      // chain budgets above 128 only
#pragma unroll
                for (uint32_t t = 0; t < 4; t++) {
                    if (kb + t * 32u >= o.n_vis) break;          // warp-uniform: the owner has no visits in this slot
                    const uint32_t k = kb + t * 32u + lane;
                    const bool val = k < o.n_vis, own = k < o.n_own;
                    const uint32_t ep = entry_pos(e[t].y);
                    const bool ended = val && !own && ep < pl;   // beyond the window (matching.rs:102-106); so is everything older
                    const bool eq = val && !ended && e[t].x == o.lo && (((e[t].y ^ o.hi) & kEntryKeyHi) == 0u);
                    const uint32_t mk = __ballot_sync(0xffffffffu, eq);
                    if (mk) {
                        if (cnt + 32u > kCandCap) { __syncwarp(); resolve_compare(A, S, cnt, rq, last_word); cnt = 0; }
                        if (eq) {
                            const uint32_t at = cnt + __popc(mk & ((1u << lane) - 1u));
                            const uint32_t cw = own ? w : w - 1u;
                            S.cand_q[at] = cw * kWindow + ep;
                            S.cand_meta[at] = (uint32_t)j | (k << 5);
                            S.cand_b8[at] = __ldg(A.K2 + (size_t)cw * kWindow + (own ? o.rank - 1u - k : o.pe - 1u - (k - o.n_own)));
                        }
                        cnt += __popc(mk);
                    }
                    if (__any_sync(0xffffffffu, ended)) { over = true; break; }
                }
            }
            if (jn < 0) break;
            j = jn; o = on;
#pragma unroll
            for (uint32_t t = 0; t < 4; t++) e[t] = en[t];
        }
        __syncwarp();
        resolve_compare(A, S, cnt, rq, last_word);
        // every owner picks up its result; the distance comes from the winning visit's entry
        if (parked) {
            const uint32_t key = S.res_key[lane];
            m_ready = 0u;
            if (key != 0u) {
                const uint32_t k = 0xffffu - (key & 0xffffu);
                const uint32_t w = rq.p >> 15;
                const uint2 e = __ldg(resolve_entry(A, w, rq.rank, rq.n_own, rq.pe, k));
                const uint32_t q = (k < rq.n_own ? w : w - 1u) * kWindow + entry_pos(e.y);
                m_ready = finalize_match(key >> 16, rq.p - q);
            }
            have_m = true;
            parked = false;
        }
        __syncwarp();
    }
}



__global__ void __launch_bounds__(kParseThreads, DFL_PARSE_CTAS) k_parse_spec(ParseArgs A) {
    __shared__ ParseShared sh[kParseWarps];
    // The lanes of a warp take segments far apart (stride = number of warps in the grid): the cost of a segment
    // depends on the kind of data, neighbouring segments are of one kind, and a warp is as slow as its slowest lane.
#ifndef DFL_PARSE_STRIDED
#define DFL_PARSE_STRIDED 1
#endif
    const uint32_t n_warps = gridDim.x * kParseWarps;
    const uint32_t s = DFL_PARSE_STRIDED ? lane_id() * n_warps + blockIdx.x * kParseWarps + warp_id()
                                         : blockIdx.x * blockDim.x + threadIdx.x;
    const bool work = s < A.n_seg;
    uint32_t a = 0, b = 0;
    ParseState st = parse_state_init(0);
    if (work) {
        a = A.begin + s * A.seg;
        b = a + A.seg < A.end ? a + A.seg : A.end;
        const uint32_t start = (s == 0) ? A.begin : (a - A.begin > A.warm ? a - A.warm : A.begin);
        st = (s == 0) ? state_from_key(A.begin, A.init_key) : parse_state_init(start);
    }
    parse_lanes(A, sh[warp_id()], work, s, st, a, b);
}

__global__ void __launch_bounds__(128) k_parse_verify(ParseArgs A, uint8_t* bad, uint32_t* start_pos,
                                                      uint32_t* start_key, DevMeta* meta) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n_seg) return;
    uint8_t is_bad = 0;
    if (s > 0) {
        uint32_t xp = A.x_pos[s - 1], xk = A.x_key[s - 1];
        if (xp != A.e_pos[s] || xk != A.e_key[s]) {
            is_bad = 1;
            start_pos[s] = xp;
            start_key[s] = xk;
            atomicAdd(&meta->n_bad, 1u);
        }
    }
    bad[s] = is_bad;
}

__global__ void __launch_bounds__(kParseThreads, DFL_PARSE_CTAS) k_parse_repair(ParseArgs A, const uint8_t* bad, const uint32_t* start_pos,
                                                                const uint32_t* start_key, DevMeta* meta) {
    __shared__ ParseShared sh[kParseWarps];
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool work = s < A.n_seg && bad[s];
    if (!__any_sync(0xffffffffu, work)) return;
    uint32_t a = 0, b = 0;
    ParseState st = parse_state_init(0);
    if (work) {
        a = A.begin + s * A.seg;
        b = a + A.seg < A.end ? a + A.seg : A.end;
        st = state_from_key(start_pos[s], start_key[s]);
        atomicAdd(&meta->n_repaired_par, 1u);
    }
    parse_lanes(A, sh[warp_id()], work, s, st, a, b);
}

// Sequential fallback (one warp; lane 0 parses, the others help with resolutions): walks the segments in order
// and re-parses every one whose entry does not continue its predecessor's exit.  Exact for any input; only slow
// on inputs whose speculative parses never resynchronise (e.g. megabytes of a single repeated byte).
__global__ void __launch_bounds__(32) k_parse_repair_seq(ParseArgs A, DevMeta* meta) {
    __shared__ ParseShared sh;
    if (meta->n_bad == 0) return;
    const uint32_t lane = lane_id();
    // The hand-offs are checked 32 at a time (on an input of a few byte values, issue_44-like, a handful of bad
    // segments per MiB survive the parallel rounds: walking 131 072 hand-offs one by one took 64 ms per 256 MiB).
    // A repair changes the exit of its segment, so the scan goes on right behind it.  (__ldcg: what lane 0 has just
    // written must be seen by the other lanes.)
    uint32_t s = 1;
    while (s < A.n_seg) {
        const uint32_t t = s + lane;
        const bool isbad = t < A.n_seg && (__ldcg(A.x_pos + t - 1) != __ldcg(A.e_pos + t) || __ldcg(A.x_key + t - 1) != __ldcg(A.e_key + t));
        const uint32_t m = __ballot_sync(0xffffffffu, isbad);
        if (m == 0u) { s += 32u; continue; }
        s += (uint32_t)__ffs((int)m) - 1u;
        const uint32_t xp = __ldcg(A.x_pos + s - 1), xk = __ldcg(A.x_key + s - 1);
        const uint32_t a = A.begin + s * A.seg;
        const uint32_t b = a + A.seg < A.end ? a + A.seg : A.end;
        parse_lanes(A, sh, lane == 0, s, state_from_key(xp, xk), a, b);
        if (lane == 0) meta->n_repaired_seq++;
        __threadfence();
        __syncwarp();
        s += 1u;
    }
    if (lane == 0) meta->n_bad = 0;
}

__global__ void k_reset_bad(DevMeta* meta) { meta->n_bad = 0; }

// Repair start states for chains of bad segments.  A speculative parse resynchronises with the true one at the
// end of a match -- except inside a chain of maximum-length matches (runs of one byte, short periods): every
// match is cut at 258 bytes (huffman_table.rs:21), so where a match ends depends on where the chain began, and
// the hand-off check fails for every segment of the run.  But then the true parser is in a blank state every 258
// bytes, so the entry of *every* segment of the chain follows from the exit of the last good segment in front of
// it: the first blank position at or after the segment start with the same phase.  This kernel rewrites the
// repair start of every bad segment whose predecessor is bad too; the next verify checks the prediction like
// any other hand-off (a wrong guess stays bad and is repaired the ordinary way).  One CTA.
__global__ void __launch_bounds__(1024) k_chain_predict(ParseArgs A, const uint8_t* bad, uint32_t* start_pos, uint32_t* start_key) {
    __shared__ int carry[1024];
    const uint32_t per = (A.n_seg + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per < A.n_seg ? threadIdx.x * per : A.n_seg;
    const uint32_t hi = lo + per < A.n_seg ? lo + per : A.n_seg;
    int last = -1;
    for (uint32_t i = lo; i < hi; i++) if (!bad[i]) last = (int)i;
    carry[threadIdx.x] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = -1;
        for (uint32_t t = 0; t < blockDim.x; t++) { const int v = carry[t]; carry[t] = run; if (v >= 0) run = v; }
    }
    __syncthreads();
    int g = carry[threadIdx.x];                      // last good segment in front of this thread's range
    for (uint32_t i = lo; i < hi; i++) {
        if (!bad[i]) { g = (int)i; continue; }
        if (i == 0 || !bad[i - 1] || g < 0) continue;   // the head of a chain starts from its good predecessor's exit
        if (A.x_key[g] != 0u) continue;              // only a blank state repeats with the stride of the length cap
        const uint32_t xg = A.x_pos[g], a = A.begin + i * A.seg;
        const uint32_t steps = a > xg ? (a - xg + kMaxMatch - 1u) / kMaxMatch : 0u;
        start_pos[i] = xg + steps * kMaxMatch;
        start_key[i] = 0u;
    }
}

// =====================================================================================
// k_lz77_seq: MatchingType::Lazy with lazy_if_less_than < 3 (compression_options.rs:93 documents 0 as "never lazy
// match"; 1 and 2 are legal too).  For these values the reference's output depends on things the parallel
// pipeline deliberately leaves out because no other option set can observe them:
//   * longest_match may return length 2 (matching.rs:161-165) -- only through a *spurious* candidate, a chain
//     entry that still holds its initial / slid-out self index (chained_hash_table.rs:34-51,197-219): two
//     positions with equal 3-byte hash and two equal bytes have the third byte equal as well.  A length-2 result
//     is never emitted, but it satisfies `match_len >= lazy_if_less_than` (lz77.rs:374-377) and becomes the floor
//     of the next search;
//   * ignore_next is re-derived as prev_length >= lazy_if_less_than at the start of every process_chunk_lazy call
//     (lz77.rs:331), i.e. at every 32 KiB window and after every full token buffer; with 0 that is always true.
// So this kernel is the reference's own loop, one thread, with the head/prev chains of chained_hash_table.rs
// kept in global memory in absolute positions.  A chain value is either an absolute position (stale once it
// lies below the buffer origin of the window being processed: `slide` would have turned it into the slot's own
// index) or "the hash itself" -- what `prev[pos] = head[hash]` copies while head[hash] still holds its own
// index -- valid as buffer-relative position `hash` until the next slide.  Slow by construction (about the speed
// of one CPU core); it exists so that every legal CompressionOptions value gives the reference's bytes.
// One-shot streams only (begin == 0, closed piece).
// =====================================================================================
struct SeqTables { uint32_t* head; uint32_t* prev_val; uint32_t* prev_org; };   // 32768 entries each

__global__ void __launch_bounds__(32) k_lz77_seq_init(SeqTables t) {
    for (uint32_t i = threadIdx.x + blockIdx.x * blockDim.x; i < kWindow; i += blockDim.x * gridDim.x) {
        t.head[i] = kSeqNone; t.prev_val[i] = kSeqNone; t.prev_org[i] = kSeqNone;
    }
}

__global__ void __launch_bounds__(32) k_lz77_seq(const uint8_t* __restrict__ in, uint32_t n, Params prm, SeqTables t,
                                                uint32_t* __restrict__ tok, DevMeta* meta) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long nt = lz77_sequential(in, n, prm, t.head, t.prev_val, t.prev_org, tok);   // dfl_core.h
    meta->n_tokens = nt;
    meta->n_blocks = (uint32_t)(nt / kBlockTokens) + 1u;
    meta->end_pos = n;
    meta->end_key = 0;
}

// =====================================================================================
// token layout: exclusive scan of per-segment token counts (single CTA), then compaction
// =====================================================================================
__global__ void __launch_bounds__(1024) k_seg_scan(uint32_t n_seg, const uint32_t* e_tok, const uint32_t* x_tok,
                                                   uint32_t* seg_cnt, unsigned long long* seg_off, DevMeta* meta,
                                                   uint32_t n_carry, int open_piece, const uint32_t* x_pos,
                                                   const uint32_t* x_key, uint32_t begin, uint32_t init_key) {
    __shared__ unsigned long long part[1024];
    uint32_t per = (n_seg + blockDim.x - 1) / blockDim.x;
    uint32_t lo = threadIdx.x * per, hi = lo + per < n_seg ? lo + per : n_seg;
    unsigned long long s = 0;
    for (uint32_t i = lo; i < hi; i++) {
        uint32_t c = x_tok[i] - e_tok[i];
        seg_cnt[i] = c;
        s += c;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = n_carry;            // tokens carried over from the previous piece come first
        for (uint32_t t = 0; t < blockDim.x; t++) { unsigned long long v = part[t]; part[t] = run; run += v; }
        meta->n_tokens = run;
        // an open piece codes complete blocks only; a closed one always ends with a (possibly empty) last block
        meta->n_blocks = (uint32_t)(run / kBlockTokens) + (open_piece ? 0u : 1u);
        meta->end_pos = n_seg ? x_pos[n_seg - 1] : begin;
        meta->end_key = n_seg ? x_key[n_seg - 1] : init_key;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; i++) { seg_off[i] = run; run += seg_cnt[i]; }
}

// Copies every segment's tokens to their place in the stream and fills in what the parser left open: the byte of a
// literal (kTokLitAt) and the distance of a match whose record is final (kTokMatchAt).
__global__ void __launch_bounds__(128) k_compact(ParseArgs A, const uint32_t* __restrict__ seg_cnt,
                                                 const unsigned long long* __restrict__ seg_off, uint32_t* __restrict__ tok,
                                                 DevMeta* meta) {
    const uint32_t s = blockIdx.x;
    const uint32_t c = seg_cnt[s], e0 = A.e_tok[s];
    if (e0 + c > A.tok_cap) { if (threadIdx.x == 0) meta->err = 1; return; }
    const uint32_t* src = A.segtok + (size_t)s * A.tok_cap + e0;
    uint32_t* dst = tok + seg_off[s];
    const long long origin = (long long)A.begin + (long long)s * A.seg - (long long)parse_origin_back(A.warm);
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
        uint32_t t = src[i];
        if (t & kTokLitAt) t = tok_literal(A.in[origin + (long long)(t & kTokRelMask)]);
        else if (t & kTokMatchAt) {
            const long long q = origin + (long long)(t & kTokRelMask);
            const uint32_t rec = (t & kTokQuarter) ? A.Mq[q] : A.Mf[q];
            t = tok_match((t >> 14) & 0x1ffu, match_dist(rec));
        }
        dst[i] = t;
    }
}

__global__ void k_set_tokens(DevMeta* meta, unsigned long long n_tokens) {
    meta->n_tokens = n_tokens;
    meta->n_blocks = (uint32_t)(n_tokens / kBlockTokens) + 1u;
}

// =====================================================================================
// k_block_stats: literal/length and distance histograms + input bytes of one deflate block
// =====================================================================================
constexpr uint32_t kHistStride = 320;

__global__ void __launch_bounds__(256) k_block_stats(const uint32_t* __restrict__ tok, const DevMeta* meta,
                                                     uint32_t* __restrict__ hist, BlockCost* __restrict__ cost) {
    __shared__ uint32_t h[kHistStride];
    __shared__ uint32_t ws[33];
    const uint32_t b = blockIdx.x;
    if (b >= meta->n_blocks) return;
    const unsigned long long T = meta->n_tokens;
    const unsigned long long t0 = (unsigned long long)b * kBlockTokens;
    const uint32_t ntok = (uint32_t)((T - t0) < kBlockTokens ? (T - t0) : kBlockTokens);
    for (uint32_t i = threadIdx.x; i < kHistStride; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t bytes = 0;
    for (uint32_t i = threadIdx.x; i < ntok; i += blockDim.x) {
        uint32_t t = tok[t0 + i];
        uint32_t d = tok_dist(t);
        if (d) {
            uint32_t c, ne, ev;
            length_symbol(tok_lo(t), c, ne, ev);
            atomicAdd(&h[c], 1u);
            dist_symbol(d, c, ne, ev);
            atomicAdd(&h[kNumLL + c], 1u);
            bytes += tok_lo(t);
        } else {
            atomicAdd(&h[tok_lo(t)], 1u);
            bytes += 1u;
        }
    }
    uint32_t total;
    block_excl_scan(bytes, ws, total);
    if (threadIdx.x == 0) {
        h[kEob] = 1;   // output_writer.rs:83,107: exactly one end-of-block symbol per block
        cost[b].input_bytes = total;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kHistStride; i += blockDim.x) hist[(size_t)b * kHistStride + i] = h[i];
}

// =====================================================================================
// k_block_codes: one thread per block builds the three Huffman codes and the cost summary
// =====================================================================================
// One deflate block per warp.  The construction itself (sort, Moffat-Katajainen, limiter, header RLE:
// dfl_core.h build_block_codes) is a chain of short data-dependent loops and runs on lane 0, with every
// array it touches in shared memory -- as one thread per block with the arrays in local memory the stage
// took 2.1 ms for any input size (32 divergent lanes per warp, local-memory latency); loading the
// histograms, copying the tables out and summing the body size are done by all 32 lanes.
constexpr uint32_t kCodesWarps = 4;
struct CodesScratch {
    uint32_t ll[288];
    uint32_t dd[32];
    uint32_t scratch[288];
    uint32_t key_ll[288];      // leaves (freq << 9 | symbol) in ascending order
    uint32_t key_d[32];
    BlockCodes bc;
};

// Leaves of one alphabet, sorted: every lane ranks its own leaves against all of them (keys are unique, so
// the rank is the number of smaller keys).  freq[] has n_sym entries (a multiple of 32, zero padded).
__device__ __forceinline__ uint32_t warp_sorted_leaves(const uint32_t* freq, uint32_t n_sym, uint32_t* tmp, uint32_t* out) {
    const uint32_t lane = lane_id();
    uint32_t n = 0;
    for (uint32_t base = 0; base < n_sym; base += 32) {          // compact the non-zero ones, symbol order
        const uint32_t f = freq[base + lane];
        const uint32_t m = __ballot_sync(0xffffffffu, f != 0u);
        if (f != 0u) tmp[n + __popc(m & ((1u << lane) - 1u))] = (f << 9) | (base + lane);
        n += __popc(m);
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t k = tmp[i];
        uint32_t r = 0;
        for (uint32_t j = 0; j < n; j++) r += tmp[j] < k ? 1u : 0u;
        out[r] = k;
    }
    __syncwarp();
    return n;
}

__global__ void __launch_bounds__(32 * kCodesWarps) k_block_codes(const DevMeta* meta, const uint32_t* __restrict__ hist,
                                                                  BlockCost* __restrict__ cost,
                                                                  BlockTables* __restrict__ tables) {
    __shared__ CodesScratch sm[kCodesWarps];
    const uint32_t b = blockIdx.x * kCodesWarps + warp_id();
    if (b >= meta->n_blocks) return;                       // warp-uniform
    CodesScratch& S = sm[warp_id()];
    const uint32_t lane = lane_id();
    for (uint32_t i = lane; i < 288; i += 32) S.ll[i] = i < kNumLL ? hist[(size_t)b * kHistStride + i] : 0u;
    S.dd[lane] = lane < kNumDist ? hist[(size_t)b * kHistStride + kNumLL + lane] : 0u;
    __syncwarp();
    const uint32_t n_ll = warp_sorted_leaves(S.ll, 288, S.scratch, S.key_ll);
    const uint32_t n_d = warp_sorted_leaves(S.dd, 32, S.scratch, S.key_d);
    if (lane == 0) build_block_codes(S.ll, S.dd, cost[b].input_bytes, S.bc, S.scratch, S.key_ll, (int)n_ll, S.key_d, (int)n_d);
    __syncwarp();
    const BlockCodes& bc = S.bc;
    BlockTables& t = tables[b];
    unsigned long long body = 0;
    if (!bc.tiny) {
        for (uint32_t i = lane; i < 288; i += 32) { t.ll_code[i] = bc.ll_code[i]; t.ll_len[i] = bc.ll_len[i]; }
        t.d_code[lane] = bc.d_code[lane]; t.d_len[lane] = bc.d_len[lane];
        if (lane < 19) { t.cl_code[lane] = bc.cl_code[lane]; t.cl_len[lane] = bc.cl_len[lane]; }
        for (uint32_t i = lane; i < bc.n_hdr_sym; i += 32) t.hdr_sym[i] = bc.hdr_sym[i];
        for (uint32_t i = lane; i < bc.hlit; i += 32) {
            uint32_t eb = i >= 257u ? length_extra_bits_of_code(i - 257u) : 0u;
            body += (unsigned long long)S.ll[i] * (bc.ll_len[i] + eb);
        }
        if (lane < bc.hdist) body += (unsigned long long)S.dd[lane] * (bc.d_len[lane] + dist_extra_bits_of_code(lane));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) body += __shfl_xor_sync(0xffffffffu, body, d);
    if (lane == 0) {
        BlockCost c;
        c.dynamic_cost = bc.dynamic_cost; c.static_cost = bc.static_cost; c.stored_cost = bc.stored_cost;
        c.dynamic_bits = bc.dynamic_bits; c.fixed_bits = bc.fixed_bits; c.input_bytes = bc.input_bytes;
        c.tiny = bc.tiny;
        t.n_hdr_sym = bc.n_hdr_sym; t.hlit = bc.hlit; t.hdist = bc.hdist; t.used_hclens = bc.used_hclens;
        c.hdr_bits = bc.tiny ? 0u : (uint32_t)(bc.dynamic_bits - body);
        cost[b] = c;
    }
}

// =====================================================================================
// k_block_scan: block type for every possible bit alignment, then an exclusive scan of block sizes.
// The Stored candidate's cost depends on the bit position of the block start modulo 8
// (huffman_lengths.rs:113-124,262), so each thread first summarises its run of blocks as a map
// alignment -> bits consumed, the maps are composed across the CTA, and the run is replayed.
// =====================================================================================
__device__ __forceinline__ int choose_block_dev(const BlockCost& c, uint32_t pending, unsigned long long& bits) {
    if (c.tiny) { bits = 3ull + c.fixed_bits; return kFixed; }
    unsigned long long stored_len = c.stored_cost + stored_padding(pending);
    unsigned long long used = c.dynamic_cost < c.static_cost ? c.dynamic_cost : c.static_cost;
    if (stored_len < used) used = stored_len;
    if (used == c.static_cost) { bits = 3ull + c.fixed_bits; return kFixed; }
    if (used == stored_len) { bits = 3ull + stored_len; return kStored; }
    bits = 3ull + c.dynamic_bits;
    return kDynamic;
}

__global__ void __launch_bounds__(256) k_block_scan(DevMeta* meta, const BlockCost* __restrict__ cost, int* blk_type,
                                                    unsigned long long* blk_bit, unsigned long long* blk_in,
                                                    int sync_marker, uint32_t in_begin, uint32_t carry_bits_n,
                                                    uint32_t* out32, unsigned long long out_bit_base, unsigned long long out_cap) {
    __shared__ unsigned long long tot[256 * 8];   // bits consumed by a thread's run per entry alignment
    __shared__ unsigned long long inb[256];
    __shared__ unsigned long long entry_bit[256];
    const uint32_t nb = meta->n_blocks;
    const uint32_t per = (nb + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per < nb ? threadIdx.x * per : nb;
    const uint32_t hi = lo + per < nb ? lo + per : nb;
    unsigned long long ib = 0;
    unsigned long long abit[8];
#pragma unroll
    for (uint32_t a = 0; a < 8; a++) abit[a] = a;
    for (uint32_t b = lo; b < hi; b++) {
        const BlockCost c = cost[b];
        ib += c.input_bytes;
#pragma unroll
        for (uint32_t a = 0; a < 8; a++) {
            unsigned long long bits;
            choose_block_dev(c, (uint32_t)(abit[a] & 7ull), bits);
            abit[a] += bits;
        }
    }
#pragma unroll
    for (uint32_t a = 0; a < 8; a++) tot[threadIdx.x * 8 + a] = abit[a] - a;
    inb[threadIdx.x] = ib;
    __syncthreads();
    if (threadIdx.x == 0) {
        // a piece that continues a stream starts inside the byte the previous piece left incomplete
        unsigned long long bit = carry_bits_n, in_off = in_begin;
        for (uint32_t t = 0; t < blockDim.x; t++) {
            entry_bit[t] = bit;
            bit += tot[t * 8 + (bit & 7ull)];
            unsigned long long v = inb[t];
            inb[t] = in_off;
            in_off += v;
        }
        meta->in_coded_end = in_off;
        if (sync_marker) {            // compress.rs:258-261: empty stored block 00 00 FF FF, byte aligned
            bit += 3ull;
            bit = (bit + 7ull) & ~7ull;
            bit += 32ull;
        }
        meta->stream_bits = bit;
        meta->stream_bytes = (bit + 7ull) >> 3;
    }
    __syncthreads();
    // k_pack writes whole words; the words two blocks share (and what lies behind the last block: sync marker,
    // padding) are combined with atomicOr and must start out as zero.  Nothing else is cleared.
    const bool fits = out32 != nullptr && (out_bit_base >> 3) + meta->stream_bytes + 16ull <= out_cap;
    if (fits && threadIdx.x == 0) {
        const unsigned long long w0 = (out_bit_base + entry_bit[0]) >> 5;   // first word of the stream (may hold container bytes too)
        out32[w0] = 0u;
    }
    unsigned long long bit = entry_bit[threadIdx.x], in_off = inb[threadIdx.x];
    uint32_t n_st = 0, n_fx = 0;
    for (uint32_t b = lo; b < hi; b++) {
        unsigned long long bits;
        const BlockCost c = cost[b];
        int t = choose_block_dev(c, (uint32_t)(bit & 7ull), bits);
        blk_type[b] = t;
        blk_bit[b] = bit;
        blk_in[b] = in_off;
        bit += bits;
        in_off += c.input_bytes;
        n_st += (t == kStored);
        n_fx += (t == kFixed);
        if (fits) {
            const unsigned long long w = (out_bit_base + bit) >> 5;       // the word this block shares with the next one
            out32[w] = 0u;
            if (b + 1 == nb) { out32[w + 1] = 0u; out32[w + 2] = 0u; }   // 3 header bits, padding and 00 00 of a sync marker
        }
        if (b + 1 == nb) blk_bit[nb] = bit;
    }
    if (n_st) atomicAdd(&meta->n_stored, n_st);
    if (n_fx) atomicAdd(&meta->n_fixed, n_fx);
}

// =====================================================================================
// k_pack: one CTA per deflate block.  The block's bits are assembled in shared memory, a tile of 1024 items
// (header fields, tokens, end-of-block) at a time: per-item bit widths -> CTA prefix sum -> every item ORed into
// a shared word buffer -> the completed words leave as coalesced 16-byte stores.  A block owns every output
// word that lies completely inside its bit range; only the word it shares with the block in front of it and
// the one it shares with the block behind it are combined with atomicOr (k_block_scan has zeroed those).
// Stored blocks are a byte-shifted copy of the input, one output word per thread.
// (encoder_state.rs:58-105, huffman_lengths.rs:290-369, bitstream.rs:76-106, stored_block.rs:13-40, compress.rs:59-77)
// =====================================================================================
constexpr uint32_t kPackThreads = 256;
constexpr uint32_t kPackItems = 4;                                // consecutive items per thread and tile
constexpr uint32_t kPackTile = kPackThreads * kPackItems;
constexpr uint32_t kPackWords = kPackTile * 48u / 32u + 8u;       // an item is at most 48 bits

// word `g` of the output: plain store if the block owns it, atomicOr if a neighbouring block writes into it too
__device__ __forceinline__ void pack_store_word(uint32_t* out32, unsigned long long g, uint32_t v, unsigned long long bs,
                                                unsigned long long be) {
    const bool shared = (g == (bs >> 5) && (bs & 31ull)) || (g == (be >> 5) && (be & 31ull));
    if (shared) { if (v) atomicOr(&out32[g], v); }
    else out32[g] = v;
}

__global__ void __launch_bounds__(kPackThreads)
k_pack(const uint8_t* __restrict__ in, const uint32_t* __restrict__ tok, DevMeta* meta, const BlockCost* __restrict__ cost,
       const BlockTables* __restrict__ tables, const int* __restrict__ blk_type,
       const unsigned long long* __restrict__ blk_bit, const unsigned long long* __restrict__ blk_in,
       uint32_t* __restrict__ out32, unsigned long long out_bit_base, int final_block, unsigned long long out_cap) {
    __shared__ uint32_t ll_cl[288];   // code | len << 16
    __shared__ uint32_t d_cl[32];
    __shared__ uint32_t c_cl[19];
    __shared__ uint32_t ws[33];
    __shared__ uint32_t sw[kPackWords];
    const uint32_t b = blockIdx.x;
    const uint32_t nb = meta->n_blocks;
    if (b >= nb) return;
    if ((out_bit_base >> 3) + meta->stream_bytes + 16ull > out_cap) {   // would not fit: write nothing
        if (threadIdx.x == 0) meta->err = 100;
        return;
    }
    const unsigned long long T = meta->n_tokens;
    const unsigned long long t0 = (unsigned long long)b * kBlockTokens;
    const uint32_t ntok = (uint32_t)((T - t0) < kBlockTokens ? (T - t0) : kBlockTokens);
    const int type = blk_type[b];
    const int last = (b + 1 == nb) && final_block;
    const unsigned long long bs = blk_bit[b] + out_bit_base, be = blk_bit[b + 1] + out_bit_base;

    if (type == kStored) {
        // bytes of the block: [B0] the 3 header bits (BFINAL only ever sets one of them) and padding, then per chunk
        // LEN, ~LEN, data; every chunk after the first is preceded by its own header byte (stored_block.rs:13-40)
        const unsigned long long pos0 = blk_in[b], nbytes = cost[b].input_bytes;
        const unsigned long long k = (nbytes - 1ull) / kMaxStored + 1ull;           // chunks (nbytes > 0 for a stored block)
        const unsigned long long B0 = bs >> 3, D0 = (bs + 3ull + 7ull) >> 3, Bend = be >> 3;
        const uint32_t stride = kMaxStored + 5u;
        for (unsigned long long g = (bs >> 5) + threadIdx.x; g <= ((be - 1ull) >> 5); g += blockDim.x) {
            uint32_t word = 0;
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) {
                const unsigned long long jb = g * 4ull + i;
                uint32_t v = 0;
                if (jb >= D0 && jb < Bend) {
                    const unsigned long long rel = jb - D0, c = rel / stride;
                    const uint32_t r = (uint32_t)(rel - c * stride);
                    const uint32_t len_c = c + 1ull < k ? kMaxStored : (uint32_t)(nbytes - c * kMaxStored);
                    if (r < 4u) { const uint32_t lv = r < 2u ? len_c : ~len_c; v = (lv >> (8u * (r & 1u))) & 0xffu; }
                    else if (r < 4u + len_c) v = in[pos0 + c * kMaxStored + (r - 4u)];
                    else v = (last && c + 2ull == k) ? 1u : 0u;                      // header byte of chunk c + 1
                } else if (jb == B0 && jb < D0) {
                    v = (last && k == 1ull) ? (1u << (uint32_t)(bs & 7ull)) : 0u;
                }
                word |= v << (8u * i);
            }
            pack_store_word(out32, g, word, bs, be);
        }
        return;
    }

    const BlockTables& tb = tables[b];
    if (type == kFixed) {
        // huffman_table.rs:32-42 fixed lengths -> canonical codes, bit-reversed (RFC 1951 3.2.6)
        for (uint32_t s = threadIdx.x; s < 288; s += blockDim.x) {
            uint32_t len = fixed_ll_length(s);
            uint32_t code = s < 144u ? 0x30u + s : (s < 256u ? 0x190u + (s - 144u) : (s < 280u ? s - 256u : 0xc0u + (s - 280u)));
            ll_cl[s] = reverse_bits(code, len) | (len << 16);
        }
        if (threadIdx.x < 32) d_cl[threadIdx.x] = reverse_bits(threadIdx.x, 5) | (5u << 16);
    } else {
        for (uint32_t s = threadIdx.x; s < 288; s += blockDim.x) ll_cl[s] = tb.ll_code[s] | ((uint32_t)tb.ll_len[s] << 16);
        if (threadIdx.x < 32) d_cl[threadIdx.x] = tb.d_code[threadIdx.x] | ((uint32_t)tb.d_len[threadIdx.x] << 16);
        if (threadIdx.x < 19) c_cl[threadIdx.x] = tb.cl_code[threadIdx.x] | ((uint32_t)tb.cl_len[threadIdx.x] << 16);
    }
    // items of the block: header fields (encoder_state.rs:85-99, huffman_lengths.rs:290-369), tokens, end-of-block
    const uint32_t hclens = type == kDynamic ? tb.used_hclens : 0u;
    const uint32_t n_hsym = type == kDynamic ? tb.n_hdr_sym : 0u;
    const uint32_t n_hdr = type == kDynamic ? 4u + hclens + n_hsym : 1u;
    const uint32_t n_items = n_hdr + ntok + 1u;
    if (threadIdx.x == 0) sw[0] = 0u;
    __syncthreads();

    unsigned long long bitpos = bs;
    for (uint32_t base = 0; base < n_items; base += kPackTile) {
        unsigned long long v[kPackItems];
        uint32_t nbv[kPackItems];
        uint32_t sum = 0;
#pragma unroll
        for (uint32_t q = 0; q < kPackItems; q++) {
            const uint32_t i = base + threadIdx.x * kPackItems + q;
            unsigned long long x = 0;
            uint32_t nbits = 0;
            if (i < n_hdr) {
                if (type == kFixed) { x = last ? 3ull : 2ull; nbits = 3; }
                else if (i == 0u) { x = last ? 5ull : 4ull; nbits = 3; }
                else if (i == 1u) { x = tb.hlit - 257u; nbits = 5; }
                else if (i == 2u) { x = tb.hdist - 1u; nbits = 5; }
                else if (i == 3u) { x = hclens - 4u; nbits = 4; }
                else if (i < 4u + hclens) {
                    const uint32_t order = 0u;   // 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15, 5 bits each, packed below
                    (void)order;
                    const unsigned long long ord_lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) |
                                                      (9ull << 30) | (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
                    const unsigned long long ord_hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
                    const uint32_t oi = i - 4u;
                    const uint32_t sym = (uint32_t)((oi < 12u ? ord_lo >> (5u * oi) : ord_hi >> (5u * (oi - 12u))) & 31ull);
                    x = c_cl[sym] >> 16; nbits = 3;
                } else {
                    const uint32_t hs = tb.hdr_sym[i - 4u - hclens];
                    const uint32_t sym = hs & 31u, rep = hs >> 8;
                    const uint32_t cl = c_cl[sym];
                    x = cl & 0xffffu; nbits = cl >> 16;
                    if (sym == 16u) { x |= (unsigned long long)(rep - 3u) << nbits; nbits += 2; }
                    else if (sym == 17u) { x |= (unsigned long long)(rep - 3u) << nbits; nbits += 3; }
                    else if (sym == 18u) { x |= (unsigned long long)(rep - 11u) << nbits; nbits += 7; }
                }
            } else if (i < n_hdr + ntok) {
                const uint32_t t = tok[t0 + (i - n_hdr)];
                const uint32_t d = tok_dist(t);
                if (d) {
                    uint32_t c, ne, ev;
                    length_symbol(tok_lo(t), c, ne, ev);
                    uint32_t cl = ll_cl[c];
                    x = cl & 0xffffu; nbits = cl >> 16;
                    x |= (unsigned long long)ev << nbits; nbits += ne;
                    dist_symbol(d, c, ne, ev);
                    cl = d_cl[c];
                    x |= (unsigned long long)(cl & 0xffffu) << nbits; nbits += cl >> 16;
                    x |= (unsigned long long)ev << nbits; nbits += ne;
                } else {
                    const uint32_t cl = ll_cl[tok_lo(t)];
                    x = cl & 0xffffu; nbits = cl >> 16;
                }
            } else if (i == n_hdr + ntok) {
                const uint32_t cl = ll_cl[kEob];
                x = cl & 0xffffu; nbits = cl >> 16;
            }
            v[q] = x; nbv[q] = nbits; sum += nbits;
        }
        uint32_t total;
        const uint32_t ex = block_excl_scan(sum, ws, total);
        const unsigned long long wbase = bitpos >> 5, end = bitpos + total;
        const uint32_t n_words = (uint32_t)(((end + 31ull) >> 5) - wbase);
        for (uint32_t i = threadIdx.x + 1u; i <= n_words; i += blockDim.x) sw[i] = 0u;   // word 0 carries the previous tile's tail
        __syncthreads();
        uint32_t off = (uint32_t)(bitpos & 31ull) + ex;
#pragma unroll
        for (uint32_t q = 0; q < kPackItems; q++) {
            if (nbv[q]) {
                const uint32_t wi = off >> 5, sh = off & 31u;
                const unsigned long long lo = v[q] << sh;
                const uint32_t x0 = (uint32_t)lo, x1 = (uint32_t)(lo >> 32), x2 = sh ? (uint32_t)(v[q] >> (64u - sh)) : 0u;
                if (x0) atomicOr(&sw[wi], x0);
                if (x1) atomicOr(&sw[wi + 1], x1);
                if (x2) atomicOr(&sw[wi + 2], x2);
                off += nbv[q];
            }
        }
        __syncthreads();
        // completed words leave; 16 bytes at a time where the output address allows it
        const uint32_t n_done = (uint32_t)((end >> 5) - wbase);
        {
            const uint32_t head0 = (uint32_t)((4ull - (wbase & 3ull)) & 3ull);
            const uint32_t head = head0 < n_done ? head0 : n_done;
            const uint32_t n_vec = (n_done - head) >> 2, tail = head + 4u * n_vec;
            if (threadIdx.x < head) pack_store_word(out32, wbase + threadIdx.x, sw[threadIdx.x], bs, be);
            if (threadIdx.x >= 32u && threadIdx.x - 32u < n_done - tail)
                pack_store_word(out32, wbase + tail + (threadIdx.x - 32u), sw[tail + (threadIdx.x - 32u)], bs, be);
            uint4* out4 = reinterpret_cast<uint4*>(out32 + wbase + head);
            for (uint32_t vi = threadIdx.x; vi < n_vec; vi += blockDim.x) {
                const uint32_t i0 = head + 4u * vi;
                const unsigned long long g = wbase + i0;
                if (g == (bs >> 5) && (bs & 31ull)) {          // the block's first word sits at a 16-byte boundary and is shared
                    pack_store_word(out32, g, sw[i0], bs, be);
                    out32[g + 1] = sw[i0 + 1]; out32[g + 2] = sw[i0 + 2]; out32[g + 3] = sw[i0 + 3];
                } else {
                    out4[vi] = make_uint4(sw[i0], sw[i0 + 1], sw[i0 + 2], sw[i0 + 3]);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) sw[0] = (end & 31ull) ? sw[n_done] : 0u;
        __syncthreads();
        bitpos = end;
    }
    if (threadIdx.x == 0) {
        if (bitpos != be) meta->err = 4;                        // the scan and the packer must agree
        else if (bitpos & 31ull) pack_store_word(out32, bitpos >> 5, sw[0], bs, be);
    }
}

// =====================================================================================
// Adler-32 (RFC 1950): per 64 KiB chunk sums, then an ordered combine in one CTA
// =====================================================================================
__global__ void __launch_bounds__(256) k_adler32_chunks(const uint8_t* __restrict__ in, unsigned long long n,
                                                        unsigned long long* __restrict__ part) {
    __shared__ unsigned long long sa[256], sb[256];
    const unsigned long long c0 = (unsigned long long)blockIdx.x * kAdlerChunk;
    const uint32_t L = (uint32_t)((n - c0) < kAdlerChunk ? (n - c0) : kAdlerChunk);
    const uint8_t* p = in + c0;
    unsigned long long a = 0, bsum = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
    uint32_t vec_end = aligned ? (L & ~15u) : 0u;
    for (uint32_t i = threadIdx.x * 16u; i < vec_end; i += blockDim.x * 16u) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(p + i));
        uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t byte = (wv[k] >> (8 * q)) & 0xffu;
                a += byte;
                bsum += (unsigned long long)(L - (i + 4 * k + q)) * byte;
            }
        }
    }
    for (uint32_t i = vec_end + threadIdx.x; i < L; i += blockDim.x) {
        uint32_t byte = p[i];
        a += byte;
        bsum += (unsigned long long)(L - i) * byte;
    }
    sa[threadIdx.x] = a; sb[threadIdx.x] = bsum;
    __syncthreads();
    for (uint32_t s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t A = (uint32_t)((1ull + sa[0]) % kAdlerMod);
        uint32_t B = (uint32_t)(((unsigned long long)L + sb[0]) % kAdlerMod);
        part[2ull * blockIdx.x] = ((unsigned long long)B << 16) | A;
        part[2ull * blockIdx.x + 1] = L;
    }
}

__global__ void __launch_bounds__(1024) k_adler32_combine(const unsigned long long* __restrict__ part, uint32_t n_chunks,
                                                          DevMeta* meta) {
    __shared__ uint32_t ad[1024];
    __shared__ unsigned long long ln[1024];
    uint32_t per = (n_chunks + blockDim.x - 1) / blockDim.x;
    uint32_t lo = threadIdx.x * per, hi = lo + per < n_chunks ? lo + per : n_chunks;
    uint32_t a = 1;
    unsigned long long len = 0;
    for (uint32_t i = lo; i < hi; i++) {
        a = adler32_combine(a, (uint32_t)part[2ull * i], part[2ull * i + 1]);
        len += part[2ull * i + 1];
    }
    ad[threadIdx.x] = a; ln[threadIdx.x] = len;
    __syncthreads();
    for (uint32_t s = 1; s < blockDim.x; s <<= 1) {
        uint32_t i = threadIdx.x;
        if ((i & (2 * s - 1)) == 0 && i + s < blockDim.x) {
            ad[i] = adler32_combine(ad[i], ad[i + s], ln[i + s]);
            ln[i] += ln[i + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) meta->adler = ad[0];
}

// =====================================================================================
// CRC-32 (gzip trailer; lib.rs:257-265, writer.rs:408-426): per 64 KiB chunk every thread checksums
// 256 bytes with a byte table, the 256 results are combined in a tree (the right operand of a full
// subtree has a power-of-two length, so its shift operator is one table entry), then one CTA
// combines the chunks in order.
// =====================================================================================
__global__ void __launch_bounds__(256) k_crc32_chunks(const uint8_t* __restrict__ in, unsigned long long n,
                                                      unsigned long long* __restrict__ part) {
    __shared__ uint32_t tab[256];
    __shared__ uint32_t sc[256];
    __shared__ uint32_t sl[256];
    {
        uint32_t c = threadIdx.x;
#pragma unroll
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ kCrcPoly : c >> 1;
        tab[threadIdx.x] = c;
    }
    __syncthreads();
    const unsigned long long c0 = (unsigned long long)blockIdx.x * kAdlerChunk;
    const uint32_t L = (uint32_t)((n - c0) < kAdlerChunk ? (n - c0) : kAdlerChunk);
    const uint32_t lo = threadIdx.x * 256u < L ? threadIdx.x * 256u : L;
    const uint32_t hi = lo + 256u < L ? lo + 256u : L;
    const uint8_t* p = in + c0;
    uint32_t c = 0xffffffffu;
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) && hi - lo == 256u;
    if (aligned) {
#pragma unroll 4
        for (uint32_t i = lo; i < hi; i += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + i));
            const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int q = 0; q < 4; q++) c = tab[(c ^ (wv[k] >> (8 * q))) & 0xffu] ^ (c >> 8);
            }
        }
    } else {
        for (uint32_t i = lo; i < hi; i++) c = tab[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    }
    sc[threadIdx.x] = c ^ 0xffffffffu;
    sl[threadIdx.x] = hi - lo;
    __syncthreads();
    for (uint32_t s = 1; s < 256; s <<= 1) {
        const uint32_t i = threadIdx.x;
        if ((i & (2 * s - 1)) == 0) {
            sc[i] = crc32_combine(sc[i], sc[i + s], sl[i + s]);
            sl[i] += sl[i + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part[2ull * blockIdx.x] = sc[0];
        part[2ull * blockIdx.x + 1] = L;
    }
}

__global__ void __launch_bounds__(1024) k_crc32_combine(const unsigned long long* __restrict__ part, uint32_t n_chunks,
                                                        DevMeta* meta) {
    __shared__ uint32_t cr[1024];
    __shared__ unsigned long long ln[1024];
    uint32_t per = (n_chunks + blockDim.x - 1) / blockDim.x;
    uint32_t lo = threadIdx.x * per, hi = lo + per < n_chunks ? lo + per : n_chunks;
    uint32_t c = 0;
    unsigned long long len = 0;
    for (uint32_t i = lo; i < hi; i++) {
        c = crc32_combine(c, (uint32_t)part[2ull * i], part[2ull * i + 1]);
        len += part[2ull * i + 1];
    }
    cr[threadIdx.x] = c; ln[threadIdx.x] = len;
    __syncthreads();
    for (uint32_t s = 1; s < blockDim.x; s <<= 1) {
        uint32_t i = threadIdx.x;
        if ((i & (2 * s - 1)) == 0 && i + s < blockDim.x) {
            cr[i] = crc32_combine(cr[i], cr[i + s], ln[i + s]);
            ln[i] += ln[i + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) meta->crc = cr[0];
}

// =====================================================================================
// k_finalize: sync marker bytes, container header and trailer (zlib.rs:59-62, lib.rs:192-196)
// =====================================================================================
__global__ void k_finalize(DevMeta* meta, uint8_t* out, unsigned long long out_cap, uint32_t hdr_bytes, int wrap,
                           int sync_marker, int write_trailer, uint32_t isize, uint32_t carry_bits_n, uint32_t carry_bits_v) {
    unsigned long long end = hdr_bytes + meta->stream_bytes;
    const unsigned long long trailer = write_trailer ? (wrap == 1 ? 4ull : (wrap == 2 ? 8ull : 0ull)) : 0ull;
    // the packer writes whole 16-byte units: the buffer must hold the container rounded up plus one unit of slack;
    // on overflow that is the size reported, so that a caller who retries with exactly it succeeds
    if (end + 16ull > out_cap) { meta->err = 100; meta->out_bytes = ((end + trailer + 16ull) + 15ull) & ~15ull; return; }
    if (carry_bits_n) out[hdr_bytes] |= (uint8_t)carry_bits_v;   // the bits the previous piece left in its last byte
    if (sync_marker && end >= 4 && end <= out_cap) {
        out[end - 2] = 0xff;
        out[end - 1] = 0xff;
    }
    if (wrap == 1) {   // zlib
        if (hdr_bytes == 2) { out[0] = 0x78; out[1] = 0x9c; }
        if (write_trailer && end + 4 <= out_cap) {
            uint32_t a = meta->adler;
            out[end] = (uint8_t)(a >> 24); out[end + 1] = (uint8_t)(a >> 16); out[end + 2] = (uint8_t)(a >> 8); out[end + 3] = (uint8_t)a;
        }
        if (write_trailer) end += 4;
    }
    if (wrap == 2 && write_trailer) {   // gzip: CRC-32 then ISIZE, little endian (lib.rs:260-265, writer.rs:408-426)
        if (end + 8 <= out_cap) {
            const uint32_t c = meta->crc;
#pragma unroll
            for (int k = 0; k < 4; k++) { out[end + k] = (uint8_t)(c >> (8 * k)); out[end + 4 + k] = (uint8_t)(isize >> (8 * k)); }
        }
        end += 8;
    }
    meta->out_bytes = end;
}

// =====================================================================================
// host-side launchers
// =====================================================================================
uint32_t max_blocks_for(uint32_t n_payload) { return n_payload / kBlockTokens + 2u; }

static ParseArgs make_parse_args(const EncodeJob& j, Buffers& b) {
    ParseArgs A;
    A.in = j.d_in; A.n = j.n; A.begin = j.begin; A.prm = j.prm; A.Mf = b.Mf; A.Mq = b.Mq; A.segtok = b.segtok;
    A.K = b.K; A.off = b.off; A.K2 = b.K2;
    A.e_pos = b.seg_e_pos; A.e_key = b.seg_e_key; A.e_tok = b.seg_e_tok;
    A.x_pos = b.seg_x_pos; A.x_key = b.seg_x_key; A.x_tok = b.seg_x_tok;
    const ParseGeom g = parse_geom(j.n - j.begin, j.prm.mode);
    A.seg = g.seg; A.warm = g.warm; A.tok_cap = parse_tok_cap(g);
    A.end = j.parse_end; A.init_key = j.init_key;
    A.n_seg = j.parse_end > j.begin ? (uint32_t)parse_n_seg(j.parse_end - j.begin, g) : 0u;
    return A;
}

// cudaFuncSetAttribute applies to the current device: once per device that is used
static cudaError_t ensure_attrs() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(k_window_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_window_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_match<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_match<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) done[dev] = true;
    return cudaSuccess;
}

// Window ranges: the sort needs windows [first_sort_window(j), n_windows(j)), the match stage
// [first_match_window(j), n_windows(j)).  Both can be issued in pieces [w_lo, w_hi) as the input
// arrives (dfl_compress overlaps the host-to-device copy with them); matching window w needs the
// sorted lists of w - 1 and w.
uint32_t n_windows(const EncodeJob& j) { return (j.n + kWindow - 1) / kWindow; }
uint32_t first_match_window(const EncodeJob& j) { return j.begin / kWindow; }
uint32_t first_sort_window(const EncodeJob& j) { uint32_t w = j.begin / kWindow; return w > 0 ? w - 1 : 0; }

cudaError_t launch_window_sort(const EncodeJob& j, Buffers& b, cudaStream_t st, uint32_t w_lo, uint32_t w_hi) {
    cudaError_t e = ensure_attrs();
    if (e != cudaSuccess) return e;
    if (w_hi <= w_lo) return cudaSuccess;
    if (one_candidate(j.prm))
        k_window_sort<true><<<w_hi - w_lo, kSortThreads, kSortSmem, st>>>(j.d_in, j.n, w_lo, b.K, b.K2, b.off, b.Mf);
    else
        k_window_sort<false><<<w_hi - w_lo, kSortThreads, kSortSmem, st>>>(j.d_in, j.n, w_lo, b.K, b.K2, b.off, b.Mf);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_match(const EncodeJob& j, Buffers& b, cudaStream_t st, uint32_t w_lo, uint32_t w_hi) {
    cudaError_t e = ensure_attrs();
    if (e != cudaSuccess) return e;
    if (w_hi <= w_lo) return cudaSuccess;
    if (one_candidate(j.prm)) {   // the sort has settled everything but the first position of every bucket
        k_match_first<<<(w_hi - w_lo) * 4u, 256, 0, st>>>(j.d_in, j.n, w_lo, b.off, b.Mf);
        DFL_LAUNCH_CHECK();
        return cudaSuccess;
    }
    // enough CTAs to fill the chip even for a handful of windows
    const uint32_t n_w = w_hi - w_lo;
    uint32_t parts = (148u * DFL_MATCH_CTAS + n_w - 1) / n_w / (j.peers ? j.peers : 1u);
    parts = parts > 16u ? 16u : (parts < 1u ? 1u : parts);
    const dim3 grid(n_w, parts);
    if (j.prm.need_quarter)
        k_match<true><<<grid, kMatchThreads, kMatchSmem, st>>>(j.d_in, j.n, j.begin, w_lo, j.prm, b.K, b.off, b.Mf, b.Mq);
    else
        k_match<false><<<grid, kMatchThreads, kMatchSmem, st>>>(j.d_in, j.n, j.begin, w_lo, j.prm, b.K, b.off, b.Mf, b.Mq);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_parse(const EncodeJob& j, Buffers& b, cudaStream_t st) {
    ParseArgs A = make_parse_args(j, b);
    if (A.n_seg == 0) return cudaSuccess;
    uint32_t grid = (A.n_seg + kParseThreads - 1) / kParseThreads;
    k_parse_spec<<<grid, kParseThreads, 0, st>>>(A);
    DFL_LAUNCH_CHECK();
    if (A.n_seg == 1) return cudaSuccess;
    for (uint32_t r = 0, nr = repair_rounds(j.n - j.begin); r < nr; r++) {
        k_reset_bad<<<1, 1, 0, st>>>(b.meta);
        DFL_LAUNCH_CHECK();
        k_parse_verify<<<grid, 128, 0, st>>>(A, b.seg_bad, b.seg_start_pos, b.seg_start_key, b.meta);
        DFL_LAUNCH_CHECK();
        if (r & 1u) {   // the round before has given every chain of bad segments a true head; now the rest of each chain
            k_chain_predict<<<1, 1024, 0, st>>>(A, b.seg_bad, b.seg_start_pos, b.seg_start_key);
            DFL_LAUNCH_CHECK();
        }
        k_parse_repair<<<grid, kParseThreads, 0, st>>>(A, b.seg_bad, b.seg_start_pos, b.seg_start_key, b.meta);
        DFL_LAUNCH_CHECK();
    }
    k_reset_bad<<<1, 1, 0, st>>>(b.meta);
    DFL_LAUNCH_CHECK();
    k_parse_verify<<<grid, 128, 0, st>>>(A, b.seg_bad, b.seg_start_pos, b.seg_start_key, b.meta);
    DFL_LAUNCH_CHECK();
    k_parse_repair_seq<<<1, 32, 0, st>>>(A, b.meta);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_lz77_seq(const EncodeJob& j, Buffers& b, cudaStream_t st) {
    SeqTables t{b.seq_tab, b.seq_tab + kWindow, b.seq_tab + 2 * kWindow};
    k_lz77_seq_init<<<32, 32, 0, st>>>(t);
    DFL_LAUNCH_CHECK();
    k_lz77_seq<<<1, 32, 0, st>>>(j.d_in, j.n, j.prm, t, b.tok, b.meta);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_token_layout(const EncodeJob& j, Buffers& b, cudaStream_t st) {
    const ParseGeom g = parse_geom(j.n - j.begin, j.prm.mode);
    uint32_t n_seg = j.parse_end > j.begin ? (uint32_t)parse_n_seg(j.parse_end - j.begin, g) : 0u;
    k_seg_scan<<<1, 1024, 0, st>>>(n_seg, b.seg_e_tok, b.seg_x_tok, b.seg_cnt, b.seg_off, b.meta, j.n_carry_tok, j.open_piece,
                                   b.seg_x_pos, b.seg_x_key, j.begin, j.init_key);
    DFL_LAUNCH_CHECK();
    if (n_seg > 0) {
        ParseArgs A = make_parse_args(j, b);
        k_compact<<<n_seg, 128, 0, st>>>(A, b.seg_cnt, b.seg_off, b.tok, b.meta);
        DFL_LAUNCH_CHECK();
    }
    return cudaSuccess;
}

cudaError_t launch_block_stats(const EncodeJob& j, Buffers& b, cudaStream_t st) {
    const uint32_t* tok = j.d_tokens_override ? j.d_tokens_override : b.tok;
    if (j.d_tokens_override) {
        k_set_tokens<<<1, 1, 0, st>>>(b.meta, j.n_tokens_override);
        DFL_LAUNCH_CHECK();
    }
    k_block_stats<<<max_blocks_for(j.n - j.begin + j.n_carry_tok), 256, 0, st>>>(tok, b.meta, b.hist, b.cost);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_block_codes(const EncodeJob& j, Buffers& b, cudaStream_t st) {
    uint32_t mb = max_blocks_for(j.n - j.begin + j.n_carry_tok);
    k_block_codes<<<(mb + kCodesWarps - 1) / kCodesWarps, 32 * kCodesWarps, 0, st>>>(b.meta, b.hist, b.cost, b.tables);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_block_scan(const EncodeJob& j, Buffers& b, cudaStream_t st) {
    k_block_scan<<<1, 256, 0, st>>>(b.meta, b.cost, b.blk_type, b.blk_bit, b.blk_in, j.sync_marker,
                                    j.n_carry_tok ? j.carry_in_pos : j.begin, j.carry_bits_n, reinterpret_cast<uint32_t*>(j.d_out),
                                    (unsigned long long)j.hdr_bytes * 8ull, (unsigned long long)j.out_cap);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_pack(const EncodeJob& j, Buffers& b, cudaStream_t st) {
    const uint32_t* tok = j.d_tokens_override ? j.d_tokens_override : b.tok;
    k_pack<<<max_blocks_for(j.n - j.begin + j.n_carry_tok), kPackThreads, 0, st>>>(j.d_in, tok, b.meta, b.cost, b.tables, b.blk_type, b.blk_bit,
                                                                    b.blk_in, reinterpret_cast<uint32_t*>(j.d_out),
                                                                    (unsigned long long)j.hdr_bytes * 8ull, j.final_block,
                                                                    (unsigned long long)j.out_cap);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_adler32(const uint8_t* d_in, size_t n, Buffers& b, cudaStream_t st) {
    uint32_t n_chunks = (uint32_t)((n + kAdlerChunk - 1) / kAdlerChunk);
    if (n_chunks > 0) {
        k_adler32_chunks<<<n_chunks, 256, 0, st>>>(d_in, (unsigned long long)n, b.adler_part);
        DFL_LAUNCH_CHECK();
    }
    k_adler32_combine<<<1, 1024, 0, st>>>(b.adler_part, n_chunks, b.meta);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_crc32(const uint8_t* d_in, size_t n, Buffers& b, cudaStream_t st) {
    uint32_t n_chunks = (uint32_t)((n + kAdlerChunk - 1) / kAdlerChunk);
    if (n_chunks > 0) {
        k_crc32_chunks<<<n_chunks, 256, 0, st>>>(d_in, (unsigned long long)n, b.adler_part);   // zlib and gzip never coincide
        DFL_LAUNCH_CHECK();
    }
    k_crc32_combine<<<1, 1024, 0, st>>>(b.adler_part, n_chunks, b.meta);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_finalize(const EncodeJob& j, Buffers& b, int wrap, cudaStream_t st) {
    k_finalize<<<1, 1, 0, st>>>(b.meta, j.d_out, (unsigned long long)j.out_cap, j.hdr_bytes, wrap, j.sync_marker,
                                j.final_block, j.isize, j.carry_bits_n, j.carry_bits_v);
    DFL_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace dfl
