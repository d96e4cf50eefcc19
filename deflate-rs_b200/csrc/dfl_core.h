// dfl_core.h -- position-independent logic of the B200 DEFLATE encoder, written once as
// host+device inline functions.  The CUDA kernels (dfl_kernels.cu) call these from device code;
// tests/model/dfl_model.cpp calls the same functions from a sequential CPU harness so that the
// algorithms can be checked against the oracle in a container without a GPU.  Nothing here touches
// memory it is not handed, and nothing here is a CPU fallback: the product library only reaches
// these functions from kernels.
//
// Reference behaviour being restated (file:line under /root/reference/src) is cited per function.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DFL_HD __host__ __device__ __forceinline__
#else
#define DFL_HD inline
#endif

namespace dfl {

// ---------------------------------------------------------------- constants
constexpr uint32_t kWindow = 32768;          // chained_hash_table.rs:1, huffman_table.rs:25
constexpr uint32_t kWindowMask = kWindow - 1;
constexpr uint32_t kMinMatch = 3;            // huffman_table.rs:20
constexpr uint32_t kMaxMatch = 258;          // huffman_table.rs:21
constexpr uint32_t kBlockTokens = 31744;     // output_writer.rs:19 (MAX_BUFFER_LENGTH)
constexpr uint32_t kNumLL = 286;             // huffman_table.rs:13
constexpr uint32_t kNumDist = 30;            // huffman_table.rs:9
constexpr uint32_t kEob = 256;               // huffman_table.rs:28
constexpr uint32_t kTooFar = 8192;           // lz77.rs:274-278
constexpr uint32_t kMaxStored = 32767;       // stored_block.rs:11

enum Mode : int { kGreedy = 0, kLazy = 1, kRle = 2 };   // lz77.rs:192-232 process_chunk dispatch
enum BlockType : int { kStored = 0, kFixed = 1, kDynamic = 2 };

// Parameters derived from CompressionOptions (compression_options.rs:78-120) exactly as
// DeflateState::new does (deflate_state.rs:100-109: lazy_if_less_than clamped to 32768).
struct Params {
    uint32_t checks;        // max_hash_checks
    uint32_t checks_quarter;// max_hash_checks >> 2, used when the pending match is >= 32 (lz77.rs:351-355)
    uint32_t lazy;          // lazy_if_less_than (clamped)
    int mode;               // Mode
    int need_quarter;       // 1 iff a search with the quarter budget can happen (lazy > 32)
};

DFL_HD Params make_params(uint16_t max_hash_checks, uint16_t lazy_if_less_than, uint8_t matching_type) {
    Params p;
    p.checks = max_hash_checks;
    p.checks_quarter = max_hash_checks >> 2;
    p.lazy = lazy_if_less_than < 32768u ? lazy_if_less_than : 32768u;
    if (matching_type == 0) p.mode = kGreedy;
    else p.mode = (max_hash_checks > 0) ? kLazy : kRle;
    p.need_quarter = (p.mode == kLazy && p.lazy > 32) ? 1 : 0;
    return p;
}

// ---------------------------------------------------------------- hashing
// chained_hash_table.rs:55-62: three applications of h = ((h << 5) ^ b) & 0x7fff leave exactly
// ((b0 & 31) << 10) ^ (b1 << 5) ^ b2 -- the contribution of older bytes is shifted out.
DFL_HD uint32_t hash3(uint32_t b0, uint32_t b1, uint32_t b2) {
    return ((b0 & 31u) << 10) ^ (b1 << 5) ^ b2;
}
// Nine bits that, together with an equal hash3, prove the three bytes are equal.
DFL_HD uint32_t tag9(uint32_t b0, uint32_t b1) {
    return (b0 >> 5) | ((b0 & 7u) << 3) | ((b1 & 7u) << 6);
}
// One entry of a per-window candidate list (windows are sorted by hash3, then by position).  64 bits:
//   lo = bytes p+3 .. p+6 (little endian), hi = byte p+7 | tag9 << 8 | position-in-window << 17.
// Together with an equal hash3 an entry therefore proves the length of the common prefix of two
// positions exactly up to 8 bytes without touching the data (bytes past the end of input are 0).
struct Entry { uint32_t lo, hi; };
constexpr uint32_t kEntryKeyHi = 0x1ffffu;   // byte 7 + tag: the part of `hi` that takes part in compares
constexpr uint32_t kEntryTagHi = 0x1ff00u;
constexpr uint32_t kEntryBytes = 8;          // prefix length an entry can prove
DFL_HD Entry make_entry(uint32_t pos_local, const uint8_t b[8]) {
    Entry e;
    e.lo = b[3] | (b[4] << 8) | (b[5] << 16) | ((uint32_t)b[6] << 24);
    e.hi = b[7] | (tag9(b[0], b[1]) << 8) | (pos_local << 17);
    return e;
}
DFL_HD uint32_t entry_pos(uint32_t hi) { return hi >> 17; }
// Masks selecting what a candidate must share with the target to be LONGER than best_len
// (best_len in {1 = nothing yet, 3..7}; from 8 on every key bit must agree).
DFL_HD void entry_mask(uint32_t best_len, uint32_t& mlo, uint32_t& mhi) {
    if (best_len < kMinMatch) { mlo = 0u; mhi = kEntryTagHi; return; }
    uint32_t nb = best_len - 2u;             // bytes 3 .. best_len must match
    mlo = nb >= 4u ? 0xffffffffu : ((1u << (8u * nb)) - 1u);
    mhi = nb >= 5u ? kEntryKeyHi : kEntryTagHi;
}
// Common prefix length (3..8) of two positions with equal hash3 and equal tag, from the xor of their entries.
DFL_HD uint32_t entry_lcp(uint32_t xlo, uint32_t xhi) {
    if (xlo != 0u) {
#if defined(__CUDA_ARCH__)
        return 3u + ((uint32_t)(__ffs((int)xlo) - 1) >> 3);
#else
        return 3u + ((uint32_t)__builtin_ctz(xlo) >> 3);
#endif
    }
    return (xhi & 0xffu) ? 7u : 8u;
}

// ---------------------------------------------------------------- candidate walk
// One longest_match call (matching.rs:87-166) visits its candidates most recent first and keeps the first one
// that is strictly longer than the running best, so the nearest candidate wins ties (:148-157).  The pipeline
// splits that work in two:
//   * the match stage (kernel k_match) walks the candidate *entries* only.  An entry proves common prefixes of
//     up to 8 bytes exactly, so every result shorter than 8 bytes is final there; a position whose best
//     candidate shares all 8 entry bytes is recorded as "long": only its rank in the window's sorted list and
//     the visit index of the nearest such candidate are kept.
//   * the parse stage resolves a long record on the data -- but only at the positions the reference's parser
//     actually searches (lz77.rs:340-377 skips everything inside an emitted match).
// best_len == 1 means "nothing yet" (matching.rs:108 floors the running best at 1; results shorter than
// MIN_MATCH are discarded by both parsers, so they are never recorded).
struct EntryWalk {
    uint32_t best_len;    // 1, or 3..8 as proven by entries (clamped to the bytes left in the input)
    uint32_t best_pos;    // position-in-window field of the best candidate
    uint32_t best_k;      // its visit index (0 = nearest candidate)
    uint32_t mlo, mhi;    // entry_mask(best_len)
    uint32_t stop;        // nothing can be longer as far as entries can tell (8 bytes, or the end of the input)
};
DFL_HD EntryWalk ewalk_init() {
    EntryWalk s; s.best_len = 1; s.best_pos = 0; s.best_k = 0; s.stop = 0; entry_mask(1, s.mlo, s.mhi); return s;
}
// Visit number k: candidate entry `ce` against target entry `me`; maxl = min(258, bytes left at the target).
DFL_HD void ewalk_visit(EntryWalk& s, Entry me, Entry ce, uint32_t k, uint32_t maxl) {
    if (s.stop) return;
    if ((((ce.lo ^ me.lo) & s.mlo) | ((ce.hi ^ me.hi) & s.mhi)) != 0u) return;
    uint32_t l = entry_lcp(ce.lo ^ me.lo, ce.hi ^ me.hi);
    if (l > maxl) l = maxl;
    if (l > s.best_len) {
        s.best_len = l; s.best_pos = entry_pos(ce.hi); s.best_k = k;
        entry_mask(l, s.mlo, s.mhi);
        if (l >= kEntryBytes || l == maxl) s.stop = 1;
    }
}

// ---------------------------------------------------------------- per-position match record
// Final:  len in bits 0..8 (0 or 3..258), dist-1 in bits 9..23; 0 == "no usable match".
// Long (bit 31): the best candidate shares >= 8 bytes and more bytes are left to compare.  Bits 0..14 = rank
// of the position in its window's sorted list, bits 15..30 = visit index of the nearest candidate that shares
// 8 bytes (every nearer one shares fewer, so a resolution starts there).
constexpr uint32_t kRecLong = 0x80000000u;
constexpr uint32_t kLenLong = 0xffu;     // length code (one byte per position beside the record): the record is long
DFL_HD uint32_t pack_match(uint32_t len, uint32_t dist) { return len | ((dist - 1u) << 9); }
DFL_HD uint32_t match_len(uint32_t m) { return m & 0x1ffu; }
DFL_HD uint32_t match_dist(uint32_t m) { return ((m >> 9) & 0x7fffu) + 1u; }
DFL_HD uint32_t rec_long(uint32_t rank, uint32_t k8) { return kRecLong | rank | (k8 << 15); }
DFL_HD bool rec_is_long(uint32_t m) { return (m & kRecLong) != 0u; }
DFL_HD uint32_t rec_rank(uint32_t m) { return m & 0x7fffu; }
DFL_HD uint32_t rec_k8(uint32_t m) { return (m >> 15) & 0xffffu; }
// lz77.rs:274-278 match_too_far applied to the result of matching.rs:87-166; results shorter than
// MIN_MATCH are never used by either parser and are recorded as "no match".
DFL_HD uint32_t finalize_match(uint32_t len, uint32_t dist) {
    if (len < kMinMatch) return 0u;
    if (len == kMinMatch && dist > kTooFar) return 0u;
    return pack_match(len, dist);
}
// Length code of a record, one byte: what the parser needs in order to decide (the distance only matters once a
// token is written).  0 = no match, 1..253 = length - 2, kLenSeeRecord = a final length of 256..258 (read it from
// the record), kLenLong = the record is long.
constexpr uint32_t kLenSeeRecord = 0xfeu;
DFL_HD uint32_t rec_len_code(uint32_t rec) {
    if (rec_is_long(rec)) return kLenLong;
    const uint32_t l = match_len(rec);
    return l == 0u ? 0u : (l <= 255u ? l - 2u : kLenSeeRecord);
}
// Record of a finished entry walk.  `rank` = index of the target in its window's sorted list, `dist` = distance
// of the best candidate.
DFL_HD uint32_t ewalk_record(const EntryWalk& s, uint32_t rank, uint32_t dist, uint32_t maxl) {
    if (s.best_len >= kEntryBytes && maxl > kEntryBytes) return rec_long(rank, s.best_k);
    return finalize_match(s.best_len, dist);
}

// ---------------------------------------------------------------- tokens
// bits 0..8: literal byte, or match length (3..258); bits 9..24: distance (0 = literal).
DFL_HD uint32_t tok_literal(uint32_t b) { return b; }
DFL_HD uint32_t tok_match(uint32_t len, uint32_t dist) { return len | (dist << 9); }
DFL_HD uint32_t tok_dist(uint32_t t) { return t >> 9; }
DFL_HD uint32_t tok_lo(uint32_t t) { return t & 0x1ffu; }
DFL_HD uint32_t tok_input_len(uint32_t t) { return (t >> 9) ? (t & 0x1ffu) : 1u; }

// ---------------------------------------------------------------- symbol arithmetic
// Closed forms of the lookup tables in huffman_table.rs:45-111 (LENGTH_CODE/BASE_LENGTH/
// LENGTH_EXTRA_BITS_LENGTH, DISTANCE_CODES/DISTANCE_BASE and num_extra_bits_for_distance_code);
// tests/test_model.py checks them against the tables parsed from the reference source.
DFL_HD uint32_t ilog2(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return 31u - (uint32_t)__clz((int)x);
#else
    return 31u - (uint32_t)__builtin_clz(x);
#endif
}
// length 3..258 -> (code 257..285, extra bit count, extra value)
DFL_HD void length_symbol(uint32_t len, uint32_t& code, uint32_t& nextra, uint32_t& extra) {
    uint32_t l = len - kMinMatch;
    if (l < 8u) { code = 257u + l; nextra = 0; extra = 0; return; }
    if (l == 255u) { code = 285u; nextra = 0; extra = 0; return; }
    uint32_t nb = ilog2(l);                     // 3..7
    nextra = nb - 2u;
    code = 257u + 4u * (nb - 1u) + ((l >> nextra) & 3u);
    extra = l & ((1u << nextra) - 1u);
}
// distance 1..32768 -> (code 0..29, extra bit count, extra value)
DFL_HD void dist_symbol(uint32_t dist, uint32_t& code, uint32_t& nextra, uint32_t& extra) {
    uint32_t x = dist - 1u;
    if (x < 4u) { code = x; nextra = 0; extra = 0; return; }
    uint32_t nb = ilog2(x);                     // 2..14
    nextra = nb - 1u;
    code = 2u * nb + ((x >> nextra) & 1u);
    extra = x & ((1u << nextra) - 1u);
}
DFL_HD uint32_t length_extra_bits_of_code(uint32_t c /*0..28*/) {
    return (c < 8u || c == 28u) ? 0u : ((c - 4u) >> 2);
}
DFL_HD uint32_t dist_extra_bits_of_code(uint32_t c /*0..29*/) {
    uint32_t h = c >> 1;                        // huffman_table.rs:120-126
    return h - (h != 0u ? 1u : 0u);
}
DFL_HD uint32_t fixed_ll_length(uint32_t s) {   // huffman_table.rs:32-42
    return s < 144u ? 8u : (s < 256u ? 9u : (s < 280u ? 7u : 8u));
}
// bit_reverse.rs:3-10
DFL_HD uint32_t reverse_bits(uint32_t v, uint32_t nbits) {
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32u - nbits);
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < nbits; i++) r |= ((v >> i) & 1u) << (nbits - 1u - i);
    return r;
#endif
}

// ---------------------------------------------------------------- parser state machines
// State of the reference's parsers *between* loop iterations, in absolute positions.
// (lz77.rs:162-173 ChunkState + the locals prev_length/prev_distance/ignore_next of
// process_chunk_lazy, lz77.rs:305-486.)  `pos` is the position the next iteration examines.
// The pending match is (prev_len, reference to its distance): the parser's decisions only depend on lengths,
// so a caller may hand in matches whose distance it has not looked up yet (kRefFull / kRefQuarter: prev_ref is
// the position whose full- / quarter-budget match record holds it) and resolve them when the token is written.
enum RefKind : uint32_t { kRefDist = 0, kRefFull = 1, kRefQuarter = 2 };
struct ParseState {
    uint32_t pos;
    uint32_t prev_len;   // pending match found at pos-1 (0 = none)
    uint32_t prev_ref;   // its distance (kRefDist) or the position of its record
    uint32_t prev_kind;  // RefKind
    uint32_t add;        // byte pos-1 is still to be emitted as a literal
    uint32_t ign;        // ignore_next
};
DFL_HD ParseState parse_state_init(uint32_t pos) {
    ParseState s; s.pos = pos; s.prev_len = 0; s.prev_ref = 0; s.prev_kind = kRefDist; s.add = 0; s.ign = 0; return s;
}
// Hand-off key of a state whose pending distance is known (prev_dist <= 32768: 16 bits).
DFL_HD uint32_t parse_state_key(const ParseState& s, uint32_t prev_dist) {
    return s.prev_len | (prev_dist << 9) | (s.add << 25) | (s.ign << 26);
}
DFL_HD ParseState parse_state_from_key(uint32_t pos, uint32_t key) {
    ParseState s;
    s.pos = pos; s.prev_len = key & 0x1ffu; s.prev_ref = (key >> 9) & 0xffffu; s.prev_kind = kRefDist;
    s.add = (key >> 25) & 1u; s.ign = (key >> 26) & 1u;
    return s;
}

// One iteration of process_chunk_lazy (lz77.rs:340-480).  (m_len, m_ref, m_kind): the longest match among the
// candidates the reference would visit at this position (matching.rs:87-166; 0 = none), whatever the floor --
// the step applies "only if longer than prev_length" itself (matching.rs:161-165).  Only looked at when the
// reference searches (p + 2 < n and !ign).  Tokens go to `out`: out.literal(position), out.match(len, ref, kind).
template <class Sink>
DFL_HD void lazy_step(ParseState& s, uint32_t n, uint32_t m_len, uint32_t m_ref, uint32_t m_kind, uint32_t lazy, Sink& out) {
    const uint32_t p = s.pos;
    if (p + 2u < n) {                                   // hash_it.next() is Some
        uint32_t cur_len = 0;
        if (!s.ign) {
            const uint32_t floor_len = s.prev_len > 1u ? s.prev_len : 1u;
            if (m_len > floor_len) cur_len = m_len;
            if (cur_len >= lazy) s.ign = 1;             // lz77.rs:374-377
        } else {
            s.ign = 0;                                  // lz77.rs:380-386
        }
        if (s.prev_len >= cur_len && s.prev_len >= kMinMatch) {   // lz77.rs:388-426
            out.match(s.prev_len, s.prev_ref, s.prev_kind);
            s.pos = p - 1u + s.prev_len;
            s.add = 0; s.prev_len = 0; s.prev_ref = 0; s.prev_kind = kRefDist; s.ign = 0;
        } else {
            if (s.add) out.literal(p - 1u);                       // lz77.rs:427-431
            else s.add = 1;                                        // lz77.rs:432-434
            s.prev_len = cur_len; s.prev_ref = m_ref; s.prev_kind = m_kind;
            s.pos = p + 1u;
        }
    } else {                                            // last two bytes, lz77.rs:440-482
        if (s.prev_len >= kMinMatch) {
            out.match(s.prev_len, s.prev_ref, s.prev_kind);
            s.pos = n; s.add = 0; s.prev_len = 0; s.prev_ref = 0; s.prev_kind = kRefDist;
        } else {
            if (s.add) { s.add = 0; out.literal(p - 1u); }
            out.literal(p);
            s.pos = p + 1u;
        }
    }
}

// One iteration of process_chunk_greedy (lz77.rs:502-544).
template <class Sink>
DFL_HD void greedy_step(ParseState& s, uint32_t n, uint32_t m_len, uint32_t m_ref, uint32_t m_kind, Sink& out) {
    const uint32_t p = s.pos;
    if (p + 2u < n && m_len >= kMinMatch) {
        out.match(m_len, m_ref, m_kind);
        s.pos = p + m_len;
    } else {
        out.literal(p);
        s.pos = p + 1u;
    }
}

// One iteration of process_chunk_greedy_rle (rle.rs:23-71): distance-1 runs only.
template <class Sink>
DFL_HD void rle_step(ParseState& s, uint32_t n, const uint8_t* data, Sink& out) {
    const uint32_t p = s.pos;
    if (p == 0u) { out.literal(0u); s.pos = 1; return; }
    const uint8_t prev = data[p - 1u];
    uint32_t len = 0;
    if (data[p] == prev) {
        uint32_t maxl = n - p < kMaxMatch ? n - p : kMaxMatch;
        while (len < maxl && data[p + len] == prev) len++;
    }
    if (len >= kMinMatch) { out.match(len, 1u, kRefDist); s.pos = p + len; }
    else { out.literal(p); s.pos = p + 1u; }
}

// ---------------------------------------------------------------- the reference's loop, sequentially
// process_chunk_lazy (lz77.rs:305-486) over the reference's own head/prev chains (chained_hash_table.rs), for
// MatchingType::Lazy with lazy_if_less_than < 3 -- the option values for which two things become observable
// that the parallel pipeline leaves out (kernel k_lz77_seq has the full story): length-2 results from spurious
// chain entries, and ignore_next being re-derived at the start of every process_chunk_lazy call.
// Chains in absolute positions: head[hash] = last inserted position (kSeqNone = still its own index);
// prev_val/prev_org[pos & 0x7fff] = what head[hash] held when pos was inserted: an absolute position (prev_org ==
// kSeqNone; stale once below the buffer origin of the window being processed, where `slide` would have reset it to
// the slot's own index, chained_hash_table.rs:197-219), or the hash itself (prev_org = buffer origin at insertion:
// head still held its own index; valid as buffer-relative position until the next slide).
// The caller initialises all three tables (32768 entries each) to kSeqNone.  Returns the number of tokens.
constexpr uint32_t kSeqNone = 0xffffffffu;
DFL_HD uint32_t seq_origin(uint32_t pos) {          // buffer origin while pos's window is processed: windows 0 and 1
    const uint32_t w = pos >> 15;                   // share the unslid buffer (lz77.rs:650-667,745-756)
    return w <= 1u ? 0u : (w - 1u) << 15;
}
DFL_HD void seq_insert_upto(const uint8_t* in, uint32_t n, uint32_t& next_insert, uint32_t upto, uint32_t* head,
                            uint32_t* prev_val, uint32_t* prev_org) {
    for (; next_insert < upto; next_insert++) {     // chained_hash_table.rs:148-158, every position once, in order
        const uint32_t q = next_insert;
        if (q + 2u >= n) continue;                  // hash_it.next() is None: not inserted
        const uint32_t h = hash3(in[q], in[q + 1], in[q + 2]);
        const uint32_t hv = head[h];
        const uint32_t qorg = seq_origin(q);
        if (hv == kSeqNone || hv < qorg) { prev_val[q & kWindowMask] = h; prev_org[q & kWindowMask] = qorg; }
        else { prev_val[q & kWindowMask] = hv; prev_org[q & kWindowMask] = kSeqNone; }
        head[h] = q;
    }
}
DFL_HD unsigned long long lz77_sequential(const uint8_t* in, uint32_t n, const Params& prm, uint32_t* head, uint32_t* prev_val,
                                          uint32_t* prev_org, uint32_t* tok) {
    const uint32_t lazy = prm.lazy, checks = prm.checks;
    unsigned long long nt = 0;
    uint32_t prev_len = 0, prev_dist = 0, add = 0, ign = 0;
    uint32_t cur_w = kSeqNone, org = 0, next_insert = 0;
    bool chunk_start = true;
    uint32_t p = 0;
    while (p < n) {
        const uint32_t w = p >> 15;
        if (w != cur_w) { cur_w = w; org = seq_origin(p); chunk_start = true; }       // a new process_chunk_lazy call
        if (chunk_start) { ign = prev_len >= lazy ? 1u : 0u; chunk_start = false; }   // lz77.rs:331
        if (p + 2u < n) {
            seq_insert_upto(in, n, next_insert, p + 1u, head, prev_val, prev_org);
            uint32_t cur_len = 0, cur_dist = 0;
            if (!ign) {
                const uint32_t budget = prev_len >= 32u ? checks >> 2 : checks;       // lz77.rs:351-355
                if (!(prev_len >= kMaxMatch || p + prev_len >= n)) {                  // longest_match, matching.rs:87-166
                    const uint32_t pos_rel = p - org;
                    const uint32_t limit = pos_rel > kWindow ? pos_rel - kWindow : 0u;
                    const uint32_t floor_len = prev_len > 1u ? prev_len : 1u;
                    const uint32_t max_len = (n - p) < kMaxMatch ? (n - p) : kMaxMatch;
                    uint32_t best = floor_len, best_dist = 0, cur = pos_rel;
                    for (uint32_t c = 0; c < budget; c++) {
                        const uint32_t prev_head = cur;
                        const uint32_t slot = cur & kWindowMask;
                        const uint32_t v = prev_val[slot], vo = prev_org[slot];
                        if (vo == kSeqNone) {                         // an absolute position
                            if (v == kSeqNone || v < org) break;      // its own index by now: the chain ends (matching.rs:127-132)
                            cur = v - org;
                        } else {                                      // the hash itself, until the next slide
                            if (vo != org) break;
                            cur = v;
                        }
                        if (cur >= prev_head || cur < limit) break;
                        const uint32_t ca = org + cur;
                        if (in[p + best - 1u] == in[ca + best - 1u] && in[p + best] == in[ca + best]) {
                            uint32_t l = 0;
                            while (l < max_len && in[p + l] == in[ca + l]) l++;
                            if (l > best) { best = l; best_dist = p - ca; if (l == max_len) break; }
                        }
                    }
                    if (best > floor_len) { cur_len = best; cur_dist = best_dist; }
                }
                if (cur_len == kMinMatch && cur_dist > kTooFar) cur_len = 0;          // lz77.rs:274-278
                if (cur_len >= lazy) ign = 1;                                         // lz77.rs:374-377
            } else {
                ign = 0;
            }
            if (prev_len >= cur_len && prev_len >= kMinMatch) {
                tok[nt++] = tok_match(prev_len, prev_dist);
                const uint32_t np = p - 1u + prev_len;
                seq_insert_upto(in, n, next_insert, np < n ? np : n, head, prev_val, prev_org);
                add = 0; prev_len = 0; prev_dist = 0;
                if (nt % kBlockTokens == 0ull) chunk_start = true; else ign = 0;      // BufferFull returns before `ignore_next = false`
                p = np;
                continue;
            }
            if (add) {
                tok[nt++] = tok_literal(in[p - 1u]);
                if (nt % kBlockTokens == 0ull) chunk_start = true;                    // BufferFull: the next call re-derives ignore_next
            } else add = 1;
            prev_len = cur_len; prev_dist = cur_dist;
            p++;
        } else {                                                                     // last two bytes, lz77.rs:440-482
            if (prev_len >= kMinMatch) {
                tok[nt++] = tok_match(prev_len, prev_dist);
                break;
            }
            if (add) { add = 0; tok[nt++] = tok_literal(in[p - 1u]); }
            tok[nt++] = tok_literal(in[p]);
            p++;
        }
    }
    return nt;
}

// ---------------------------------------------------------------- Huffman code construction
// length_encode.rs:347-415 in_place_lengths: stable sort by frequency, Moffat-Katajainen in-place
// (step_1/step_2, :218-278), miniz-style limiter (:290-327), then lengths handed out from the
// highest-frequency leaf down (:402-408).  `key` scratch must hold n_freq entries.  A leaf is
// (freq << 9) | symbol, so an ordinary sort of the keys is the reference's stable sort.
// `presorted` >= 0: key[] already holds that many leaves in ascending order (the kernel sorts them with the
// whole warp); < 0: collect and sort them here.
DFL_HD void huffman_lengths(const uint32_t* freqs, uint32_t n_freq, uint32_t max_len,
                            uint8_t* lens /*n_lens*/, uint32_t n_lens, uint32_t* key, int presorted = -1) {
    for (uint32_t i = 0; i < n_lens; i++) lens[i] = 0;
    uint32_t n = 0;
    if (presorted >= 0) {
        n = (uint32_t)presorted;
    } else {
        for (uint32_t i = 0; i < n_freq; i++)
            if (freqs[i] > 0) key[n++] = (freqs[i] << 9) | i;
    }
    if (n == 0) return;
    if (n == 1) { lens[key[0] & 0x1ffu] = 1; return; }
    // insertion sort on unique keys == stable sort by freq (symbols were appended in index order)
    if (presorted < 0) {
        for (uint32_t i = 1; i < n; i++) {
            uint32_t k = key[i];
            uint32_t j = i;
            while (j > 0 && key[j - 1] > k) { key[j] = key[j - 1]; j--; }
            key[j] = k;
        }
    }
    // From here `key[i] >> 9` plays the role of leaves[i].value and `key[i] & 511` of .symbol.
#define DFL_VAL(i) (key[(i)] >> 9)
#define DFL_SETVAL(i, v) (key[(i)] = ((uint32_t)(v) << 9) | (key[(i)] & 0x1ffu))
    {   // step_1
        uint32_t root = 0, leaf = 2;
        DFL_SETVAL(0, DFL_VAL(0) + DFL_VAL(1));
        for (uint32_t next = 1; next + 1 < n; next++) {
            uint32_t v;
            if (leaf >= n || DFL_VAL(root) < DFL_VAL(leaf)) { v = DFL_VAL(root); DFL_SETVAL(root, next); root++; }
            else { v = DFL_VAL(leaf); leaf++; }
            if (leaf >= n || (root < next && DFL_VAL(root) < DFL_VAL(leaf))) { v += DFL_VAL(root); DFL_SETVAL(root, next); root++; }
            else { v += DFL_VAL(leaf); leaf++; }
            DFL_SETVAL(next, v);
        }
    }
    {   // step_2
        DFL_SETVAL(n - 2, 0);
        for (uint32_t t = n - 2; t-- > 0;) DFL_SETVAL(t, DFL_VAL(DFL_VAL(t)) + 1);
        uint32_t available = 1, used = 0, depth = 0;
        int root = (int)n - 2, next = (int)n - 1;
        while (available > 0) {
            while (root >= 0 && DFL_VAL(root) == depth) { used++; root--; }
            while (available > used) { DFL_SETVAL(next, depth); next--; available--; }
            available = 2 * used; depth++; used = 0;
        }
    }
    uint32_t num_codes[33];
    for (int i = 0; i < 33; i++) num_codes[i] = 0;
    for (uint32_t i = 0; i < n; i++) { uint32_t d = DFL_VAL(i); num_codes[d < 32u ? d : 32u]++; }
    {   // enforce_max_code_lengths
        uint32_t above = 0;
        for (uint32_t i = max_len + 1; i < 33; i++) above += num_codes[i];
        num_codes[max_len] += above;
        uint32_t total = 0;
        for (uint32_t i = max_len; i >= 1; i--) total += num_codes[i] << (max_len - i);
        while (total != (1u << max_len)) {
            num_codes[max_len]--;
            for (uint32_t i = max_len - 1; i >= 1; i--)
                if (num_codes[i] != 0) { num_codes[i]--; num_codes[i + 1] += 2; break; }
            total--;
        }
    }
    uint32_t it = n;
    for (uint32_t len = 1; len <= max_len; len++)
        for (uint32_t c = 0; c < num_codes[len]; c++) { it--; lens[key[it] & 0x1ffu] = (uint8_t)len; }
#undef DFL_VAL
#undef DFL_SETVAL
}
// NB: the weight of an internal node can reach the sum of all frequencies (<= 31744 + 1 per
// block, output_writer.rs:19), far below 2^23, so packing value and symbol in 32 bits is exact.

// huffman_table.rs:253-278 create_codes_in_place: canonical codes, bit-reversed for LSB-first output.
DFL_HD void canonical_codes(const uint8_t* lens, uint32_t n, uint16_t* codes) {
    uint32_t cnt[16];
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (uint32_t i = 0; i < n; i++) if (lens[i]) cnt[lens[i]]++;
    uint32_t next_code[16];
    uint32_t code = 0;
    next_code[0] = 0;
    for (uint32_t b = 1; b < 16; b++) { code = (code + cnt[b - 1]) << 1; next_code[b] = code; }
    for (uint32_t i = 0; i < n; i++) {
        uint32_t l = lens[i];
        codes[i] = l ? (uint16_t)reverse_bits(next_code[l]++ & 0xffffu, l) : (uint16_t)0;
    }
}

// length_encode.rs:82-155 encode_lengths_m, restated as explicit run handling but emitting exactly
// the reference's symbol sequence (tests compare against the oracle's literal port on random and
// adversarial inputs).  Output symbols: low 5 bits = symbol 0..18, bits 8.. = repeat count.
// Returns the number of symbols; freqs19 is incremented.
DFL_HD uint32_t rle_not_max(uint32_t l, uint32_t repeats) { return (l == 0u && repeats < 138u) || repeats < 6u; }
DFL_HD uint32_t encode_lengths(const uint8_t* lengths, uint32_t n, uint16_t* out, uint32_t* freqs19) {
    uint32_t no = 0;
    uint32_t repeat = 0;
    uint32_t prev = (uint32_t)(uint8_t)~lengths[0];
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t l = lengths[i];
        const bool at_end = (i + 1 == n);
        if (l == prev && rle_not_max(l, repeat)) repeat++;
        if (l != prev || at_end || !rle_not_max(l, repeat)) {
            if (repeat >= 3u) {
                uint32_t sym = (prev == 0u) ? (repeat <= 10u ? 17u : 18u) : 16u;
                out[no++] = (uint16_t)(sym | (repeat << 8)); freqs19[sym]++;
                repeat = 0;
                if (l != prev) {
                    if (l != 0u || at_end) { out[no++] = (uint16_t)l; freqs19[l]++; repeat = 0; }
                    else repeat = 1;
                }
            } else {
                uint32_t extra_skip = (at_end && l == prev) ? 1u : 0u;
                uint32_t skip = i + extra_skip - repeat;
                uint32_t extra = (l != 0u || at_end) ? 1u : 0u;
                uint32_t take = repeat + extra;
                for (uint32_t k = skip; k < n && k < skip + take; k++) { out[no++] = lengths[k]; freqs19[lengths[k]]++; }
                repeat = 1u - extra;
            }
        }
        prev = l;
    }
    return no;
}

// Everything gen_huffman_lengths (huffman_lengths.rs:167-287) derives from one block's histograms,
// except the final Stored decision which needs the bit position of the block (see choose_block).
struct BlockCodes {
    uint8_t ll_len[288];
    uint8_t d_len[32];
    uint16_t ll_code[288];
    uint16_t d_code[32];
    uint8_t cl_len[19];
    uint16_t cl_code[19];
    uint16_t hdr_sym[320];      // encoded lengths of ll||dist (symbol | repeat << 8)
    uint32_t n_hdr_sym;
    uint32_t hlit, hdist;       // number of ll / dist lengths transmitted
    uint32_t used_hclens;
    uint32_t tiny;              // num_input_bytes <= 4  -> Fixed without any cost computation
    uint64_t dynamic_cost;      // the reference's *estimate* (symbol 16 counted with 3 extra bits)
    uint64_t static_cost;       // ditto (distance symbols costed with the literal table)
    uint64_t stored_cost;       // stored_length(n) without the position-dependent padding
    uint64_t dynamic_bits;      // bits actually emitted after the 3-bit block header
    uint64_t fixed_bits;
    uint64_t input_bytes;
};

DFL_HD uint64_t stored_padding(uint32_t pending_bits /*0..7*/) {   // huffman_lengths.rs:113-124
    uint32_t free_space = 8u - pending_bits;
    return free_space >= 3u ? free_space - 3u : 8u - (3u - free_space);
}
DFL_HD uint64_t stored_length(uint64_t n) {                         // huffman_lengths.rs:132-143
    uint64_t k = (n - 1u) / kMaxStored + 1u;
    return (n + 4u * k + (k - 1u)) * 8u;
}

// ll_sorted / d_sorted (optional): the non-zero (freq << 9 | symbol) leaves of the two alphabets in ascending
// order, n_ll / n_d of them, in buffers this function may overwrite.
DFL_HD void build_block_codes(const uint32_t* ll_freq /*286*/, const uint32_t* d_freq /*30*/,
                              uint64_t input_bytes, BlockCodes& bc, uint32_t* scratch /*>=288*/,
                              uint32_t* ll_sorted = nullptr, int n_ll = -1, uint32_t* d_sorted = nullptr, int n_d = -1) {
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    bc.input_bytes = input_bytes;
    bc.tiny = input_bytes <= 4u ? 1u : 0u;
    // fixed-code body size is needed for every block type decision and for tiny blocks
    {
        uint64_t fb = 0;
        for (uint32_t c = 0; c < kNumLL; c++) {
            uint32_t eb = c >= 257u ? length_extra_bits_of_code(c - 257u) : 0u;
            fb += (uint64_t)ll_freq[c] * (fixed_ll_length(c) + eb);
        }
        for (uint32_t c = 0; c < kNumDist; c++) fb += (uint64_t)d_freq[c] * (5u + dist_extra_bits_of_code(c));
        bc.fixed_bits = fb;
    }
    if (bc.tiny) {
        bc.dynamic_cost = bc.static_cost = bc.stored_cost = bc.dynamic_bits = 0;
        bc.n_hdr_sym = 0; bc.hlit = 257; bc.hdist = 1; bc.used_hclens = 4;
        return;
    }
    uint32_t nl = kNumLL; while (nl > 257u && ll_freq[nl - 1] == 0) nl--;   // remove_trailing_zeroes
    uint32_t nd = kNumDist; while (nd > 1u && d_freq[nd - 1] == 0) nd--;
    bc.hlit = nl; bc.hdist = nd;
    huffman_lengths(ll_freq, nl, 15, bc.ll_len, 288, ll_sorted ? ll_sorted : scratch, ll_sorted ? n_ll : -1);
    huffman_lengths(d_freq, nd, 15, bc.d_len, 32, d_sorted ? d_sorted : scratch, d_sorted ? n_d : -1);
    uint32_t f19[19];
    for (int i = 0; i < 19; i++) f19[i] = 0;
    uint8_t chained[288 + 32];
    for (uint32_t i = 0; i < nl; i++) chained[i] = bc.ll_len[i];
    for (uint32_t i = 0; i < nd; i++) chained[nl + i] = bc.d_len[i];
    bc.n_hdr_sym = encode_lengths(chained, nl + nd, bc.hdr_sym, f19);
    huffman_lengths(f19, 19, 7, bc.cl_len, 19, scratch);
    uint32_t trailing = 0;
    while (trailing < 19u && bc.cl_len[order[18u - trailing]] == 0) trailing++;
    bc.used_hclens = 19u - trailing;
    canonical_codes(bc.ll_len, 288, bc.ll_code);
    canonical_codes(bc.d_len, 32, bc.d_code);
    canonical_codes(bc.cl_len, 19, bc.cl_code);
    // calculate_block_length (:75-107): the distance pass zips with the literal FIXED table too.
    uint64_t d_ll = 0, s_ll = 0, d_dist = 0, s_dist = 0;
    for (uint32_t c = 0; c < nl; c++) {
        uint64_t eb = c >= 257u ? length_extra_bits_of_code(c - 257u) : 0u;
        d_ll += (uint64_t)ll_freq[c] * (bc.ll_len[c] + eb);
        s_ll += (uint64_t)ll_freq[c] * (fixed_ll_length(c) + eb);
    }
    for (uint32_t c = 0; c < nd; c++) {
        uint64_t eb = dist_extra_bits_of_code(c);
        d_dist += (uint64_t)d_freq[c] * (bc.d_len[c] + eb);
        s_dist += (uint64_t)d_freq[c] * (fixed_ll_length(c) + eb);
    }
    uint64_t table_est = 0, table_bits = 0;
    for (uint32_t i = 0; i < 19; i++) {
        uint32_t est_extra = (i == 16u || i == 17u) ? 3u : (i == 18u ? 7u : 0u);   // :50-56
        uint32_t real_extra = i == 16u ? 2u : (i == 17u ? 3u : (i == 18u ? 7u : 0u));
        table_est += (uint64_t)f19[i] * (bc.cl_len[i] + est_extra);
        table_bits += (uint64_t)f19[i] * (bc.cl_len[i] + real_extra);
    }
    bc.dynamic_cost = d_ll + d_dist + table_est + (uint64_t)bc.used_hclens * 3u + 14u;
    bc.static_cost = s_ll + s_dist;
    bc.stored_cost = stored_length(input_bytes);
    bc.dynamic_bits = d_ll + d_dist + table_bits + (uint64_t)bc.used_hclens * 3u + 14u;
}

// The tail of gen_huffman_lengths (huffman_lengths.rs:265-286): pick the cheapest representation,
// ties Fixed > Stored > Dynamic.  `pending` = bit position of the block start modulo 8.
// Returns the type and the number of bits the block occupies including its 3-bit header.
DFL_HD int choose_block(const BlockCodes& bc, uint32_t pending, uint64_t& total_bits) {
    if (bc.tiny) { total_bits = 3u + bc.fixed_bits; return kFixed; }
    uint64_t stored_len = bc.stored_cost + stored_padding(pending);
    uint64_t used = bc.dynamic_cost < bc.static_cost ? bc.dynamic_cost : bc.static_cost;
    if (stored_len < used) used = stored_len;
    if (used == bc.static_cost) { total_bits = 3u + bc.fixed_bits; return kFixed; }
    if (used == stored_len) { total_bits = 3u + stored_len; return kStored; }
    total_bits = 3u + bc.dynamic_bits;
    return kDynamic;
}

// ---------------------------------------------------------------- Adler-32 (RFC 1950)
constexpr uint32_t kAdlerMod = 65521;
// combine(adler of A, adler of B, len(B)) -> adler of A||B
DFL_HD uint32_t adler32_combine(uint32_t ad1, uint32_t ad2, uint64_t len2) {
    uint32_t a1 = ad1 & 0xffffu, b1 = ad1 >> 16, a2 = ad2 & 0xffffu, b2 = ad2 >> 16;
    uint32_t rem = (uint32_t)(len2 % kAdlerMod);
    uint32_t a = (a1 + a2 + kAdlerMod - 1u) % kAdlerMod;
    uint64_t b = (uint64_t)b1 + b2 + (uint64_t)rem * ((a1 + kAdlerMod - 1u) % kAdlerMod);
    return (uint32_t)((b % kAdlerMod) << 16) | a;
}

// ---------------------------------------------------------------- CRC-32 (RFC 1952 section 8)
// The reference takes CRC-32 from the `gzip-header` crate (lib.rs:257-258, writer.rs:411-412); it is
// the standard reflected CRC with polynomial 0xEDB88320.  Pieces are checksummed independently and
// combined: crc(A || B) = crc(A) * x^(8 len(B)) mod P  xor  crc(B) (carry-less arithmetic).
constexpr uint32_t kCrcPoly = 0xedb88320u;
DFL_HD uint32_t crc32_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1u)) == 0u) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ kCrcPoly : b >> 1;
    }
    return p;
}
// x^(2^k) mod P for k = 0..31 (the exponent sequence repeats with period 32 beyond that)
DFL_HD uint32_t crc32_x2n(uint32_t k) {
    const uint32_t t[32] = {0x40000000u, 0x20000000u, 0x08000000u, 0x00800000u, 0x00008000u, 0xedb88320u, 0xb1e6b092u,
                            0xa06a2517u, 0xed627daeu, 0x88d14467u, 0xd7bbfe6au, 0xec447f11u, 0x8e7ea170u, 0x6427800eu,
                            0x4d47bae0u, 0x09fe548fu, 0x83852d0fu, 0x30362f1au, 0x7b5a9cc3u, 0x31fec169u, 0x9fec022au,
                            0x6c8dedc4u, 0x15d6874du, 0x5fde7a4eu, 0xbad90e37u, 0x2e4e5eefu, 0x4eaba214u, 0xa8a472c0u,
                            0x429a969eu, 0x148d302au, 0xc40ba6d0u, 0xc4e22c3cu};
    return t[k & 31u];
}
// combine(crc of A, crc of B, len(B)) -> crc of A || B
DFL_HD uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
    if (len2 == 0) return crc1;
    uint32_t p = 1u << 31;          // x^0
    uint32_t k = 3;                 // bytes -> bits
    for (uint64_t nn = len2; nn; nn >>= 1, k++)
        if (nn & 1u) p = crc32_multmodp(crc32_x2n(k), p);
    return crc32_multmodp(p, crc1) ^ crc2;
}

}  // namespace dfl
