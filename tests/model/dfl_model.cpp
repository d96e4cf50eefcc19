// dfl_model.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Sequential CPU walk-through of the *GPU pipeline's* algorithms (not of the reference): it calls
// the same host+device functions from deflate-rs_b200/csrc/dfl_core.h that the CUDA kernels call,
// in the same stage order and with the same data layouts (per-window sorted candidate lists,
// per-position match records, speculative parse segments with hand-off verification, fixed
// 31744-token blocks, per-block code construction, bit-offset scan, scatter bit packing).
// tests/test_model.py checks its output byte-for-byte against the oracle, which lets the
// algorithmic claims of DESIGN.md be verified in a container that has no GPU.  It is never linked
// into the product library.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../deflate-rs_b200/csrc/dfl_core.h"

using namespace dfl;

namespace {

struct Cfg {
    uint32_t pseg;   // parse segment length
    uint32_t warm;   // speculative warm-up
    uint32_t rounds; // parallel repair rounds before the sequential fallback
};

// ---- stage 1: per-window stable sort by hash3 into 64-bit entries (kernel k_window_sort)
void window_sort(const uint8_t* d, uint32_t n, std::vector<Entry>& K, std::vector<uint16_t>& off,
                 std::vector<uint32_t>& cnt) {
    uint32_t nseg = (n + kWindow - 1) / kWindow;
    K.assign((size_t)nseg * kWindow, Entry{0, 0});
    off.assign((size_t)nseg * kWindow, 0);
    cnt.assign(nseg, 0);
    uint32_t hashable = n >= 2 ? n - 2 : 0;   // positions p with p + 2 < n
    for (uint32_t s = 0; s < nseg; s++) {
        uint32_t base = s * kWindow;
        uint32_t c = hashable > base ? std::min(kWindow, hashable - base) : 0;
        cnt[s] = c;
        std::vector<uint32_t> hist(kWindow + 1, 0);
        for (uint32_t i = 0; i < c; i++) hist[hash3(d[base + i], d[base + i + 1], d[base + i + 2]) + 1]++;
        for (uint32_t h = 0; h < kWindow; h++) hist[h + 1] += hist[h];
        for (uint32_t h = 0; h < kWindow; h++) off[(size_t)s * kWindow + h] = (uint16_t)hist[h];
        std::vector<uint32_t> cur(hist.begin(), hist.end() - 1);
        for (uint32_t i = 0; i < c; i++) {
            uint32_t p = base + i;
            uint32_t h = hash3(d[p], d[p + 1], d[p + 2]);
            uint8_t b[8];
            for (uint32_t k = 0; k < 8; k++) b[k] = p + k < n ? d[p + k] : 0;   // bytes past the end are 0
            K[(size_t)s * kWindow + cur[h]++] = make_entry(i, b);
        }
    }
}

struct HostBytes {
    const uint8_t* d;
    uint32_t byte(uint32_t i) const { return d[i]; }
    uint32_t common_prefix(uint32_t a, uint32_t c, uint32_t from, uint32_t maxl) const {
        uint32_t l = from;
        while (l < maxl && d[a + l] == d[c + l]) l++;
        return l;
    }
};

// ---- stage 2: candidate walk per sorted entry (kernel k_match)
void match_all(const uint8_t* d, uint32_t n, const Params& prm, const std::vector<Entry>& K,
               const std::vector<uint16_t>& off, const std::vector<uint32_t>& cnt,
               std::vector<uint32_t>& Mf, std::vector<uint32_t>& Mq) {
    Mf.assign(n, 0);
    if (prm.need_quarter) Mq.assign(n, 0);
    uint32_t nseg = (uint32_t)cnt.size();
    HostBytes data{d};
    for (uint32_t s = 0; s < nseg; s++) {
        const Entry* Kw = &K[(size_t)s * kWindow];
        const uint16_t* ow = &off[(size_t)s * kWindow];
        for (uint32_t i = 0; i < cnt[s]; i++) {
            Entry me = Kw[i];
            uint32_t pl = entry_pos(me.hi);
            uint32_t p = s * kWindow + pl;
            uint32_t h = hash3(d[p], d[p + 1], d[p + 2]);
            uint32_t maxl = std::min(kMaxMatch, n - p);
            uint32_t budget = prm.checks;
            // own window, most recent first; then the previous window's bucket, positions at
            // distance <= 32768 only (matching.rs:102-106,127)
            uint32_t s0 = ow[h];
            uint32_t n_own = std::min(budget, i - s0), n_tot = n_own, pe = 0;
            const Entry* Kp = nullptr;
            if (s > 0 && n_own < budget) {
                Kp = &K[(size_t)(s - 1) * kWindow];
                const uint16_t* op = &off[(size_t)(s - 1) * kWindow];
                uint32_t ps = op[h];
                pe = (h + 1 < kWindow) ? op[h + 1] : cnt[s - 1];
                uint32_t rem = budget - n_own;
                uint32_t lo = (pe - ps > rem) ? pe - rem : ps;
                while (lo < pe && entry_pos(Kp[lo].hi) < pl) lo++;
                n_tot = n_own + (pe - lo);
            }
            WalkState st = walk_init();
            uint32_t q_len = 1, q_dist = 0;
            for (uint32_t k = 0; k < n_tot; k++) {
                Entry ce = k < n_own ? Kw[i - 1 - k] : Kp[pe - 1 - (k - n_own)];
                if (walk_passes(st, me, ce)) {
                    uint32_t q = (k < n_own ? s * kWindow : (s - 1) * kWindow) + entry_pos(ce.hi);
                    walk_consider(st, data, p, q, me, ce, maxl);
                    if (st.done) n_tot = k + 1;
                }
                if (prm.need_quarter && k + 1 == prm.checks_quarter) { q_len = st.best_len; q_dist = st.best_dist; }
            }
            if (prm.need_quarter && n_tot < prm.checks_quarter) { q_len = st.best_len; q_dist = st.best_dist; }
            Mf[p] = finalize_match(st.best_len, st.best_dist);
            if (prm.need_quarter) Mq[p] = finalize_match(q_len, q_dist);
        }
    }
}

// ---- stage 1b/2b: the span path (kernels k_window_sort<items>, k_span_scatter, k_match_chains)
// Sorted positions per window -> per-span merged entry lists -> multi-level chains walked per target.
uint64_t g_chain_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // targets, level queries, chain steps, deep compares

void window_sort_items(const uint8_t* d, uint32_t n, std::vector<uint32_t>& items, std::vector<uint16_t>& off,
                       std::vector<uint32_t>& cnt) {
    uint32_t nwin = (n + kWindow - 1) / kWindow;
    items.assign((size_t)nwin * kWindow, 0);
    off.assign((size_t)nwin * kWindow, 0);
    cnt.assign(nwin, 0);
    uint32_t hashable = n >= 2 ? n - 2 : 0;
    for (uint32_t v = 0; v < nwin; v++) {
        uint32_t base = v * kWindow;
        uint32_t c = hashable > base ? std::min(kWindow, hashable - base) : 0;
        cnt[v] = c;
        std::vector<uint32_t> hist(kWindow + 1, 0);
        for (uint32_t i = 0; i < c; i++) hist[hash3(d[base + i], d[base + i + 1], d[base + i + 2]) + 1]++;
        for (uint32_t h = 0; h < kWindow; h++) hist[h + 1] += hist[h];
        for (uint32_t h = 0; h < kWindow; h++) off[(size_t)v * kWindow + h] = (uint16_t)hist[h];
        std::vector<uint32_t> cur(hist.begin(), hist.end() - 1);
        for (uint32_t i = 0; i < c; i++) {
            uint32_t h = hash3(d[base + i], d[base + i + 1], d[base + i + 2]);
            items[(size_t)v * kWindow + cur[h]++] = (h << 15) | i;
        }
    }
}

void span_scatter(const uint8_t* d, uint32_t n, const std::vector<uint32_t>& items, const std::vector<uint16_t>& off,
                  const std::vector<uint32_t>& cnt, std::vector<Entry>& M, std::vector<uint32_t>& span_cnt) {
    uint32_t nwin = (uint32_t)cnt.size();
    uint32_t nspan = (nwin + kSpanWin - 1) / kSpanWin;
    M.assign((size_t)nspan * kSpanSlots, Entry{0, 0});
    span_cnt.assign(nspan, 0);
    auto offx = [&](uint32_t u, uint32_t h) { return h < kWindow ? (uint32_t)off[(size_t)u * kWindow + h] : cnt[u]; };
    for (uint32_t s = 0; s < nspan; s++) {
        uint32_t w_lo = s * kSpanWin > 0 ? s * kSpanWin - 1 : 0, w_hi = std::min(nwin, s * kSpanWin + kSpanWin);
        for (uint32_t v = w_lo; v < w_hi; v++) {
            span_cnt[s] += cnt[v];
            for (uint32_t r = 0; r < cnt[v]; r++) {
                uint32_t it = items[(size_t)v * kWindow + r];
                uint32_t h = it >> 15, pl = it & kWindowMask, p = v * kWindow + pl;
                uint32_t idx = r;
                bool first = (r == offx(v, h));
                for (uint32_t u = w_lo; u < w_hi; u++) {
                    if (u < v) { idx += offx(u, h + 1); if (offx(u, h + 1) != offx(u, h)) first = false; }
                    else if (u > v) idx += offx(u, h);
                }
                uint8_t b[7];
                for (uint32_t k = 0; k < 7; k++) b[k] = p + k < n ? d[p + k] : 0;
                uint32_t pos_in_span = p - s * kSpanWin * kWindow + kWindow;
                M[(size_t)s * kSpanSlots + idx] = make_span_entry(pos_in_span, b, first);
            }
        }
    }
}

constexpr uint32_t kModelChainChunk = 2048;   // entries per warp chunk
constexpr uint32_t kModelChainCtx = 128;      // entries re-inserted in front of a chunk

void match_all_chains(const uint8_t* d, uint32_t n, const Params& prm, const std::vector<Entry>& M,
                      const std::vector<uint32_t>& span_cnt, std::vector<uint32_t>& Mf) {
    Mf.assign(n, 0);
    HostBytes data{d};
    const uint32_t c = prm.checks;
    for (uint32_t s = 0; s < span_cnt.size(); s++) {
        const Entry* E = &M[(size_t)s * kSpanSlots];
        const uint32_t ne = span_cnt[s];
        const uint32_t span_base = s * kSpanWin * kWindow;
        for (uint32_t a = 0; a < ne; a += kModelChainChunk) {
            const uint32_t ctx = a >= kModelChainCtx ? a - kModelChainCtx : 0;
            const uint32_t b = std::min(ne, a + kModelChainChunk);
            uint16_t head[kChainLevels][128];
            uint8_t prevd[kChainLevels][256];
            Entry ring[256];
            memset(head, 0, sizeof(head));
            memset(prevd, 0, sizeof(prevd));
            uint32_t bstart = 256;                       // local index of the current bucket's first entry
            for (uint32_t m = ctx; m < b; m++) {
                const uint32_t i = m - ctx + 256;       // local index: never 0, so a zero head means "none"
                const Entry me = E[m];
                ring[i & 255] = me;
                if (me.hi & kSpanFirstBit) bstart = i;
                for (uint32_t lv = 0; lv < kChainLevels; lv++) {
                    uint32_t sg = span_sig(me.lo, me.hi, lv) >> 1;   // 7-bit slot, as in the kernel
                    uint32_t dl = i - head[lv][sg];
                    prevd[lv][i & 255] = (uint8_t)(dl < 256 ? dl : 0);
                    head[lv][sg] = (uint16_t)i;
                }
                const uint32_t pos = span_entry_pos(me.hi);
                if (m < a || pos < kWindow) continue;   // context or history entry: not a target here
                const uint32_t p = span_base + pos - kWindow;
                const uint32_t maxl = std::min(kMaxMatch, n - p);
                const uint32_t lb = std::max(bstart, i > c ? i - c : 0u);
                uint32_t best_len = 1, best_q = 0, level = 0;
                bool deep = false, done = false;
                g_chain_stats[0]++;
                while (level < kChainLevels && !done && !deep) {
                    g_chain_stats[1]++;
                    uint32_t j = i;
                    bool found = false;
                    Entry ce{0, 0};
                    for (;;) {
                        uint32_t dl = prevd[level][j & 255];
                        if (dl == 0) break;
                        j -= dl;
                        if (j < lb) break;
                        g_chain_stats[2]++;
                        ce = ring[j & 255];
                        if (span_key_equal(me, ce, level)) { found = true; break; }
                        g_chain_stats[4]++;
                    }
                    if (!found) break;
                    if (pos - span_entry_pos(ce.hi) > kWindow) break;      // matching.rs:102-106: beyond the window
                    uint32_t l = span_entry_lcp(ce.lo ^ me.lo);
                    uint32_t q = span_base + span_entry_pos(ce.hi) - kWindow;
                    if (l >= maxl) { best_len = maxl; best_q = q; done = true; break; }
                    if (l == kSpanEntryBytes) { deep = true; break; }
                    best_len = l; best_q = q; level = l - 2;
                }
                if (deep) {
                    // every candidate sharing 7+ bytes, nearest first: the reference's quick reject on the byte
                    // that would extend the best match (matching.rs:141-143), then the real length
                    g_chain_stats[7]++;
                    uint32_t j = i;
                    const uint32_t lv = kChainLevels - 1;
                    for (;;) {
                        uint32_t dl = prevd[lv][j & 255];
                        if (dl == 0) break;
                        j -= dl;
                        if (j < lb) break;
                        g_chain_stats[5]++;
                        Entry ce = ring[j & 255];
                        if (!span_key_equal(me, ce, lv)) { g_chain_stats[6]++; continue; }
                        if (pos - span_entry_pos(ce.hi) > kWindow) break;
                        uint32_t q = span_base + span_entry_pos(ce.hi) - kWindow;
                        if (best_len >= kSpanEntryBytes && d[q + best_len] != d[p + best_len]) continue;
                        g_chain_stats[3]++;
                        uint32_t l = data.common_prefix(p, q, kSpanEntryBytes, maxl);
                        if (l > best_len) { best_len = l; best_q = q; if (l == maxl) break; }
                    }
                }
                Mf[p] = finalize_match(best_len, p - best_q);
            }
        }
    }
}

int g_match_impl = 1;   // 0 = candidate walk (k_match), 1 = span chains (k_match_chains) when the options allow it
bool use_chains(const Params& prm) { return g_match_impl == 1 && prm.checks <= kChainMaxChecks && !prm.need_quarter; }

void find_matches(const uint8_t* in, uint32_t n, const Params& prm, std::vector<uint32_t>& Mf, std::vector<uint32_t>& Mq) {
    std::vector<uint32_t> cnt;
    std::vector<uint16_t> off;
    if (use_chains(prm)) {
        std::vector<uint32_t> items, span_cnt;
        std::vector<Entry> M;
        window_sort_items(in, n, items, off, cnt);
        span_scatter(in, n, items, off, cnt, M, span_cnt);
        match_all_chains(in, n, prm, M, span_cnt, Mf);
    } else {
        std::vector<Entry> S;
        window_sort(in, n, S, off, cnt);
        match_all(in, n, prm, S, off, cnt, Mf, Mq);
    }
}

// ---- stage 3: speculative segment parse + hand-off verification + repair (k_parse*, k_verify)
struct SegRec {
    uint32_t e_pos, e_key, e_tok;   // first iteration at or after the segment start
    uint32_t x_pos, x_key, x_tok;   // first iteration at or after the segment end (exclusive count)
    std::vector<uint32_t> toks;
};

int step(const Params& prm, ParseState& st, uint32_t n, const uint8_t* d, const std::vector<uint32_t>& Mf,
         const std::vector<uint32_t>& Mq, uint32_t out[2]) {
    uint32_t p = st.pos;
    uint32_t mf = (prm.mode != kRle && p + 2 < n && !Mf.empty()) ? Mf[p] : 0;
    if (prm.mode == kLazy) {
        uint32_t mq = (prm.need_quarter && p + 2 < n) ? Mq[p] : 0;
        return lazy_step(st, n, d, mf, mq, prm.lazy, out);
    }
    if (prm.mode == kGreedy) return greedy_step(st, n, d, mf, out);
    return rle_step(st, n, d, out);
}

void parse_segment(const Params& prm, const uint8_t* d, uint32_t n, const std::vector<uint32_t>& Mf,
                   const std::vector<uint32_t>& Mq, ParseState st, uint32_t a, uint32_t b, SegRec& r) {
    // runs from `st` until the first iteration position >= b (or the end of data)
    r.toks.clear();
    bool have_e = false;
    uint32_t out[2];
    while (st.pos < n) {
        if (!have_e && st.pos >= a) { r.e_pos = st.pos; r.e_key = parse_state_key(st); r.e_tok = (uint32_t)r.toks.size(); have_e = true; }
        if (st.pos >= b) break;
        int ne = step(prm, st, n, d, Mf, Mq, out);
        for (int i = 0; i < ne; i++) r.toks.push_back(out[i]);
    }
    if (!have_e) { r.e_pos = st.pos; r.e_key = parse_state_key(st); r.e_tok = (uint32_t)r.toks.size(); }
    r.x_pos = st.pos; r.x_key = parse_state_key(st); r.x_tok = (uint32_t)r.toks.size();
}

ParseState state_from(uint32_t pos, uint32_t key) {
    ParseState s; s.pos = pos; s.prev_len = key & 0x1ff; s.prev_dist = (key >> 9) & 0xffff; s.add = (key >> 25) & 1; s.ign = (key >> 26) & 1;
    return s;
}

uint32_t g_last_repairs = 0, g_last_seq_repairs = 0;

void parse_all(const Params& prm, const Cfg& cfg, const uint8_t* d, uint32_t n, const std::vector<uint32_t>& Mf,
               const std::vector<uint32_t>& Mq, std::vector<uint32_t>& tokens, uint32_t begin = 0) {
    tokens.clear();
    g_last_repairs = g_last_seq_repairs = 0;
    if (n <= begin) return;
    uint32_t nseg = (n - begin + cfg.pseg - 1) / cfg.pseg;
    std::vector<SegRec> seg(nseg);
    for (uint32_t s = 0; s < nseg; s++) {
        uint32_t a = begin + s * cfg.pseg, b = std::min(n, a + cfg.pseg);
        uint32_t start = a - begin > cfg.warm ? a - cfg.warm : begin;
        parse_segment(prm, d, n, Mf, Mq, parse_state_init(start), a, b, seg[s]);
    }
    auto bad = [&](uint32_t s) {
        return s > 0 && (seg[s - 1].x_pos != seg[s].e_pos || seg[s - 1].x_key != seg[s].e_key);
    };
    for (uint32_t round = 0; round < cfg.rounds; round++) {
        std::vector<uint32_t> list;
        for (uint32_t s = 1; s < nseg; s++) if (bad(s)) list.push_back(s);
        if (list.empty()) break;
        // all repairs of one round read the *previous* round's exits, like a parallel kernel would
        std::vector<SegRec> fixed(list.size());
        for (size_t i = 0; i < list.size(); i++) {
            uint32_t s = list[i];
            uint32_t a = begin + s * cfg.pseg, b = std::min(n, a + cfg.pseg);
            parse_segment(prm, d, n, Mf, Mq, state_from(seg[s - 1].x_pos, seg[s - 1].x_key), a, b, fixed[i]);
        }
        for (size_t i = 0; i < list.size(); i++) { seg[list[i]] = fixed[i]; g_last_repairs++; }
    }
    for (uint32_t s = 1; s < nseg; s++) {   // sequential fallback
        if (bad(s)) {
            uint32_t a = begin + s * cfg.pseg, b = std::min(n, a + cfg.pseg);
            SegRec r;
            parse_segment(prm, d, n, Mf, Mq, state_from(seg[s - 1].x_pos, seg[s - 1].x_key), a, b, r);
            seg[s] = r; g_last_seq_repairs++;
        }
    }
    for (uint32_t s = 0; s < nseg; s++)
        tokens.insert(tokens.end(), seg[s].toks.begin() + seg[s].e_tok, seg[s].toks.begin() + seg[s].x_tok);
}

// ---- stage 4..7: blocks
void or_bits(std::vector<uint8_t>& out, uint64_t bitpos, uint64_t v, uint32_t nbits) {
    for (uint32_t i = 0; i < nbits; i++)
        if ((v >> i) & 1) out[(bitpos + i) >> 3] |= (uint8_t)(1u << ((bitpos + i) & 7));
}

void emit_blocks(const uint8_t* d, uint32_t n, const std::vector<uint32_t>& tokens, std::vector<uint8_t>& out,
                 int final_stream) {
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint64_t T = tokens.size();
    uint32_t nblocks = (uint32_t)(T / kBlockTokens) + 1;
    std::vector<BlockCodes> bc(nblocks);
    std::vector<uint64_t> in_start(nblocks + 1, 0);
    std::vector<uint32_t> scratch(512);
    for (uint32_t b = 0; b < nblocks; b++) {
        uint32_t ll[kNumLL] = {0}, dd[kNumDist] = {0};
        ll[kEob] = 1;
        uint64_t bytes = 0;
        uint64_t t0 = (uint64_t)b * kBlockTokens, t1 = std::min<uint64_t>(T, t0 + kBlockTokens);
        for (uint64_t t = t0; t < t1; t++) {
            uint32_t tk = tokens[t];
            if (tok_dist(tk)) {
                uint32_t c, ne, ev;
                length_symbol(tok_lo(tk), c, ne, ev); ll[c]++;
                dist_symbol(tok_dist(tk), c, ne, ev); dd[c]++;
            } else ll[tok_lo(tk)]++;
            bytes += tok_input_len(tk);
        }
        in_start[b + 1] = in_start[b] + bytes;
        build_block_codes(ll, dd, bytes, bc[b], scratch.data());
    }
    // sequential bit-offset scan (kernel k_block_scan)
    std::vector<uint64_t> bitpos(nblocks + 1, 0);
    std::vector<int> type(nblocks);
    for (uint32_t b = 0; b < nblocks; b++) {
        uint64_t bits;
        type[b] = choose_block(bc[b], (uint32_t)(bitpos[b] & 7), bits);
        bitpos[b + 1] = bitpos[b] + bits;
    }
    out.assign((size_t)((bitpos[nblocks] + 7) / 8), 0);
    (void)n;
    for (uint32_t b = 0; b < nblocks; b++) {
        uint64_t bp = bitpos[b];
        int last = (b + 1 == nblocks) && final_stream;
        uint64_t t0 = (uint64_t)b * kBlockTokens, t1 = std::min<uint64_t>(T, t0 + kBlockTokens);
        if (type[b] == kStored) {
            uint64_t pos = in_start[b], left = bc[b].input_bytes;
            while (left > 0) {
                uint32_t chunk = (uint32_t)std::min<uint64_t>(left, kMaxStored);
                int lastchunk = (left == chunk);
                or_bits(out, bp, (last && lastchunk) ? 1 : 0, 3); bp += 3;
                bp = (bp + 7) & ~7ull;
                or_bits(out, bp, chunk, 16); bp += 16;
                or_bits(out, bp, (~chunk) & 0xffff, 16); bp += 16;
                for (uint32_t i = 0; i < chunk; i++) out[(size_t)(bp >> 3) + i] = d[pos + i];
                bp += 8ull * chunk; pos += chunk; left -= chunk;
            }
            continue;
        }
        const BlockCodes& c = bc[b];
        uint8_t fl[288], fd[32]; uint16_t fcl[288], fcd[32];
        const uint8_t* ll_len = c.ll_len; const uint8_t* d_len = c.d_len;
        const uint16_t* ll_code = c.ll_code; const uint16_t* d_code = c.d_code;
        if (type[b] == kFixed) {
            for (int i = 0; i < 288; i++) fl[i] = (uint8_t)fixed_ll_length(i);
            for (int i = 0; i < 32; i++) fd[i] = 5;
            canonical_codes(fl, 288, fcl); canonical_codes(fd, 32, fcd);
            ll_len = fl; d_len = fd; ll_code = fcl; d_code = fcd;
            or_bits(out, bp, last ? 3 : 2, 3); bp += 3;
        } else {
            or_bits(out, bp, last ? 5 : 4, 3); bp += 3;
            or_bits(out, bp, c.hlit - 257, 5); bp += 5;
            or_bits(out, bp, c.hdist - 1, 5); bp += 5;
            or_bits(out, bp, c.used_hclens - 4, 4); bp += 4;
            for (uint32_t i = 0; i < c.used_hclens; i++) { or_bits(out, bp, c.cl_len[order[i]], 3); bp += 3; }
            for (uint32_t i = 0; i < c.n_hdr_sym; i++) {
                uint32_t sym = c.hdr_sym[i] & 31, rep = c.hdr_sym[i] >> 8;
                or_bits(out, bp, c.cl_code[sym], c.cl_len[sym]); bp += c.cl_len[sym];
                if (sym == 16) { or_bits(out, bp, rep - 3, 2); bp += 2; }
                else if (sym == 17) { or_bits(out, bp, rep - 3, 3); bp += 3; }
                else if (sym == 18) { or_bits(out, bp, rep - 11, 7); bp += 7; }
            }
        }
        for (uint64_t t = t0; t < t1; t++) {
            uint32_t tk = tokens[t];
            if (tok_dist(tk)) {
                uint32_t cc, ne, ev;
                length_symbol(tok_lo(tk), cc, ne, ev);
                or_bits(out, bp, ll_code[cc], ll_len[cc]); bp += ll_len[cc];
                or_bits(out, bp, ev, ne); bp += ne;
                dist_symbol(tok_dist(tk), cc, ne, ev);
                or_bits(out, bp, d_code[cc], d_len[cc]); bp += d_len[cc];
                or_bits(out, bp, ev, ne); bp += ne;
            } else {
                or_bits(out, bp, ll_code[tok_lo(tk)], ll_len[tok_lo(tk)]); bp += ll_len[tok_lo(tk)];
            }
        }
        or_bits(out, bp, ll_code[kEob], ll_len[kEob]); bp += ll_len[kEob];
        if (bp != bitpos[b + 1]) abort();   // the scan and the packer must agree
    }
}

}  // namespace

extern "C" {

// Raw deflate stream for `in` as the GPU pipeline's algorithm produces it.
int dflm_compress(const uint8_t* in, uint32_t n, uint16_t checks, uint16_t lazy, uint8_t mtype, uint32_t pseg,
                  uint32_t warm, uint32_t rounds, uint8_t** out, size_t* out_len, uint32_t* stats /*[4]*/) {
    Params prm = make_params(checks, lazy, mtype);
    Cfg cfg{pseg, warm, rounds};
    std::vector<uint32_t> Mf, Mq, tokens;
    if (prm.mode != kRle && prm.checks > 0) find_matches(in, n, prm, Mf, Mq);
    parse_all(prm, cfg, in, n, Mf, Mq, tokens);
    std::vector<uint8_t> o;
    emit_blocks(in, n, tokens, o, 1);
    *out = (uint8_t*)malloc(o.size() ? o.size() : 1);
    memcpy(*out, o.data(), o.size());
    *out_len = o.size();
    if (stats) { stats[0] = g_last_repairs; stats[1] = g_last_seq_repairs; stats[2] = (uint32_t)tokens.size(); stats[3] = 0; }
    return 0;
}

// tokens only (for token-level comparison with the oracle)
int dflm_tokens(const uint8_t* in, uint32_t n, uint16_t checks, uint16_t lazy, uint8_t mtype, uint32_t pseg,
                uint32_t warm, uint32_t rounds, uint32_t** toks, size_t* ntoks) {
    Params prm = make_params(checks, lazy, mtype);
    Cfg cfg{pseg, warm, rounds};
    std::vector<uint32_t> Mf, Mq, tokens;
    if (prm.mode != kRle && prm.checks > 0) find_matches(in, n, prm, Mf, Mq);
    parse_all(prm, cfg, in, n, Mf, Mq, tokens);
    *toks = (uint32_t*)malloc(tokens.size() * 4 + 4);
    memcpy(*toks, tokens.data(), tokens.size() * 4);
    *ntoks = tokens.size();
    return 0;
}

// tokens of in[begin..n) with in[0..begin) as dictionary (the GPU pipeline's `begin` semantics)
int dflm_tokens_from(const uint8_t* in, uint32_t n, uint32_t begin, uint16_t checks, uint16_t lazy, uint8_t mtype,
                     uint32_t** toks, size_t* ntoks) {
    Params prm = make_params(checks, lazy, mtype);
    Cfg cfg{8192, 1024, 3};
    std::vector<uint32_t> Mf, Mq, tokens;
    if (prm.mode != kRle && prm.checks > 0) find_matches(in, n, prm, Mf, Mq);
    parse_all(prm, cfg, in, n, Mf, Mq, tokens, begin);
    *toks = (uint32_t*)malloc(tokens.size() * 4 + 4);
    memcpy(*toks, tokens.data(), tokens.size() * 4);
    *ntoks = tokens.size();
    return 0;
}

void dflm_symbols(uint32_t len, uint32_t dist, uint32_t* o /*[6]*/) {
    length_symbol(len, o[0], o[1], o[2]);
    dist_symbol(dist, o[3], o[4], o[5]);
}

void dflm_free(void* p) { free(p); }
uint32_t dflm_crc32_combine(uint32_t c1, uint32_t c2, uint64_t len2) { return crc32_combine(c1, c2, len2); }
uint32_t dflm_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2) { return adler32_combine(a1, a2, len2); }
void dflm_set_match_impl(int impl) { g_match_impl = impl; }
void dflm_chain_stats(uint64_t* o /*[8]*/, int reset) {
    for (int i = 0; i < 8; i++) { o[i] = g_chain_stats[i]; if (reset) g_chain_stats[i] = 0; }
}
}
