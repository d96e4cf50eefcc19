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

// Candidates of sorted entry i of window s, most recent first (matching.rs:102-106,127): the entries in front of
// it in its own bucket, then the tail of the same bucket of the previous window at distance <= 32768, at most
// `budget` in total.  Visit k < n_own is Kw[i - 1 - k], visit k >= n_own is Kp[pe - 1 - (k - n_own)].
struct CandRange { uint32_t n_own, n_tot, pe; const Entry* Kw; const Entry* Kp; };
CandRange cand_range(const std::vector<Entry>& K, const std::vector<uint16_t>& off, const std::vector<uint32_t>& cnt,
                     uint32_t s, uint32_t i, uint32_t h, uint32_t pl, uint32_t budget) {
    CandRange c;
    c.Kw = &K[(size_t)s * kWindow];
    c.Kp = nullptr;
    const uint16_t* ow = &off[(size_t)s * kWindow];
    uint32_t s0 = ow[h];
    c.n_own = std::min(budget, i - s0);
    c.n_tot = c.n_own;
    c.pe = 0;
    if (s > 0 && c.n_own < budget) {
        c.Kp = &K[(size_t)(s - 1) * kWindow];
        const uint16_t* op = &off[(size_t)(s - 1) * kWindow];
        uint32_t ps = op[h];
        c.pe = (h + 1 < kWindow) ? op[h + 1] : cnt[s - 1];
        uint32_t rem = budget - c.n_own;
        uint32_t lo = (c.pe - ps > rem) ? c.pe - rem : ps;
        while (lo < c.pe && entry_pos(c.Kp[lo].hi) < pl) lo++;
        c.n_tot = c.n_own + (c.pe - lo);
    }
    return c;
}

// ---- stage 2: entry walk per sorted entry (kernel k_match): final records below 8 bytes, "long" records otherwise
void match_all(const uint8_t* d, uint32_t n, const Params& prm, const std::vector<Entry>& K,
               const std::vector<uint16_t>& off, const std::vector<uint32_t>& cnt,
               std::vector<uint32_t>& Mf, std::vector<uint32_t>& Mq) {
    Mf.assign(n, 0);
    if (prm.need_quarter) Mq.assign(n, 0);
    uint32_t nseg = (uint32_t)cnt.size();
    for (uint32_t s = 0; s < nseg; s++) {
        const Entry* Kw = &K[(size_t)s * kWindow];
        for (uint32_t i = 0; i < cnt[s]; i++) {
            Entry me = Kw[i];
            uint32_t pl = entry_pos(me.hi);
            uint32_t p = s * kWindow + pl;
            uint32_t h = hash3(d[p], d[p + 1], d[p + 2]);
            uint32_t maxl = std::min(kMaxMatch, n - p);
            CandRange c = cand_range(K, off, cnt, s, i, h, pl, prm.checks);
            EntryWalk st = ewalk_init(), sq = st;
            uint32_t dist = 0, qdist = 0;
            for (uint32_t k = 0; k < c.n_tot; k++) {
                if (prm.need_quarter && k == prm.checks_quarter) { sq = st; qdist = dist; }
                Entry ce = k < c.n_own ? c.Kw[i - 1 - k] : c.Kp[c.pe - 1 - (k - c.n_own)];
                uint32_t before = st.best_len;
                ewalk_visit(st, me, ce, k, maxl);
                if (st.best_len != before) dist = (k < c.n_own ? pl : pl + kWindow) - entry_pos(ce.hi);
            }
            if (prm.need_quarter && c.n_tot <= prm.checks_quarter) { sq = st; qdist = dist; }
            Mf[p] = ewalk_record(st, i, dist, maxl);
            if (prm.need_quarter) Mq[p] = ewalk_record(sq, i, qdist, maxl);
        }
    }
}

// ---- stage 1' + 2': max_hash_checks == 1 (kernels k_window_sort<true>, k_match_first).  The one candidate of a position is
// its predecessor in the bucket: the neighbouring item of the sorted order if that has the same hash, else the last
// position of the same bucket in the previous window, if it is within the window.
void match_one(const uint8_t* d, uint32_t n, const std::vector<Entry>& K, const std::vector<uint16_t>& off,
               const std::vector<uint32_t>& cnt, std::vector<uint32_t>& Mf) {
    Mf.assign(n, 0);
    const uint32_t nseg = (uint32_t)cnt.size();
    auto lcp = [&](uint32_t p, uint32_t q) { uint32_t maxl = std::min(kMaxMatch, n - p), l = 0; while (l < maxl && d[p + l] == d[q + l]) l++; return l; };
    std::vector<uint16_t> last_prev, last_cur;       // last position of every bucket of a window, 0xffff = none
    for (uint32_t s = 0; s < nseg; s++) {
        const Entry* Kw = &K[(size_t)s * kWindow];
        const uint16_t* ow = &off[(size_t)s * kWindow];
        last_cur.assign(kWindow, 0xffff);
        for (uint32_t h = 0; h < kWindow; h++) {
            const uint32_t lo = ow[h], hi = h + 1 < kWindow ? ow[h + 1] : cnt[s];
            for (uint32_t i = lo; i < hi; i++) {
                const uint32_t pl = entry_pos(Kw[i].hi), p = s * kWindow + pl;
                uint32_t rec = 0;
                if (i > lo) {                                    // the neighbour in the sorted order, same bucket
                    const uint32_t q = s * kWindow + entry_pos(Kw[i - 1].hi);
                    rec = finalize_match(lcp(p, q), p - q);
                } else if (s > 0 && last_prev[h] != 0xffff && last_prev[h] >= pl) {   // k_match_first
                    const uint32_t q = (s - 1) * kWindow + last_prev[h];
                    rec = finalize_match(lcp(p, q), p - q);
                }
                Mf[p] = rec;
            }
            if (hi > lo) last_cur[h] = (uint16_t)entry_pos(Kw[hi - 1].hi);
        }
        last_prev.swap(last_cur);
    }
}

// ---- stage 3a: resolution of a long record at a position the parser searches (warp-cooperative in k_parse*)
// Every candidate from visit k8 on that shares the target's 8 entry bytes is compared on the data; the first
// strictly longer one wins (matching.rs:148-157), starting from max(floor, 7): the caller only uses a result
// longer than `floor` (= prev_length, matching.rs:161-165) and the record proves a length of at least 8.
uint64_t g_resolve_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // resolutions, candidates compared, bytes compared, visits scanned
uint32_t resolve_long(const uint8_t* d, uint32_t n, const std::vector<Entry>& K, const std::vector<uint16_t>& off,
                      const std::vector<uint32_t>& cnt, uint32_t p, uint32_t rec, uint32_t floor, uint32_t budget,
                      uint32_t full_budget) {
    const uint32_t s = p / kWindow, pl = p % kWindow, i = rec_rank(rec);
    const Entry me = K[(size_t)s * kWindow + i];
    const uint32_t h = hash3(d[p], d[p + 1], d[p + 2]);
    const uint32_t maxl = std::min(kMaxMatch, n - p);
    CandRange c = cand_range(K, off, cnt, s, i, h, pl, full_budget);
    const uint32_t n_vis = std::min(c.n_tot, budget);
    uint32_t best = std::max(floor, kEntryBytes - 1u), best_q = 0;
    g_resolve_stats[0]++;
    for (uint32_t k = rec_k8(rec); k < n_vis; k++) {
        g_resolve_stats[3]++;
        Entry ce = k < c.n_own ? c.Kw[i - 1 - k] : c.Kp[c.pe - 1 - (k - c.n_own)];
        if (ce.lo != me.lo || ((ce.hi ^ me.hi) & kEntryKeyHi) != 0u) continue;
        uint32_t q = (k < c.n_own ? s * kWindow : (s - 1) * kWindow) + entry_pos(ce.hi);
        g_resolve_stats[1]++;
        if (best < maxl && d[q + best] != d[p + best]) continue;   // cannot be longer than the running best
        uint32_t l = kEntryBytes;
        while (l < maxl && d[p + l] == d[q + l]) l++;
        g_resolve_stats[2] += l - kEntryBytes;
        if (l > best) { best = l; best_q = q; if (l == maxl) break; }
    }
    {   // statistics: is the nearest candidate that shares the 8 entry bytes already the answer?
        uint32_t k = rec_k8(rec);
        Entry ce = k < c.n_own ? c.Kw[i - 1 - k] : c.Kp[c.pe - 1 - (k - c.n_own)];
        uint32_t q = (k < c.n_own ? s * kWindow : (s - 1) * kWindow) + entry_pos(ce.hi);
        uint32_t l = 0; while (l < maxl && d[p + l] == d[q + l]) l++;
        uint32_t want = best > std::max(floor, kEntryBytes - 1u) ? best : 0, got = l > std::max(floor, kEntryBytes - 1u) ? l : 0;
        if (want == got) g_resolve_stats[4]++;
        if (floor >= 8) g_resolve_stats[5]++;
        if (want == 0) g_resolve_stats[6]++;
    }
    return best > std::max(floor, kEntryBytes - 1u) ? finalize_match(best, p - best_q) : 0u;
}

int g_force_generic = 0;   // test hook: 1 = the generic walk also for max_hash_checks == 1
void find_matches(const uint8_t* in, uint32_t n, const Params& prm, std::vector<Entry>& S, std::vector<uint16_t>& off,
                  std::vector<uint32_t>& cnt, std::vector<uint32_t>& Mf, std::vector<uint32_t>& Mq) {
    window_sort(in, n, S, off, cnt);
    if (prm.checks == 1u && prm.mode != kRle && !prm.need_quarter && !g_force_generic) match_one(in, n, S, off, cnt, Mf);   // dfl_internal.h one_candidate()
    else match_all(in, n, prm, S, off, cnt, Mf, Mq);
}

// ---- stage 3: speculative segment parse + hand-off verification + repair (k_parse*, k_verify)
struct SegRec {
    uint32_t e_pos, e_key, e_tok;   // first iteration at or after the segment start
    uint32_t x_pos, x_key, x_tok;   // first iteration at or after the segment end (exclusive count)
    std::vector<uint32_t> toks;
};

// What the match stage leaves for the parser: the records and the sorted lists they refer to.
struct MatchData {
    std::vector<Entry> K;
    std::vector<uint16_t> off;
    std::vector<uint32_t> cnt, Mf, Mq;
};

// Token sink of the model: distances and literal bytes are looked up as the token is written (the kernels defer
// exactly these two look-ups to k_compact).
struct ModelSink {
    const uint8_t* d; const MatchData* md; std::vector<uint32_t>* toks;
    uint32_t dist_of(uint32_t ref, uint32_t kind) const {
        return kind == kRefDist ? ref : match_dist((kind == kRefQuarter ? md->Mq : md->Mf)[ref]);
    }
    void literal(uint32_t pos) { toks->push_back(tok_literal(d[pos])); }
    void match(uint32_t len, uint32_t ref, uint32_t kind) { toks->push_back(tok_match(len, dist_of(ref, kind))); }
};

void step(const Params& prm, ParseState& st, uint32_t n, const uint8_t* d, const MatchData& md, ModelSink& out) {
    uint32_t p = st.pos;
    const bool has_m = prm.mode != kRle && p + 2 < n && !md.Mf.empty();
    uint32_t m_len = 0, m_ref = 0, m_kind = kRefDist;
    auto take = [&](uint32_t rec, bool quarter, uint32_t floor) {
        if (rec_is_long(rec)) {      // resolved now: the distance is known
            uint32_t m = resolve_long(d, n, md.K, md.off, md.cnt, p, rec, floor, quarter ? prm.checks_quarter : prm.checks, prm.checks);
            m_len = match_len(m); m_ref = m_len ? match_dist(m) : 0; m_kind = kRefDist;
        } else {                     // final record: only its length is needed to decide; the distance stays behind `p`
            m_len = match_len(rec); m_ref = p; m_kind = quarter ? kRefQuarter : kRefFull;
        }
    };
    if (prm.mode == kLazy) {
        if (has_m && !st.ign) {                       // the only case in which the reference searches (lz77.rs:347)
            const bool quarter = st.prev_len >= 32u;  // lz77.rs:351-355
            if (!quarter || prm.need_quarter) take(quarter ? md.Mq[p] : md.Mf[p], quarter, st.prev_len);
        }
        lazy_step(st, n, m_len, m_ref, m_kind, prm.lazy, out);
    } else if (prm.mode == kGreedy) {
        if (has_m) take(md.Mf[p], false, 0);
        greedy_step(st, n, m_len, m_ref, m_kind, out);
    } else {
        rle_step(st, n, d, out);
    }
}

void parse_segment(const Params& prm, const uint8_t* d, uint32_t n, const MatchData& md, ParseState st, uint32_t a,
                   uint32_t b, SegRec& r) {
    // runs from `st` until the first iteration position >= b (or the end of data)
    r.toks.clear();
    bool have_e = false;
    ModelSink out{d, &md, &r.toks};
    auto key = [&](const ParseState& s) { return parse_state_key(s, s.prev_len ? out.dist_of(s.prev_ref, s.prev_kind) : 0u); };
    while (st.pos < n) {
        if (!have_e && st.pos >= a) { r.e_pos = st.pos; r.e_key = key(st); r.e_tok = (uint32_t)r.toks.size(); have_e = true; }
        if (st.pos >= b) break;
        step(prm, st, n, d, md, out);
    }
    if (!have_e) { r.e_pos = st.pos; r.e_key = key(st); r.e_tok = (uint32_t)r.toks.size(); }
    r.x_pos = st.pos; r.x_key = key(st); r.x_tok = (uint32_t)r.toks.size();
}

ParseState state_from(uint32_t pos, uint32_t key) { return parse_state_from_key(pos, key); }

uint32_t g_last_repairs = 0, g_last_seq_repairs = 0;

void parse_all(const Params& prm, const Cfg& cfg, const uint8_t* d, uint32_t n, const MatchData& md,
               std::vector<uint32_t>& tokens, uint32_t begin = 0) {
    tokens.clear();
    g_last_repairs = g_last_seq_repairs = 0;
    if (n <= begin) return;
    uint32_t nseg = (n - begin + cfg.pseg - 1) / cfg.pseg;
    std::vector<SegRec> seg(nseg);
    for (uint32_t s = 0; s < nseg; s++) {
        uint32_t a = begin + s * cfg.pseg, b = std::min(n, a + cfg.pseg);
        uint32_t start = a - begin > cfg.warm ? a - cfg.warm : begin;
        parse_segment(prm, d, n, md, parse_state_init(start), a, b, seg[s]);
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
        std::vector<uint8_t> isbad(nseg, 0);
        for (uint32_t s : list) isbad[s] = 1;
        for (size_t i = 0; i < list.size(); i++) {
            uint32_t s = list[i];
            uint32_t a = begin + s * cfg.pseg, b = std::min(n, a + cfg.pseg);
            ParseState st = state_from(seg[s - 1].x_pos, seg[s - 1].x_key);
            if ((round & 1u) && isbad[s - 1]) {
                // kernel k_chain_predict: inside a chain of maximum-length matches the parser is blank every 258
                // bytes, so the entry follows from the exit of the last good segment in front of the chain
                int g = (int)s - 1;
                while (g >= 0 && isbad[g]) g--;
                if (g >= 0 && seg[g].x_key == 0) {
                    uint32_t xg = seg[g].x_pos;
                    uint32_t steps = a > xg ? (a - xg + kMaxMatch - 1) / kMaxMatch : 0;
                    st = state_from(xg + steps * kMaxMatch, 0);
                }
            }
            parse_segment(prm, d, n, md, st, a, b, fixed[i]);
        }
        for (size_t i = 0; i < list.size(); i++) { seg[list[i]] = fixed[i]; g_last_repairs++; }
    }
    for (uint32_t s = 1; s < nseg; s++) {   // sequential fallback
        if (bad(s)) {
            uint32_t a = begin + s * cfg.pseg, b = std::min(n, a + cfg.pseg);
            SegRec r;
            parse_segment(prm, d, n, md, state_from(seg[s - 1].x_pos, seg[s - 1].x_key), a, b, r);
            seg[s] = r; g_last_seq_repairs++;
        }
    }
    for (uint32_t s = 0; s < nseg; s++)
        tokens.insert(tokens.end(), seg[s].toks.begin() + seg[s].e_tok, seg[s].toks.begin() + seg[s].x_tok);
}

// ---- Lazy with lazy_if_less_than < 3: the reference's loop itself (kernel k_lz77_seq, dfl_core.h lz77_sequential)
void sequential_tokens(const uint8_t* in, uint32_t n, const Params& prm, std::vector<uint32_t>& tokens) {
    std::vector<uint32_t> head(kWindow, kSeqNone), pv(kWindow, kSeqNone), po(kWindow, kSeqNone);
    tokens.assign((size_t)n + 4, 0);
    tokens.resize((size_t)lz77_sequential(in, n, prm, head.data(), pv.data(), po.data(), tokens.data()));
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
    std::vector<uint32_t> tokens;
    MatchData md;
    if (prm.mode == kLazy && prm.lazy < 3) {          // the library's dispatch: k_lz77_seq
        sequential_tokens(in, n, prm, tokens);
    } else {
        if (prm.mode != kRle && prm.checks > 0) find_matches(in, n, prm, md.K, md.off, md.cnt, md.Mf, md.Mq);
        parse_all(prm, cfg, in, n, md, tokens);
    }
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
    std::vector<uint32_t> tokens;
    MatchData md;
    if (prm.mode == kLazy && prm.lazy < 3) {          // the library's dispatch: k_lz77_seq
        sequential_tokens(in, n, prm, tokens);
    } else {
        if (prm.mode != kRle && prm.checks > 0) find_matches(in, n, prm, md.K, md.off, md.cnt, md.Mf, md.Mq);
        parse_all(prm, cfg, in, n, md, tokens);
    }
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
    std::vector<uint32_t> tokens;
    MatchData md;
    if (prm.mode != kRle && prm.checks > 0) find_matches(in, n, prm, md.K, md.off, md.cnt, md.Mf, md.Mq);
    parse_all(prm, cfg, in, n, md, tokens, begin);
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
void dflm_force_generic_match(int on) { g_force_generic = on; }
void dflm_resolve_stats(uint64_t* o /*[8]*/, int reset) {
    for (int i = 0; i < 8; i++) { o[i] = g_resolve_stats[i]; if (reset) g_resolve_stats[i] = 0; }
}

}
