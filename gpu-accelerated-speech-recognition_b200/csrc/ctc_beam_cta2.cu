// ctc_beam_cta2.cu -- CTC prefix beam search, one CTA per utterance for WHOLE sequences (beam <= 32, vocabulary <= 32): the
// lowest-latency decoder (cfg2 single batch, the streaming pipeline).  Stands behind CTCBeamSearch::decode (reference
// CTCBeamSearch.cu:262-312).
// =====================================================================================================
// Latency path, second generation: 4 main warps + 1 auxiliary warp per utterance (beam <= 32, vocabulary <= 32,
// whole sequence in one launch).  Same algorithm, same bits as the kernels above; what changed is the critical
// path of a frame:
//   * the trie (global memory: child lookup, node allocation) is owned by the AUX warp and runs one frame BEHIND
//     the beam: nothing on the main path needs node ids any more -- twin / parent / absorbed relations of the new
//     beam are derived from the prefix-relation matrix while it is updated (REL_EQ = twin, proper prefix with
//     depth + 1 = parent), so the global-memory latency of the trie is off the critical path;
//   * the aux warp also fetches (and, in the streaming pipeline, waits for) the next frame's log-probabilities
//     and ranks them, one frame ahead, into a shared-memory ring;
//   * "stay" candidates are produced in the candidate phase itself (every lane sums up to three addends in
//     canonical order; missing addends are the merge's neutral element, which it returns bit-exactly);
//   * probe cells are written while the candidates are produced; every survivor's rank is counted by two
//     threads and written straight to its slot -- 5 block-wide barriers per frame instead of 9.
// =====================================================================================================
#include "ctc_beam.cuh"

namespace gasr {

template <int BMAX>
struct Cta2Beam {
    float sc[2][BMAX];
    int node[2][BMAX];
    int depth[2][BMAX];
    int pk[2][BMAX];
    int tw[BMAX], p0[BMAX], p1[BMAX];
    unsigned abs0[BMAX], abs1[BMAX];
    unsigned selkey[2][BMAX];
    int seli[2][BMAX], selv[2][BMAX];
    int sel_m[2];
    unsigned char rel[2][BMAX][BMAX];
    unsigned cand[BMAX][32];
    alignas(16) unsigned ckey[128];
    unsigned surv_key[128];
    unsigned short surv_iv[128];
    float lpring[2][32];
    int order[2][32], rankof[2][32];
    unsigned theta;
    int ns, nodes;
    int sanc[2][BMAX];                  // trie warp: skip pointer of each kept state's node
    unsigned char cellmap[BMAX * 32];   // (parent rank, score rank) -> probe cell, 255 = none (copied from the parameters)
};

template <int MW> __device__ __forceinline__ void cta2_bar_main() { asm volatile("bar.sync 1, %0;" ::"n"(MW * 32) : "memory"); }
template <int MW> __device__ __forceinline__ void cta2_bar_all() { asm volatile("bar.sync 2, %0;" ::"n"(MW * 32 + 64) : "memory"); }

// MW = number of main warps; warp MW fetches and ranks the log-probabilities, warp MW + 1 owns the trie
template <int DOMAIN, int BMAX, int MW>
__global__ void __launch_bounds__(MW * 32 + 64) ctc_beam_cta2_kernel(const CtcParams p) {
    constexpr int MT = MW * 32;                          // main threads
    __shared__ __align__(16) Cta2Beam<BMAX> cb;
    __shared__ char vch_s[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int utt = blockIdx.x;
    const int V = p.V, B = p.beam, blank = p.blank, Vp = p.Vp, T = utt_frames(p, (int)blockIdx.x);   // this utterance's frames
    constexpr unsigned FULL = 0xffffffffu;
    const bool active = lane < V;
    const char *vch = vch_s;
    const float NEUTRAL = DOMAIN ? -INFINITY : 0.0f;      // merge's neutral element (returned bit-exactly)

    int *parent = p.parent + (size_t)utt * p.cap;
    int *meta = p.meta + (size_t)utt * p.cap;
    int *born = p.born ? p.born + (size_t)utt * p.cap : nullptr;
    int *anc = p.anc + (size_t)utt * p.cap;
    int *child = p.child + (size_t)utt * p.cap * Vp;
    const float *S = p.scores + (size_t)utt * p.ld;
    const size_t frame_stride = (size_t)p.frame_rows * p.ld;

    if (tid < V) vch_s[tid] = p.vocab[tid];
    if (tid < Vp) child[tid] = 0;
    if (tid < BMAX) { cb.tw[tid] = kNone; cb.p0[tid] = kNone; cb.p1[tid] = kNone; cb.abs0[tid] = 0u; cb.abs1[tid] = 0u; }
    if (tid < 128) cb.ckey[tid] = 0u;
    for (int i = tid; i < BMAX * 32; i += MT + 64) cb.cellmap[i] = p.cellmap[i];
    if (tid == 0) {
        parent[0] = -1; meta[0] = 0xff; anc[0] = 0;
        cb.sc[0][0] = DOMAIN ? 0.0f : 1.0f;
        cb.node[0][0] = 0; cb.depth[0][0] = 0; cb.pk[0][0] = 0xff | (1 << 8);
        cb.rel[0][0][0] = REL_EQ;
        cb.nodes = 1; cb.theta = 0u; cb.ns = 0; cb.sanc[0][0] = 0;
    }


    if (w == MW) {
        // =============================== fetch warp: log-probabilities one frame ahead ===============================
        const volatile unsigned *lpr = p.lp_ready;
        int ready_frames = lpr != nullptr ? 0 : T;
        auto fetch_frame = [&](int t) {                  // log-probabilities of frame t -> ring slot t & 1, ranked
            while (t >= ready_frames) {                  // streaming: wait for the producer's block counter
                const int blk = ready_frames / p.lp_fpb;
                unsigned v = 0;
                unsigned long long t_start = 0;
                do {
                    if (lane == 0) v = lpr[blk];
                    v = __shfl_sync(FULL, v, 0);
                    if (v < (unsigned)p.lp_need) {
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t_start == 0) t_start = now;
                        if ((p.abort && *p.abort) || now - t_start > 2000000000ull) {
                            if (p.abort) *p.abort = 1u;
                            if (p.error) { *reinterpret_cast<volatile int *>(p.error) = 4; __threadfence_system(); }
                            break;
                        }
                        __nanosleep(100);
                    }
                } while (v < (unsigned)p.lp_need);
                ready_frames = (blk + 1) * p.lp_fpb;
                __threadfence();                         // acquire side of the counter: the rows are read after this fence
            }
            const float lpv = active ? __ldcg(S + (size_t)t * frame_stride + lane) : 0.0f;
            const unsigned mine = active ? f2ord(lpv) : 0u;
            int lr = 0;
#pragma unroll
            for (int u = 0; u < 32; u++) {
                const unsigned x = __shfl_sync(FULL, mine, u);
                lr += (x > mine || (x == mine && u < lane)) ? 1 : 0;
            }
            cb.lpring[t & 1][lane] = lpv;
            cb.rankof[t & 1][lane] = lr;
            cb.order[t & 1][lr] = lane;
        };
        fetch_frame(0);
        __syncthreads();
        for (int t = 0; t < T; t++) {
            if (t + 1 < T) fetch_frame(t + 1);
            cta2_bar_all<MW>();
        }
        cta2_bar_all<MW>();
        return;
    }
    if (w == MW + 1) {
        // =============================== trie warp: one frame behind the beam ===============================
        __syncthreads();
        int cur = 0, kept = 1;
        for (int t = 0; t < T; t++) {
            cta2_bar_all<MW>();                              // frame t's selection is visible; my previous trie work is done
            // ---- trie: the selected candidates become nodes (child-table lookup, allocation on a miss); this warp has
            // a whole frame for the global-memory round trip
            const int nxt = cur ^ 1, sb = t & 1;
            const int m = cb.sel_m[sb];
            bool need_new = false;
            int nd = 0, pn = 0, dp = 0, v = 0, an = 0;
            if (lane < m) {
                const int i = cb.seli[sb][lane];
                v = cb.selv[sb][lane];
                const int pki = cb.pk[cur][i];
                const int ebi = (pki >> 8) & 1, lasti = pki & 0xff;
                if (v == blank || (ebi == 0 && v == lasti)) { nd = cb.node[cur][i]; an = cb.sanc[cur][i]; }
                else {
                    pn = cb.node[cur][i]; dp = cb.depth[cur][i] + 1;
                    nd = child[(size_t)pn * Vp + v];
                    need_new = nd == 0;
                    an = ((dp - 1) & 31) == 0 ? pn : cb.sanc[cur][i];      // a parent at a multiple-of-32 depth starts a new block
                    if (!need_new) an = anc[nd];                            // re-created prefix (rare): its own record
                }
            }
            const unsigned nb = __ballot_sync(FULL, need_new);
            const int nodes = cb.nodes;
            if (need_new) {
                nd = nodes + __popc(nb & ((1u << lane) - 1u));
                parent[nd] = pn;
                meta[nd] = (dp << 8) | v;
                if (born) born[nd] = t;
                anc[nd] = an;
                child[(size_t)pn * Vp + v] = nd;
                int4 *row = reinterpret_cast<int4 *>(child + (size_t)nd * Vp);
                for (int q = 0; q < Vp / 4; q++) row[q] = make_int4(0, 0, 0, 0);
            }
            if (lane < m) { cb.node[nxt][lane] = nd; cb.sanc[nxt][lane] = an; }
            if (lane == 0) cb.nodes = nodes + __popc(nb);
            __syncwarp();
            kept = m;
            cur = nxt;
        }
        __syncwarp();
        // ---- result (CTCBeamSearch.cu:290-298): the aux warp owns the trie, so it writes the paths ----
        cta2_bar_all<MW>();                                  // final scores / depths / pk are in place
        if (lane == 0 && p.out_counts) p.out_counts[utt] = kept;
        // the path of a kept state is read off the trie leaf-to-root; the skip pointers cut the chain of dependent loads
        // from depth to depth / 32 + 32: lane 0 collects the 32-block end nodes, then every lane walks one block
        int *ends = reinterpret_cast<int *>(&cb.cand[0][0]);          // the beam is final: the candidate matrix is free
        constexpr int ENDS_CAP = BMAX * 32;
        for (int r = 0; r < p.nbest; r++) {
            char *out = p.out_paths + ((size_t)utt * p.nbest + r) * p.max_len;
            int *ts = p.out_ts ? p.out_ts + ((size_t)utt * p.nbest + r) * p.max_len : nullptr;
            int len = 0;
            float scv = 0.0f;
            if (r < kept) {
                const int nd0 = cb.node[cur][r];
                const int dpt = cb.depth[cur][r];
                len = dpt;
                if (T == 1 && ((cb.pk[cur][r] >> 8) & 1)) { if (lane == 0 && len < p.max_len) { out[len] = vch[blank]; if (ts) ts[len] = 0; } len += 1; }
                const int nblk = (dpt + 31) >> 5;
                if (nblk <= ENDS_CAP) {
                    if (lane == 0) {
                        int nd = nd0;
                        for (int j = nblk - 1; j >= 0; j--) { ends[j] = nd; nd = anc[nd]; }
                    }
                    __syncwarp();
                    for (int j = lane; j < nblk; j += 32) {
                        int nd = ends[j];
                        const int top = min(dpt, 32 * (j + 1));
                        for (int pos = top - 1; pos >= 32 * j; pos--) {
                            if (pos < p.max_len) { out[pos] = vch[meta[nd] & 0xff]; if (ts) ts[pos] = born[nd]; }
                            nd = parent[nd];
                        }
                    }
                    __syncwarp();
                } else if (lane == 0) {
                    int nd = nd0;
                    for (int pos = dpt - 1; pos >= 0; pos--) {
                        if (pos < p.max_len) { out[pos] = vch[meta[nd] & 0xff]; if (ts) ts[pos] = born[nd]; }
                        nd = parent[nd];
                    }
                }
                scv = cb.sc[cur][r];
            }
            if (lane == 0) {
                p.out_lens[(size_t)utt * p.nbest + r] = len;
                p.out_scores[(size_t)utt * p.nbest + r] = scv;
            }
        }
        return;
    }

    // =============================== main warps ===============================
    __syncthreads();
    int cur = 0, kept = 1;
    int stat_surv = 0, stat_fallback = 0;
    constexpr int NC = BMAX <= 16 ? 64 : 128;
    constexpr int TPC = MT / NC;                         // threads per probe cell
    constexpr int TPS = MT / 64;                         // threads per survivor in the ranking phase (halved above 64 survivors)
    static_assert(TPS >= 2, "the ranking phase needs at least two threads per survivor slot");

    for (int t = 0; t < T; t++) {
        const int k = kept, slot = t & 1, sb = t & 1, nxt = cur ^ 1;
        const bool last_frame = (t == T - 1) && (t > 0);
        const float lp = cb.lpring[slot][lane];
        const float *sc = cb.sc[cur];
        const int *pk = cb.pk[cur], *depth = cb.depth[cur];
        const unsigned char (*rel)[BMAX] = cb.rel[cur];
        const int *order = cb.order[slot];

        // ================= phase B: merged candidates (+ probe cells) =================
        if (!last_frame) {
            const int jr = cb.rankof[slot][lane];
            if (w < MW - 1) {
                // ---- row warps: parent rows w, w + (MW-1), ... processed together, stage by stage, branch-free.  A plain
                // candidate is one add; only rows that own a twin pair run the merge.  The "stay" cell of a row is left to
                // the stay warp below.
                constexpr int RW = MW - 1;
                constexpr int RPWB = (BMAX + RW - 1) / RW;
                int ri[RPWB], pki[RPWB], twi[RPWB];
                unsigned ab0[RPWB];
                bool have[RPWB];
#pragma unroll
                for (int q = 0; q < RPWB; q++) {
                    const int i = w + RW * q;
                    have[q] = i < k;
                    ri[q] = have[q] ? i : 0;
                    pki[q] = pk[ri[q]]; twi[q] = cb.tw[ri[q]]; ab0[q] = cb.abs0[ri[q]];
                }
                float acc[RPWB], x1[RPWB];
                bool keep[RPWB], stay[RPWB];
                bool need1 = false;
#pragma unroll
                for (int q = 0; q < RPWB; q++) {
                    const int i = ri[q];
                    const int ebi = (pki[q] >> 8) & 1, lasti = pki[q] & 0xff;
                    stay[q] = (ebi == 0 && lane == lasti);
                    const bool is_blank = (lane == blank);
                    const bool member = twi[q] >= 0 && (is_blank || ebi == 0 || lane != lasti);
                    const bool dead = (member && twi[q] < i) || (!is_blank && ((ab0[q] >> lane) & 1u));
                    const bool twin_owner = have[q] && member && twi[q] > i && !stay[q];
                    acc[q] = comb<DOMAIN>(sc[i], lp);
                    x1[q] = twin_owner ? comb<DOMAIN>(sc[twin_owner ? twi[q] : 0], lp) : NEUTRAL;
                    need1 |= twin_owner;
                    keep[q] = have[q] && active && !dead;
                }
                if (__any_sync(FULL, need1)) {            // merging the neutral element returns the other operand bit-exactly
#pragma unroll
                    for (int q = 0; q < RPWB; q++) acc[q] = mrg_bf<DOMAIN>(acc[q], x1[q]);
                }
#pragma unroll
                for (int q = 0; q < RPWB; q++) {
                    if (have[q] && !stay[q]) {
                        const unsigned key = keep[q] ? f2ord(acc[q]) : 0u;
                        cb.cand[ri[q]][lane] = key;
                        const int c = cb.cellmap[ri[q] * 32 + jr];
                        if (c != 255) cb.ckey[c] = key;
                    }
                }
            } else {
                // ---- stay warp: lane = kept state (X, 0); its "stay" candidate sums up to three addends -- (P,0)+last if
                // last(P) != last, (P,1)+last, (X,0)+last -- in ascending parent rank (kNone = -1 sorts first)
                for (int i0 = 0; i0 < k; i0 += 32) {
                    const int i = i0 + lane;
                    const bool on = i < k;
                    const int ii = on ? i : 0;
                    const int pkI = pk[ii];
                    const int lasti = pkI & 0xff;
                    const bool is_state0 = on && ((pkI >> 8) & 1) == 0 && lasti < 32;
                    const int q0 = cb.p0[ii], q1 = cb.p1[ii];
                    const int pq0 = pk[q0 >= 0 ? q0 : 0];
                    int s0 = (q0 >= 0 && (pq0 & 0xff) == lasti) ? kNone : q0, s1 = q1, s2 = ii;
                    int lo = min(s0, s1), hi = max(s0, s1);
                    s0 = lo; s1 = hi;
                    lo = min(s1, s2); hi = max(s1, s2);
                    s1 = lo; s2 = hi;
                    lo = min(s0, s1); hi = max(s0, s1);
                    s0 = lo; s1 = hi;
                    const int t0 = s0 >= 0 ? s0 : (s1 >= 0 ? s1 : s2);
                    const int t1 = s0 >= 0 ? s1 : (s1 >= 0 ? s2 : kNone);
                    const int t2 = s0 >= 0 ? s2 : kNone;
                    const float lpv = cb.lpring[slot][is_state0 ? lasti : 0];
                    float acc = comb<DOMAIN>(sc[t0], lpv);
                    const bool n1 = is_state0 && t1 >= 0, n2 = is_state0 && t2 >= 0;
                    if (__any_sync(FULL, n1)) {
                        acc = mrg_bf<DOMAIN>(acc, n1 ? comb<DOMAIN>(sc[t1 >= 0 ? t1 : 0], lpv) : NEUTRAL);
                        if (__any_sync(FULL, n2)) acc = mrg_bf<DOMAIN>(acc, n2 ? comb<DOMAIN>(sc[t2 >= 0 ? t2 : 0], lpv) : NEUTRAL);
                    }
                    if (is_state0) {
                        const unsigned key = lasti < V ? f2ord(acc) : 0u;
                        cb.cand[ii][lasti] = key;
                        const int c = cb.cellmap[ii * 32 + cb.rankof[slot][lasti]];
                        if (c != 255) cb.ckey[c] = key;
                    }
                }
            }
        } else if (w == 0) {
            const float lpb = __shfl_sync(FULL, lp, blank);
            for (int i = 0; i < k; i++) {
                const int pki = pk[i], twi = cb.tw[i];
                const int ebi = (pki >> 8) & 1, lasti = pki & 0xff;
                const float s = comb<DOMAIN>(sc[i], lp);
                float acc = s;
                bool dead = false;
                const bool is_stay = (ebi == 0 && lane == lasti);
                const bool is_blank = (lane == blank);
                const bool member = twi >= 0 && (is_blank || ebi == 0 || lane != lasti);
                if (!is_blank) {
                    if (twi >= 0) {
                        if (twi < i) dead = member;
                        else {
                            const float mm = mrg<DOMAIN>(s, comb<DOMAIN>(sc[twi], lp));
                            acc = member ? mm : s;
                        }
                    }
                    if ((cb.abs0[i] >> lane) & 1u) dead = true;
                    if (!dead && !is_stay && ((cb.abs1[i] >> lane) & 1u)) {
                        // kept (X.v, 1): on the last frame its blank candidate strips to X.v as well
                        for (int j = 0; j < k; j++)
                            if ((ebi ? cb.p1[j] : cb.p0[j]) == i && (pk[j] & 0xff) == lane && ((pk[j] >> 8) & 1))
                                acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[j], lpb));
                    }
                } else {
                    if (ebi == 0 || twi >= 0) dead = true;
                    else if (lasti != 0xff) {
                        const int q0 = cb.p0[i], q1 = cb.p1[i];
                        if (q1 >= 0 || (q0 >= 0 && (pk[q0] & 0xff) != lasti)) dead = true;
                    }
                }
                if (is_stay) {
                    int m0 = cb.p0[i], m1 = cb.p1[i], m2 = i, tmp;
                    if (m0 >= 0 && (pk[m0] & 0xff) == lasti) m0 = kNone;
                    if (m0 > m1) { tmp = m0; m0 = m1; m1 = tmp; }
                    if (m1 > m2) { tmp = m1; m1 = m2; m2 = tmp; }
                    if (m0 > m1) { tmp = m0; m0 = m1; m1 = tmp; }
                    acc = comb<DOMAIN>(sc[m0 >= 0 ? m0 : (m1 >= 0 ? m1 : m2)], lp);
                    if (m0 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[m1], lp));
                    if (m1 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[m2], lp));
                    int b0 = i, b1 = twi;
                    if (b1 >= 0 && b1 < b0) { b0 = twi; b1 = i; }
                    acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[b0], lpb));
                    if (b1 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[b1], lpb));
                    dead = false;
                }
                const unsigned key = (active && !dead) ? f2ord(acc) : 0u;
                cb.cand[i][lane] = key;
            }
            __syncwarp();
            // probe cells of the last frame are gathered after the fact (one warp computed everything)
            for (int c = lane; c < NC; c += 32) {
                const int ci = p.cell_i[c];
                cb.ckey[c] = ci < k ? cb.cand[ci][order[p.cell_j[c]]] : 0u;
            }
        }
        cta2_bar_main<MW>();

        // ================= phase C: lower bound of the beam-th largest merged key =================
        {
            const int c = tid / TPC;
            const unsigned mine = cb.ckey[c];
            constexpr int span = NC / TPC;
            const int ob = (tid % TPC) * span;
            int cnt = 0;
            static_assert(span % 4 == 0, "probe cells are scanned four at a time");
#pragma unroll
            for (int o = 0; o < span; o += 4) {
                const uint4 x = *reinterpret_cast<const uint4 *>(&cb.ckey[ob + o]);
                cnt += (x.x > mine || (x.x == mine && ob + o < c)) ? 1 : 0;
                cnt += (x.y > mine || (x.y == mine && ob + o + 1 < c)) ? 1 : 0;
                cnt += (x.z > mine || (x.z == mine && ob + o + 2 < c)) ? 1 : 0;
                cnt += (x.w > mine || (x.w == mine && ob + o + 3 < c)) ? 1 : 0;
            }
#pragma unroll
            for (int off = 1; off < TPC; off <<= 1) cnt += __shfl_xor_sync(FULL, cnt, off);
            if (cnt == B - 1 && mine != 0u && (tid % TPC) == 0) atomicMax(&cb.theta, mine);
            // second bound: the best parent's beam best-ranked candidates, if none of them was absorbed
            if (w == MW - 1 && B <= V) {
                const unsigned mn = __reduce_min_sync(FULL, lane < B ? cb.cand[0][order[lane]] : 0xffffffffu);
                if (lane == 0 && mn != 0u) atomicMax(&cb.theta, mn);
            }
        }
        cta2_bar_main<MW>();
        const unsigned theta = cb.theta;

        // ================= phase D: survivors =================
        {
            constexpr int RPW = (BMAX + MW - 1) / MW;             // rows per warp
            unsigned keys4[RPW], masks4[RPW];
            int total = 0;
#pragma unroll
            for (int q = 0; q < RPW; q++) {
                const int i = w + MW * q;
                const unsigned key = i < k ? cb.cand[i][lane] : 0u;
                const bool sv = key != 0u && key >= theta;
                keys4[q] = key;
                masks4[q] = __ballot_sync(FULL, sv);
                total += __popc(masks4[q]);
            }
            int base = 0;
            if (lane == 0 && total) base = atomicAdd(&cb.ns, total);
            base = __shfl_sync(FULL, base, 0);
#pragma unroll
            for (int q = 0; q < RPW; q++) {
                const int pos = base + __popc(masks4[q] & ((1u << lane) - 1u));
                if (((masks4[q] >> lane) & 1u) && pos < 128) {
                    cb.surv_key[pos] = keys4[q]; cb.surv_iv[pos] = (unsigned short)(((w + MW * q) << 8) | lane);
                }
                base += __popc(masks4[q]);
            }
            // housekeeping for the next frame (read again only after the barriers below)
            if (tid < NC) cb.ckey[tid] = 0u;
        }
        cta2_bar_main<MW>();
        const int ns = cb.ns;
        if (tid == 0) { stat_surv += ns; stat_fallback += ns > 128; }

        // ================= phase E: exact order of the survivors, straight into their slots =================
        int m = 0;
        if (ns <= 128) {
            m = ns < B ? ns : B;
            // TPS threads share a survivor (fewer when the bound was loose and survivors are many)
            // threads per survivor: as many as fit (a power of two, at most 16, at least MT / 128)
            constexpr int TS_MIN = MT == 128 ? 0 : MT == 256 ? 1 : 2;
            int tshift = TS_MIN;
            for (int cap = 64; cap >= ns && tshift < 4; cap >>= 1) tshift++;
            const int tps = 1 << tshift;
            const int sidx = tid >> tshift, half = tid & (tps - 1);
            int rank = 0;
            unsigned key = 0u;
            int iv = 0;
            if (sidx < ns) {
                key = cb.surv_key[sidx];
                iv = cb.surv_iv[sidx];
                for (int o = half; o < ns; o += tps) {
                    const unsigned ok = cb.surv_key[o];
                    rank += ok > key ? 1 : 0;
                    if (ok == key && o != sidx) {             // exact tie: raw-string order (rare)
                        const int oiv = cb.surv_iv[o];
                        if (t == 0) rank += oiv < iv;
                        else {
                            const int mi = iv >> 8, mv = iv & 0xff;
                            const int ms = cand_suffix_id(mv, blank, pk[mi]);
                            const int oi = oiv >> 8, ov = oiv & 0xff;
                            rank += cand_less_rel(rel[oi][mi], cand_suffix_id(ov, blank, pk[oi]), ms, vch) ? 1 : 0;
                        }
                    }
                }
            }
#pragma unroll
            for (int off = 1; off < 16; off <<= 1) {
                const int other = __shfl_xor_sync(FULL, rank, off);
                if (off < tps) rank += other;
            }
            if (sidx < ns && half == 0 && rank < B) {
                cb.selkey[sb][rank] = key; cb.seli[sb][rank] = iv >> 8; cb.selv[sb][rank] = iv & 0xff;
            }
            if (tid == 0) cb.sel_m[sb] = m;
        } else {
            // more than 128 survivors (very loose bound): beam rounds of warp-max extraction on warp 0
            if (w == 0) {
                unsigned lmax = 0u;
                for (int i = 0; i < k; i++) lmax = max(lmax, cb.cand[i][lane]);
                for (m = 0; m < B; m++) {
                    const unsigned gmax = __reduce_max_sync(FULL, lmax);
                    if (gmax == 0u) break;
                    const unsigned any = __ballot_sync(FULL, lmax == gmax);
                    int wl = __ffs(any) - 1;
                    unsigned x = lane < k ? cb.cand[lane][wl] : 0u;
                    const unsigned colmask = __ballot_sync(FULL, x == gmax);
                    int wi = __ffs(colmask) - 1;
                    if (t > 0 && (__popc(any) > 1 || __popc(colmask) > 1)) {
                        int bi = -1, bs = 0;
                        if (lmax == gmax) {
                            for (int i = 0; i < k; i++) {
                                if (cb.cand[i][lane] != gmax) continue;
                                const int si = cand_suffix_id(lane, blank, pk[i]);
                                if (bi < 0 || cand_less_rel(rel[i][bi], si, bs, vch)) { bi = i; bs = si; }
                            }
                        }
                        int bl = lane;
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            const int oi = __shfl_xor_sync(FULL, bi, off), os = __shfl_xor_sync(FULL, bs, off);
                            const int ol = __shfl_xor_sync(FULL, bl, off);
                            if (oi >= 0 && (bi < 0 || cand_less_rel(rel[oi][bi], os, bs, vch))) { bi = oi; bs = os; bl = ol; }
                        }
                        wi = bi; wl = bl;
                        x = lane < k ? cb.cand[lane][wl] : 0u;
                    }
                    if (lane == wi) { cb.cand[wi][wl] = 0u; x = 0u; }
                    const unsigned cmax = __reduce_max_sync(FULL, x);
                    if (lane == wl) lmax = cmax;
                    if (lane == 0) { cb.selkey[sb][m] = gmax; cb.seli[sb][m] = wi; cb.selv[sb][m] = wl; }
                    __syncwarp();
                }
                if (lane == 0) cb.sel_m[sb] = m;
            }
        }
        // relations of the next beam are rebuilt below: clear them (last read in phase B)
        if (tid < BMAX) { cb.tw[tid] = kNone; cb.p0[tid] = kNone; cb.p1[tid] = kNone; cb.abs0[tid] = 0u; cb.abs1[tid] = 0u; }
        cta2_bar_all<MW>();                                  // selection visible to everybody (aux warp: trie update may start)
        m = cb.sel_m[sb];

        // ================= phase F: next beam: scores, labels, prefix relations and the relations derived from them ====
        {
            const int *seli = cb.seli[sb], *selv = cb.selv[sb];
            if (tid < m) {
                const int i = seli[tid], v = selv[tid];
                const int pki = pk[i];
                const int ebi = (pki >> 8) & 1, lasti = pki & 0xff;
                int dp, npk;
                if (v == blank) { dp = depth[i]; npk = lasti | (1 << 8); }
                else if (ebi == 0 && v == lasti) { dp = depth[i]; npk = lasti; }
                else { dp = depth[i] + 1; npk = v; }
                cb.sc[nxt][tid] = ord2f(cb.selkey[sb][tid]);
                cb.depth[nxt][tid] = dp; cb.pk[nxt][tid] = npk;
            }
            if (tid == 0) { cb.theta = 0u; cb.ns = 0; }
            for (int e = tid; e < BMAX * BMAX; e += MT) {
                const int r = e / BMAX, q = e % BMAX;
                if (r >= m || q >= m) continue;
                const int ar = seli[r], aq = seli[q];
                const int vr = selv[r], vq = selv[q];
                const int pkr = pk[ar], pkq = pk[aq];
                const int er = cand_ext_id(vr, blank, pkr), eq2 = cand_ext_id(vq, blank, pkq);
                const int R = rel_child(rel[ar][aq], er, eq2, depth[ar], depth[aq], cb.node[cur][ar], cb.node[cur][aq], vch, parent, meta);
                cb.rel[nxt][r][q] = (unsigned char)R;
                if (r == q) continue;
                if (R == REL_EQ) cb.tw[r] = q;                               // same prefix, other ends-in-blank flag
                else if (R >= REL_PFX && R < REL_RPFX) {
                    const int dr = depth[ar] + (er >= 0 ? 1 : 0), dq = depth[aq] + (eq2 >= 0 ? 1 : 0);
                    if (dq == dr + 1) {                                      // X_q = X_r + last(q): r is q's parent prefix
                        const int eb_r = (vr == blank) ? 1 : 0;
                        const int eb_q = (vq == blank) ? 1 : 0;
                        const int last_q = eq2 >= 0 ? eq2 : (pkq & 0xff);
                        if (eb_r) cb.p1[q] = r; else cb.p0[q] = r;
                        if (eb_q) atomicOr(&cb.abs1[r], 1u << last_q); else atomicOr(&cb.abs0[r], 1u << last_q);
                    }
                }
            }
        }
        cta2_bar_main<MW>();
        kept = m;
        cur = nxt;
    }
    if (tid == 0) { p.out_stats[2 * utt] = stat_fallback; p.out_stats[2 * utt + 1] = stat_surv; }
    cta2_bar_all<MW>();                                      // final beam complete: the aux warp writes the result
}

int ctc_launch_cta2(const CtcParams &p, int domain, int utterances, int mw, size_t pad, cudaStream_t st) {
#define GASR_CTA2(DOM, BM)                                                                                  \
    do {                                                                                                    \
        if (mw == 4) ctc_beam_cta2_kernel<DOM, BM, 4><<<utterances, 192, pad, st>>>(p);                     \
        else if (mw == 16) ctc_beam_cta2_kernel<DOM, BM, 16><<<utterances, 576, pad, st>>>(p);              \
        else ctc_beam_cta2_kernel<DOM, BM, 8><<<utterances, 320, pad, st>>>(p);                             \
    } while (0)
    if (domain == GASR_DOMAIN_LOG) {
        if (p.beam <= 16) GASR_CTA2(1, 16); else GASR_CTA2(1, 32);
    } else {
        if (p.beam <= 16) GASR_CTA2(0, 16); else GASR_CTA2(0, 32);
    }
#undef GASR_CTA2
    return GASR_OK;
}

}  // namespace gasr
