// ctc_beam_warp.cu -- CTC prefix beam search, one WARP per utterance (beam <= 32, vocabulary <= 32): the throughput decoder of
// the wave engine.  Stands behind CTCBeamSearch::decode (reference CTCBeamSearch.cu:262-312); the contract and the shared helpers
// are described in ctc_beam_general.cu / ctc_beam.cuh.
#include "ctc_beam.cuh"

namespace gasr {

template <int DOMAIN, int BMAX>
__global__ void __launch_bounds__(256) ctc_beam_warp_kernel(const CtcParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int utt = blockIdx.x * W + warp;
    if (utt >= p.N) return;
    WarpBeam<BMAX> &wb = reinterpret_cast<WarpBeam<BMAX> *>(smem_raw)[warp];
    __shared__ char vch_s[32];
    const int V = p.V, B = p.beam, blank = p.blank, Vp = p.Vp;
    constexpr unsigned FULL = 0xffffffffu;
    const bool active = lane < V;

    int *parent = p.parent + (size_t)utt * p.cap;
    int *meta = p.meta + (size_t)utt * p.cap;
    int *born = p.born ? p.born + (size_t)utt * p.cap : nullptr;
    int *child = p.child + (size_t)utt * p.cap * Vp;
    const float *S = p.scores + (size_t)utt * p.ld;
    const size_t frame_stride = (size_t)p.frame_rows * p.ld;

    // every warp writes the same bytes; only __syncwarp ordering is needed for its own reads
    if (active) vch_s[lane] = p.vocab[lane];
    const char *vch = vch_s;
    int kept = 1, nodes = 1, cur = 0;
    int stat_surv = 0, stat_fallback = 0;
    int4 *gstate = reinterpret_cast<int4 *>(p.state + (size_t)utt * p.state_stride);
    constexpr int kStateVec = (int)(sizeof(WarpBeam<BMAX>) / sizeof(int4));
    const int Tu = utt_frames(p, utt);
    if (p.t0 > 0 && p.t0 >= Tu) return;                  // this utterance ended in an earlier chunk (its result is written)
    const int t_end = p.t1 < Tu ? p.t1 : Tu;
    if (p.t0 == 0) {
        if (lane < Vp) child[lane] = 0;
        if (lane == 0) {
            parent[0] = -1; meta[0] = 0xff;
            wb.sc[0][0] = DOMAIN ? 0.0f : 1.0f;
            wb.node[0][0] = 0; wb.depth[0][0] = 0; wb.pk[0][0] = 0xff | (1 << 8);
            wb.rel[0][0][0] = REL_EQ;
        }
    } else {
        // resume: the previous chunk launch left the beam in HBM
        int4 *dst = reinterpret_cast<int4 *>(&wb);
        for (int i = lane; i < kStateVec; i += 32) dst[i] = gstate[i];
        const int4 hdr = gstate[kStateVec];
        kept = hdr.x; nodes = hdr.y; cur = hdr.z;
    }
    float lp_next = active ? S[(size_t)p.t0 * frame_stride + lane] : 0.0f;
    __syncwarp();

    for (int t = p.t0; t < t_end; t++) {
        const float lp = lp_next;
        if (t + 1 < t_end && active) lp_next = S[(size_t)(t + 1) * frame_stride + lane];
        const bool last_frame = (t == Tu - 1) && (t > 0);
        const int k = kept;
        const float *sc = wb.sc[cur];
        const int *node = wb.node[cur], *pk = wb.pk[cur], *depth = wb.depth[cur];
        const unsigned char (*rel)[BMAX] = wb.rel[cur];
        const float lpb = __shfl_sync(FULL, lp, blank);

        // ---- rank of this frame's scores over the vocabulary (independent of the beam) ------------------------
        {
            const unsigned mine = active ? f2ord(lp) : 0u;
            int lr = 0;
            // (partially unrolled on purpose, here and in the two probe-cell loops below: fully unrolled the kernel was 7008 SASS
            // instructions and stalled on instruction fetch with 4096 warps in different phases of the frame; 4424 now (with the rolled tie loop),
            // decoder alone -9 %, cfg5 step -2.5 %, same-box A/B)
#pragma unroll 4
            for (int u = 0; u < 32; u++) {
                const unsigned x = __shfl_sync(FULL, mine, u);
                lr += (x > mine || (x == mine && u < lane)) ? 1 : 0;
            }
            wb.order[lr] = lane;
        }
        // ---- relations among kept states, read off the prefix-relation matrix (lane r owns state r) ---------
        int my_last = 0xff, my_eb = 1, my_tw = kNone, my_p0 = kNone, my_p1 = kNone;
        if (lane < k) {
            const int dr = depth[lane];
            my_last = pk[lane] & 0xff; my_eb = (pk[lane] >> 8) & 1;
            unsigned a0 = 0, a1 = 0;
            for (int j = 0; j < k; j++) {
                const int R = rel[lane][j], pkj = pk[j], dj = depth[j];
                if (R == REL_EQ && j != lane) my_tw = j;
                if (R >= REL_RPFX && dr == dj + 1) { if ((pkj >> 8) & 1) my_p1 = j; else my_p0 = j; }   // X_j = parent(X_r)
                if (R >= REL_PFX && R < REL_RPFX && dj == dr + 1) {                                        // X_j = X_r . y
                    const unsigned bit = 1u << (R - REL_PFX);
                    if ((pkj >> 8) & 1) a1 |= bit; else a0 |= bit;
                }
            }
            wb.tw[lane] = my_tw; wb.p0[lane] = my_p0; wb.p1[lane] = my_p1; wb.abs0[lane] = a0; wb.abs1[lane] = a1;
        }
        // twin pairs (i < twin): their V merged scores are computed once, by the pair loop below
        const unsigned pair_mask = __ballot_sync(FULL, lane < k && my_tw > lane);
        const int npairs = __popc(pair_mask);
        if (lane < k) {
            int pidx = -1;
            if (my_tw > lane) {
                pidx = __popc(pair_mask & ((1u << lane) - 1u));
                wb.pair_i[pidx] = lane; wb.pair_tw[pidx] = my_tw;
            }
            wb.pinfo[lane] = make_int4(__float_as_int(sc[lane]), pidx, pk[lane] | ((my_tw + 1) << 9), (int)wb.abs0[lane]);
        }
        // ---- "stay" candidates, one per (X,0) state, all lanes in parallel ------------------------------
        {
            const bool do_stay = lane < k && my_eb == 0;
            const float lpv = __shfl_sync(FULL, lp, do_stay ? my_last : 0);
            if (do_stay) {
                int m0 = my_p0, m1 = my_p1, m2 = lane, tmp;
                if (m0 >= 0 && (pk[m0] & 0xff) == my_last) m0 = kNone;    // (P,0)+v with last(P)==v stays on P
                if (m0 > m1) { tmp = m0; m0 = m1; m1 = tmp; }
                if (m1 > m2) { tmp = m1; m1 = m2; m2 = tmp; }
                if (m0 > m1) { tmp = m0; m0 = m1; m1 = tmp; }
                // chain in ascending state index; absent members (-1) sorted to the front, m2 is always present
                float acc = comb<DOMAIN>(sc[m0 >= 0 ? m0 : (m1 >= 0 ? m1 : m2)], lpv);
                if (m0 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[m1], lpv));
                if (m1 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[m2], lpv));
                if (last_frame) {
                    int b0 = lane, b1 = my_tw;
                    if (b1 >= 0 && b1 < b0) { b0 = my_tw; b1 = lane; }
                    acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[b0], lpb));
                    if (b1 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[b1], lpb));
                }
                wb.stay[lane] = acc;
            }
        }
        __syncwarp();

        // ---- merged candidate scores: val[i] of lane v  <->  candidate i*V + v -----------------------------
        if (!last_frame) {
            // twin merges first: branch-free and unrolled so that independent pairs interleave
#pragma unroll 2
            for (int q = 0; q < npairs; q++) {
                const float sa = comb<DOMAIN>(sc[wb.pair_i[q]], lp), sb = comb<DOMAIN>(sc[wb.pair_tw[q]], lp);
                wb.pairmm[q][lane] = mrg_bf<DOMAIN>(sa, sb);
            }
            __syncwarp();
#pragma unroll 4
            for (int i = 0; i < k; i++) {
                const int4 pi = wb.pinfo[i];
                const int pki = pi.z & 0x1ff, twi = ((pi.z >> 9) & 0x3f) - 1;
                const int ebi = (pki >> 8) & 1, lasti = pki & 0xff;
                const float s = comb<DOMAIN>(__int_as_float(pi.x), lp);
                const bool is_stay = (ebi == 0 && lane == lasti);
                const bool is_blank = (lane == blank);
                const bool member = twi >= 0 && (is_blank || ebi == 0 || lane != lasti);
                const bool dead = (member && twi < i) || (!is_blank && (((unsigned)pi.w >> lane) & 1u));
                float acc = s;
                if (pi.y >= 0) { const float mm = wb.pairmm[pi.y][lane]; acc = member ? mm : s; }
                const float sv = wb.stay[i];
                acc = is_stay ? sv : acc;
                wb.cand[i][lane] = (active && (is_stay || !dead)) ? f2ord(acc) : 0u;
            }
        } else {
            for (int i = 0; i < k; i++) {
                const int4 pi = wb.pinfo[i];
                const float sci = __int_as_float(pi.x);
                const int pki = pi.z & 0x1ff, twi = ((pi.z >> 9) & 0x3f) - 1;
                const int ebi = (pki >> 8) & 1, lasti = pki & 0xff;
                const float s = comb<DOMAIN>(sci, lp);
                float acc = s;
                bool dead = false;
                const bool is_stay = (ebi == 0 && lane == lasti);
                const bool is_blank = (lane == blank);
                const bool member = twi >= 0 && (is_blank || ebi == 0 || lane != lasti);
                if (!last_frame || !is_blank) {
                    if (twi >= 0) {
                        if (twi < i) dead = member;
                        else {
                            const float mm = mrg<DOMAIN>(s, comb<DOMAIN>(sc[twi], lp));
                            acc = member ? mm : s;
                        }
                    }
                    if (!is_blank && (((unsigned)pi.w >> lane) & 1u)) dead = true;   // kept (X.v, 0) hosts this extend
                    if (last_frame && !dead && !is_stay && ((wb.abs1[i] >> lane) & 1u)) {
                        // kept (X.v, 1): on the last frame its blank candidate strips to X.v as well
                        for (int j = 0; j < k; j++) {
                            const int R = rel[i][j];
                            if (R == REL_PFX + lane && depth[j] == depth[i] + 1 && ((pk[j] >> 8) & 1))
                                acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[j], lpb));
                        }
                    }
                } else {
                    // blank candidate on the last frame: it strips to X, so the (X,0) stay slot or an extend slot
                    // that spells X hosts it; otherwise it stands alone
                    if (ebi == 0 || twi >= 0) dead = true;
                    else if (lasti != 0xff) {
                        const int q0 = wb.p0[i], q1 = wb.p1[i];
                        if (q1 >= 0 || (q0 >= 0 && (pk[q0] & 0xff) != lasti)) dead = true;
                    }
                }
                if (is_stay) { acc = wb.stay[i]; dead = false; }
                const unsigned key = (active && !dead) ? f2ord(acc) : 0u;
                wb.cand[i][lane] = key;
            }
        }
        __syncwarp();

        // ---- prune (reference: stable descending prob sort on top of the ascending string sort, keep beam) --------
        // (1) lower bound: parents are in score order and order[] ranks this frame's scores, so unmerged candidates
        //     form a matrix sorted along both axes whose top-beam lies in the "staircase" (i+1)(j+1) <= beam.  The
        //     beam-th largest key among probe cells of that staircase is a valid lower bound of the beam-th largest
        //     merged key overall (merging only raises keys), and usually a tight one.
        // (2) survivors = candidates >= bound, compacted in candidate-index order (a few more than beam);
        // (3) exact rank of each survivor by all-pairs counting with the full order (score desc, raw string asc via
        //     rel[][]; at t = 0 ties keep vocabulary order, CTCBeamSearch.cu:390): rank r < beam -> kept state r.
        // If more than 64 candidates survive, fall back to beam rounds of warp-max extraction (same order).
        int m = 0;
        unsigned theta = 0u;
        {
            unsigned ck[2];
            const int cpl = p.n_cells > 32 ? 2 : 1;          // probe cells per lane
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int ci = p.cell_i[lane + 32 * q];
                ck[q] = (q < cpl && ci < k) ? wb.cand[ci][wb.order[p.cell_j[lane + 32 * q]]] : 0u;
            }
            int cnt0 = 0, cnt1 = 0;
            if (cpl > 1) {
#pragma unroll 4
                for (int u = 0; u < 32; u++) {
                    const unsigned x0 = __shfl_sync(FULL, ck[0], u);
                    cnt0 += (x0 > ck[0] || (x0 == ck[0] && u < lane)) ? 1 : 0;
                    const unsigned x1 = __shfl_sync(FULL, ck[1], u);
                    cnt0 += (x1 > ck[0]) ? 1 : 0;
                    cnt1 += (x0 >= ck[1]) ? 1 : 0;
                    cnt1 += (x1 > ck[1] || (x1 == ck[1] && u < lane)) ? 1 : 0;
                }
            } else {
#pragma unroll 4
                for (int u = 0; u < 32; u++) {
                    const unsigned x0 = __shfl_sync(FULL, ck[0], u);
                    cnt0 += (x0 > ck[0] || (x0 == ck[0] && u < lane)) ? 1 : 0;
                }
            }
            unsigned th = (cnt0 == B - 1) ? ck[0] : 0u;
            if (cpl > 1 && cnt1 == B - 1) th = ck[1];
            theta = __reduce_max_sync(FULL, th);
            // second bound: the best parent's beam best-ranked candidates, if none of them was absorbed
            if (B <= 32 && B <= V) {
                const unsigned r0 = lane < B ? wb.cand[0][wb.order[lane]] : 0xffffffffu;
                const unsigned mn = __reduce_min_sync(FULL, r0);
                theta = max(theta, mn);
            }
        }
        int ns = 0;
#pragma unroll 4
        for (int i = 0; i < k; i++) {
            const unsigned key = wb.cand[i][lane];
            const bool sv = key != 0u && key >= theta;
            const unsigned mask = __ballot_sync(FULL, sv);
            const int pos = ns + __popc(mask & ((1u << lane) - 1u));
            if (sv && pos < 64) { wb.surv_key[pos] = key; wb.surv_iv[pos] = (i << 8) | lane; }
            ns += __popc(mask);
        }
        if (lane < 4 && ns <= 64) wb.surv_key[ns + lane] = 0u;   // pad for the four-at-a-time ranking below
        __syncwarp();
        stat_surv += ns;
        stat_fallback += ns > 64;
        if (ns <= 64) {
            m = ns < B ? ns : B;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int sidx = lane + 32 * q;
                if (sidx < ns) {
                    const unsigned key = wb.surv_key[sidx];
                    const int iv = wb.surv_iv[sidx];
                    const int mi = iv >> 8, mv = iv & 0xff;
                    const int ms = cand_suffix_id(mv, blank, pk[mi]);
                    // four keys per iteration (the list is zero-padded to a multiple of four: real keys are > 0); an exact tie
                    // (rare) takes the raw-string order from the relation matrix
                    int rank = 0;
                    for (int o = 0; o < ns; o += 4) {
                        const uint4 k4 = *reinterpret_cast<const uint4 *>(&wb.surv_key[o]);
                        rank += (k4.x > key) + (k4.y > key) + (k4.z > key) + (k4.w > key);
                        if (k4.x == key || k4.y == key || k4.z == key || k4.w == key) {
#pragma unroll 1
                            for (int j = 0; j < 4; j++) {
                                if (wb.surv_key[o + j] != key || o + j == sidx || o + j >= ns) continue;
                                if (t == 0) rank += o + j < sidx;
                                else {
                                    const int oiv = wb.surv_iv[o + j];
                                    const int oi = oiv >> 8, ov = oiv & 0xff;
                                    rank += cand_less_rel(rel[oi][mi], cand_suffix_id(ov, blank, pk[oi]), ms, vch) ? 1 : 0;
                                }
                            }
                        }
                    }
                    if (rank < B) { wb.selkey[rank] = key; wb.seli[rank] = mi; wb.selv[rank] = mv; }
                }
            }
        } else {
            unsigned lmax = 0u;
            for (int i = 0; i < k; i++) lmax = max(lmax, wb.cand[i][lane]);
            for (m = 0; m < B; m++) {
                const unsigned gmax = __reduce_max_sync(FULL, lmax);
                if (gmax == 0u) break;
                const unsigned any = __ballot_sync(FULL, lmax == gmax);
                int wl = __ffs(any) - 1;
                unsigned x = lane < k ? wb.cand[lane][wl] : 0u;          // column wl: row `lane`
                const unsigned colmask = __ballot_sync(FULL, x == gmax);
                int wi = __ffs(colmask) - 1;
                if (t > 0 && (__popc(any) > 1 || __popc(colmask) > 1)) {
                    // exact tie: smallest raw string wins
                    int bi = -1, bs = 0;
                    if (lmax == gmax) {
                        for (int i = 0; i < k; i++) {
                            if (wb.cand[i][lane] != gmax) continue;
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
                    x = lane < k ? wb.cand[lane][wl] : 0u;
                }
                if (lane == wi) { wb.cand[wi][wl] = 0u; x = 0u; }
                const unsigned cmax = __reduce_max_sync(FULL, x);         // new maximum of the winning column
                if (lane == wl) lmax = cmax;
                if (lane == 0) { wb.selkey[m] = gmax; wb.seli[m] = wi; wb.selv[m] = wl; }
                __syncwarp();
            }
        }
        __syncwarp();

        // ---- the selected candidates become the next kept states (lane r builds state r) ---------------------
        const int nxt = cur ^ 1;
        {
            bool need_new = false;
            int i = 0, v = 0, nd = 0, pn = 0, dp = 0, npk = 0;
            if (lane < m) {
                i = wb.seli[lane]; v = wb.selv[lane];
                const int pki = pk[i];
                const int ebi = (pki >> 8) & 1, lasti = pki & 0xff;
                if (v == blank) { nd = node[i]; dp = depth[i]; npk = lasti | (1 << 8); }
                else if (ebi == 0 && v == lasti) { nd = node[i]; dp = depth[i]; npk = lasti; }
                else {
                    pn = node[i]; dp = depth[i] + 1; npk = v;
                    nd = child[(size_t)pn * Vp + v];
                    need_new = nd == 0;
                }
            }
            const unsigned nb = __ballot_sync(FULL, need_new);
            if (need_new) {
                nd = nodes + __popc(nb & ((1u << lane) - 1u));
                parent[nd] = pn;
                meta[nd] = (dp << 8) | v;
                if (born) born[nd] = t;
                child[(size_t)pn * Vp + v] = nd;
                int4 *row = reinterpret_cast<int4 *>(child + (size_t)nd * Vp);
                for (int q = 0; q < Vp / 4; q++) row[q] = make_int4(0, 0, 0, 0);
            }
            nodes += __popc(nb);
            if (lane < m) {
                wb.sc[nxt][lane] = ord2f(wb.selkey[lane]);
                wb.node[nxt][lane] = nd; wb.depth[nxt][lane] = dp; wb.pk[nxt][lane] = npk;
            }
        }
        // ---- prefix relations of the new kept states from the old ones ----------------------------------------
        // The relation is antisymmetric (rel[q][r] = mirror of rel[r][q]): every unordered pair is evaluated once -- pair
        // (r, (r + d) mod BMAX) for d = 1 .. BMAX/2 (d = BMAX/2 only from the lower half) -- and written to both cells.
        if (lane < m) wb.rel[nxt][lane][lane] = REL_EQ;
        for (int e = lane; e < BMAX * (BMAX / 2); e += 32) {
            const int r = e / (BMAX / 2), d = e % (BMAX / 2) + 1;
            const int q = (r + d) & (BMAX - 1);
            if (r >= m || q >= m || (d == BMAX / 2 && r >= BMAX / 2)) continue;
            const int ar = wb.seli[r], aq = wb.seli[q];
            const int er = cand_ext_id(wb.selv[r], blank, pk[ar]);
            const int eq2 = cand_ext_id(wb.selv[q], blank, pk[aq]);
            const int R = rel[ar][aq];
            const int dA = depth[ar], dB = depth[aq];
            int out;
            if (R == REL_EQ) {
                if (er < 0 && eq2 < 0) out = REL_EQ;
                else if (er < 0) out = REL_PFX + eq2;
                else if (eq2 < 0) out = REL_RPFX + er;
                else if (er == eq2) out = REL_EQ;
                else out = ch_less(vch, er, eq2) ? REL_LT : REL_GT;
            } else if (R == REL_LT || R == REL_GT) {
                out = R;
            } else if (R < REL_RPFX) {                         // A proper prefix of B, B = A.y...
                const int y = R - REL_PFX;
                if (er < 0) out = R;
                else if (er != y) out = ch_less(vch, er, y) ? REL_LT : REL_GT;
                else if (dB == dA + 1) out = eq2 < 0 ? REL_EQ : REL_PFX + eq2;
                else out = REL_PFX + trie_char_at(parent, meta, node[aq], dA + 1);
            } else {                                           // B proper prefix of A, A = B.y...
                const int y = R - REL_RPFX;
                if (eq2 < 0) out = R;
                else if (eq2 != y) out = ch_less(vch, y, eq2) ? REL_LT : REL_GT;
                else if (dA == dB + 1) out = er < 0 ? REL_EQ : REL_RPFX + er;
                else out = REL_RPFX + trie_char_at(parent, meta, node[ar], dB + 1);
            }
            wb.rel[nxt][r][q] = (unsigned char)out;
            wb.rel[nxt][q][r] = (unsigned char)(out < REL_PFX ? (out == REL_EQ ? REL_EQ : (REL_LT + REL_GT) - out) : (out < REL_RPFX ? out + 32 : out - 32));
        }
        kept = m;
        cur = nxt;
        __syncwarp();
    }

    if (p.t1 < Tu) {
        // more chunks follow: park the beam in HBM
        const int4 *src = reinterpret_cast<const int4 *>(&wb);
        for (int i = lane; i < kStateVec; i += 32) gstate[i] = src[i];
        if (lane == 0) {
            gstate[kStateVec] = make_int4(kept, nodes, cur, 0);
            if (p.t0 == 0) { p.out_stats[2 * utt] = stat_fallback; p.out_stats[2 * utt + 1] = stat_surv; }
            else { p.out_stats[2 * utt] += stat_fallback; p.out_stats[2 * utt + 1] += stat_surv; }
        }
        return;
    }
    if (lane == 0) {
        if (p.t0 == 0) { p.out_stats[2 * utt] = stat_fallback; p.out_stats[2 * utt + 1] = stat_surv; }
        else { p.out_stats[2 * utt] += stat_fallback; p.out_stats[2 * utt + 1] += stat_surv; }
    }
    // ---- result (CTCBeamSearch.cu:290-298): kept states best first, path = labels of X -------------------------
    if (lane == 0 && p.out_counts) p.out_counts[utt] = kept;
    for (int r = lane; r < p.nbest; r += 32) {
        char *out = p.out_paths + ((size_t)utt * p.nbest + r) * p.max_len;
        int *ts = p.out_ts ? p.out_ts + ((size_t)utt * p.nbest + r) * p.max_len : nullptr;
        int len = 0;
        float scv = 0.0f;
        if (r < kept) {
            int nd = wb.node[cur][r];
            const int dpt = wb.depth[cur][r];
            len = dpt;
            if (Tu == 1 && ((wb.pk[cur][r] >> 8) & 1)) { if (len < p.max_len) { out[len] = vch[blank]; if (ts) ts[len] = 0; } len += 1; }
            for (int pos = dpt - 1; pos >= 0; pos--) {
                if (pos < p.max_len) { out[pos] = vch[meta[nd] & 0xff]; if (ts) ts[pos] = born[nd]; }
                nd = parent[nd];
            }
            scv = wb.sc[cur][r];
        }
        p.out_lens[(size_t)utt * p.nbest + r] = len;
        p.out_scores[(size_t)utt * p.nbest + r] = scv;
    }
}

int ctc_launch_warp(const CtcParams &p, int domain, int blocks, int W, cudaStream_t st) {
#define GASR_CTCW_LAUNCH(DOM, BM)                                                                                \
    do {                                                                                                          \
        GASR_CUDA(cudaFuncSetAttribute(ctc_beam_warp_kernel<DOM, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)(sizeof(WarpBeam<BM>) * 8)));                                         \
        ctc_beam_warp_kernel<DOM, BM><<<blocks, W * 32, sizeof(WarpBeam<BM>) * W, st>>>(p);                       \
    } while (0)
    if (domain == GASR_DOMAIN_LOG) {
        if (p.beam <= 16) GASR_CTCW_LAUNCH(1, 16); else GASR_CTCW_LAUNCH(1, 32);
    } else {
        if (p.beam <= 16) GASR_CTCW_LAUNCH(0, 16); else GASR_CTCW_LAUNCH(0, 32);
    }
#undef GASR_CTCW_LAUNCH
    return GASR_OK;
}

}  // namespace gasr
