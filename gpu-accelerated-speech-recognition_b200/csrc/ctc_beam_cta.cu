// ctc_beam_cta.cu -- CTC prefix beam search, one 128-thread CTA per utterance (beam <= 32, vocabulary <= 32), resumable per time
// chunk: the decoder of the time-chunked pipeline for small batches.  Stands behind CTCBeamSearch::decode (reference
// CTCBeamSearch.cu:262-312); phases of a frame and the beam record: ctc_beam.cuh.
#include "ctc_beam.cuh"

namespace gasr {

template <int DOMAIN, int BMAX>
__global__ void __launch_bounds__(128) ctc_beam_cta_kernel(const CtcParams p) {
    __shared__ __align__(16) CtaBeam<BMAX> cb;
    __shared__ char vch_s[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int utt = blockIdx.x;
    const int V = p.V, B = p.beam, blank = p.blank, Vp = p.Vp;
    constexpr unsigned FULL = 0xffffffffu;
    const bool active = lane < V;
    const char *vch = vch_s;

    int *parent = p.parent + (size_t)utt * p.cap;
    int *meta = p.meta + (size_t)utt * p.cap;
    int *born = p.born ? p.born + (size_t)utt * p.cap : nullptr;
    int *child = p.child + (size_t)utt * p.cap * Vp;
    const float *S = p.scores + (size_t)utt * p.ld;
    const size_t frame_stride = (size_t)p.frame_rows * p.ld;
    int4 *gstate = reinterpret_cast<int4 *>(p.state + (size_t)utt * p.state_stride);
    constexpr int kStateVec = (int)(sizeof(CtaBeam<BMAX>) / sizeof(int4));

    const int Tu = utt_frames(p, utt);
    if (p.t0 > 0 && p.t0 >= Tu) return;                  // this utterance ended in an earlier chunk (its result is written)
    const int t_end = p.t1 < Tu ? p.t1 : Tu;
    if (tid < V) vch_s[tid] = p.vocab[tid];
    int cur = 0;
    int stat_surv = 0, stat_fallback = 0;
    if (p.t0 == 0) {
        if (tid < Vp) child[tid] = 0;
        if (tid == 0) {
            parent[0] = -1; meta[0] = 0xff;
            cb.sc[0][0] = DOMAIN ? 0.0f : 1.0f;
            cb.node[0][0] = 0; cb.pnode[0][0] = kNone; cb.depth[0][0] = 0; cb.pk[0][0] = 0xff | (1 << 8);
            cb.rel[0][0][0] = REL_EQ;
            cb.kept = 1; cb.nodes = 1;
        }
    } else {
        int4 *dst = reinterpret_cast<int4 *>(&cb);
        for (int i = tid; i < kStateVec; i += 128) dst[i] = gstate[i];
        cur = gstate[kStateVec].x;
    }
    // streaming: the log-probabilities are produced while this kernel runs; every warp tracks how many frames are
    // known complete (ready_frames) and samples the next block's counter one block early (flag_next)
    const volatile unsigned *lpr = p.lp_ready;
    const bool streaming = lpr != nullptr;
    int ready_frames = streaming ? 0 : p.T;
    unsigned flag_next = 0;
    auto frames_ready = [&](int t) {                     // returns once frame t may be read
        while (t >= ready_frames) {
            const int blk = ready_frames / p.lp_fpb;
            unsigned v = __shfl_sync(FULL, flag_next, 0);
            if (v < (unsigned)p.lp_need) {
                unsigned long long t_start = 0;
                do {
                    if (lane == 0) v = lpr[blk];
                    v = __shfl_sync(FULL, v, 0);
                    if (v < (unsigned)p.lp_need) {
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (t_start == 0) t_start = now;
                        if ((p.abort && *p.abort) || now - t_start > 2000000000ull) {
                            if (p.abort) *p.abort = 1u; if (p.error) { *reinterpret_cast<volatile int *>(p.error) = 4; __threadfence_system(); } break; }   // watchdog: give up waiting
                        __nanosleep(200);
                    }
                } while (v < (unsigned)p.lp_need);
            }
            ready_frames = (blk + 1) * p.lp_fpb;
            __threadfence();                             // acquire side of the counter: the rows are read after this fence
            flag_next = 0;
            if (ready_frames < p.T && lane == 0) flag_next = lpr[blk + 1];
        }
    };
    if (streaming) frames_ready(p.t0);
    float lp_next = active ? __ldcg(S + (size_t)p.t0 * frame_stride + lane) : 0.0f;
    __syncthreads();
    int kept = cb.kept;

    for (int t = p.t0; t < t_end; t++) {
        const float lp = lp_next;
        if (t + 1 < t_end) {
            if (streaming) frames_ready(t + 1);
            if (active) lp_next = __ldcg(S + (size_t)(t + 1) * frame_stride + lane);
        }
        const bool last_frame = (t == Tu - 1) && (t > 0);
        const int k = kept;
        const float *sc = cb.sc[cur];
        const int *node = cb.node[cur], *pnode = cb.pnode[cur], *pk = cb.pk[cur], *depth = cb.depth[cur];
        const unsigned char (*rel)[BMAX] = cb.rel[cur];
        const float lpb = __shfl_sync(FULL, lp, blank);

        // ================= phase A =================
        if (w == 0) {
            // relations from node ids; with beam <= 16 two lanes share a state and split the scan
            constexpr int HALVES = BMAX <= 16 ? 2 : 1;
            const int r = HALVES == 2 ? (lane & 15) : lane, half = HALVES == 2 ? (lane >> 4) : 0;
            int my_tw = kNone, my_p0 = kNone, my_p1 = kNone;
            unsigned a0 = 0, a1 = 0;
            int my_last = 0xff, my_eb = 1;
            if (r < k) {
                const int nd = node[r], pn = pnode[r];
                my_last = pk[r] & 0xff; my_eb = (pk[r] >> 8) & 1;
                const int jb = HALVES == 2 ? half * 8 : 0, je = HALVES == 2 ? min(k, jb + 8) : k;
#pragma unroll 4
                for (int j = jb; j < je; j++) {
                    const int nj = node[j], pnj = pnode[j], pkj = pk[j];
                    if (nj == nd && j != r) my_tw = j;
                    if (nj == pn) { if ((pkj >> 8) & 1) my_p1 = j; else my_p0 = j; }
                    if (pnj == nd) { const unsigned bit = 1u << (pkj & 0xff); if ((pkj >> 8) & 1) a1 |= bit; else a0 |= bit; }
                }
            }
            if (HALVES == 2) {
                my_tw = max(my_tw, __shfl_xor_sync(FULL, my_tw, 16));
                my_p0 = max(my_p0, __shfl_xor_sync(FULL, my_p0, 16));
                my_p1 = max(my_p1, __shfl_xor_sync(FULL, my_p1, 16));
                a0 |= __shfl_xor_sync(FULL, a0, 16);
                a1 |= __shfl_xor_sync(FULL, a1, 16);
            }
            const bool owner = r < k && half == 0;
            if (owner) {
                cb.tw[r] = my_tw; cb.p0[r] = my_p0; cb.p1[r] = my_p1; cb.abs0[r] = a0; cb.abs1[r] = a1;
                cb.pinfo[r] = make_int4(__float_as_int(sc[r]), __float_as_int(my_tw >= 0 ? sc[my_tw] : 0.0f),
                                        pk[r] | ((my_tw + 1) << 9), (int)a0);
            }
            // "stay" candidates, one per (X,0) state
            const bool do_stay = owner && my_eb == 0;
            const float lpv = __shfl_sync(FULL, lp, do_stay ? my_last : 0);
            if (do_stay) {
                int m0 = my_p0, m1 = my_p1, m2 = r, tmp;
                if (m0 >= 0 && (pk[m0] & 0xff) == my_last) m0 = kNone;
                if (m0 > m1) { tmp = m0; m0 = m1; m1 = tmp; }
                if (m1 > m2) { tmp = m1; m1 = m2; m2 = tmp; }
                if (m0 > m1) { tmp = m0; m0 = m1; m1 = tmp; }
                float acc = comb<DOMAIN>(sc[m0 >= 0 ? m0 : (m1 >= 0 ? m1 : m2)], lpv);
                if (m0 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[m1], lpv));
                if (m1 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[m2], lpv));
                if (last_frame) {
                    int b0 = r, b1 = my_tw;
                    if (b1 >= 0 && b1 < b0) { b0 = my_tw; b1 = r; }
                    acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[b0], lpb));
                    if (b1 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[b1], lpb));
                }
                cb.stay[r] = acc;
            }
            if (lane == 0) { cb.ns = 0; cb.theta = 0u; }
        } else if (w == 1) {
            if (t == p.t0) {   // later frames: ranked at the end of the previous frame, in the shadow of phase F
                const unsigned mine = active ? f2ord(lp) : 0u;
                int lr = 0;
#pragma unroll
                for (int u = 0; u < 32; u++) {
                    const unsigned x = __shfl_sync(FULL, mine, u);
                    lr += (x > mine || (x == mine && u < lane)) ? 1 : 0;
                }
                cb.order[lr] = lane;
            }
            cb.rankc[lane] = 0; cb.rankc[lane + 32] = 0;
        }
        __syncthreads();

        // ================= phase B: merged candidates =================
        if (!last_frame) {
            for (int i = w; i < k; i += 4) {
                const int4 pi = cb.pinfo[i];
                const int pki = pi.z & 0x1ff, twi = ((pi.z >> 9) & 0x3f) - 1;
                const int ebi = (pki >> 8) & 1, lasti = pki & 0xff;
                const float s = comb<DOMAIN>(__int_as_float(pi.x), lp);
                const bool is_stay = (ebi == 0 && lane == lasti);
                const bool is_blank = (lane == blank);
                const bool member = twi >= 0 && (is_blank || ebi == 0 || lane != lasti);
                const bool dead = (member && twi < i) || (!is_blank && (((unsigned)pi.w >> lane) & 1u));
                float acc = s;
                if (twi > i) {   // uniform per warp: this parent hosts the twin pair
                    const float mm = mrg_bf<DOMAIN>(s, comb<DOMAIN>(__int_as_float(pi.y), lp));
                    acc = member ? mm : s;
                }
                const float sv = cb.stay[i];
                acc = is_stay ? sv : acc;
                cb.cand[i][lane] = (active && (is_stay || !dead)) ? f2ord(acc) : 0u;
            }
        } else if (w == 0) {
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
                        const int nd = node[i];
                        for (int j = 0; j < k; j++)
                            if (pnode[j] == nd && (pk[j] & 0xff) == lane && ((pk[j] >> 8) & 1))
                                acc = mrg<DOMAIN>(acc, comb<DOMAIN>(sc[j], lpb));
                    }
                } else {
                    if (ebi == 0 || twi >= 0) dead = true;
                    else if (lasti != 0xff) {
                        const int q0 = cb.p0[i], q1 = cb.p1[i];
                        if (q1 >= 0 || (q0 >= 0 && (pk[q0] & 0xff) != lasti)) dead = true;
                    }
                }
                if (is_stay) { acc = cb.stay[i]; dead = false; }
                cb.cand[i][lane] = (active && !dead) ? f2ord(acc) : 0u;
            }
        }
        __syncthreads();

        // ================= phase C: lower bound of the beam-th largest merged key =================
        // probe cells -> shared memory; then every cell counts how many probe keys precede it (two threads per cell
        // when 64 cells are probed), and the cell of rank beam-1 is the bound
        {
            constexpr int NC = BMAX <= 16 ? 64 : 128;
            constexpr int TPC = 128 / NC;                     // threads per cell
            const int c = tid / TPC;
            const int ci = p.cell_i[c];
            const unsigned mine = ci < k ? cb.cand[ci][cb.order[p.cell_j[c]]] : 0u;
            if (TPC == 1 || (tid & 1) == 0) cb.ckey[c] = mine;
            if (tid == 0 && B <= V) cb.theta = 0u;
            __syncthreads();
            const int span = NC / TPC, ob = (tid % TPC) * span;
            int cnt = 0;
#pragma unroll 8
            for (int o = ob; o < ob + span; o++) {
                const unsigned x = cb.ckey[o];
                cnt += (x > mine || (x == mine && o < c)) ? 1 : 0;
            }
            if (TPC == 2) cnt += __shfl_xor_sync(FULL, cnt, 1);
            if (cnt == B - 1 && mine != 0u && (TPC == 1 || (tid & 1) == 0)) atomicMax(&cb.theta, mine);
            // second bound: the best parent's beam best-ranked candidates, if none of them was absorbed
            if (w == 3 && B <= V) {
                const unsigned mn = __reduce_min_sync(FULL, lane < B ? cb.cand[0][cb.order[lane]] : 0xffffffffu);
                if (lane == 0 && mn != 0u) atomicMax(&cb.theta, mn);
            }
        }
        __syncthreads();
        const unsigned theta = cb.theta;

        // ================= phase D: survivors =================
        {
            unsigned keys4[(BMAX + 3) / 4], masks4[(BMAX + 3) / 4];
            int total = 0;
#pragma unroll
            for (int q = 0; q < (BMAX + 3) / 4; q++) {
                const int i = w + 4 * q;
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
            for (int q = 0; q < (BMAX + 3) / 4; q++) {
                const int pos = base + __popc(masks4[q] & ((1u << lane) - 1u));
                if (((masks4[q] >> lane) & 1u) && pos < 64) {
                    cb.surv_key[pos] = keys4[q]; cb.surv_iv[pos] = ((w + 4 * q) << 8) | lane;
                }
                base += __popc(masks4[q]);
            }
        }
        __syncthreads();
        const int ns = cb.ns;
        if (tid == 0) { stat_surv += ns; stat_fallback += ns > 64; }

        // ================= phase E: exact order of the survivors =================
        int m = 0;
        if (ns <= 64) {
            m = ns < B ? ns : B;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int sidx = lane + 32 * q;
                if (sidx < ns) {
                    const unsigned key = cb.surv_key[sidx];
                    const int iv = cb.surv_iv[sidx];
                    const int mi = iv >> 8, mv = iv & 0xff;
                    const int ms = cand_suffix_id(mv, blank, pk[mi]);
                    int rank = 0;
                    for (int o = w; o < ns; o += 4) {
                        const unsigned ok = cb.surv_key[o];
                        if (ok > key) rank++;
                        else if (ok == key && o != sidx) {
                            const int oiv = cb.surv_iv[o];
                            if (t == 0) rank += oiv < iv;
                            else {
                                const int oi = oiv >> 8, ov = oiv & 0xff;
                                rank += cand_less_rel(rel[oi][mi], cand_suffix_id(ov, blank, pk[oi]), ms, vch) ? 1 : 0;
                            }
                        }
                    }
                    if (rank) atomicAdd(&cb.rankc[sidx], rank);
                }
            }
            __syncthreads();
            if (w == 0) {
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int sidx = lane + 32 * q;
                    if (sidx < ns) {
                        const int rank = cb.rankc[sidx];
                        if (rank < B) {
                            const int iv = cb.surv_iv[sidx];
                            cb.selkey[rank] = cb.surv_key[sidx]; cb.seli[rank] = iv >> 8; cb.selv[rank] = iv & 0xff;
                        }
                    }
                }
            }
        } else {
            // more than 64 survivors (loose bound): beam rounds of warp-max extraction on warp 0
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
                    if (lane == 0) { cb.selkey[m] = gmax; cb.seli[m] = wi; cb.selv[m] = wl; }
                    __syncwarp();
                }
                if (lane == 0) cb.sel_m = m;
            }
            __syncthreads();
            m = cb.sel_m;
        }
        __syncthreads();

        // ================= phase F: the selected candidates become the next kept states =================
        const int nxt = cur ^ 1;
        if (w == 0) {
            bool need_new = false;
            int i = 0, v = 0, nd = 0, pn = 0, dp = 0, npk = 0;
            if (lane < m) {
                i = cb.seli[lane]; v = cb.selv[lane];
                const int pki = pk[i];
                const int ebi = (pki >> 8) & 1, lasti = pki & 0xff;
                if (v == blank) { nd = node[i]; pn = pnode[i]; dp = depth[i]; npk = lasti | (1 << 8); }
                else if (ebi == 0 && v == lasti) { nd = node[i]; pn = pnode[i]; dp = depth[i]; npk = lasti; }
                else {
                    pn = node[i]; dp = depth[i] + 1; npk = v;
                    nd = child[(size_t)pn * Vp + v];
                    need_new = nd == 0;
                }
            }
            const unsigned nb = __ballot_sync(FULL, need_new);
            const int nodes = cb.nodes;
            if (need_new) {
                nd = nodes + __popc(nb & ((1u << lane) - 1u));
                parent[nd] = pn;
                meta[nd] = (dp << 8) | v;
                if (born) born[nd] = t;
                child[(size_t)pn * Vp + v] = nd;
                int4 *row = reinterpret_cast<int4 *>(child + (size_t)nd * Vp);
                for (int q = 0; q < Vp / 4; q++) row[q] = make_int4(0, 0, 0, 0);
            }
            if (lane < m) {
                cb.sc[nxt][lane] = ord2f(cb.selkey[lane]);
                cb.node[nxt][lane] = nd; cb.pnode[nxt][lane] = pn; cb.depth[nxt][lane] = dp; cb.pk[nxt][lane] = npk;
            }
            __syncwarp();
            if (lane == 0) { cb.nodes = nodes + __popc(nb); cb.kept = m; }
        } else if (w == 1) {
            // rank of the NEXT frame's scores (independent of the beam)
            if (t + 1 < p.t1) {
                const unsigned mine = active ? f2ord(lp_next) : 0u;
                int lr = 0;
#pragma unroll
                for (int u = 0; u < 32; u++) {
                    const unsigned x = __shfl_sync(FULL, mine, u);
                    lr += (x > mine || (x == mine && u < lane)) ? 1 : 0;
                }
                cb.order[lr] = lane;
            }
        } else {
            // prefix relations of the new beam from the current one and this frame's choices (only tie-breaks read them)
            for (int e = tid - 64; e < BMAX * BMAX; e += 64) {
                const int r = e / BMAX, q = e % BMAX;
                if (r >= m || q >= m) continue;
                const int ar = cb.seli[r], aq = cb.seli[q];
                const int er = cand_ext_id(cb.selv[r], blank, pk[ar]);
                const int eq2 = cand_ext_id(cb.selv[q], blank, pk[aq]);
                cb.rel[nxt][r][q] = (unsigned char)rel_child(rel[ar][aq], er, eq2, depth[ar], depth[aq], node[ar], node[aq], vch,
                                                             parent, meta);
            }
        }
        __syncthreads();
        kept = m;
        cur = nxt;
    }

    if (tid == 0) {
        if (p.t0 == 0) { p.out_stats[2 * utt] = stat_fallback; p.out_stats[2 * utt + 1] = stat_surv; }
        else { p.out_stats[2 * utt] += stat_fallback; p.out_stats[2 * utt + 1] += stat_surv; }
    }
    if (p.t1 < Tu) {
        const int4 *src = reinterpret_cast<const int4 *>(&cb);
        for (int i = tid; i < kStateVec; i += 128) gstate[i] = src[i];
        if (tid == 0) gstate[kStateVec] = make_int4(cur, 0, 0, 0);
        return;
    }
    // ---- result (CTCBeamSearch.cu:290-298) ----
    if (tid == 0 && p.out_counts) p.out_counts[utt] = kept;
    for (int r = tid; r < p.nbest; r += 128) {
        char *out = p.out_paths + ((size_t)utt * p.nbest + r) * p.max_len;
        int *ts = p.out_ts ? p.out_ts + ((size_t)utt * p.nbest + r) * p.max_len : nullptr;
        int len = 0;
        float scv = 0.0f;
        if (r < kept) {
            int nd = cb.node[cur][r];
            const int dpt = cb.depth[cur][r];
            len = dpt;
            if (Tu == 1 && ((cb.pk[cur][r] >> 8) & 1)) { if (len < p.max_len) { out[len] = vch[blank]; if (ts) ts[len] = 0; } len += 1; }
            for (int pos = dpt - 1; pos >= 0; pos--) {
                if (pos < p.max_len) { out[pos] = vch[meta[nd] & 0xff]; if (ts) ts[pos] = born[nd]; }
                nd = parent[nd];
            }
            scv = cb.sc[cur][r];
        }
        p.out_lens[(size_t)utt * p.nbest + r] = len;
        p.out_scores[(size_t)utt * p.nbest + r] = scv;
    }
}

int ctc_launch_cta(const CtcParams &p, int domain, int utterances, size_t pad, cudaStream_t st) {
    if (domain == GASR_DOMAIN_LOG) {
        if (p.beam <= 16) ctc_beam_cta_kernel<1, 16><<<utterances, 128, pad, st>>>(p); else ctc_beam_cta_kernel<1, 32><<<utterances, 128, pad, st>>>(p);
    } else {
        if (p.beam <= 16) ctc_beam_cta_kernel<0, 16><<<utterances, 128, pad, st>>>(p); else ctc_beam_cta_kernel<0, 32><<<utterances, 128, pad, st>>>(p);
    }
    return GASR_OK;
}

}  // namespace gasr
