// ctc_beam.cu -- CTC prefix beam search, one CTA per utterance, candidates resident in shared memory.
//
// Stands behind CTCBeamSearch::decode (reference CTCBeamSearch.cu:262-312) and implements the CTC-REF
// contract of SURVEY.md 8c / DESIGN.md: the reference's extension rules (CTCBeamSearch.cu:404-458), merge of
// equal paths (:460-489) with the summation order fixed to ascending (raw string, candidate index), stable
// descending prune (:174-196, :103-112), result = rank-0 state (:290-298).
//
// What is different from the reference's ~40 launches + Thrust sorts per frame:
//   * a kept state is (X, eb) = (label prefix, ends-in-blank); X is a node of a per-utterance prefix trie in HBM
//     (parent / char / depth + a child table so node ids are canonical over time) -- no 264-byte BeamState,
//     no string sort, no 31-hash (equal paths merge by identity, never by hash collision);
//   * duplicates are found structurally: a candidate can only coincide with its twin state's candidate
//     ((X,0) and (X,1)) or with the "stay" candidate of a kept child state, so every merged candidate is
//     produced once, by one thread, with its <=3 (<=5 on the last frame) addends summed in canonical order;
//   * prune = one in-shared-memory bitonic sort of 64-bit keys (ordered score | ~candidate index); exact score
//     ties are re-ordered by raw-string order (trie walk to the lowest common ancestor), as the reference's
//     stable sort on top of the string sort does.
// This file: the general kernel (any beam / vocabulary); the fast paths for beam <= 32 and vocabulary <= 32 are
// ctc_beam_warp.cu / ctc_beam_cta.cu / ctc_beam_cta2.cu, the host side and the dispatch rules are in ctc_beam.cu.
#include "ctc_beam.cuh"

namespace gasr {

template <int DOMAIN, int MAXT>
__global__ void __launch_bounds__(MAXT) ctc_beam_kernel(const CtcParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int utt = blockIdx.x;
    const int V = p.V, B = p.beam, blank = p.blank, Vp = p.Vp, n_pad = p.n_pad;

    // ---- shared-memory carve-up --------------------------------------------------------------------
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned char *sp = smem_raw + sizeof(unsigned long long) * n_pad;
    float *lp = reinterpret_cast<float *>(sp); sp += sizeof(float) * Vp;
    float *sc2 = reinterpret_cast<float *>(sp); sp += sizeof(float) * 2 * B;
    int *node2 = reinterpret_cast<int *>(sp); sp += sizeof(int) * 2 * B;
    int *pnode2 = reinterpret_cast<int *>(sp); sp += sizeof(int) * 2 * B;
    int *newflag = reinterpret_cast<int *>(sp); sp += sizeof(int) * B;
    short *twin = reinterpret_cast<short *>(sp); sp += sizeof(short) * B;
    short *P0 = reinterpret_cast<short *>(sp); sp += sizeof(short) * B;
    short *P1 = reinterpret_cast<short *>(sp); sp += sizeof(short) * B;
    short *last2 = reinterpret_cast<short *>(sp); sp += sizeof(short) * 2 * B;
    uint16_t *redir0 = reinterpret_cast<uint16_t *>(sp); sp += sizeof(uint16_t) * B * V;
    uint16_t *redir1 = reinterpret_cast<uint16_t *>(sp); sp += sizeof(uint16_t) * B * V;
    char *vch = reinterpret_cast<char *>(sp); sp += (V + 3) / 4 * 4;
    unsigned char *eb2 = sp; sp += 2 * B;
    sp = smem_raw + (((size_t)(sp - smem_raw) + 7) & ~(size_t)7);
    int *depth2 = reinterpret_cast<int *>(sp); sp += sizeof(int) * 2 * B;
    int *sel_i = reinterpret_cast<int *>(sp); sp += sizeof(int) * B;
    int *sel_v = reinterpret_cast<int *>(sp); sp += sizeof(int) * B;
    unsigned short *relw = reinterpret_cast<unsigned short *>(sp);      // [2][B][B], only when p.use_rel
    const bool use_rel = p.use_rel != 0;
    // the two beam buffers (current / next) are halves of the arrays above; no dynamically indexed struct array
    auto beam_view = [&](int w) {
        BeamView v;
        v.score = sc2 + w * B; v.node = node2 + w * B; v.pnode = pnode2 + w * B; v.last = last2 + w * B;
        v.eb = eb2 + w * B;
        return v;
    };
    __shared__ int s_kept, s_nodes, s_m;
    __shared__ unsigned s_hist[256], s_prefix, s_need, s_lo, s_hi, s_nv;
    __shared__ int s_wcnt[32];
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned vinv = 0xffffffffu / (unsigned)p.V + 1u;             // ceil(2^32 / V): exact quotients for c * V < 2^32

    int *parent = p.parent + (size_t)utt * p.cap;
    int *meta = p.meta + (size_t)utt * p.cap;
    int *born = p.born ? p.born + (size_t)utt * p.cap : nullptr;
    int *child = p.child + (size_t)utt * p.cap * Vp;
    const float *S = p.scores + (size_t)utt * p.ld;
    const size_t frame_stride = (size_t)p.frame_rows * p.ld;

    // ---- init: one virtual parent (empty prefix, "ends in blank", unit score); frame 0 then yields the
    //      reference's t = 0 states (kernelInitialPath, CTCBeamSearch.cu:337-364) -------------------------
    for (int v = tid; v < V; v += NT) vch[v] = p.vocab[v];
    for (int v = tid; v < Vp; v += NT) child[v] = 0;   // root's child row
    if (tid == 0) {
        parent[0] = -1; meta[0] = 0 | 0xff;
        sc2[0] = DOMAIN ? 0.0f : 1.0f;
        node2[0] = 0; pnode2[0] = kNone; last2[0] = -1; eb2[0] = 1;
        depth2[0] = 0;
        if (use_rel) relw[0] = RW_EQ;
        s_kept = 1; s_nodes = 1;
    }
    float next_lp = 0.0f;
    if (tid < V) next_lp = S[tid];
    __syncthreads();

    int cur = 0;
    const int Tu = utt_frames(p, utt);
    for (int t = 0; t < Tu; t++) {
        const BeamView st = beam_view(cur), nx = beam_view(cur ^ 1);
        const int k = s_kept;
        const int ncand = k * V;
        const bool last_frame = (t == Tu - 1) && (t > 0);

        // ---- A: this frame's scores to smem, prefetch the next row, beam-level relations ----------------
        if (tid < V) {
            lp[tid] = next_lp;
            if (t + 1 < Tu) next_lp = S[(size_t)(t + 1) * frame_stride + tid];
        }
        for (int c = tid; c < ncand; c += NT) { redir0[c] = kNoRedir; redir1[c] = kNoRedir; }
        if (tid == 0) { s_m = 0; s_lo = 0xffffffffu; s_hi = 0u; s_nv = 0u; }
        {
            // twin (same prefix, other "ends in blank" flag) and parent states of every kept state: each is unique if it
            // exists, so slices of the scan (2^tsh threads per state) combine with a max
            int tsh = 0;
            while (tsh < 5 && (2 << tsh) * k <= NT) tsh++;
            const int i = tid >> tsh, sub = tid & ((1 << tsh) - 1);
            int tw = kNone, p0 = kNone, p1 = kNone;
            if (i < k) {
                const int nd = st.node[i], pn = st.pnode[i];
                for (int j = sub; j < k; j += 1 << tsh) {
                    const int nj = st.node[j];
                    if (nj == nd && j != i) tw = j;
                    if (nj == pn) { if (st.eb[j]) p1 = j; else p0 = j; }
                }
            }
            for (int off = 1; off < (1 << tsh); off <<= 1) {
                tw = max(tw, __shfl_xor_sync(0xffffffffu, tw, off));
                p0 = max(p0, __shfl_xor_sync(0xffffffffu, p0, off));
                p1 = max(p1, __shfl_xor_sync(0xffffffffu, p1, off));
            }
            if (i < k && sub == 0) { twin[i] = (short)tw; P0[i] = (short)p0; P1[i] = (short)p1; }
        }
        __syncthreads();
        // ---- B: kept child states claim the extend candidates that land on them --------------------------
        if (tid < k && st.last[tid] >= 0) {
            const int lv = st.last[tid];
            uint16_t *rd = st.eb[tid] ? redir1 : redir0;
            if (P0[tid] >= 0) rd[P0[tid] * V + lv] = (uint16_t)tid;
            if (P1[tid] >= 0) rd[P1[tid] * V + lv] = (uint16_t)tid;
        }
        __syncthreads();
        // ---- C: merged candidates -> sort keys ------------------------------------------------------------
        for (int c = tid; c < n_pad; c += NT) {
            unsigned long long key = 0ull;
            if (c < ncand) {
                const int i = V == 1 ? c : (int)__umulhi((unsigned)c, vinv), v = c - i * V;      // c / V (c < 2^16, V <= 255)
                const float pv = lp[v];
                const float s = comb<DOMAIN>(st.score[i], pv);
                const int tw = twin[i];
                const int ebi = st.eb[i], lasti = st.last[i];
                bool host = true;
                float acc = s;
                if (v == blank) {
                    if (!last_frame) {
                        if (tw >= 0) {
                            if (tw < i) host = false;
                            else acc = mrg<DOMAIN>(s, comb<DOMAIN>(st.score[tw], pv));
                        }
                    } else {
                        if (ebi == 0 || tw >= 0) host = false;   // the (X,0) "stay" slot hosts the whole group
                        else if (lasti >= 0) {
                            const int p0 = P0[i], p1 = P1[i];
                            if (p1 >= 0 || (p0 >= 0 && st.last[p0] != lasti)) host = false;  // an extend slot hosts
                        }
                    }
                } else if (ebi == 0 && v == lasti) {
                    // stay on X: plus the extends of X's parent states that spell X again
                    int m0 = P0[i], m1 = P1[i], m2 = i;
                    if (m0 >= 0 && st.last[m0] == v) m0 = kNone;   // (P,0)+v with last(P)==v stays on P
                    // ascending state index == ascending candidate index (same v)
                    int a0 = m0, a1 = m1, a2 = m2, tmp;
                    if (a0 > a1) { tmp = a0; a0 = a1; a1 = tmp; }
                    if (a1 > a2) { tmp = a1; a1 = a2; a2 = tmp; }
                    if (a0 > a1) { tmp = a0; a0 = a1; a1 = tmp; }
                    bool have = false;
                    acc = 0.0f;
                    const int order[3] = {a0, a1, a2};
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        const int j = order[q];
                        if (j < 0) continue;
                        const float sj = comb<DOMAIN>(st.score[j], pv);
                        acc = have ? mrg<DOMAIN>(acc, sj) : sj;
                        have = true;
                    }
                    if (last_frame) {
                        const float pb = lp[blank];
                        int b0 = i, b1 = tw;
                        if (b1 >= 0 && b1 < b0) { b0 = tw; b1 = i; }
                        acc = mrg<DOMAIN>(acc, comb<DOMAIN>(st.score[b0], pb));
                        if (b1 >= 0) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(st.score[b1], pb));
                    }
                } else {
                    // extend to X.v
                    if (redir0[c] != kNoRedir) host = false;         // kept state (X.v, 0) hosts it
                    else {
                        const bool tw_member = (tw >= 0) && (st.eb[tw] == 1 || v != st.last[tw]);
                        if (tw_member) {
                            if (tw < i) host = false;
                            else acc = mrg<DOMAIN>(s, comb<DOMAIN>(st.score[tw], pv));
                        }
                        if (host && last_frame) {
                            const int j = redir1[c];                 // kept (X.v, 1): its blank candidate strips to X.v
                            if (j != kNoRedir) acc = mrg<DOMAIN>(acc, comb<DOMAIN>(st.score[j], lp[blank]));
                        }
                    }
                }
                if (host) key = ((unsigned long long)f2ord(acc) << 32) | (unsigned)(0xffffffffu - (unsigned)c);
            }
            keys[c] = key;
        }
        __syncthreads();
        // ---- D: prune.  Only the kept window matters, so instead of sorting all n_pad keys: radix-select the beam-th
        //      largest score (four 8-bit passes over the order-preserving score bits, warp-aggregated shared-memory
        //      histogram), move every candidate not below it to the front of keys[], and order that short list.
        uint32_t thr = 0;
        {
            // the candidates' scores span a narrow band (a few thousand fp32 steps): select on (score - minimum), whose
            // leading zero bytes need no pass -- usually 2 passes instead of 4, and the digits are spread over the bins
            uint32_t lo = 0xffffffffu, hi = 0u;
            unsigned nv = 0u;
            for (int c = tid; c < ncand; c += NT) {
                const unsigned long long key = keys[c];
                if (key != 0ull) { const uint32_t o = (uint32_t)(key >> 32); lo = o < lo ? o : lo; hi = o > hi ? o : hi; nv++; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const uint32_t l2 = __shfl_xor_sync(0xffffffffu, lo, off), h2 = __shfl_xor_sync(0xffffffffu, hi, off);
                lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
                nv += __shfl_xor_sync(0xffffffffu, nv, off);
            }
            if (lane == 0 && nv) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); atomicAdd(&s_nv, nv); }
            __syncthreads();
            lo = s_lo; hi = s_hi;
            const uint32_t span = hi - lo;
            const int passes = ((int)s_nv <= B) ? 0 : (39 - __clz(span | 1u)) >> 3;      // ceil(bits(span) / 8); span 0 -> 1
            uint32_t prefix = 0;
            unsigned need = (unsigned)B;
            for (int shift = 8 * (passes - 1); shift >= 0; shift -= 8) {
                for (int i = tid; i < 256; i += NT) s_hist[i] = 0u;
                __syncthreads();
                const bool top = shift == 8 * (passes - 1);
                for (int base = 0; base < ncand; base += NT) {           // uniform trip count (warp votes below)
                    const int c = base + tid;
                    unsigned digit = 256u;
                    if (c < ncand) {
                        const unsigned long long key = keys[c];
                        const uint32_t o = (uint32_t)(key >> 32) - lo;
                        if (key != 0ull && (top || (o >> (shift + 8)) == (prefix >> (shift + 8)))) digit = (o >> shift) & 255u;
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, digit < 256u);
                    if (bal == 0u) continue;
                    const int first = __ffs(bal) - 1;
                    const unsigned d0 = __shfl_sync(0xffffffffu, digit, first);
                    if (__all_sync(0xffffffffu, digit >= 256u || digit == d0)) {       // one bin for the whole warp
                        if (lane == first) atomicAdd(&s_hist[d0], (unsigned)__popc(bal));
                    } else if (digit < 256u) atomicAdd(&s_hist[digit], 1u);
                }
                __syncthreads();
                if (warp == 0) {
                    unsigned cnt[8], sum = 0u;
#pragma unroll
                    for (int j = 0; j < 8; j++) { cnt[j] = s_hist[lane * 8 + j]; sum += cnt[j]; }
                    unsigned suf = sum;                                   // inclusive suffix sum over lanes (lane .. 31)
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const unsigned o = __shfl_down_sync(0xffffffffu, suf, off);
                        if (lane + off < 32) suf += o;
                    }
                    const unsigned above = suf - sum;
                    if (above < need && suf >= need) {                    // the beam-th largest has its digit in my 8 bins
                        unsigned acc = above;
#pragma unroll
                        for (int j = 7; j >= 0; j--) {
                            if (acc < need && acc + cnt[j] >= need) { s_prefix = prefix | ((uint32_t)(lane * 8 + j) << shift); s_need = need - acc; }
                            acc += cnt[j];
                        }
                    }
                }
                __syncthreads();
                prefix = s_prefix; need = s_need;
            }
            thr = passes == 0 ? 0u : lo + prefix;                        // no more candidates than the beam: keep all
        }
        // compaction in place, NT keys per round: a round's keys are all read before its survivors are written, and
        // the survivors land below the end of that round's range
        for (int base = 0; base < ncand; base += NT) {
            const int c = base + tid;
            unsigned long long key = 0ull;
            if (c < ncand) key = keys[c];
            const bool keep = key != 0ull && (uint32_t)(key >> 32) >= thr;
            __syncthreads();
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            int wbase = 0;
            if (lane == 0 && bal) wbase = atomicAdd(&s_m, __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (keep) keys[wbase + __popc(bal & ((1u << lane) - 1u))] = key;
        }
        __syncthreads();
        const int M = s_m;                       // >= min(beam, #candidates); larger only by ties at the threshold
        // raw-string order of two merged candidates (exact score ties, t > 0): O(1) through the relation matrix
        auto cand_before = [&](unsigned long long ka, unsigned long long kb) -> bool {
            const int ca = (int)(0xffffffffu - (uint32_t)ka), cb = (int)(0xffffffffu - (uint32_t)kb);
            if (t == 0) return ca < cb;
            const int ia = ca / V, va = ca - ia * V, ib = cb / V, vb = cb - ib * V;
            const bool staya = (va != blank) && (st.eb[ia] == 0 && va == st.last[ia]);
            const bool stayb = (vb != blank) && (st.eb[ib] == 0 && vb == st.last[ib]);
            if (use_rel)
                return candw_less(relw[((size_t)cur * B + ia) * B + ib], staya ? -1 : va, stayb ? -1 : vb, vch);
            return raw_less(parent, meta, vch, st.node[ia], staya ? 0 : vch[va], st.node[ib], stayb ? 0 : vch[vb]);
        };
        if (M <= NT) {
            // rank sort of the survivors, TPE threads per survivor (each scans a slice of the list, shuffle-reduced);
            // equal scores rank by raw string (CTC-REF step 4: ties keep the ascending string order of step 3; t = 0:
            // label order).  The survivors' (parent state, suffix label) are decoded once into the redirect tables,
            // which are dead after phase C.
            int tpe = 1;
            while (tpe < 32 && 2 * tpe * M <= NT) tpe <<= 1;
            if (tid < M) {
                const int c = (int)(0xffffffffu - (uint32_t)keys[tid]);
                const int i = c / V, v = c - i * V;
                const bool stay = (v != blank) && (st.eb[i] == 0 && v == st.last[i]);
                redir0[tid] = (uint16_t)i;
                redir1[tid] = (uint16_t)(stay ? 0 : v + 1);               // suffix label + 1, 0 = none
            }
            __syncthreads();
            const int x = tid / tpe, sub = tid & (tpe - 1);
            unsigned long long mykey = 0ull;
            int rank = 0;
            if (x < M) {
                mykey = keys[x];
                const uint32_t ms = (uint32_t)(mykey >> 32);
                const int ix = redir0[x], sx = (int)redir1[x] - 1;
                for (int y = sub; y < M; y += tpe) {
                    const unsigned long long ky = keys[y];
                    const uint32_t ys = (uint32_t)(ky >> 32);
                    if (ys > ms) rank++;
                    else if (ys == ms && y != x) {
                        bool before;
                        if (t == 0) before = ky > mykey;                 // smaller candidate index first
                        else {
                            const int iy = redir0[y], sy = (int)redir1[y] - 1;
                            if (use_rel) before = candw_less(relw[((size_t)cur * B + iy) * B + ix], sy, sx, vch);
                            else before = raw_less(parent, meta, vch, st.node[iy], sy < 0 ? 0 : vch[sy], st.node[ix], sx < 0 ? 0 : vch[sx]);
                        }
                        rank += before ? 1 : 0;
                    }
                }
            }
            for (int off = 1; off < tpe; off <<= 1) rank += __shfl_xor_sync(0xffffffffu, rank, off);
            __syncthreads();
            if (x < M && sub == 0) keys[rank] = mykey;
            for (int c = M + tid; c < B; c += NT) keys[c] = 0ull;
            __syncthreads();
        } else {
            // (more survivors than threads: massive ties) bitonic sort of the survivors, descending, then the tied runs
            // that reach into the kept window are put into raw-string order by one thread
            int n_sort = 32;
            while (n_sort < M) n_sort <<= 1;
            for (int c = M + tid; c < n_sort || c < B; c += NT) keys[c] = 0ull;
            __syncthreads();
            for (int size = 2; size <= n_sort; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int idx = tid; idx < (n_sort >> 1); idx += NT) {
                        const int pos = 2 * idx - (idx & (stride - 1));
                        const unsigned long long a = keys[pos], b = keys[pos + stride];
                        const bool desc = (pos & size) == 0;
                        if ((a < b) == desc) { keys[pos] = b; keys[pos + stride] = a; }
                    }
                    __syncthreads();
                }
            }
            if (t > 0 && tid == 0) {
                int r = 0;
                while (r < B && r < M) {
                    const uint32_t sc = (uint32_t)(keys[r] >> 32);
                    int e = r + 1;
                    while (e < M && (uint32_t)(keys[e] >> 32) == sc) e++;
                    for (int x = r + 1; x < e; x++) {                    // insertion sort of keys[r:e), ascending raw string
                        const unsigned long long kx = keys[x];
                        int y = x - 1;
                        while (y >= r && cand_before(kx, keys[y])) { keys[y + 1] = keys[y]; y--; }
                        keys[y + 1] = kx;
                    }
                    r = e;
                }
            }
            __syncthreads();
        }
        // ---- F: the top-B merged candidates become the next kept states -----------------------------------
        int my_new = 0, my_i = 0, my_v = 0;
        bool valid = false;
        if (tid < B) {
            const unsigned long long key = keys[tid];
            valid = key != 0ull;
            if (valid) {
                const int c = (int)(0xffffffffu - (uint32_t)key);
                my_i = c / V; my_v = c - my_i * V;
                nx.score[tid] = ord2f((uint32_t)(key >> 32));
                const bool stay = (my_v == blank) || (st.eb[my_i] == 0 && my_v == st.last[my_i]);
                if (stay) {
                    nx.node[tid] = st.node[my_i]; nx.pnode[tid] = st.pnode[my_i]; nx.last[tid] = st.last[my_i];
                    nx.eb[tid] = (my_v == blank) ? 1 : 0;
                } else {
                    const int pn = st.node[my_i];
                    const int nd = child[(size_t)pn * Vp + my_v];
                    nx.pnode[tid] = pn; nx.last[tid] = (short)my_v; nx.eb[tid] = 0;
                    nx.node[tid] = nd;            // 0 = not created yet
                    my_new = (nd == 0);
                }
            }
            newflag[tid] = my_new;
            sel_i[tid] = valid ? my_i : -1;
            sel_v[tid] = (valid && !((my_v == blank) || (st.eb[my_i] == 0 && my_v == st.last[my_i]))) ? my_v : -1;   // appended label
            if (valid) {
                const bool stay2 = (my_v == blank) || (st.eb[my_i] == 0 && my_v == st.last[my_i]);
                depth2[(cur ^ 1) * B + tid] = depth2[cur * B + my_i] + (stay2 ? 0 : 1);
            }
        }
        const unsigned new_bal = __ballot_sync(0xffffffffu, my_new != 0);
        if (lane == 0) s_wcnt[warp] = __popc(new_bal);
        __syncthreads();
        if (use_rel) {
            // prefix relations of the new beam from the current one and this frame's choices (old node ids still in st)
            const unsigned short *rc = relw + (size_t)cur * B * B;
            unsigned short *rn = relw + (size_t)(cur ^ 1) * B * B;
            int r = tid / B, q = tid - r * B;
            const int dr = NT / B, dq = NT - dr * B;
            for (int e = tid; e < B * B; e += NT) {
                const int ar = sel_i[r], aq = sel_i[q];
                if (ar >= 0 && aq >= 0)
                    rn[e] = (unsigned short)relw_child(rc[(size_t)ar * B + aq], sel_v[r], sel_v[q], depth2[cur * B + ar],
                                                       depth2[cur * B + aq], st.node[ar], st.node[aq], vch, parent, meta);
                r += dr; q += dq;
                if (q >= B) { q -= B; r++; }
            }
        }
        if (tid < B && valid) {
            if (my_new) {
                int off = __popc(new_bal & ((1u << lane) - 1u));
                for (int w = 0; w < warp; w++) off += s_wcnt[w];
                const int nd = s_nodes + off;
                const int pn = st.node[my_i];
                parent[nd] = pn;
                meta[nd] = (((meta[pn] >> 8) + 1) << 8) | my_v;
                if (born) born[nd] = t;
                child[(size_t)pn * Vp + my_v] = nd;
                int4 *row = reinterpret_cast<int4 *>(child + (size_t)nd * Vp);
                for (int q = 0; q < Vp / 4; q++) row[q] = make_int4(0, 0, 0, 0);
                nx.node[tid] = nd;
            }
        }
        const int created = __syncthreads_count(my_new != 0);
        if (tid == 0) { s_nodes += created; s_kept = M < B ? M : B; }
        cur ^= 1;
        __syncthreads();
    }

    // ---- result: kept states best first; path = labels of X (blank stripped), CTCBeamSearch.cu:290-298 ------
    const BeamView st = beam_view(cur);
    const int kept = s_kept;
    if (tid == 0 && p.out_counts) p.out_counts[utt] = kept;
    for (int r = tid; r < p.nbest; r += NT) {
        char *out = p.out_paths + ((size_t)utt * p.nbest + r) * p.max_len;
        int *ts = p.out_ts ? p.out_ts + ((size_t)utt * p.nbest + r) * p.max_len : nullptr;
        int len = 0;
        float sc = 0.0f;
        if (r < kept) {
            int nd = st.node[r];
            const int depth = meta[nd] >> 8;
            len = depth;
            // T == 1: the reference returns the initial path as is, blank included
            if (Tu == 1 && st.eb[r]) { if (len < p.max_len) { out[len] = vch[blank]; if (ts) ts[len] = 0; } len += 1; }
            for (int pos = depth - 1; pos >= 0; pos--) {
                if (pos < p.max_len) { out[pos] = vch[meta[nd] & 0xff]; if (ts) ts[pos] = born[nd]; }
                nd = parent[nd];
            }
            sc = st.score[r];
        }
        p.out_lens[(size_t)utt * p.nbest + r] = len;
        p.out_scores[(size_t)utt * p.nbest + r] = sc;
    }
}

int ctc_launch_general(const CtcParams &p, int domain, int utterances, int threads, size_t smem, cudaStream_t st) {
#define GASR_CTC_LAUNCH(DOM, MT)                                                                                  \
    do {                                                                                                          \
        GASR_CUDA(cudaFuncSetAttribute(ctc_beam_kernel<DOM, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                       (int)smem));                                                               \
        ctc_beam_kernel<DOM, MT><<<utterances, threads, smem, st>>>(p);                                           \
    } while (0)
    if (domain == GASR_DOMAIN_LOG) {
        if (threads <= 256) GASR_CTC_LAUNCH(1, 256); else GASR_CTC_LAUNCH(1, 1024);
    } else {
        if (threads <= 256) GASR_CTC_LAUNCH(0, 256); else GASR_CTC_LAUNCH(0, 1024);
    }
#undef GASR_CTC_LAUNCH
    return GASR_OK;
}

}  // namespace gasr
