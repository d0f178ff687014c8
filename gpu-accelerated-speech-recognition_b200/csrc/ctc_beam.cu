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
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gasr {

// Deterministic fp32 log-add-exp (DESIGN.md "log-add-exp"): only correctly rounded IEEE operations, so the CPU
// oracle evaluates the same bits.
__device__ __forceinline__ float logaddexp_det(float a, float b) {
    const float mx = a > b ? a : b;
    const float mn = a > b ? b : a;
    if (mn == -INFINITY) return mx;
    const float d = __fsub_rn(mn, mx);
    if (d < -17.5f) return mx;
    const float n = rintf(__fmul_rn(d, 1.44269504088896341f));
    float r = __fmaf_rn(n, -0.693359375f, d);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float r2 = __fmul_rn(r, r);
    const float ex = __fadd_rn(__fmaf_rn(p, r2, r), 1.0f);
    const float scale = __int_as_float(((int)n + 127) << 23);
    const float e = __fmul_rn(ex, scale);
    const float t = __fdiv_rn(e, __fadd_rn(2.0f, e));
    const float w = __fmul_rn(t, t);
    float q = 1.0f / 13.0f;
    q = __fmaf_rn(q, w, 1.0f / 11.0f);
    q = __fmaf_rn(q, w, 1.0f / 9.0f);
    q = __fmaf_rn(q, w, 1.0f / 7.0f);
    q = __fmaf_rn(q, w, 1.0f / 5.0f);
    q = __fmaf_rn(q, w, 1.0f / 3.0f);
    q = __fmaf_rn(q, w, 1.0f);
    const float l = __fmul_rn(__fmul_rn(2.0f, t), q);
    return __fadd_rn(mx, l);
}

// Same function, same bits, without data-dependent branches (selects instead of early returns) so that several
// independent evaluations interleave in one warp.
__device__ __forceinline__ float logaddexp_det_bf(float a, float b) {
    const float mx = a > b ? a : b;
    const float mn = a > b ? b : a;
    const float d0 = __fsub_rn(mn, mx);
    const bool skip = !(d0 >= -17.5f);          // d < -17.5, mn = -inf (d = -inf) or both -inf (d = NaN)
    const float d = skip ? 0.0f : d0;
    const float n = rintf(__fmul_rn(d, 1.44269504088896341f));
    float r = __fmaf_rn(n, -0.693359375f, d);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float r2 = __fmul_rn(r, r);
    const float ex = __fadd_rn(__fmaf_rn(p, r2, r), 1.0f);
    const float scale = __int_as_float(((int)n + 127) << 23);
    const float e = __fmul_rn(ex, scale);
    const float t = __fdiv_rn(e, __fadd_rn(2.0f, e));
    const float w = __fmul_rn(t, t);
    float q = 1.0f / 13.0f;
    q = __fmaf_rn(q, w, 1.0f / 11.0f);
    q = __fmaf_rn(q, w, 1.0f / 9.0f);
    q = __fmaf_rn(q, w, 1.0f / 7.0f);
    q = __fmaf_rn(q, w, 1.0f / 5.0f);
    q = __fmaf_rn(q, w, 1.0f / 3.0f);
    q = __fmaf_rn(q, w, 1.0f);
    const float l = __fmul_rn(__fmul_rn(2.0f, t), q);
    return skip ? mx : __fadd_rn(mx, l);
}

// order-preserving float -> uint32 (larger float => larger key); every real score maps to a key > 0
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct CtcParams {
    const float *scores;
    int T, N, V, ld, beam, blank;
    int frame_rows; // rows of `scores` per frame (>= N; the wave engine pads the batch to whole groups of 128)
    int *born;       // [N, cap] frame at which a trie node was created (null: per-token timesteps not wanted)
    int *out_ts;     // [N, nbest, max_len] frame at which each output token's prefix first entered the beam (null: not wanted)
    const int *lens; // per-utterance frame counts (device, N entries, clamped to 1..T); null = every utterance has T frames
    int Vp;        // child-table row pitch (ints)
    int n_pad;     // power of two >= beam * V
    int cap;       // trie nodes per utterance
    int max_len, nbest;
    const char *vocab;   // device copy
    int *parent;         // [N, cap]
    int *meta;           // [N, cap]  depth << 8 | vocab id of the node's last label
    int *child;          // [N, cap, Vp] 0 = absent
    int *anc;            // [N, cap] skip pointer: the ancestor at the last multiple-of-32 depth below the node's own (cta2 kernel)
    char *out_paths;     // [N, nbest, max_len]
    int *out_lens;       // [N, nbest]
    float *out_scores;   // [N, nbest]
    int *out_counts;     // [N]
    int *out_stats;      // [N, 2]: frames that took the prune fallback, sum of prune survivors (diagnostics)
    unsigned char cell_i[128], cell_j[128]; // prune lower-bound probe cells (parent rank, score rank), by rising (i+1)(j+1)
    unsigned char cellmap[32 * 32];          // (parent rank, score rank) -> probe cell index, 255 = not probed
    int n_cells;         // 32 (one per lane) or 64
    int use_rel;         // general kernel: prefix-relation matrix in shared memory (O(1) tie-breaks)
    int t0, t1;          // frames [t0, t1) are decoded by this launch (time chunking; warp kernel only)
    unsigned char *state;   // [N, state_stride] saved beam state between chunk launches
    size_t state_stride;
    // streaming (CTA kernel only): frame t may be read once lp_ready[t / lp_fpb] >= lp_need (null: everything is ready)
    const unsigned *lp_ready;
    int lp_need, lp_fpb;
    int *error;
    volatile unsigned *abort;
};

// frames of utterance `utt` (baseline/main.py:45-46 passes out_lens to its decoder): decoding stops after Tu frames and the
// last-frame rule (trailing blank stripped, CTCBeamSearch.cu:452-456) applies at frame Tu - 1
__device__ __forceinline__ int utt_frames(const CtcParams &p, int utt) {
    if (p.lens == nullptr) return p.T;
    const int n = p.lens[utt];
    return n < 1 ? 1 : (n > p.T ? p.T : n);
}

constexpr int kNone = -1;
constexpr uint16_t kNoRedir = 0xffffu;

struct BeamView {
    float *score;
    int *node;
    int *pnode;
    short *last;          // vocab id of the last label of X, -1 for the empty prefix
    unsigned char *eb;    // 1 = raw path ends in the blank
};

template <int DOMAIN>
__device__ __forceinline__ float comb(float s, float p) {
    return DOMAIN ? __fadd_rn(s, p) : __fmul_rn(s, p);
}
template <int DOMAIN>
__device__ __forceinline__ float mrg(float a, float b) {
    return DOMAIN ? logaddexp_det(a, b) : __fadd_rn(a, b);
}
template <int DOMAIN>
__device__ __forceinline__ float mrg_bf(float a, float b) {
    return DOMAIN ? logaddexp_det_bf(a, b) : __fadd_rn(a, b);
}

// raw-string order of two candidates = (trie node, optional suffix char): walk both up to the lowest common
// ancestor and compare the first characters after it (reference operator<, CTCBeamSearch.cu:137-147).
__device__ __noinline__ bool raw_less(const int *__restrict__ parent, const int *__restrict__ meta, const char *vocab, int na,
                         int sufa, int nb, int sufb) {
    int da = meta[na] >> 8, db = meta[nb] >> 8;
    const int lena = da + (sufa ? 1 : 0), lenb = db + (sufb ? 1 : 0);
    int a = na, b = nb, la = 0, lb = 0;  // la/lb: char stepped over last (0 = never stepped)
    while (da > db) { la = vocab[meta[a] & 0xff]; a = parent[a]; da--; }
    while (db > da) { lb = vocab[meta[b] & 0xff]; b = parent[b]; db--; }
    while (a != b) {
        la = vocab[meta[a] & 0xff]; a = parent[a];
        lb = vocab[meta[b] & 0xff]; b = parent[b];
        da--;
    }
    const int ca = la ? la : sufa, cb = lb ? lb : sufb;   // 0 = end of string
    if (ca != cb) return (signed char)ca < (signed char)cb;
    if (ca == 0) return false;
    // same char right after the common ancestor: the string that ends there is a prefix of the other
    const int end = da + 1;
    return lena == end && lenb > end;
}

// ---- prefix relations for the general kernel (vocabulary up to 255) -----------------------------------------------
// rel[a][b] of two kept states' label prefixes: 0 equal, 1 X_a < X_b with the first difference inside both, 2 the
// reverse, 3 + y: X_a is a proper prefix of X_b and y is X_b's next label, 3 + 256 + y: the mirror image.  Updated in
// O(1) per pair and frame (children append one label); makes the raw-string tie-break O(1) instead of a trie walk.
constexpr int RW_EQ = 0, RW_LT = 1, RW_GT = 2, RW_PFX = 3, RW_RPFX = 3 + 256;
__device__ int trie_char_at(const int *__restrict__ parent, const int *__restrict__ meta, int nd, int pos);   // below
__device__ __forceinline__ bool chw_less(const char *vch, int a, int b) { return (signed char)vch[a] < (signed char)vch[b]; }
// raw-string order of candidates (a, suffix sa) and (b, suffix sb); suffix < 0 = none ("stay")
__device__ __forceinline__ bool candw_less(int R, int sa, int sb, const char *vch) {
    if (R == RW_EQ) {
        if (sa < 0) return sb >= 0;
        if (sb < 0 || sa == sb) return false;
        return chw_less(vch, sa, sb);
    }
    if (R == RW_LT) return true;
    if (R == RW_GT) return false;
    if (R < RW_RPFX) {
        const int y = R - RW_PFX;
        if (sa < 0 || sa == y) return true;
        return chw_less(vch, sa, y);
    }
    const int y = R - RW_RPFX;
    if (sb < 0 || sb == y) return false;
    return chw_less(vch, y, sb);
}
// relation of the children (A + er, B + eq2; e < 0 = nothing appended) from the relation R of A and B
__device__ __forceinline__ int relw_child(int R, int er, int eq2, int dA, int dB, int nodeA, int nodeB, const char *vch,
                                          const int *parent, const int *meta) {
    if (R == RW_EQ) {
        if (er < 0 && eq2 < 0) return RW_EQ;
        if (er < 0) return RW_PFX + eq2;
        if (eq2 < 0) return RW_RPFX + er;
        if (er == eq2) return RW_EQ;
        return chw_less(vch, er, eq2) ? RW_LT : RW_GT;
    }
    if (R == RW_LT || R == RW_GT) return R;
    if (R < RW_RPFX) {
        const int y = R - RW_PFX;
        if (er < 0) return R;
        if (er != y) return chw_less(vch, er, y) ? RW_LT : RW_GT;
        if (dB == dA + 1) return eq2 < 0 ? RW_EQ : RW_PFX + eq2;
        return RW_PFX + trie_char_at(parent, meta, nodeB, dA + 1);
    }
    const int y = R - RW_RPFX;
    if (eq2 < 0) return R;
    if (eq2 != y) return chw_less(vch, y, eq2) ? RW_LT : RW_GT;
    if (dA == dB + 1) return er < 0 ? RW_EQ : RW_RPFX + er;
    return RW_RPFX + trie_char_at(parent, meta, nodeA, dB + 1);
}

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

// =====================================================================================================
// Fast path: ONE WARP PER UTTERANCE (beam <= 32, vocabulary <= 32).  lane = vocabulary id, so the V candidates of a
// parent state are evaluated by one warp instruction stream with the parent's fields warp-uniform; the merged
// candidate scores stay in registers (val[i] of lane v = candidate i*V+v), and the prune is beam rounds of
// "warp max" (REDUX) extraction, which yields the kept states already in rank order.  There is no block-level
// barrier at all: warps of a CTA decode different utterances and only use __syncwarp().
//
// Exact score ties are common (fp32 spacing is ~2.4e-4 at |score| ~ 3000), so the raw-string tie-break must be
// O(1): the warp keeps rel[i][j], the lexicographic relation between the label prefixes of kept states i and j
// (equal / first difference inside both / one is a proper prefix of the other + the next character), and updates
// it incrementally when the beam moves -- children only append one character, so the new relation is a function
// of the old one and the two appended characters.  No trie walk on the hot path.
// Same CTC-REF semantics, bit for bit, as ctc_beam_kernel below.
// =====================================================================================================
constexpr int REL_EQ = 0, REL_LT = 1, REL_GT = 2, REL_PFX = 3, REL_RPFX = 3 + 32;   // PFX + y / RPFX + y (y < 32)

template <int BMAX>
struct WarpBeam {
    float sc[2][BMAX];
    int node[2][BMAX];
    int depth[2][BMAX];
    int pk[2][BMAX];          // last label (0xff = none) | eb << 8
    int4 pinfo[BMAX];         // per kept state: {score, twin's score, pk | (twin + 1) << 9, abs0} for the candidate loop
    int tw[BMAX], p0[BMAX], p1[BMAX];
    unsigned abs0[BMAX], abs1[BMAX];
    float stay[BMAX];
    unsigned selkey[BMAX];
    int seli[BMAX], selv[BMAX];
    unsigned char rel[2][BMAX][BMAX];
    unsigned cand[BMAX][32];  // merged candidate keys staged [parent rank][vocab id]; 0 = absorbed / absent
    float pairmm[BMAX / 2][32];   // merged scores of twin pairs (X,0)+(X,1), [pair][vocab id]
    int pair_i[BMAX / 2], pair_tw[BMAX / 2];
    int order[32];            // order[j] = vocab id with the j-th largest score this frame
    unsigned surv_key[68];    // prune survivors (candidates >= the lower bound), in candidate-index order (+ zero padding)
    int surv_iv[64];          // parent rank << 8 | vocab id
    int pad_[3];              // keeps sizeof a multiple of 16 (the beam is parked with int4 copies)
};

// the character at 0-based position pos of the label string of trie node nd (depth(nd) > pos); rare path
__device__ __noinline__ int trie_char_at(const int *__restrict__ parent, const int *__restrict__ meta, int nd, int pos) {
    while ((meta[nd] >> 8) > pos + 1) nd = parent[nd];
    return meta[nd] & 0xff;
}

__device__ __forceinline__ bool ch_less(const char *vch, int a, int b) { return (signed char)vch[a] < (signed char)vch[b]; }

// label appended to the prefix when candidate (state pk, vocab id v) is kept: -1 = none (stay / blank)
__device__ __forceinline__ int cand_ext_id(int v, int blank, int pki) {
    if (v == blank) return -1;
    if (((pki >> 8) & 1) == 0 && v == (pki & 0xff)) return -1;
    return v;
}

// suffix of candidate (state with pk, vocab id v): -1 = none ("stay"), otherwise the appended vocab id
__device__ __forceinline__ int cand_suffix_id(int v, int blank, int pki) {
    if (v == blank) return blank;
    if (((pki >> 8) & 1) == 0 && v == (pki & 0xff)) return -1;
    return v;
}

// raw-string order of candidates (i, sa) and (j, sb) from the relation R = rel[i][j] of their label prefixes
__device__ __forceinline__ bool cand_less_rel(int R, int sa, int sb, const char *vch) {
    if (R == REL_EQ) {
        if (sa < 0) return sb >= 0;
        if (sb < 0 || sa == sb) return false;
        return ch_less(vch, sa, sb);
    }
    if (R == REL_LT) return true;
    if (R == REL_GT) return false;
    if (R < REL_RPFX) {                 // X_i is a proper prefix of X_j, next char y
        const int y = R - REL_PFX;
        if (sa < 0 || sa == y) return true;
        return ch_less(vch, sa, y);
    }
    const int y = R - REL_RPFX;         // X_j is a proper prefix of X_i
    if (sb < 0 || sb == y) return false;
    return ch_less(vch, y, sb);
}

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

// =====================================================================================================
// Latency path: ONE 128-THREAD CTA PER UTTERANCE (beam <= 32, vocabulary <= 32).  Same algorithm and data layout as
// the warp kernel above, but the phases of a frame are spread over four warps so that the serial critical path
// is short when utterances are scarce (cfg2: 64 per GPU):
//   A  warp 0: twin / parent relations from node ids + "stay" candidates | warp 1: rank of the frame's scores |
//      warps 2-3: prefix-relation matrix of the beam chosen in the previous frame (only tie-breaks need it)
//   B  all warps: merged candidates, parent i on warp i % 4 (lane = vocab id)
//   C  warps 0-1: staircase lower bound (32 probe cells each), warp 2: best-parent bound
//   D  all warps: filter rows i % 4 == warp, survivors appended through a shared counter
//   E  all warps: rank counting, "others" o % 4 == warp, partial ranks summed in shared memory
//   F  warp 0: rank < beam -> kept state; trie lookup / allocation
// =====================================================================================================
template <int BMAX>
struct CtaBeam {
    float sc[2][BMAX];
    int node[2][BMAX];
    int pnode[2][BMAX];
    int depth[2][BMAX];
    int pk[2][BMAX];
    int4 pinfo[BMAX];
    int tw[BMAX], p0[BMAX], p1[BMAX];
    unsigned abs0[BMAX], abs1[BMAX];
    float stay[BMAX];
    unsigned selkey[BMAX];
    int seli[BMAX], selv[BMAX];
    unsigned char rel[2][BMAX][BMAX];
    unsigned cand[BMAX][32];
    int order[32];
    unsigned surv_key[64];
    int surv_iv[64];
    int rankc[64];
    unsigned ckey[128];
    unsigned theta;
    int ns, kept, nodes, sel_m;
};

// new prefix relation of kept states r, q (chosen from old states ar, aq with appended labels er, eq2; -1 = none)
__device__ __forceinline__ int rel_child(int R, int er, int eq2, int dA, int dB, int nodeA, int nodeB, const char *vch,
                                         const int *parent, const int *meta) {
    if (R == REL_EQ) {
        if (er < 0 && eq2 < 0) return REL_EQ;
        if (er < 0) return REL_PFX + eq2;
        if (eq2 < 0) return REL_RPFX + er;
        if (er == eq2) return REL_EQ;
        return ch_less(vch, er, eq2) ? REL_LT : REL_GT;
    }
    if (R == REL_LT || R == REL_GT) return R;
    if (R < REL_RPFX) {                          // A proper prefix of B, B = A.y...
        const int y = R - REL_PFX;
        if (er < 0) return R;
        if (er != y) return ch_less(vch, er, y) ? REL_LT : REL_GT;
        if (dB == dA + 1) return eq2 < 0 ? REL_EQ : REL_PFX + eq2;
        return REL_PFX + trie_char_at(parent, meta, nodeB, dA + 1);
    }
    const int y = R - REL_RPFX;                  // B proper prefix of A, A = B.y...
    if (eq2 < 0) return R;
    if (eq2 != y) return ch_less(vch, y, eq2) ? REL_LT : REL_GT;
    if (dA == dB + 1) return er < 0 ? REL_EQ : REL_RPFX + er;
    return REL_RPFX + trie_char_at(parent, meta, nodeA, dB + 1);
}

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

static size_t ctc_smem_bytes(int B, int V, int Vp, int n_pad) {
    size_t s = sizeof(unsigned long long) * n_pad + sizeof(float) * Vp;
    s += (sizeof(float) + 2 * sizeof(int)) * 2 * B;   // score, node, pnode (x2 buffers)
    s += sizeof(int) * B;                             // newflag
    s += 3 * sizeof(short) * B;                       // twin, P0, P1
    s += sizeof(short) * 2 * B;                       // last (x2)
    s += 2 * sizeof(uint16_t) * (size_t)B * V;        // redir0/1
    s += (V + 3) / 4 * 4;                             // vocab chars
    s += 2 * B;                                       // eb (x2)
    s = (s + 7) / 8 * 8;
    s += sizeof(int) * 4 * B;                         // depth (x2), selected parent / label
    return s + 16;
}
static size_t ctc_rel_bytes(int B) { return sizeof(unsigned short) * 2 * (size_t)B * B; }

struct CtcLayout {
    int Vp, n_pad, cap, threads;
    size_t smem, off_vocab, off_parent, off_meta, off_anc, off_child, off_state, state_stride, off_paths, off_lens, off_scores,
        off_counts, off_stats, off_born, off_ts, total;
    size_t out_bytes;
};

static int ctc_layout(const CtcArgs &a, CtcLayout &L) {
    L.Vp = (a.V + 3) / 4 * 4;
    L.n_pad = 32;
    while (L.n_pad < a.beam * a.V) L.n_pad <<= 1;
    L.cap = 1 + a.beam * a.T;
    L.threads = L.n_pad / 2;
    if (L.threads < 128) L.threads = 128;
    if (L.threads > 1024) L.threads = 1024;
    L.smem = ctc_smem_bytes(a.beam, a.V, L.Vp, L.n_pad);
    size_t o = 0;
    L.off_vocab = o; o = align_up(o + a.V, 256);
    L.off_parent = o; o = align_up(o + sizeof(int) * (size_t)a.N * L.cap, 256);
    L.off_meta = o; o = align_up(o + sizeof(int) * (size_t)a.N * L.cap, 256);
    L.off_anc = o; o = align_up(o + sizeof(int) * (size_t)a.N * L.cap, 256);
    L.off_child = o; o = align_up(o + sizeof(int) * (size_t)a.N * L.cap * L.Vp, 256);
    L.state_stride = align_up((sizeof(CtaBeam<32>) > sizeof(WarpBeam<32>) ? sizeof(CtaBeam<32>) : sizeof(WarpBeam<32>)) + sizeof(int4), 256);
    L.off_state = o; o = align_up(o + L.state_stride * (size_t)a.N, 256);
    L.off_born = o; if (a.out_timesteps) o = align_up(o + sizeof(int) * (size_t)a.N * L.cap, 256);
    L.total = o;
    size_t q = 0;
    L.off_paths = q; q = align_up(q + (size_t)a.N * a.nbest * a.max_len, 256);
    L.off_lens = q; q = align_up(q + sizeof(int) * (size_t)a.N * a.nbest, 256);
    L.off_scores = q; q = align_up(q + sizeof(float) * (size_t)a.N * a.nbest, 256);
    L.off_counts = q; q = align_up(q + sizeof(int) * (size_t)a.N, 256);
    L.off_stats = q; q = align_up(q + 2 * sizeof(int) * (size_t)a.N, 256);
    L.off_ts = q; if (a.out_timesteps) q = align_up(q + sizeof(int) * (size_t)a.N * a.nbest * a.max_len, 256);
    L.out_bytes = q;
    return GASR_OK;
}

// Grow the decoder's workspaces for this problem size (allocation synchronises the device: the streaming pipeline
// calls this before its persistent kernels start).
int ctc_decode_reserve(gasr_ctx *ctx, const CtcArgs &a) {
    if (a.N == 0) return GASR_OK;
    CtcLayout L;
    ctc_layout(a, L);
    GASR_TRY(ws_reserve(ctx, ctx->ws_ctc, L.total));
    GASR_TRY(ws_reserve(ctx, ctx->ws_out, L.out_bytes));
    GASR_TRY(pinned_reserve(ctx, L.out_bytes));
    return GASR_OK;
}

// Upload the vocabulary ahead of the launch (CtcArgs::vocab_resident then skips it): a small pageable copy issued at
// launch time would queue behind whatever the copy engine is busy with.
int ctc_decode_upload_vocab(gasr_ctx *ctx, const CtcArgs &a, cudaStream_t st) {
    if (a.N == 0) return GASR_OK;
    CtcLayout L;
    ctc_layout(a, L);
    GASR_CUDA(cudaMemcpyAsync(static_cast<unsigned char *>(ctx->ws_ctc.ptr) + L.off_vocab, a.vocab_host, a.V, cudaMemcpyHostToDevice, st));
    return GASR_OK;
}

int ctc_decode_launch(gasr_ctx *ctx, const CtcArgs &a, cudaStream_t st) {
    GASR_CHECK(a.scores != nullptr && a.vocab_host != nullptr, "ctc_decode: null scores/vocab");
    GASR_CHECK(a.T >= 1 && a.N >= 0, "ctc_decode: T must be >= 1 and N >= 0 (T=%d N=%d)", a.T, a.N);
    GASR_CHECK(a.V >= 1 && a.V <= 255 && a.ld >= a.V, "ctc_decode: vocabulary size %d / ld %d unsupported", a.V, a.ld);
    GASR_CHECK(a.beam >= 1 && a.beam <= 1024, "ctc_decode: beam %d out of range", a.beam);
    GASR_CHECK(a.blank >= 0 && a.blank < a.V, "ctc_decode: blank id %d outside the vocabulary", a.blank);
    GASR_CHECK(a.nbest >= 1 && a.nbest <= a.beam && a.max_len >= 0, "ctc_decode: nbest/max_len out of range");
    GASR_CHECK(a.domain == GASR_DOMAIN_PROB || a.domain == GASR_DOMAIN_LOG, "ctc_decode: unknown score domain");
    for (int v = 0; v < a.V; v++) {
        GASR_CHECK(a.vocab_host[v] > 0, "ctc_decode: vocab chars must be in 1..127 (entry %d)", v);
        for (int u = 0; u < v; u++)
            GASR_CHECK(a.vocab_host[u] != a.vocab_host[v], "ctc_decode: duplicate vocab char at %d and %d", u, v);
    }
    if (a.N == 0) return GASR_OK;
    CtcLayout L;
    ctc_layout(a, L);
    bool general_rel = false;
    if (!(a.beam <= 32 && a.V <= 32) && L.smem + ctc_rel_bytes(a.beam) <= (size_t)ctx->max_smem_optin) {
        general_rel = true;                           // the prefix-relation matrix fits: O(1) raw-string tie-breaks
        L.smem += ctc_rel_bytes(a.beam);
    }
    if (!(a.beam <= 32 && a.V <= 32) && L.smem > (size_t)ctx->max_smem_optin) {
        set_error("ctc_decode: beam %d x vocab %d needs %zu B of shared memory (> %d)", a.beam, a.V, L.smem,
                  ctx->max_smem_optin);
        return GASR_ERR_UNSUPPORTED;
    }
    GASR_TRY(ws_reserve(ctx, ctx->ws_ctc, L.total));
    GASR_TRY(ws_reserve(ctx, ctx->ws_out, L.out_bytes));
    GASR_TRY(pinned_reserve(ctx, L.out_bytes));
    unsigned char *ws = static_cast<unsigned char *>(ctx->ws_ctc.ptr);
    unsigned char *wo = static_cast<unsigned char *>(ctx->ws_out.ptr);
    const int t0 = a.t0, t1 = a.t1 > 0 ? a.t1 : a.T;
    const bool fast = a.beam <= 32 && a.V <= 32;
    GASR_CHECK(t0 >= 0 && t0 < t1 && t1 <= a.T, "ctc_decode: bad frame range [%d, %d)", t0, t1);
    GASR_CHECK(fast || (t0 == 0 && t1 == a.T), "ctc_decode: time-chunked decoding needs beam <= 32 and vocab <= 32");
    if (t0 == 0 && !a.vocab_resident) GASR_CUDA(cudaMemcpyAsync(ws + L.off_vocab, a.vocab_host, a.V, cudaMemcpyHostToDevice, st));

    CtcParams p;
    p.lens = a.lens_dev;
    p.born = nullptr; p.out_ts = nullptr;
    p.scores = a.scores; p.T = a.T; p.N = a.N; p.V = a.V; p.ld = a.ld; p.beam = a.beam; p.blank = a.blank;
    p.frame_rows = a.frame_rows > 0 ? a.frame_rows : a.N;
    GASR_CHECK(p.frame_rows >= a.N, "ctc_decode: frame_rows %d < N %d", p.frame_rows, a.N);
    p.Vp = L.Vp; p.n_pad = L.n_pad; p.cap = L.cap; p.max_len = a.max_len; p.nbest = a.nbest;
    p.vocab = reinterpret_cast<const char *>(ws + L.off_vocab);
    p.parent = reinterpret_cast<int *>(ws + L.off_parent);
    p.meta = reinterpret_cast<int *>(ws + L.off_meta);
    p.anc = reinterpret_cast<int *>(ws + L.off_anc);
    p.child = reinterpret_cast<int *>(ws + L.off_child);
    p.out_paths = reinterpret_cast<char *>(wo + L.off_paths);
    p.out_lens = reinterpret_cast<int *>(wo + L.off_lens);
    p.out_scores = reinterpret_cast<float *>(wo + L.off_scores);
    p.out_counts = reinterpret_cast<int *>(wo + L.off_counts);
    p.out_stats = reinterpret_cast<int *>(wo + L.off_stats);
    if (a.out_timesteps) { p.born = reinterpret_cast<int *>(ws + L.off_born); p.out_ts = reinterpret_cast<int *>(wo + L.off_ts); }
    if (!fast) GASR_CUDA(cudaMemsetAsync(p.out_stats, 0, 2 * sizeof(int) * (size_t)a.N, st));
    p.t0 = t0; p.t1 = t1; p.use_rel = general_rel ? 1 : 0;
    p.state = ws + L.off_state; p.state_stride = L.state_stride;
    p.lp_ready = a.lp_ready; p.lp_need = a.lp_need; p.lp_fpb = a.lp_fpb > 0 ? a.lp_fpb : 1; p.error = a.error; p.abort = a.abort;
    const char force_kc = ctx->opt.ctc_kernel;
    const char *force_k = force_kc ? &force_kc : nullptr;
    const bool use_cta = fast && (a.lp_ready != nullptr || (force_k ? (force_k[0] == 'c' || force_k[0] == 'd') : a.N <= 2 * ctx->sm_count));
    GASR_CHECK(a.lp_ready == nullptr || (fast && t0 == 0 && t1 == a.T), "ctc_decode: streaming needs beam <= 32, vocab <= 32, whole sequence");
    {
        // probe cells of the prune lower bound: the (parent rank, score rank) pairs with the smallest (i+1)(j+1)
        p.n_cells = (use_cta && a.beam > 16) ? 128 : 64;
        // warp kernel, beam <= 16: one probe cell per lane.  The bound is looser (21 instead of 16 survivors on peaky / random
        // log-probabilities, 26 instead of 19 on flat ones) but the all-pairs counting of 64 cells was 17 % of the kernel's
        // instructions: decoder alone 24.8 -> 23.6 ms per 4096 utterances, cfg5 step 118.4 -> 115.9 ms (same-box A/B, 3 runs each)
        if (!use_cta && fast && a.beam <= 16 && ctx->opt.ctc_cells != 64) p.n_cells = 32;
        if (!use_cta && fast && ctx->opt.ctc_cells == 32) p.n_cells = 32;      // GASR_CTC_CELLS = 32 | 64 forces either
        int taken = 0;
        for (int prod = 1; taken < p.n_cells && prod <= a.beam * a.V; prod++)
            for (int i = 0; i < a.beam && taken < p.n_cells; i++) {
                if (prod % (i + 1)) continue;
                const int j = prod / (i + 1) - 1;
                if (j >= a.V) continue;
                p.cell_i[taken] = (unsigned char)i; p.cell_j[taken] = (unsigned char)j; taken++;
            }
        for (; taken < 128; taken++) { p.cell_i[taken] = 255; p.cell_j[taken] = 0; }
        memset(p.cellmap, 255, sizeof(p.cellmap));
        for (int c = 0; c < p.n_cells; c++)
            if (p.cell_i[c] < 32 && p.cell_j[c] < 32) p.cellmap[p.cell_i[c] * 32 + p.cell_j[c]] = (unsigned char)c;
    }

    if (t0 == 0) GASR_CUDA(cudaMemsetAsync(wo + L.off_paths, 0, (size_t)a.N * a.nbest * a.max_len, st));
    if (t0 == 0 && a.out_timesteps) GASR_CUDA(cudaMemsetAsync(wo + L.off_ts, 0, sizeof(int) * (size_t)a.N * a.nbest * a.max_len, st));
    const bool whole = t0 == 0 && t1 == a.T;
    if (use_cta && whole && !(force_k && force_k[0] == 'c')) {
        // latency path, second generation: 8 main warps + fetch warp + trie warp per utterance, whole sequence.  No
        // shared-memory padding: in the streaming pipeline these CTAs may share an SM with a GEMM CTA (whose few
        // threads leave the issue slots free); the recurrence CTAs fill their SMs' register file, so nothing lands there.
        const size_t pad = (size_t)ctx->opt.ctc_pad;
        const int mw = ctx->opt.ctc_mw;
#define GASR_CTA2(DOM, BM)                                                                          \
    do {                                                                                            \
        if (mw == 4) ctc_beam_cta2_kernel<DOM, BM, 4><<<a.N, 192, pad, st>>>(p);                    \
        else if (mw == 16) ctc_beam_cta2_kernel<DOM, BM, 16><<<a.N, 576, pad, st>>>(p);             \
        else ctc_beam_cta2_kernel<DOM, BM, 8><<<a.N, 320, pad, st>>>(p);                            \
    } while (0)
        if (a.domain == GASR_DOMAIN_LOG) {
            if (a.beam <= 16) GASR_CTA2(1, 16); else GASR_CTA2(1, 32);
        } else {
            if (a.beam <= 16) GASR_CTA2(0, 16); else GASR_CTA2(0, 32);
        }
#undef GASR_CTA2
    } else if (use_cta) {
        // latency path: a 128-thread CTA per utterance (state parked per utterance needs sizeof(CtaBeam) <= stride)
        // 20 KB of (unused) dynamic shared memory keeps these CTAs off the SMs whose shared memory is filled by a
        // recurrence CTA (208 KB) when the pipeline runs both at once
        const size_t pad = 20480;
        if (a.domain == GASR_DOMAIN_LOG) {
            if (a.beam <= 16) ctc_beam_cta_kernel<1, 16><<<a.N, 128, pad, st>>>(p); else ctc_beam_cta_kernel<1, 32><<<a.N, 128, pad, st>>>(p);
        } else {
            if (a.beam <= 16) ctc_beam_cta_kernel<0, 16><<<a.N, 128, pad, st>>>(p); else ctc_beam_cta_kernel<0, 32><<<a.N, 128, pad, st>>>(p);
        }
    } else if (fast) {
        // warp-per-utterance fast path; few warps per CTA when utterances are scarce (latency), 8 when plentiful
        int W = ceil_div(a.N, ctx->sm_count);
        if (W > 8) W = 8;
        if (a.warps_per_cta >= 1 && a.warps_per_cta <= 8) W = a.warps_per_cta;
        const int blocks = ceil_div(a.N, W);
#define GASR_CTCW_LAUNCH(DOM, BM)                                                                                \
    do {                                                                                                          \
        GASR_CUDA(cudaFuncSetAttribute(ctc_beam_warp_kernel<DOM, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)(sizeof(WarpBeam<BM>) * 8)));                                         \
        ctc_beam_warp_kernel<DOM, BM><<<blocks, W * 32, sizeof(WarpBeam<BM>) * W, st>>>(p);                       \
    } while (0)
        if (a.domain == GASR_DOMAIN_LOG) {
            if (a.beam <= 16) GASR_CTCW_LAUNCH(1, 16); else GASR_CTCW_LAUNCH(1, 32);
        } else {
            if (a.beam <= 16) GASR_CTCW_LAUNCH(0, 16); else GASR_CTCW_LAUNCH(0, 32);
        }
#undef GASR_CTCW_LAUNCH
    } else {
#define GASR_CTC_LAUNCH(DOM, MT)                                                                                  \
    do {                                                                                                          \
        GASR_CUDA(cudaFuncSetAttribute(ctc_beam_kernel<DOM, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                       (int)L.smem));                                                             \
        ctc_beam_kernel<DOM, MT><<<a.N, L.threads, L.smem, st>>>(p);                                              \
    } while (0)
        if (a.domain == GASR_DOMAIN_LOG) {
            if (L.threads <= 256) GASR_CTC_LAUNCH(1, 256); else GASR_CTC_LAUNCH(1, 1024);
        } else {
            if (L.threads <= 256) GASR_CTC_LAUNCH(0, 256); else GASR_CTC_LAUNCH(0, 1024);
        }
#undef GASR_CTC_LAUNCH
    }
    GASR_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (t1 == a.T) GASR_CUDA(cudaMemcpyAsync(ctx->pinned_out, wo, L.out_bytes, cudaMemcpyDeviceToHost, st));
    return GASR_OK;
}

int ctc_decode_finish(gasr_ctx *ctx, const CtcArgs &a) {
    if (a.N == 0) return GASR_OK;
    CtcLayout L;
    ctc_layout(a, L);
    const unsigned char *h = static_cast<const unsigned char *>(ctx->pinned_out);
    memcpy(a.out_paths, h + L.off_paths, (size_t)a.N * a.nbest * a.max_len);
    memcpy(a.out_lens, h + L.off_lens, sizeof(int) * (size_t)a.N * a.nbest);
    memcpy(a.out_scores, h + L.off_scores, sizeof(float) * (size_t)a.N * a.nbest);
    if (a.out_counts) memcpy(a.out_counts, h + L.off_counts, sizeof(int) * (size_t)a.N);
    if (a.out_timesteps) memcpy(a.out_timesteps, h + L.off_ts, sizeof(int) * (size_t)a.N * a.nbest * a.max_len);
    {
        const int *stt = reinterpret_cast<const int *>(h + L.off_stats);
        ctx->ctc_fallback_frames = 0; ctx->ctc_survivors = 0;
        for (int n = 0; n < a.N; n++) { ctx->ctc_fallback_frames += stt[2 * n]; ctx->ctc_survivors += stt[2 * n + 1]; }
    }
    int status = GASR_OK;
    for (size_t i = 0; i < (size_t)a.N * a.nbest; i++)
        if (a.out_lens[i] > a.max_len) status = GASR_ERR_TRUNCATED;
    if (status != GASR_OK) set_error("ctc_decode: at least one path is longer than max_len=%d", a.max_len);
    return status;
}

}  // namespace gasr
