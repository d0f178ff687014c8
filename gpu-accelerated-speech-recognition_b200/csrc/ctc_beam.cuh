// ctc_beam.cuh -- what the decoder kernels share: the deterministic log-add-exp, the kernel parameter block, the raw-string
// order helpers, the per-utterance beam records of the warp / CTA kernels, and the launchers of the four kernels.
//
//   ctc_beam.cu           host side: workspace layout, argument checks, DISPATCH (see ctc_decode_launch), result unpacking
//   ctc_beam_general.cu   any beam <= 1024 / vocabulary <= 255: one CTA per utterance, bitonic / radix-select prune
//   ctc_beam_warp.cu      beam <= 32, vocabulary <= 32, many utterances: one WARP per utterance, resumable per time chunk
//   ctc_beam_cta.cu       beam <= 32, vocabulary <= 32, few utterances: one 128-thread CTA per utterance, resumable per time chunk
//   ctc_beam_cta2.cu      the same for whole sequences: main warps + fetch warp + trie warp (lowest latency; streaming pipeline)
// All four implement the same CTC-REF semantics bit for bit (tests/test_gpu_parity.py, test_gpu_wide.py, test_gpu_lengths.py).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

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
static __device__ __noinline__ bool raw_less(const int *__restrict__ parent, const int *__restrict__ meta, const char *vocab, int na,
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
static __device__ int trie_char_at(const int *__restrict__ parent, const int *__restrict__ meta, int nd, int pos);   // below
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
// Same CTC-REF semantics, bit for bit, as ctc_beam_kernel (ctc_beam_general.cu).
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
static __device__ __noinline__ int trie_char_at(const int *__restrict__ parent, const int *__restrict__ meta, int nd, int pos) {
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

// launchers (one per kernel translation unit); p is complete, the caller has reserved the workspaces
int ctc_launch_general(const CtcParams &p, int domain, int utterances, int threads, size_t smem, cudaStream_t st);
int ctc_launch_warp(const CtcParams &p, int domain, int blocks, int warps_per_cta, cudaStream_t st);
int ctc_launch_cta(const CtcParams &p, int domain, int utterances, size_t pad_smem, cudaStream_t st);
int ctc_launch_cta2(const CtcParams &p, int domain, int utterances, int main_warps, size_t pad_smem, cudaStream_t st);

}  // namespace gasr
