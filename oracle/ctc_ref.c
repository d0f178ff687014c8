/*
 * oracle/ctc_ref.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the reference CTC beam search ("CTC-REF").  It follows the
 * reference GPU decoder kernel by kernel, on explicit path strings:
 *
 *   t = 0 initial states ........ CTCBeamSearch.cu:337-364 (kernelInitialPath), :366-401
 *   candidate extension rules ... CTCBeamSearch.cu:404-458 (kernelGenNextPaths)
 *   31-hash of path[0:len] ...... CTCBeamSearch.cu:50-58   (genHashCode)
 *   raw C-string order .......... CTCBeamSearch.cu:137-147 (operator<), :159 (stable sort)
 *   merge of equal neighbours ... CTCBeamSearch.cu:460-489 (kernelTestDifferentPaths, kernelMergeSamePaths)
 *   stable desc sort + prune .... CTCBeamSearch.cu:174-196 (batchSortbyKey), :103-112
 *   result = rank-0 state ....... CTCBeamSearch.cu:290-298
 *
 * Intended-semantics fixes w.r.t. the literal reference: unused beam slots never
 * take part (the reference faults for beam > vocab), no 256-byte path cap, no reliance on
 * cudaMalloc returning zeroed memory, and the racy atomicAdd merge order is fixed to the
 * sorted (raw string, candidate index) order.
 *
 * merge_mode 0 ("identity"): candidates merge iff path[0:len] is the same string.
 * merge_mode 1 ("hash31")  : literal reference rule -- sorted neighbours merge iff their
 *                            Java-style 31-hashes are equal (collisions merge different strings).
 *
 * domain 0: scores are probabilities, combine = fp32 multiply, merge = fp32 add (reference).
 * domain 1: scores are log-probabilities, combine = fp32 add, merge = oracle_logaddexp (logadd_ref.h).
 *
 * Path storage: a candidate is held as (parent state, kept prefix length, optional extra char)
 * instead of the reference's 256-byte memcpy (CTCBeamSearch.cu:428); the string it denotes is
 * exactly the reference's NUL-terminated path buffer.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "logadd_ref.h"

typedef struct {
    float score;
    int len;     /* BeamState::len                                    */
    int rawlen;  /* strlen(BeamState::path); > len only after the final-frame strip */
    char *raw;   /* BeamState::path (owned, capacity T + 2)           */
} state_t;

typedef struct {
    int parent;  /* rank of the parent state          */
    int base;    /* chars kept from the parent's raw  */
    char extra;  /* appended char or 0                */
    int len;     /* len after the optional final strip */
    float score;
} cand_t;

typedef struct {
    const state_t *st;
    const cand_t *cd;
} cmp_ctx_t;

static inline int cand_rawlen(const cand_t *c) { return c->base + (c->extra ? 1 : 0); }

static inline char cand_at(const cmp_ctx_t *cx, const cand_t *c, int p) {
    if (p < c->base) return cx->st[c->parent].raw[p];
    if (p == c->base) return c->extra; /* 0 when there is no extra char */
    return 0;
}

/* operator< of CTCBeamSearch.cu:137-147 on the two NUL-terminated raw strings (plain `char` compare). */
static int cand_less(const cmp_ctx_t *cx, int ia, int ib) {
    const cand_t *a = &cx->cd[ia], *b = &cx->cd[ib];
    int common = a->base < b->base ? a->base : b->base;
    const char *ra = cx->st[a->parent].raw, *rb = cx->st[b->parent].raw;
    int i = 0;
    if (ra != rb) {
        while (i < common && ra[i] == rb[i]) i++;
    } else {
        i = common;
    }
    for (;; i++) {
        char ca = cand_at(cx, a, i), cb = cand_at(cx, b, i);
        if (ca && cb && ca == cb) continue;
        return ca < cb;
    }
}

/* identity = path[0:len] */
static int cand_same_identity(const cmp_ctx_t *cx, int ia, int ib) {
    const cand_t *a = &cx->cd[ia], *b = &cx->cd[ib];
    if (a->len != b->len) return 0;
    for (int i = 0; i < a->len; i++)
        if (cand_at(cx, a, i) != cand_at(cx, b, i)) return 0;
    return 1;
}

/* does raw(ib) start with path[0:len] of ia ? */
static int cand_has_prefix(const cmp_ctx_t *cx, int ia, int ib) {
    const cand_t *a = &cx->cd[ia], *b = &cx->cd[ib];
    if (cand_rawlen(b) < a->len) return 0;
    for (int i = 0; i < a->len; i++)
        if (cand_at(cx, a, i) != cand_at(cx, b, i)) return 0;
    return 1;
}

/* genHashCode, CTCBeamSearch.cu:50-58: int32 wrap-around, signed chars */
static int32_t cand_hash31(const cmp_ctx_t *cx, int ia) {
    const cand_t *a = &cx->cd[ia];
    uint32_t h = 0;
    for (int i = 0; i < a->len; i++) h = 31u * h + (uint32_t)(int32_t)cand_at(cx, a, i);
    return (int32_t)h;
}

/* stable bottom-up merge sort of idx[0:n] by cand_less (thrust::stable_sort_by_key, .cu:159) */
static void sort_by_string(const cmp_ctx_t *cx, int *idx, int *tmp, int n) {
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                if (cand_less(cx, idx[j], idx[i])) tmp[k++] = idx[j++];
                else tmp[k++] = idx[i++];
            }
            while (i < mid) tmp[k++] = idx[i++];
            while (j < hi) tmp[k++] = idx[j++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int));
    }
}

/* stable descending sort of idx by key[idx] (thrust::greater<float>, .cu:184) */
static void sort_by_score_desc(const float *key, int *idx, int *tmp, int n) {
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                if (key[idx[j]] > key[idx[i]]) tmp[k++] = idx[j++];
                else tmp[k++] = idx[i++];
            }
            while (i < mid) tmp[k++] = idx[i++];
            while (j < hi) tmp[k++] = idx[j++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int));
    }
}

static inline float combine(int domain, float s, float p) { return domain ? s + p : s * p; }
static inline float merge2(int domain, float a, float b) { return domain ? oracle_logaddexp(a, b) : a + b; }

typedef struct {
    int n_beams;    /* states kept after the last frame */
    /* optional trace sink: per frame, per rank: score + raw string */
} utt_out_t;

/*
 * Decode one utterance.  S points at row (t=0, this utterance); consecutive frames are
 * `stride` floats apart (time-major [T, N, V]: stride = N * ld).
 * Outputs the kept beams after the last frame, best first:
 *   out_paths[r * max_len ...], out_lens[r], out_scores[r]  for r < nbest_cap
 * trace (optional): trace_scores[t * beam + r], trace_lens[t * beam + r], trace_paths[(t*beam + r) * max_len]
 * (raw strings incl. a trailing blank) and trace_counts[t].
 */
static int decode_one(const float *S, long stride, int T, int V, const char *vocab, int blank, int beam,
                      int domain, int merge_mode, int max_len, int nbest_cap, char *out_paths, int *out_lens,
                      float *out_scores, int *out_count, float *trace_scores, int *trace_lens,
                      char *trace_paths, int *trace_counts) {
    const char blank_ch = vocab[blank];
    int cap = beam * V > V ? beam * V : V;
    state_t *st = (state_t *)calloc((size_t)beam, sizeof(state_t));
    state_t *nx = (state_t *)calloc((size_t)beam, sizeof(state_t));
    for (int i = 0; i < beam; i++) {
        st[i].raw = (char *)calloc((size_t)T + 2, 1);
        nx[i].raw = (char *)calloc((size_t)T + 2, 1);
    }
    cand_t *cd = (cand_t *)malloc((size_t)cap * sizeof(cand_t));
    int *idx = (int *)malloc((size_t)cap * sizeof(int));
    int *tmp = (int *)malloc((size_t)cap * sizeof(int));
    int *grp_first = (int *)malloc((size_t)cap * sizeof(int));
    float *grp_score = (float *)malloc((size_t)cap * sizeof(float));
    int *gidx = (int *)malloc((size_t)cap * sizeof(int));
    int kept;

    /* ---- t = 0: kernelInitialPath + batchSortbyKey + prune (.cu:337-401) ---- */
    {
        for (int v = 0; v < V; v++) { grp_score[v] = S[v]; idx[v] = v; }
        sort_by_score_desc(grp_score, idx, tmp, V);
        kept = beam < V ? beam : V;
        for (int r = 0; r < kept; r++) {
            st[r].raw[0] = vocab[idx[r]];
            st[r].raw[1] = 0;
            st[r].len = st[r].rawlen = 1;
            st[r].score = grp_score[idx[r]];
        }
    }
    if (trace_counts) {
        trace_counts[0] = kept;
        for (int r = 0; r < kept; r++) {
            trace_scores[r] = st[r].score;
            trace_lens[r] = st[r].rawlen;
            int n = st[r].rawlen < max_len ? st[r].rawlen : max_len;
            memcpy(trace_paths + (size_t)r * max_len, st[r].raw, (size_t)n);
        }
    }

    for (int t = 1; t < T; t++) {
        const float *P = S + (long)t * stride;
        const int last_step = (t == T - 1);
        int n = 0;
        /* ---- kernelGenNextPaths (.cu:404-458): candidate index = rank * V + v ---- */
        for (int r = 0; r < kept; r++) {
            const state_t *s = &st[r];
            const char last = s->raw[s->len - 1];
            for (int v = 0; v < V; v++) {
                cand_t *c = &cd[n];
                c->parent = r;
                c->score = combine(domain, s->score, P[v]);
                if (v == blank) {
                    if (last == blank_ch) { c->base = s->len; c->extra = 0; }
                    else { c->base = s->len; c->extra = blank_ch; }
                } else {
                    if (last == blank_ch) { c->base = s->len - 1; c->extra = vocab[v]; }
                    else if (last == vocab[v]) { c->base = s->len; c->extra = 0; }
                    else { c->base = s->len; c->extra = vocab[v]; }
                }
                c->len = cand_rawlen(c);
                if (last_step) {
                    /* strip one trailing blank from len only; the raw string keeps it (.cu:452-456) */
                    char tail = c->extra ? c->extra : s->raw[c->base - 1];
                    if (tail == blank_ch) c->len -= 1;
                }
                idx[n] = n;
                n++;
            }
        }
        cmp_ctx_t cx = {st, cd};
        /* ---- batchSortbyStr (.cu:149-172): stable ascending raw C-string order ---- */
        sort_by_string(&cx, idx, tmp, n);
        /* ---- kernelTestDifferentPaths / kernelMergeSamePaths (.cu:460-489) ---- */
        int G = 0;
        if (merge_mode == 1) {
            int32_t prev_hash = 0;
            for (int k = 0; k < n; k++) {
                int32_t h = cand_hash31(&cx, idx[k]);
                if (k == 0 || h != prev_hash) {
                    grp_first[G] = idx[k];
                    grp_score[G] = cd[idx[k]].score;
                    G++;
                } else {
                    grp_score[G - 1] = merge2(domain, grp_score[G - 1], cd[idx[k]].score);
                }
                prev_hash = h;
            }
        } else {
            for (int k = 0; k < n; k++) {
                int c = idx[k], g = -1;
                if (G > 0 && cand_same_identity(&cx, c, grp_first[G - 1])) {
                    g = G - 1;
                } else if (last_step && cd[c].len != cand_rawlen(&cd[c])) {
                    /* "X$" on the final frame: its group "X" lies earlier in the sorted order, and
                       every string between X and X$ has X as a prefix */
                    for (int q = G - 1; q >= 0; q--) {
                        if (cand_same_identity(&cx, c, grp_first[q])) { g = q; break; }
                        if (!cand_has_prefix(&cx, c, grp_first[q])) break;
                    }
                }
                if (g < 0) {
                    grp_first[G] = c;
                    grp_score[G] = cd[c].score;
                    G++;
                } else {
                    grp_score[g] = merge2(domain, grp_score[g], cd[c].score);
                }
            }
        }
        /* ---- batchSortbyKey<float> + prune (.cu:174-196, :103-112) ---- */
        for (int g = 0; g < G; g++) gidx[g] = g;
        sort_by_score_desc(grp_score, gidx, tmp, G);
        int nk = beam < G ? beam : G;
        for (int r = 0; r < nk; r++) {
            const cand_t *c = &cd[grp_first[gidx[r]]];
            state_t *d = &nx[r];
            memcpy(d->raw, st[c->parent].raw, (size_t)c->base);
            int rl = c->base;
            if (c->extra) d->raw[rl++] = c->extra;
            d->raw[rl] = 0;
            d->rawlen = rl;
            d->len = c->len;
            d->score = grp_score[gidx[r]];
        }
        state_t *sw = st; st = nx; nx = sw;
        kept = nk;
        if (trace_counts) {
            trace_counts[t] = kept;
            for (int r = 0; r < kept; r++) {
                trace_scores[(size_t)t * beam + r] = st[r].score;
                trace_lens[(size_t)t * beam + r] = st[r].rawlen;
                int m = st[r].rawlen < max_len ? st[r].rawlen : max_len;
                memcpy(trace_paths + ((size_t)t * beam + r) * max_len, st[r].raw, (size_t)m);
            }
        }
    }

    /* ---- result fetch (.cu:290-298): path[0:len] and prob of each kept state, best first ---- */
    int nout = kept < nbest_cap ? kept : nbest_cap;
    for (int r = 0; r < nout; r++) {
        int m = st[r].len < max_len ? st[r].len : max_len;
        memcpy(out_paths + (size_t)r * max_len, st[r].raw, (size_t)m);
        out_lens[r] = st[r].len;
        out_scores[r] = st[r].score;
    }
    if (out_count) *out_count = kept;

    for (int i = 0; i < beam; i++) { free(st[i].raw); free(nx[i].raw); }
    free(st); free(nx); free(cd); free(idx); free(tmp); free(grp_first); free(grp_score); free(gidx);
    return 0;
}

typedef struct {
    const float *S; int T, N, V, ld; const char *vocab; int blank, beam, domain, merge_mode, max_len, nbest;
    char *out_paths; int *out_lens; float *out_scores; int *out_counts;
    int next; pthread_mutex_t mu;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int n = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (n >= j->N) break;
        decode_one(j->S + (long)n * j->ld, (long)j->N * j->ld, j->T, j->V, j->vocab, j->blank, j->beam, j->domain,
                   j->merge_mode, j->max_len, j->nbest, j->out_paths + (size_t)n * j->nbest * j->max_len,
                   j->out_lens + (size_t)n * j->nbest, j->out_scores + (size_t)n * j->nbest,
                   j->out_counts ? j->out_counts + n : NULL, NULL, NULL, NULL, NULL);
    }
    return NULL;
}

/*
 * Batched decode of scores S[T, N, ld] (time-major, row t*N+n, first V columns used), like
 * CTCBeamSearch::decode (CTCBeamSearch.cu:262-312).  nbest = 1 reproduces the reference's
 * top-1 output; nbest > 1 also returns the lower-ranked kept states.
 * out_paths: [N, nbest, max_len] bytes (not NUL-terminated), out_lens/out_scores: [N, nbest],
 * out_counts: [N] kept states (may be NULL).
 */
int oracle_ctc_decode(const float *S, int T, int N, int V, int ld, const char *vocab, int blank, int beam,
                      int domain, int merge_mode, int max_len, int nbest, char *out_paths, int *out_lens,
                      float *out_scores, int *out_counts, int nthreads) {
    if (T < 1 || N < 0 || V < 1 || beam < 1 || blank < 0 || blank >= V || nbest < 1 || ld < V) return 1;
    job_t j = {S, T, N, V, ld, vocab, blank, beam, domain, merge_mode, max_len, nbest,
               out_paths, out_lens, out_scores, out_counts, 0, PTHREAD_MUTEX_INITIALIZER};
    memset(out_paths, 0, (size_t)N * nbest * max_len);
    memset(out_lens, 0, (size_t)N * nbest * sizeof(int));
    for (long i = 0; i < (long)N * nbest; i++) out_scores[i] = 0.0f;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > N) nthreads = N > 0 ? N : 1;
    if (nthreads == 1) { worker(&j); return 0; }
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, worker, &j);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    free(th);
    return 0;
}

/* Single utterance with the per-frame kept beams recorded (raw strings, trailing blank included). */
int oracle_ctc_trace(const float *S, int T, int V, const char *vocab, int blank, int beam, int domain,
                     int merge_mode, int max_len, float *trace_scores, int *trace_lens, char *trace_paths,
                     int *trace_counts, char *out_path, int *out_len, float *out_score) {
    if (T < 1 || V < 1 || beam < 1 || blank < 0 || blank >= V) return 1;
    memset(trace_paths, 0, (size_t)T * beam * max_len);
    return decode_one(S, V, T, V, vocab, blank, beam, domain, merge_mode, max_len, 1, out_path, out_len,
                      out_score, NULL, trace_scores, trace_lens, trace_paths, trace_counts);
}

float oracle_logaddexp_f32(float a, float b) { return oracle_logaddexp(a, b); }
