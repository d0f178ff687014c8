/*
 * oracle/am_ref.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement, in fp32, of the reference acoustic-model modules:
 *   matrixMul  z = x*y row-major ............ cuMatrix.cpp:33-70
 *   matrixAdd  z = x + lambda*y ............. cuMatrix.cpp:147-168
 *   Linear     ReLU(x*W + b), W[in,out] ..... Linear.cu:3-10, :42-49
 *   RNN_Cell   tanh(x*W_ih + h*W_hh + (b_hh + b_ih)), bias pair summed first ... RNN_Cell.cu:5-13, :65-74
 *   RNN        L stacked cells, time-major input [T*N, in], h_0 = 0, returns every layer's
 *              hidden sequence [T*N, H] ..... RNN.cu:9-30, RNN.h:8-21
 *   log-softmax over the vocabulary ......... baseline/model.py:49 (absent from the C++ code)
 *   GRU (gate order r,z,n; b_hn inside the reset product) and bidirectional stacking follow the
 *   torch.nn.GRU equations named by BASELINE.json cfg3; the reference itself has no GRU.
 * Weight layout is the reference's: W[in, out] row-major (the transpose of torch's).
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* z[m,n] = x[m,k] * y[k,n] (+ z if accumulate) -- k-outer so the inner loop streams rows of y */
static void gemm_rm(const float *x, const float *y, float *z, int m, int k, int n, int accumulate) {
    for (int i = 0; i < m; i++) {
        float *zi = z + (size_t)i * n;
        if (!accumulate) memset(zi, 0, (size_t)n * sizeof(float));
        const float *xi = x + (size_t)i * k;
        for (int p = 0; p < k; p++) {
            const float a = xi[p];
            const float *yp = y + (size_t)p * n;
            for (int j = 0; j < n; j++) zi[j] += a * yp[j];
        }
    }
}

void oracle_matmul(const float *x, const float *y, float *z, int m, int k, int n) { gemm_rm(x, y, z, m, k, n, 0); }

void oracle_matadd(const float *x, const float *y, float *z, int rows, int cols, float lambda) {
    for (size_t i = 0; i < (size_t)rows * cols; i++) z[i] = x[i] + lambda * y[i];
}

void oracle_log_softmax(const float *x, float *y, int rows, int cols) {
    for (int i = 0; i < rows; i++) {
        const float *xi = x + (size_t)i * cols;
        float *yi = y + (size_t)i * cols;
        float mx = xi[0];
        for (int j = 1; j < cols; j++) mx = xi[j] > mx ? xi[j] : mx;
        double s = 0.0;
        for (int j = 0; j < cols; j++) s += exp((double)xi[j] - (double)mx);
        float lse = (float)log(s);
        for (int j = 0; j < cols; j++) yi[j] = (xi[j] - mx) - lse;
    }
}

/* act: 0 = none, 1 = ReLU (the reference's Linear), 2 = log-softmax */
void oracle_linear(const float *x, int rows, int in, int out, const float *W, const float *b, int act, float *y) {
    gemm_rm(x, W, y, rows, in, out, 0);
    for (int i = 0; i < rows; i++) {
        float *yi = y + (size_t)i * out;
        for (int j = 0; j < out; j++) {
            yi[j] += b ? b[j] : 0.0f;                 /* data[i] += bias[j]      (Linear.cu:8) */
            if (act == 1 && yi[j] < 0) yi[j] = 0.0f;  /* if (data[i] < 0) ... 0  (Linear.cu:9) */
        }
    }
    if (act == 2) oracle_log_softmax(y, y, rows, out);
}

typedef struct {
    const float *x; int T, N, in, H, L;
    const float *const *w_ih, *const *w_hh, *const *b_ih, *const *b_hh;
    float *const *hiddens; int n0, n1;
} rnn_job_t;

/* utterances [n0, n1): all layers, all timesteps (rows of different utterances never interact) */
static void *rnn_worker(void *arg) {
    rnn_job_t *j = (rnn_job_t *)arg;
    const int T = j->T, N = j->N, H = j->H, nb = j->n1 - j->n0;
    if (nb <= 0) return NULL;
    float *ih = (float *)malloc((size_t)nb * H * sizeof(float));
    float *hh = (float *)malloc((size_t)nb * H * sizeof(float));
    float *h0 = (float *)calloc((size_t)nb * H, sizeof(float));
    for (int l = 0; l < j->L; l++) {
        const int in = l == 0 ? j->in : H;
        const float *src = l == 0 ? j->x : j->hiddens[l - 1];
        float *dst = j->hiddens[l];
        for (int t = 0; t < T; t++) {
            const float *xt = src + ((size_t)t * N + j->n0) * in;
            const float *hp = t == 0 ? h0 : dst + ((size_t)(t - 1) * N + j->n0) * H;
            float *ht = dst + ((size_t)t * N + j->n0) * H;
            gemm_rm(xt, j->w_ih[l], ih, nb, in, H, 0);          /* matrixMul(inputs, w_ih, ih_outputs)     */
            gemm_rm(hp, j->w_hh[l], hh, nb, H, H, 0);           /* matrixMul(pre_hidden, w_hh, hh_outputs) */
            for (int i = 0; i < nb; i++)
                for (int c = 0; c < H; c++) {
                    float v = ih[(size_t)i * H + c] + hh[(size_t)i * H + c];   /* matrixAdd(..., 1) */
                    v += j->b_hh[l][c] + j->b_ih[l][c];                        /* Tanh kernel       */
                    ht[(size_t)i * H + c] = tanhf(v);
                }
        }
    }
    free(ih); free(hh); free(h0);
    return NULL;
}

/*
 * RNN::forward.  x: [T*N, in] time-major.  hiddens[l]: caller-allocated [T*N, H] for every layer.
 * w_ih[l]: [in_l, H], w_hh[l]: [H, H], b_ih[l], b_hh[l]: [H].
 */
void oracle_rnn_forward(const float *x, int T, int N, int in, int H, int L, const float *const *w_ih,
                        const float *const *w_hh, const float *const *b_ih, const float *const *b_hh,
                        float *const *hiddens, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > N) nthreads = N > 0 ? N : 1;
    rnn_job_t *jobs = (rnn_job_t *)malloc((size_t)nthreads * sizeof(rnn_job_t));
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    for (int i = 0; i < nthreads; i++) {
        rnn_job_t jb = {x, T, N, in, H, L, w_ih, w_hh, b_ih, b_hh, hiddens,
                        (int)((long)N * i / nthreads), (int)((long)N * (i + 1) / nthreads)};
        jobs[i] = jb;
    }
    if (nthreads == 1) rnn_worker(&jobs[0]);
    else {
        for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, rnn_worker, &jobs[i]);
        for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    }
    free(jobs); free(th);
}

static inline float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

/*
 * One GRU direction of one layer over all timesteps (torch.nn.GRU equations, weights in [in, 3H] /
 * [H, 3H] reference layout, gate order r, z, n):
 *   r = s(x W_ir + b_ir + h W_hr + b_hr); z = s(x W_iz + b_iz + h W_hz + b_hz)
 *   n = tanh(x W_in + b_in + r * (h W_hn + b_hn)); h' = (1 - z) * n + z * h
 * out: [T*N, out_ld] written at column offset col0.
 */
static void gru_direction(const float *x, int T, int N, int in, int H, const float *w_ih, const float *w_hh,
                          const float *b_ih, const float *b_hh, int reverse, float *out, int out_ld, int col0) {
    float *gi = (float *)malloc((size_t)N * 3 * H * sizeof(float));
    float *gh = (float *)malloc((size_t)N * 3 * H * sizeof(float));
    float *h = (float *)calloc((size_t)N * H, sizeof(float));
    for (int s = 0; s < T; s++) {
        const int t = reverse ? T - 1 - s : s;
        gemm_rm(x + (size_t)t * N * in, w_ih, gi, N, in, 3 * H, 0);
        gemm_rm(h, w_hh, gh, N, H, 3 * H, 0);
        for (int i = 0; i < N; i++) {
            const float *a = gi + (size_t)i * 3 * H, *g = gh + (size_t)i * 3 * H;
            float *hi = h + (size_t)i * H;
            float *o = out + ((size_t)t * N + i) * out_ld + col0;
            for (int c = 0; c < H; c++) {
                float r = sigmoidf_((a[c] + b_ih[c]) + (g[c] + b_hh[c]));
                float z = sigmoidf_((a[H + c] + b_ih[H + c]) + (g[H + c] + b_hh[H + c]));
                float nn = tanhf((a[2 * H + c] + b_ih[2 * H + c]) + r * (g[2 * H + c] + b_hh[2 * H + c]));
                float hv = (1.0f - z) * nn + z * hi[c];
                hi[c] = hv;
                o[c] = hv;
            }
        }
    }
    free(gi); free(gh); free(h);
}

typedef struct {
    const float *x; int T, N, in, H; const float *w_ih, *w_hh, *b_ih, *b_hh; int reverse; float *out; int out_ld, col0;
} gru_job_t;
static void *gru_worker(void *arg) {
    gru_job_t *j = (gru_job_t *)arg;
    gru_direction(j->x, j->T, j->N, j->in, j->H, j->w_ih, j->w_hh, j->b_ih, j->b_hh, j->reverse, j->out, j->out_ld, j->col0);
    return NULL;
}

/*
 * Stacked (optionally bidirectional) GRU.  Parameter arrays are indexed [l * D + d] (D = 1 or 2).
 * hiddens[l]: [T*N, D*H]; layer l > 0 consumes hiddens[l-1] (forward | backward concatenated).
 */
void oracle_gru_forward(const float *x, int T, int N, int in, int H, int L, int bidir, const float *const *w_ih,
                        const float *const *w_hh, const float *const *b_ih, const float *const *b_hh,
                        float *const *hiddens) {
    const int D = bidir ? 2 : 1;
    for (int l = 0; l < L; l++) {
        const int in_l = l == 0 ? in : D * H;
        const float *src = l == 0 ? x : hiddens[l - 1];
        gru_job_t jb[2];
        pthread_t th[2];
        for (int d = 0; d < D; d++) {
            gru_job_t j = {src, T, N, in_l, H, w_ih[l * D + d], w_hh[l * D + d], b_ih[l * D + d], b_hh[l * D + d],
                           d, hiddens[l], D * H, d * H};
            jb[d] = j;
            pthread_create(&th[d], NULL, gru_worker, &jb[d]);
        }
        for (int d = 0; d < D; d++) pthread_join(th[d], NULL);
    }
}
