/*
 * oracle/ref_driver.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" entry points around the UNMODIFIED reference classes (compiled from /root/reference by
 * oracle/Makefile `ref` into oracle/_ref/libgasr_ref.so).  Used on the GPU box to pin the CPU oracle
 * against the reference's own outputs inside the envelope where the reference is defined
 * (beam <= vocab, short T before fp32 underflow, path length < 256).
 *
 * The reference decoder assumes cudaMalloc returns zeroed memory ("Assume: cudaMalloc initialize
 * memory to zero", CTCBeamSearch.cu:520); ref_prezero() makes that true for the allocations that
 * follow by zeroing and releasing a block of the same total size first.
 */
#include <string>
#include <utility>
#include <vector>

#include "cuMatrix.h"
#include "Linear.h"
#include "RNN.h"
#include "CTCBeamSearch.h"

static void ref_prezero(size_t bytes) {
    void *p = NULL;
    if (cudaMalloc(&p, bytes) == cudaSuccess) {
        cudaMemset(p, 0, bytes);
        cudaDeviceSynchronize();
        cudaFree(p);
    }
}

extern "C" int ref_linear_forward(const float *x, int rows, int in, int out, const float *W, const float *b,
                                  float *y) {
    cuMatrix<float> *inp = new cuMatrix<float>((float *)x, rows, in, 1);
    inp->toGpu();
    Linear *lin = new Linear(rows, in, out);
    lin->initParams((float *)W, (float *)b);
    cuMatrix<float> *o = lin->forward(inp);
    cudaDeviceSynchronize();
    o->toCpu();
    memcpy(y, o->getHost(), sizeof(float) * rows * out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int ref_rnn_forward(const float *x, int T, int N, int in, int H, int L, const float *const *w_ih,
                               const float *const *w_hh, const float *const *b_ih, const float *const *b_hh,
                               float *y) {
    cuMatrix<float> *inp = new cuMatrix<float>((float *)x, T * N, in, 1);
    inp->toGpu();
    RNN *rnn = new RNN(N, in, H, T, L);
    for (int l = 0; l < L; l++)
        rnn->rnn_cell[l]->initParams((float *)w_ih[l], (float *)w_hh[l], (float *)b_ih[l], (float *)b_hh[l]);
    cuMatrix<float> *o = rnn->forward(inp);
    cudaDeviceSynchronize();
    o->toCpu();
    memcpy(y, o->getHost(), sizeof(float) * T * N * H);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int ref_ctc_decode(const float *probs, int T, int N, int V, const char *vocab, int beam, int blank,
                              int max_len, char *out_paths, int *out_lens, float *out_scores) {
    cuMatrix<float> *seq = new cuMatrix<float>((float *)probs, T * N, V, 1);
    seq->toGpu();
    size_t slots = (size_t)N * beam * V;
    ref_prezero(slots * (2 * sizeof(BeamState) + 2 * sizeof(BeamState *) + 10 * sizeof(int)) + (1 << 20));
    CTCBeamSearch *dec = new CTCBeamSearch((char *)vocab, V, beam, blank);
    std::vector<std::pair<std::string, float> > res = dec->decode(seq, T, N);
    cudaDeviceSynchronize();
    for (int n = 0; n < N && n < (int)res.size(); n++) {
        int len = (int)res[n].first.size();
        out_lens[n] = len;
        memcpy(out_paths + (size_t)n * max_len, res[n].first.data(), len < max_len ? len : max_len);
        out_scores[n] = res[n].second;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
