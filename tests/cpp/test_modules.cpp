// C++ drop-in check: the reference's own drivers (nn_test.cpp:9-79, main.cpp:48-75), rewritten as assertions, compiled
// with plain g++ against include/*.h and linked to libgasr.so.  Exit code 0 = all expectations met.
#include <math.h>
#include <stdio.h>

#include "modules.h"

static int failures = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { printf("FAIL %s:%d ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); failures++; } } while (0)

int main() {
    // ---- Linear known answer (nn_test.cpp:9-32) ----
    float inp[6] = {0.0932f, 0.3362f, 0.1910f, 0.6148f, 0.5331f, 0.1238f};
    float weight[12] = {0.5699999928474426f, 0.03020000085234642f, -0.22759999334812164f, 0.1242000013589859f,
                        0.34470000863075256f, 0.49300000071525574f, 0.37700000405311584f, 0.04749999940395355f,
                        0.3377000093460083f, -0.4636000096797943f, -0.5188999772071838f, 0.09910000115633011f};
    float bias[4] = {0.37158000469207764f, -0.4036799967288971f, 0.21911999583244324f, 0.0001550900051370263f};
    cuMatrix<float> *inp_test = new cuMatrix<float>(inp, 2, 3, 1);
    inp_test->toGpu();
    Linear *mlp_test = new Linear(2, 3, 4);
    mlp_test->initParams(weight, bias);
    cuMatrix<float> *out_test = mlp_test->forward(inp_test);
    out_test->toCpu();
    const float lin_expect[8] = {0.6051f, 0.0f, 0.2255f, 0.0466f, 0.9476f, 0.0f, 0.2159f, 0.1141f};
    for (int i = 0; i < 8; i++) EXPECT(fabsf(out_test->getHost()[i] - lin_expect[i]) < 1e-4f, "linear[%d] = %f", i, out_test->getHost()[i]);

    // ---- RNN known answer (nn_test.cpp:35-79) ----
    float inp_rnn[4 * 2 * 3] = {0.1321f, 0.0296f, 0.2351f, 0.9742f, 0.7064f, 0.3638f, 0.8129f, 0.8474f, 0.7844f, 0.9279f, 0.9768f, 0.7575f,
                                0.5693f, 0.9383f, 0.6537f, 0.1245f, 0.9113f, 0.5213f, 0.2325f, 0.2616f, 0.2558f, 0.0063f, 0.3980f, 0.8896f};
    float w_ih[15] = {0.0269f, -0.1896f, 0.0500f, 0.1968f, -0.2331f, -0.1524f, -0.1069f, -0.3821f, 0.3744f, -0.0753f, -0.0177f, 0.1578f, -0.1543f, 0.0330f, 0.2318f};
    float w_hh[25] = {0.0964f, 0.3816f, 0.1670f, 0.2344f, -0.0322f, -0.3150f, 0.2676f, 0.1690f, 0.1398f, 0.0135f, -0.4383f, -0.1151f, 0.0135f, 0.2061f, -0.0159f,
                      0.2352f, -0.3320f, -0.2943f, 0.0488f, -0.0794f, 0.2098f, -0.0613f, 0.3000f, 0.2912f, -0.0485f};
    float b_ih[5] = {-0.1762f, 0.1190f, 0.3201f, -0.2779f, -0.0340f};
    float b_hh[5] = {-0.1449f, -0.0929f, 0.0448f, -0.0617f, 0.4359f};
    cuMatrix<float> *inp_test_rnn = new cuMatrix<float>(inp_rnn, 4 * 2, 3, 1);
    inp_test_rnn->toGpu();
    RNN *rnn_test = new RNN(2, 3, 5, 4, 1);
    rnn_test->rnn_cell[0]->initParams(w_ih, w_hh, b_ih, b_hh);
    cuMatrix<float> *out_rnn = rnn_test->forward(inp_test_rnn);
    out_rnn->toCpu();
    const float rnn_expect[40] = {-0.3151f, 0.0350f, 0.3130f, -0.2865f, 0.3998f, -0.3876f, -0.1749f, 0.0873f, 0.1279f, 0.2031f,
                                  -0.5402f, -0.1695f, 0.1219f, 0.2557f, 0.3270f, -0.3853f, -0.3751f, -0.1476f, 0.1991f, 0.2695f,
                                  -0.3659f, -0.4214f, -0.1590f, 0.1271f, 0.3159f, -0.2134f, -0.3147f, -0.1635f, -0.0416f, 0.3850f,
                                  -0.0956f, -0.2925f, 0.1586f, -0.2606f, 0.3544f, -0.1743f, -0.0339f, 0.1121f, -0.1758f, 0.5128f};
    for (int i = 0; i < 40; i++) EXPECT(fabsf(out_rnn->getHost()[i] - rnn_expect[i]) < 1e-4f, "rnn[%d] = %f", i, out_rnn->getHost()[i]);
    // slice view + one cell step == first timestep
    cuMatrix<float> *x0 = new cuMatrix<float>(inp_test_rnn, 0, 2, 3, 1);
    cuMatrix<float> *h1 = new cuMatrix<float>(2, 5, 1);
    rnn_test->rnn_cell[0]->forward(x0, rnn_test->h_0s[0], h1);
    h1->toCpu();
    for (int i = 0; i < 10; i++) EXPECT(fabsf(h1->getHost()[i] - rnn_expect[i]) < 1e-4f, "cell[%d] = %f", i, h1->getHost()[i]);

    // ---- CTC (main.cpp:48-72): vocab {'$','a','b','c'}, blank 0, beam 2 ----
    char vocab[] = {'$', 'a', 'b', 'c'};
    float test[] = {0.36225085f, 0.09518672f, 0.08850375f, 0.45405867f, 0.08869431f, 0.18445025f, 0.3304224f, 0.39643304f,
                    0.09951598f, 0.17646984f, 0.42063249f, 0.30338169f, 0.15361776f, 0.46521112f, 0.18132693f, 0.19984419f,
                    0.33478711f, 0.16607367f, 0.29571415f, 0.20342507f, 0.01292992f, 0.36438928f, 0.00184853f, 0.62083227f,
                    0.34142441f, 0.16742833f, 0.38500542f, 0.10614183f, 0.4443139f, 0.12738693f, 0.36856127f, 0.0597379f,
                    0.37673064f, 0.13478024f, 0.2735787f, 0.21491042f, 0.34790623f, 0.04654182f, 0.34069546f, 0.26485648f};
    CTCBeamSearch *decoder = new CTCBeamSearch(vocab, 4, 2, 0);
    cuMatrix<float> *seqProb = new cuMatrix<float>(10, 4, 1);
    for (int j = 0; j < seqProb->getLen(); j++) seqProb->getHost()[j] = test[j];
    seqProb->toGpu();
    std::vector<std::pair<std::string, float> > best = decoder->decode(seqProb, 10, 1);
    EXPECT(best.size() == 1 && best[0].first == "cbacbc", "ctc path = %s", best[0].first.c_str());
    EXPECT(best[0].second == 1.9566051e-3f, "ctc prob = %.9g", best[0].second);
    // per-utterance frame counts in, per-token timesteps out (what baseline/main.py:45-46 takes from its decoder)
    {
        std::vector<std::vector<int> > ts;
        int full_len[1] = {10}, short_len[1] = {4};
        std::vector<std::pair<std::string, float> > b2 = decoder->decode(seqProb, 10, 1, full_len, &ts);
        EXPECT(b2[0].first == best[0].first && b2[0].second == best[0].second, "lens = T changes nothing");
        EXPECT(ts.size() == 1 && ts[0].size() == 6, "one timestep per character");
        for (size_t i = 1; i < ts[0].size(); i++) EXPECT(ts[0][i - 1] < ts[0][i] && ts[0][i] < 10, "timesteps increase");
        cuMatrix<float> *first4 = new cuMatrix<float>(4, 4, 1);
        for (int j = 0; j < first4->getLen(); j++) first4->getHost()[j] = test[j];
        first4->toGpu();
        std::vector<std::pair<std::string, float> > want = decoder->decode(first4, 4, 1);
        b2 = decoder->decode(seqProb, 10, 1, short_len, NULL);
        EXPECT(b2[0].first == want[0].first && b2[0].second == want[0].second, "lens = 4 decodes the first four frames: %s vs %s",
               b2[0].first.c_str(), want[0].first.c_str());
    }

    // matrixMul / matrixAdd
    float a[4] = {1, 2, 3, 4}, b[4] = {5, 6, 7, 8};
    cuMatrix<float> A(a, 2, 2, 1), B(b, 2, 2, 1), Z(2, 2, 1);
    A.toGpu(); B.toGpu();
    matrixMul(&A, &B, &Z); Z.toCpu();
    EXPECT(Z.get(0, 0, 0) == 19 && Z.get(1, 1, 0) == 50, "matrixMul");
    matrixAdd(&A, &B, &Z, 2.0f); Z.toCpu();
    EXPECT(Z.get(0, 1, 0) == 14, "matrixAdd");
    MemoryMonitor::instance()->printGpuMemory();
    printf(failures ? "FAILED (%d)\n" : "all C++ module checks passed\n", failures);
    return failures ? 1 : 0;
}
