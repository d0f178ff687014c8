/* Linear.h -- mirror of the reference's Linear (Linear.h:3-28, Linear.cu:12-49): out = ReLU(in * W + b), W [in, out].
 * `act` widens it to no activation / fused log-softmax for the output layer (baseline/model.py:46-49). */
#ifndef GASR_LINEAR_H
#define GASR_LINEAR_H
#include <stdlib.h>

#include "cuMatrix.h"

class Linear {
public:
    Linear(cuMatrix<float> *weight, cuMatrix<float> *bias, int batch_size, int input_size, int output_size)
        : w(weight), b(bias), input_size(input_size), output_size(output_size), batch_size(batch_size), act(GASR_ACT_RELU) {
        outputs = new cuMatrix<float>(batch_size, output_size, 1);
        outputs->toGpu();
    }
    Linear(int batch_size, int input_size, int output_size, int act = GASR_ACT_RELU)
        : input_size(input_size), output_size(output_size), batch_size(batch_size), act(act) {
        initRandom();
        outputs = new cuMatrix<float>(batch_size, output_size, 1);
        outputs->toGpu();
    }
    /* U[-1,1] weights from glibc rand(), zero bias (Linear.cu:12-21) */
    void initRandom() {
        w = new cuMatrix<float>(input_size, output_size, 1);
        b = new cuMatrix<float>(output_size, 1, 1);
        for (int j = 0; j < w->getLen(); j++) w->getHost()[j] = (2.0f * rand() / RAND_MAX - 1.0f);
        w->toGpu(); b->toGpu();
    }
    void initParams(float *weight, float *bias) {
        w = new cuMatrix<float>(input_size, output_size, 1);
        b = new cuMatrix<float>(output_size, 1, 1);
        for (int j = 0; j < w->getLen(); j++) w->getHost()[j] = weight[j];
        for (int j = 0; j < b->getLen(); j++) b->getHost()[j] = bias[j];
        w->toGpu(); b->toGpu();
    }
    /* inputs [batch_size, input_size] -> borrowed outputs [batch_size, output_size] (Linear.cu:42-49), one fused kernel */
    cuMatrix<float> *forward(cuMatrix<float> *inputs) {
        if (inputs->cols != input_size || inputs->rows != batch_size) { printf("matrix mul dimension mismatch\n"); exit(1); }
        gasr_cxx::check(gasr_linear_forward(gasr_cxx::ctx(), inputs->getDev(), inputs->cols, w->getDev(), b->getDev(),
                                            outputs->getDev(), output_size, batch_size, input_size, output_size, act),
                        "Linear::forward");
        return outputs;
    }
    cuMatrix<float> *w;
    cuMatrix<float> *b;
    cuMatrix<float> *outputs;
    int input_size;
    int output_size;
    int batch_size;
    int act;
};
#endif
