/* RNN_Cell.h -- mirror of the reference's Elman cell (RNN_Cell.h:3-38, RNN_Cell.cu:15-74):
 * out = tanh(x * W_ih + h * W_hh + (b_hh + b_ih)), W_ih [in, hid], W_hh [hid, hid]. */
#ifndef GASR_RNN_CELL_H
#define GASR_RNN_CELL_H
#include <stdlib.h>

#include "cuMatrix.h"

class RNN_Cell {
public:
    RNN_Cell(int batch_size, int input_size, int hidden_size)
        : input_size(input_size), hidden_size(hidden_size), batch_size(batch_size) {
        initRandom();
    }
    void initRandom() {
        alloc();
        for (int i = 0; i < w_ih->getLen(); i++) w_ih->getHost()[i] = (2.0f * rand() / RAND_MAX - 1.0f);
        for (int j = 0; j < w_hh->getLen(); j++) w_hh->getHost()[j] = (2.0f * rand() / RAND_MAX - 1.0f);
        upload();
    }
    void initParams(float *_w_ih, float *_w_hh, float *_b_ih, float *_b_hh) {
        alloc();
        for (int j = 0; j < w_ih->getLen(); j++) w_ih->getHost()[j] = _w_ih[j];
        for (int j = 0; j < w_hh->getLen(); j++) w_hh->getHost()[j] = _w_hh[j];
        for (int j = 0; j < b_ih->getLen(); j++) b_ih->getHost()[j] = _b_ih[j];
        for (int j = 0; j < b_hh->getLen(); j++) b_hh->getHost()[j] = _b_hh[j];
        upload();
    }
    /* one step (RNN_Cell.cu:65-74): 2 Sgemm + Sgeam + Tanh there, one kernel here */
    cuMatrix<float> *forward(cuMatrix<float> *inputs, cuMatrix<float> *pre_hidden, cuMatrix<float> *outputs) {
        gasr_cxx::check(gasr_rnn_cell_forward(gasr_cxx::ctx(), inputs->getDev(), pre_hidden->getDev(), w_ih->getDev(),
                                              w_hh->getDev(), b_ih->getDev(), b_hh->getDev(), outputs->getDev(),
                                              batch_size, input_size, hidden_size), "RNN_Cell::forward");
        return outputs;
    }
    cuMatrix<float> *w_ih;
    cuMatrix<float> *w_hh;
    cuMatrix<float> *b_ih;
    cuMatrix<float> *b_hh;
    int input_size;
    int hidden_size;
    int batch_size;

private:
    void alloc() {
        w_ih = new cuMatrix<float>(input_size, hidden_size, 1);
        w_hh = new cuMatrix<float>(hidden_size, hidden_size, 1);
        b_ih = new cuMatrix<float>(hidden_size, 1, 1);
        b_hh = new cuMatrix<float>(hidden_size, 1, 1);
    }
    void upload() { w_ih->toGpu(); w_hh->toGpu(); b_ih->toGpu(); b_hh->toGpu(); }
};
#endif
