/* RNN.h -- mirror of the reference's stacked RNN (RNN.h:4-37, RNN.cu:9-30): L tanh cells, time-major input
 * [time_step * batch, input_size] (row t*N+n), h_0 = 0, returns the last layer's full hidden sequence (borrowed).
 * forward() is ONE call into the library: per layer a batched tcgen05 projection GEMM + a persistent recurrence kernel,
 * instead of the reference's host loop over (t, layer) with 4 launches and 3 syncs per iteration. */
#ifndef GASR_RNN_H
#define GASR_RNN_H
#include "RNN_Cell.h"
#include "cuMatrix.h"

class RNN {
public:
    RNN(int batch_size, int input_size, int hidden_size, int time_step, int num_layers, int precision = GASR_PREC_FP32)
        : input_size(input_size), hidden_size(hidden_size), batch_size(batch_size), time_step(time_step),
          num_layers(num_layers), precision(precision) {
        rnn_cell = new RNN_Cell *[num_layers];
        h_0s = new cuMatrix<float> *[num_layers];
        hiddens = new cuMatrix<float> *[num_layers];
        for (int i = 0; i < num_layers; i++) {
            int _input_size = i == 0 ? input_size : hidden_size;
            rnn_cell[i] = new RNN_Cell(batch_size, _input_size, hidden_size);
            h_0s[i] = new cuMatrix<float>(batch_size, hidden_size, 1);
            h_0s[i]->toGpu();
            hiddens[i] = new cuMatrix<float>(time_step * batch_size, hidden_size, 1);
            hiddens[i]->toGpu();
        }
    }
    cuMatrix<float> *forward(cuMatrix<float> *inputs) {
        const float **w_ih = new const float *[num_layers], **w_hh = new const float *[num_layers];
        const float **b_ih = new const float *[num_layers], **b_hh = new const float *[num_layers];
        float **hid = new float *[num_layers];
        for (int l = 0; l < num_layers; l++) {
            w_ih[l] = rnn_cell[l]->w_ih->getDev(); w_hh[l] = rnn_cell[l]->w_hh->getDev();
            b_ih[l] = rnn_cell[l]->b_ih->getDev(); b_hh[l] = rnn_cell[l]->b_hh->getDev();
            hid[l] = hiddens[l]->getDev();
        }
        gasr_cxx::check(gasr_rnn_forward(gasr_cxx::ctx(), GASR_CELL_TANH, 0, time_step, batch_size, input_size, hidden_size,
                                         num_layers, w_ih, w_hh, b_ih, b_hh, inputs->getDev(), hid, precision),
                        "RNN::forward");
        delete[] w_ih; delete[] w_hh; delete[] b_ih; delete[] b_hh; delete[] hid;
        return hiddens[num_layers - 1];
    }
    cuMatrix<float> **h_0s;     /* num_layers * [batch_size, hidden_size], zero (RNN.h:16-17) */
    cuMatrix<float> **hiddens;  /* num_layers * [seq_len * batch, hidden_size] */
    RNN_Cell **rnn_cell;
    int input_size;
    int hidden_size;
    int batch_size;
    int time_step;
    int num_layers;
    int precision;
};
#endif
