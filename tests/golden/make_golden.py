"""Regenerates tests/golden/*.npz and *.json.  Run in the build container (needs /root/reference and torch):

    python tests/golden/make_golden.py

Sources of the pinned values:
  * nn_test.json      -- inputs nn_test.cpp:10-12 (Linear) and :37-60 (RNN); expected outputs are the
                         4-decimal comments nn_test.cpp:29-30 and :70-77 (the reference records nothing else).
  * ctc_main.json     -- vocab / beam / blank / probabilities of main.cpp:48-64.  The reference records no
                         expected output; `expected` holds the outputs of a literal array-level emulation of
                         CTCBeamSearch.cu recorded in SURVEY.md section 4 ([derived]), which the oracle reproduces.
  * deepspeech_small.npz -- /root/reference/baseline/model.py (DeepSpeech) imported and run on CPU with a seeded
                         small config: input, every parameter transposed to the reference's [in, out] layout,
                         and the log-softmax output.
  * rnn3_torch.npz    -- torch.nn.RNN(tanh, 3 layers) == the math of RNN.cu/RNN_Cell.cu with W transposed
                         (SURVEY.md section 4 [derived]); cfg2-shaped but small.
  * bigru_torch.npz   -- torch.nn.GRU(bidirectional, 2 layers): the cfg3 definition (the reference has no GRU).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/baseline")


def main():
    torch.manual_seed(20261018)
    torch.set_num_threads(1)
    # ---- baseline/model.py ---------------------------------------------------------------------------
    from model import DeepSpeech  # noqa: E402  (the reference's own module)
    cfg = {"batch_size": 3, "input_size": 5, "n_context": 1, "linear_size": 24, "rnn_hidden_size": 16,
           "vocab_size": 6}
    T = 7
    m = DeepSpeech(cfg).eval()
    x = torch.rand((cfg["batch_size"], T, cfg["input_size"] * 3))
    with torch.no_grad():
        y = m(x)  # [T, N, V+1] log-probs
    sd = m.state_dict()
    out = {"x_bt": x.numpy(), "logp_tnv": y.numpy()}
    for i, k in enumerate(["mlp123.0", "mlp123.2", "mlp123.4", "mlp56.0", "mlp56.2"]):
        out[f"fc{i}_w"] = sd[k + ".weight"].numpy().T.copy()  # -> [in, out]
        out[f"fc{i}_b"] = sd[k + ".bias"].numpy()
    out["rnn_w_ih"] = sd["rnn.weight_ih_l0"].numpy().T.copy()
    out["rnn_w_hh"] = sd["rnn.weight_hh_l0"].numpy().T.copy()
    out["rnn_b_ih"] = sd["rnn.bias_ih_l0"].numpy()
    out["rnn_b_hh"] = sd["rnn.bias_hh_l0"].numpy()
    np.savez(os.path.join(HERE, "deepspeech_small.npz"), **out)

    # ---- 3-layer tanh RNN (cfg2 topology, small) -----------------------------------------------------
    T, N, D, H, L = 12, 4, 9, 32, 3
    rnn = torch.nn.RNN(D, H, num_layers=L).eval()
    x = torch.rand((T, N, D))
    with torch.no_grad():
        y, _ = rnn(x)
    out = {"x": x.numpy().reshape(T * N, D), "y": y.numpy().reshape(T * N, H), "T": T, "N": N}
    for l in range(L):
        out[f"w_ih{l}"] = getattr(rnn, f"weight_ih_l{l}").detach().numpy().T.copy()
        out[f"w_hh{l}"] = getattr(rnn, f"weight_hh_l{l}").detach().numpy().T.copy()
        out[f"b_ih{l}"] = getattr(rnn, f"bias_ih_l{l}").detach().numpy()
        out[f"b_hh{l}"] = getattr(rnn, f"bias_hh_l{l}").detach().numpy()
    np.savez(os.path.join(HERE, "rnn3_torch.npz"), **out)

    # ---- 2-layer bidirectional GRU (cfg3 topology, small) --------------------------------------------
    T, N, D, H, L = 10, 3, 7, 24, 2
    gru = torch.nn.GRU(D, H, num_layers=L, bidirectional=True).eval()
    x = torch.rand((T, N, D))
    with torch.no_grad():
        y, _ = gru(x)
    out = {"x": x.numpy().reshape(T * N, D), "y": y.numpy().reshape(T * N, 2 * H), "T": T, "N": N, "H": H, "L": L}
    for l in range(L):
        for d, sfx in enumerate(["", "_reverse"]):
            out[f"w_ih{l}_{d}"] = getattr(gru, f"weight_ih_l{l}{sfx}").detach().numpy().T.copy()
            out[f"w_hh{l}_{d}"] = getattr(gru, f"weight_hh_l{l}{sfx}").detach().numpy().T.copy()
            out[f"b_ih{l}_{d}"] = getattr(gru, f"bias_ih_l{l}{sfx}").detach().numpy()
            out[f"b_hh{l}_{d}"] = getattr(gru, f"bias_hh_l{l}{sfx}").detach().numpy()
    np.savez(os.path.join(HERE, "bigru_torch.npz"), **out)

    # ---- reference test fixtures ---------------------------------------------------------------------
    nn_test = {
        "source": "nn_test.cpp:10-12,29-30 (Linear) and nn_test.cpp:37-60,70-77 (RNN)",
        "linear": {
            "x": [[0.0932, 0.3362, 0.1910], [0.6148, 0.5331, 0.1238]],
            "w_in_out": [[0.5699999928474426, 0.03020000085234642, -0.22759999334812164, 0.1242000013589859],
                         [0.34470000863075256, 0.49300000071525574, 0.37700000405311584, 0.04749999940395355],
                         [0.3377000093460083, -0.4636000096797943, -0.5188999772071838, 0.09910000115633011]],
            "b": [0.37158000469207764, -0.4036799967288971, 0.21911999583244324, 0.0001550900051370263],
            "expected_4dp": [[0.6051, 0.0000, 0.2255, 0.0466], [0.9476, 0.0000, 0.2159, 0.1141]],
        },
        "rnn": {
            "T": 4, "N": 2, "in": 3, "H": 5,
            "x_time_major": [[0.1321, 0.0296, 0.2351], [0.9742, 0.7064, 0.3638], [0.8129, 0.8474, 0.7844],
                             [0.9279, 0.9768, 0.7575], [0.5693, 0.9383, 0.6537], [0.1245, 0.9113, 0.5213],
                             [0.2325, 0.2616, 0.2558], [0.0063, 0.3980, 0.8896]],
            "w_ih": [[0.0269, -0.1896, 0.0500, 0.1968, -0.2331], [-0.1524, -0.1069, -0.3821, 0.3744, -0.0753],
                     [-0.0177, 0.1578, -0.1543, 0.0330, 0.2318]],
            "w_hh": [[0.0964, 0.3816, 0.1670, 0.2344, -0.0322], [-0.3150, 0.2676, 0.1690, 0.1398, 0.0135],
                     [-0.4383, -0.1151, 0.0135, 0.2061, -0.0159], [0.2352, -0.3320, -0.2943, 0.0488, -0.0794],
                     [0.2098, -0.0613, 0.3000, 0.2912, -0.0485]],
            "b_ih": [-0.1762, 0.1190, 0.3201, -0.2779, -0.0340],
            "b_hh": [-0.1449, -0.0929, 0.0448, -0.0617, 0.4359],
            "expected_4dp": [[-0.3151, 0.0350, 0.3130, -0.2865, 0.3998], [-0.3876, -0.1749, 0.0873, 0.1279, 0.2031],
                             [-0.5402, -0.1695, 0.1219, 0.2557, 0.3270], [-0.3853, -0.3751, -0.1476, 0.1991, 0.2695],
                             [-0.3659, -0.4214, -0.1590, 0.1271, 0.3159], [-0.2134, -0.3147, -0.1635, -0.0416, 0.3850],
                             [-0.0956, -0.2925, 0.1586, -0.2606, 0.3544], [-0.1743, -0.0339, 0.1121, -0.1758, 0.5128]],
        },
    }
    json.dump(nn_test, open(os.path.join(HERE, "nn_test.json"), "w"), indent=1)

    ctc_main = {
        "source": "main.cpp:48-64 (inputs); expected = SURVEY.md section 4 [derived] literal emulation of CTCBeamSearch.cu",
        "vocab": "$abc", "blank": 0, "T": 10,
        "probs": [0.36225085, 0.09518672, 0.08850375, 0.45405867, 0.08869431, 0.18445025, 0.3304224, 0.39643304,
                  0.09951598, 0.17646984, 0.42063249, 0.30338169, 0.15361776, 0.46521112, 0.18132693, 0.19984419,
                  0.33478711, 0.16607367, 0.29571415, 0.20342507, 0.01292992, 0.36438928, 0.00184853, 0.62083227,
                  0.34142441, 0.16742833, 0.38500542, 0.10614183, 0.4443139, 0.12738693, 0.36856127, 0.0597379,
                  0.37673064, 0.13478024, 0.2735787, 0.21491042, 0.34790623, 0.04654182, 0.34069546, 0.26485648],
        "expected": {"1": ["cbacb", 1.6414496e-4], "2": ["cbacbc", 1.9566051e-3], "3": ["cbacb", 4.6938560e-3],
                     "4": ["cbacb", 4.6938560e-3], "5": ["cbacb", 5.1246714e-3], "8": ["cbacb", 1.0547487e-2]},
        "beam2_frames": [[["c", .45405868], ["$", .36225086]], [["c", .32361206], ["cb", .15003116]],
                         [["cb", .19922973], ["c", .09817797]], [["cba", .09268389], ["cb", .05392803]],
                         [["cba$", .03102937], ["cbab", .02740794]], [["cbac", .01926403], ["cbabc", .01701573]],
                         [["cbacb", .00741676], ["cbac$", .00657721]], [["cbacb", .00515763], ["cbacb$", .00329537]],
                         [["cbacb$", .00318451], ["cbacbc", .00181664]], [["cbacbc", .00195661], ["cbacb", .00110791]]],
    }
    json.dump(ctc_main, open(os.path.join(HERE, "ctc_main.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
