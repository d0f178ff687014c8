"""Benchmark harness compatible with the reference's baseline/config.json.

    python gpu-accelerated-speech-recognition_b200/harness.py config.json [--state-dict model.pt]

Reads the same list of run dicts as baseline/main.py:59-65 ({batch_size, input_size, n_context, linear_size, rnn_hidden_size,
vocab_size, seg_len, epoch, device, num_threads, beam_width}), builds the DeepSpeech topology of baseline/model.py:22-49
(3 x (Linear + ReLU) -> tanh RNN -> Linear + ReLU -> Linear -> log_softmax) from the GPU modules of gasr.py, decodes with
CTCBeamSearch (log domain, blank 0) and prints the reference's three lines per run (baseline/main.py:54-56):

    Forward: %f s / CTC Decode %f s / Overall %f s   (means over `epoch` iterations)

Weights are torch-default random (U(+-1/sqrt(fan_in)), seeded) or imported from a PyTorch state_dict of baseline/model.py's
DeepSpeech (torch keeps [out, in]; the reference's C++ modules and this library keep [in, out]: the importer transposes).
There is no CPU path: entries with device "cpu" are skipped with a note (bench.py --impl reference times the CPU port).
"""
import json
import sys
import time

import numpy as np

import gasr
import synth

N_CONTEXT_DEFAULT = 1     # baseline/config.py supplies n_context to main.py; the shipped config.json states it per run


def weights_from_state_dict(sd):
    """baseline/model.py's parameter names -> the arrays the modules take, transposed to [in, out].
    Accepts a dict of numpy arrays or torch tensors."""
    def arr(k):
        v = sd[k]
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        return np.ascontiguousarray(v, dtype=np.float32)

    out = {}
    for i, idx in enumerate((0, 2, 4)):                       # mlp123: Linear at Sequential positions 0, 2, 4
        out[f"fc{i}_w"] = np.ascontiguousarray(arr(f"mlp123.{idx}.weight").T)
        out[f"fc{i}_b"] = arr(f"mlp123.{idx}.bias")
    out["rnn_w_ih"] = np.ascontiguousarray(arr("rnn.weight_ih_l0").T)
    out["rnn_w_hh"] = np.ascontiguousarray(arr("rnn.weight_hh_l0").T)
    out["rnn_b_ih"], out["rnn_b_hh"] = arr("rnn.bias_ih_l0"), arr("rnn.bias_hh_l0")
    for i, idx in enumerate((0, 2)):                          # mlp56: Linear at positions 0, 2
        out[f"fc{3 + i}_w"] = np.ascontiguousarray(arr(f"mlp56.{idx}.weight").T)
        out[f"fc{3 + i}_b"] = arr(f"mlp56.{idx}.bias")
    return out


def random_weights(config, seed=20261018):
    """torch-default init of the same topology (nn.Linear / nn.RNN: U(+-1/sqrt(fan_in)) resp. U(+-1/sqrt(hidden)))."""
    d_in = config["input_size"] + 2 * config["input_size"] * config.get("n_context", N_CONTEXT_DEFAULT)
    lin, hid, V = config["linear_size"], config["rnn_hidden_size"], config["vocab_size"] + 1
    shapes = [(d_in, lin), (lin, lin), (lin, hid), (hid, lin), (lin, V)]
    out, off = {}, 0
    for i, (a, b) in enumerate(shapes):
        k = 1.0 / np.sqrt(a)
        out[f"fc{i}_w"] = synth.uniform(seed, (a, b), -k, k, offset=off); off += a * b
        out[f"fc{i}_b"] = synth.uniform(seed, (b,), -k, k, offset=off); off += b
    k = 1.0 / np.sqrt(hid)
    for name, shape in (("rnn_w_ih", (hid, hid)), ("rnn_w_hh", (hid, hid)), ("rnn_b_ih", (hid,)), ("rnn_b_hh", (hid,))):
        out[name] = synth.uniform(seed, shape, -k, k, offset=off); off += int(np.prod(shape))
    return out


class DeepSpeech:
    """baseline/model.py:DeepSpeech on the GPU modules; forward(x [B, T, D]) -> device cuMatrix of log-probs [T*B, V+1]."""

    def __init__(self, config, weights, ctx):
        self.ctx = ctx
        self.B, self.T = config["batch_size"], config["seg_len"]
        rows = self.T * self.B
        w = weights
        self.fc = []
        for i in range(3):
            a, b = w[f"fc{i}_w"].shape
            self.fc.append(gasr.Linear(rows, a, b, ctx=ctx).initParams(w[f"fc{i}_w"], w[f"fc{i}_b"]))
        hid = w["rnn_w_hh"].shape[0]
        self.rnn = gasr.RNN(self.B, w["rnn_w_ih"].shape[0], hid, self.T, 1, ctx=ctx)
        self.rnn.rnn_cell[0].initParams(w["rnn_w_ih"], w["rnn_w_hh"], w["rnn_b_ih"], w["rnn_b_hh"])
        a, b = w["fc3_w"].shape
        self.fc3 = gasr.Linear(rows, a, b, ctx=ctx).initParams(w["fc3_w"], w["fc3_b"])
        a, b = w["fc4_w"].shape
        self.fc4 = gasr.Linear(rows, a, b, act=gasr.ACT_LOGSOFTMAX, ctx=ctx).initParams(w["fc4_w"], w["fc4_b"])
        self.V = b

    def forward(self, x_btd):
        B, T, D = x_btd.shape
        x = gasr.cuMatrix(np.ascontiguousarray(x_btd.transpose(1, 0, 2)).reshape(T * B, D), ctx=self.ctx).toGpu()   # x.permute(1, 0, 2)
        for m in self.fc:
            x = m.forward(x)
        h = self.rnn.forward(x)
        return self.fc4.forward(self.fc3.forward(h))


def run(config, state_dict=None, out=sys.stdout):
    if str(config.get("device", "cuda")).startswith("cpu"):
        print("skipped: this library has no CPU path (bench.py --impl reference times the CPU port of the reference)", file=out)
        return None
    ctx = gasr.Context(0)
    weights = weights_from_state_dict(state_dict) if state_dict is not None else random_weights(config)
    model = DeepSpeech(config, weights, ctx)
    B, T, V = config["batch_size"], config["seg_len"], config["vocab_size"] + 1
    d_in = config["input_size"] + 2 * config["input_size"] * config.get("n_context", N_CONTEXT_DEFAULT)
    vocab = bytes(range(1, V + 1))                            # V distinct labels, blank_id = 0 (main.py:28)
    decoder = gasr.CTCBeamSearch(vocab, V, config["beam_width"], 0, domain=gasr.DOMAIN_LOG, ctx=ctx)
    fwd = dec = tot = 0.0
    n_iter = config["epoch"]
    result = None
    for i in range(n_iter):
        inp = synth.uniform(1000 + i, (B, T, d_in), 0.0, 1.0)           # torch.rand((batch, seq, features)), main.py:39
        ctx.sync(); t0 = time.perf_counter()
        logp = model.forward(inp)
        ctx.sync(); t1 = time.perf_counter()
        result = decoder.decode(logp, T, B)
        ctx.sync(); t2 = time.perf_counter()
        fwd += t1 - t0; dec += t2 - t1; tot += t2 - t0
    print("Forward: %f s" % (fwd / n_iter), file=out)
    print("CTC Decode %f s" % (dec / n_iter), file=out)
    print("Overall %f s" % (tot / n_iter), file=out)
    ctx.close()
    return result


def main(argv):
    if len(argv) < 2:
        print(__doc__)
        return 2
    sd = None
    if "--state-dict" in argv:
        import torch
        sd = torch.load(argv[argv.index("--state-dict") + 1], map_location="cpu")
    configs = json.load(open(argv[1]))
    for config in configs:
        print("====== config ======")
        print(config)
        print("====================")
        run(config, sd)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
