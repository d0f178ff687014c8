"""baseline/config.json-compatible harness (gpu-accelerated-speech-recognition_b200/harness.py): the state_dict importer on
the CPU; the DeepSpeech topology against the torch golden of baseline/model.py and a config.json run on the GPU."""
import io
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _fake_state_dict(g):
    """The golden file stores the reference layout [in, out]; a torch state_dict holds [out, in]."""
    sd = {}
    for i, idx in enumerate((0, 2, 4)):
        sd[f"mlp123.{idx}.weight"] = g[f"fc{i}_w"].T.copy(); sd[f"mlp123.{idx}.bias"] = g[f"fc{i}_b"]
    sd["rnn.weight_ih_l0"] = g["rnn_w_ih"].T.copy(); sd["rnn.weight_hh_l0"] = g["rnn_w_hh"].T.copy()
    sd["rnn.bias_ih_l0"] = g["rnn_b_ih"]; sd["rnn.bias_hh_l0"] = g["rnn_b_hh"]
    for i, idx in enumerate((0, 2)):
        sd[f"mlp56.{idx}.weight"] = g[f"fc{3 + i}_w"].T.copy(); sd[f"mlp56.{idx}.bias"] = g[f"fc{3 + i}_b"]
    return sd


def test_state_dict_importer_transposes_to_reference_layout():
    import harness
    g = np.load(os.path.join(GOLDEN, "deepspeech_small.npz"))
    w = harness.weights_from_state_dict(_fake_state_dict(g))
    for k in w:
        assert w[k].shape == g[k].shape and np.array_equal(w[k], g[k]), k
        assert w[k].flags["C_CONTIGUOUS"] and w[k].dtype == np.float32
    # torch tensors are accepted too
    import torch
    w2 = harness.weights_from_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in _fake_state_dict(g).items()})
    assert all(np.array_equal(w[k], w2[k]) for k in w)
    cfg = {"input_size": 26, "n_context": 1, "linear_size": 64, "rnn_hidden_size": 32, "vocab_size": 46}
    r = harness.random_weights(cfg)
    assert r["fc0_w"].shape == (78, 64) and r["fc2_w"].shape == (64, 32) and r["fc4_w"].shape == (64, 47)
    assert abs(r["rnn_w_hh"]).max() <= 1 / np.sqrt(32)


@pytest.mark.gpu
def test_harness_model_matches_baseline_model_golden():
    import gasr
    import harness
    g = np.load(os.path.join(GOLDEN, "deepspeech_small.npz"))
    B, T, D = g["x_bt"].shape
    cfg = {"batch_size": B, "seg_len": T}
    ctx = gasr.Context(0)
    model = harness.DeepSpeech(cfg, harness.weights_from_state_dict(_fake_state_dict(g)), ctx)
    got = model.forward(g["x_bt"]).toCpu().getHost().reshape(T, B, -1)
    assert np.abs(got - g["logp_tnv"]).max() < 1e-4
    ctx.close()


@pytest.mark.gpu
def test_harness_runs_a_config_json(tmp_path):
    import harness
    cfgs = [{"batch_size": 8, "input_size": 26, "n_context": 1, "linear_size": 256, "rnn_hidden_size": 256, "vocab_size": 46,
             "seg_len": 40, "epoch": 2, "device": "cuda", "num_threads": 4, "beam_width": 100},
            {"batch_size": 8, "input_size": 26, "n_context": 1, "linear_size": 256, "rnn_hidden_size": 256, "vocab_size": 46,
             "seg_len": 40, "epoch": 2, "device": "cpu", "num_threads": 4, "beam_width": 100}]
    path = tmp_path / "config.json"
    path.write_text(json.dumps(cfgs))
    buf = io.StringIO()
    res = harness.run(cfgs[0], out=buf)
    text = buf.getvalue()
    assert "Forward:" in text and "CTC Decode" in text and "Overall" in text
    assert len(res) == 8 and all(isinstance(p, bytes) for p, _ in res)
    buf = io.StringIO()
    assert harness.run(cfgs[1], out=buf) is None and "no CPU path" in buf.getvalue()
    assert harness.main(["harness.py", str(path)]) == 0
