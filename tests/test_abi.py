"""CPU tests: the C-ABI library loads without a GPU, exports exactly what include/gasr.h declares, and refuses
to run without a CUDA device (no CPU fallback)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200")
sys.path.insert(0, PKG)


@pytest.fixture(scope="module")
def gasr():
    if not os.path.exists(os.path.join(PKG, "libgasr.so")):
        subprocess.run(["make", "-s", "-j8", "-C", PKG], check=True)
    import gasr as g
    return g


def _declared():
    text = open(os.path.join(ROOT, "include", "gasr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gasr_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree(gasr):
    declared = _declared()
    assert declared == sorted(gasr.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", gasr.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (gasr_[a-z0-9_]+)", out)))
    assert exported == declared
    for name in declared:
        assert hasattr(gasr._lib, name)


def test_library_is_sm100a_only(gasr):
    out = subprocess.run(["cuobjdump", "-lelf", gasr.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(gasr):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gasr.GasrError) as e:
        gasr.Context(0)
    assert e.value.status == gasr.ERR_CUDA and "no CPU fallback" in str(e.value)
    assert gasr._lib.gasr_version() >= 100


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, link, open or execute it
    (comments may mention that an oracle exists)."""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|oracle/|oracle\\|liboracle|oracle\.py|oracle_[a-z]+\s*\(|-loracle)")
    for dirpath, _, files in os.walk(PKG):
        if os.path.basename(dirpath) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(text), (dirpath, f, pat.search(text).group(0))
    for f in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", f), errors="replace").read()
        assert not pat.search(text), f


def test_synth_is_deterministic_and_sharded_consistently():
    import synth
    a = synth.spectrogram_batch(1234, 7, 6, 5)
    b = synth.spectrogram_batch(1234, 7, 6, 5)
    assert (a == b).all() and a.min() >= 0 and a.max() < 1
    # utterances 2..3 generated as their own shard equal the same columns of the full batch
    part = synth.spectrogram_batch(1234, 7, 2, 5, first_utt=2).reshape(7, 2, 5)
    assert (part == a.reshape(7, 6, 5)[:, 2:4, :]).all()
    w1 = synth.rnn_weights(4321, 5, 8, 2)
    w2 = synth.rnn_weights(4321, 5, 8, 2)
    assert all((x == y).all() for l1, l2 in zip(w1, w2) for x, y in zip(l1, l2))
    assert abs(np.abs(w1[1][0]).max()) <= 1 / np.sqrt(8)
    lp = synth.random_logprobs(5, 9, 2, 29)
    assert np.allclose(np.exp(lp.astype(np.float64)).sum(-1), 1.0, atol=1e-5)


def test_cpp_headers_compile_without_cuda(tmp_path, gasr):
    """The C++ mirror of the reference classes needs only g++ and libgasr.so (no CUDA headers on the caller side)."""
    exe = str(tmp_path / "t")
    r = subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "test_modules.cpp"), "-o", exe, "-L" + PKG, "-lgasr",
                        "-Wl,-rpath," + PKG], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
