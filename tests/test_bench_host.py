"""CPU checks of bench.py's host-side helpers (no GPU, no library calls): the roofline inputs it reads from profiles/, the NUMA
binding of the pinned buffers, the peak table."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_recurrence_traffic_comes_from_the_committed_capture():
    b = _bench()
    t = json.load(open(os.path.join(ROOT, "profiles", "r2_rnn_wide2_traffic.json")))
    wave = int(t["utterances_per_launch"])
    assert b.rec_traffic(wave) == t["dram_bytes_per_launch"]
    assert b.rec_traffic(wave + 1) is None                    # a capture at another launch shape says nothing
    # the capture itself: DRAM traffic within 15 % of the algorithmic bytes of a launch (no wasted re-reads)
    assert 0.85 < t["dram_bytes_per_launch"] / t["algorithmic_bytes_per_launch"] < 1.15
    assert t["algorithmic_bytes_per_launch"] == t["steps_per_launch"] * wave * 512 * 8


def test_numa_binding_is_a_no_op_without_a_second_node_or_when_switched_off(monkeypatch):
    b = _bench()
    monkeypatch.setenv("GASR_BENCH_NUMA", "0")
    assert b.bind_host_memory_to_gpu_node(0) is None
    monkeypatch.delenv("GASR_BENCH_NUMA")
    nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()] \
        if os.path.isdir("/sys/devices/system/node") else []
    if len(nodes) < 2:
        assert b.bind_host_memory_to_gpu_node(0) is None


def test_peaks_and_workload_constants():
    b = _bench()
    hbm, tc, kind = b.measured_peaks()
    assert hbm > 1000 and tc > 100 and kind in ("measured", "fallback")
    assert b.CFG == dict(T=1000, D=161, H=512, L=3, V=29, beam=16) and b.UTTS == 8192
    assert "cfg5" in b.WORKLOAD
