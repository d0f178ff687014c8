"""CPU test of the N > 1 host path: world_size-2 gloo processes each take their contiguous shard of utterances,
decode it (the CPU oracle stands in for the GPU kernels here -- tests may use it as a checker / stand-in), and the
gathered result must equal the single-process result bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-accelerated-speech-recognition_b200"))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, T, N, V, beam, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import shard
    import synth
    from oracle import oracle as O
    lo, hi = shard.shard_range(N, world, rank)
    lp = synth.random_logprobs(7, T, N, V)[:, lo:hi, :]
    p, s = O.ctc_decode(np.ascontiguousarray(lp), synth.VOCAB29, 0, beam, domain="log")
    paths, scores = shard.gather_results(p, s)
    if rank == 0:
        ret["paths"], ret["scores"] = paths, scores
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    import shard
    for n in (0, 1, 7, 64, 8192):
        for w in (1, 2, 3, 8):
            spans = [shard.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_equals_single_process():
    import synth
    from oracle import oracle as O
    T, N, V, beam = 40, 7, 29, 8
    full_p, full_s = O.ctc_decode(synth.random_logprobs(7, T, N, V), synth.VOCAB29, 0, beam, domain="log")
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, T, N, V, beam, ret), nprocs=2, join=True)
    assert list(ret["paths"]) == full_p
    assert list(ret["scores"]) == full_s
