"""Multi-GPU host logic: utterances are independent (RNN rows never interact, RNN.cu:15-27; the decoder segments by
utterance, CTCBeamSearch.cu:416), so a batch is split into contiguous shards, one per rank / GPU, with NO collective
on the data path; only the few-hundred-byte results per utterance are gathered on the host side."""


def shard_range(n_utt, world, rank):
    """Contiguous block of utterances [lo, hi) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_utt, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local_paths, local_scores, group=None):
    """All ranks call this; every rank gets the full, utterance-ordered (paths, scores).  Uses
    torch.distributed.all_gather_object over whatever backend the process group has (gloo on CPU hosts,
    nccl under torchrun on GPUs) -- the payload is host data, a few hundred bytes per utterance."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(local_paths), list(local_scores)
    world = dist.get_world_size(group)
    bucket = [None] * world
    dist.all_gather_object(bucket, (list(local_paths), list(local_scores)), group=group)
    paths, scores = [], []
    for p, s in bucket:       # rank order == utterance order (contiguous shards)
        paths.extend(p)
        scores.extend(s)
    return paths, scores
