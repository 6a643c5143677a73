"""world_size-2 gloo tests (CPU) of the multi-rank host logic: sharding of videos across ranks and the gather of scored
moments to rank 0.  The records are produced by the oracle here (no GPU); on the GPU box the same gather runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mr_eval_oracle as mo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mraudio_b200 import mr_eval
        from mraudio_b200.sharding import shard_range
        sub, gt = mo.synth_submission(301, seed=5)
        lo, hi = shard_range(len(sub), rank, world)
        mine = sub[lo:hi]
        rec = mo.score_records(mine, gt)                      # stands in for the GPU kernel in this CPU test
        merged = mr_eval.gather_records(rec, np.arange(lo, hi))
        if rank == 0:
            full = mo.score_records(sub, gt)
            ok = np.array_equal(merged["ap"], full["ap"]) and np.array_equal(merged["iou"], full["iou"]) and \
                np.array_equal(merged["invalid"], full["invalid"])
            brief = mr_eval.eval_submission(sub, gt, verbose=False, _records=merged)["brief"]
            ref = mo.eval_submission(sub, gt)["brief"]
            ok = ok and all(brief[k] == ref[k] or (np.isnan(brief[k]) and np.isnan(ref[k])) for k in ref)
            q.put(("ok" if ok else "mismatch", len(merged["iou"])))
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_gather_of_scored_moments_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    status, n = q.get(timeout=5)
    assert status == "ok" and n == 301


@pytest.mark.parametrize("n,world", [(301, 2), (8, 8), (5, 8), (128, 8), (0, 2)])
def test_shard_range_partitions_like_a_contiguous_sampler(n, world):
    from mraudio_b200.sharding import shard_range
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def test_warmup_cosine_schedule_matches_lavis_formula():
    """LinearWarmupCosineLRScheduler(max_epoch, min_lr=0, init_lr=3e-4, warmup_steps=1000, warmup_start_lr=1e-8)
    stepped per iteration (utils/trainer.py:66,127)."""
    import math
    from mraudio_b200.training import warmup_cosine_lr
    assert warmup_cosine_lr(0, 0, 10) == pytest.approx(1e-8)
    assert warmup_cosine_lr(0, 500, 10) == pytest.approx(1e-8 + (3e-4 - 1e-8) * 0.5)
    assert warmup_cosine_lr(0, 1000, 10) == pytest.approx(3e-4)            # warm-up over: cosine at epoch 0 = init_lr
    assert warmup_cosine_lr(5, 3, 10) == pytest.approx(3e-4 * 0.5 * (1 + math.cos(math.pi * 0.5)))
    assert warmup_cosine_lr(9, 0, 10) == pytest.approx(3e-4 * 0.5 * (1 + math.cos(math.pi * 0.9)))
