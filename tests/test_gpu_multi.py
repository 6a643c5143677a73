"""Two-GPU checks (skipped on a single-GPU box): data-parallel fine-tuning step with the NCCL gradient all-reduce over
NVLink, and the NCCL gather of scored moments.  Launched as 2 ranks with torch.multiprocessing."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mraudio_b200 import mr_eval
        from mraudio_b200.sharding import shard_range
        from mraudio_b200.training import QFormerTrainer
        from mraudio_b200.xinstructblip import XInstructBLIPQFormers
        from oracle import mr_eval_oracle as mo

        # ---- DDP-equivalent training: every rank holds the same weights, sees ITS shard of the global batch
        torch.manual_seed(0)
        model = XInstructBLIPQFormers(modalities=("video", "audio"), encoder_num_features={"video": 128, "audio": 64},
                                      llm_hidden_size=128, num_hidden_layers=2).cuda()
        tr = QFormerTrainer(model, accum_grad_iters=1, warmup_steps=0, init_lr=1e-3)
        g = torch.Generator().manual_seed(3)
        B = 4
        feats = {"video": torch.randn(B, 2, 17, 128, generator=g).to(torch.bfloat16), "audio": torch.randn(B, 2, 16, 64, generator=g).to(torch.bfloat16)}
        ids = torch.randint(1000, 30000, (B, 8), generator=g)
        mask = torch.ones(B, 8, dtype=torch.long)
        sur = {m: torch.randn(B, 2 * 32, 128, generator=g) for m in feats}
        lo, hi = shard_range(B, rank, world)
        loss = tr.train_step({m: t[lo:hi].cuda() for m, t in feats.items()}, ids[lo:hi].cuda(), mask[lo:hi].cuda(),
                             surrogate={m: t[lo:hi].cuda() for m, t in sur.items()})
        torch.cuda.synchronize()
        flat = torch.cat([s.flat for s in tr.states.values()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], t) for t in gathered)      # replicas stay bit-identical after the step
        # the same step with the flat after-backward all-reduce (no overlap) gives the same parameters
        torch.manual_seed(0)
        model2 = XInstructBLIPQFormers(modalities=("video", "audio"), encoder_num_features={"video": 128, "audio": 64},
                                       llm_hidden_size=128, num_hidden_layers=2).cuda()
        tr2 = QFormerTrainer(model2, accum_grad_iters=1, warmup_steps=0, init_lr=1e-3, overlap_allreduce=False)
        tr2.train_step({m: t[lo:hi].cuda() for m, t in feats.items()}, ids[lo:hi].cuda(), mask[lo:hi].cuda(),
                       surrogate={m: t[lo:hi].cuda() for m, t in sur.items()})
        torch.cuda.synchronize()
        flat2 = torch.cat([s.flat for s in tr2.states.values()])
        same = same and torch.allclose(flat, flat2, rtol=0, atol=2e-5) and not torch.equal(flat2, torch.zeros_like(flat2))
        # ---- the defining property of DistributedDataParallel (utils/trainer.py:69): the all-reduced gradient of the ranks'
        #      shards == the gradient of the whole batch on one rank (same prompt for every video: the reference tiles prompts
        #      frame-major over the rows, so with different prompts the pairing would depend on the batch size); fp32 exchange
        #      to 1e-4 (summation order only), bf16 exchange to bf16 rounding
        ids1 = ids[:1].expand(B, -1).contiguous()
        for comm, tol in ((torch.float32, 1e-4), (torch.bfloat16, 1e-2)):
            tr.set_grad_comm_dtype(comm)
            for st in tr.states.values():
                st.zero_grad()
            tr.train_step({m: t[lo:hi].cuda() for m, t in feats.items()}, ids1[lo:hi].cuda(), mask[lo:hi].cuda(),
                          surrogate={m: t[lo:hi].cuda() for m, t in sur.items()}, apply_optimizer=False)
            torch.cuda.synchronize()
            g_ddp = {m: st.grad.clone() for m, st in tr.states.items()}
            for st in tr.states.values():
                st.zero_grad()
            tr.allreduce_enabled = False
            tr.train_step({m: t.cuda() for m, t in feats.items()}, ids1.cuda(), mask.cuda(), surrogate={m: t.cuda() for m, t in sur.items()},
                          apply_optimizer=False)
            tr.allreduce_enabled = True
            torch.cuda.synchronize()
            for m, st in tr.states.items():
                err = ((g_ddp[m] - st.grad).abs().max() / st.grad.abs().max()).item()
                same = same and err < tol and st.grad.abs().max().item() > 0
            for st in tr.states.values():
                st.zero_grad()
        # ---- scorer: NCCL gather of scored moments == single-process result
        sub, gt = mo.synth_submission(333, seed=9)
        lo, hi = shard_range(len(sub), rank, world)
        mine = [dict(d, _order=lo + i) for i, d in enumerate(sub[lo:hi])]
        rec = mr_eval.score_records_distributed(mine, gt)
        ok = same
        if rank == 0:
            full = mo.score_records(sub, gt)
            ok = ok and np.array_equal(rec["ap"], full["ap"]) and np.array_equal(rec["iou"], full["iou"])
            q.put(("ok" if ok else "mismatch", float(loss)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_training_step_and_scored_moment_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    status, _ = q.get(timeout=5)
    assert status == "ok"
