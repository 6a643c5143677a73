"""Config 4 of BASELINE.json: fine-tuning step of both Q-Formers + projections on cached encoder features, data-parallel,
per-GPU batch 8 videos x 8 frames, T = 32, surrogate loss (the frozen LLM is out of scope), fp32 NCCL gradient all-reduce.
    python tools/train_bench.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_bench.py
Prints one JSON line on rank 0: ms per optimizer step (CUDA events, max over ranks) with the overlapped bucketed
all-reduce and with the flat after-backward all-reduce."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from mraudio_b200.training import QFormerTrainer
from mraudio_b200.xinstructblip import XInstructBLIPQFormers

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = XInstructBLIPQFormers(modalities=("video", "audio")).to(dev)
g = torch.Generator().manual_seed(1 + rank)
B, F, T = 8, 8, 32
feats = {"video": torch.randn(B, F, 257, 1408, generator=g).to(torch.bfloat16).to(dev),
         "audio": torch.randn(B, F, 256, 768, generator=g).to(torch.bfloat16).to(dev)}
ids = torch.randint(1000, 30000, (B, T), generator=g).to(dev)
mask = torch.ones(B, T, dtype=torch.long, device=dev)
sur = {m: (torch.randn(B, F * 32, 4096, generator=g) * 1e-3).to(dev) for m in feats}
tr = QFormerTrainer(model, accum_grad_iters=1, warmup_steps=0, init_lr=1e-5)


def timed(n=8, warm=3):
    for _ in range(warm):
        tr.train_step(feats, ids, mask, surrogate=sur)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tr.train_step(feats, ids, mask, surrogate=sur)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {}
for rep in range(2):
    for name, ov, opt in (("overlapped_bucketed_allreduce", True, False), ("overlapped_allreduce_and_bucketed_optimizer", True, True),
                          ("flat_allreduce_after_backward", False, False)):
        tr.overlap_allreduce, tr.overlap_optimizer = ov, opt
        t = timed()
        res[name] = min(res.get(name, 1e9), t)
tr.overlap_allreduce, tr.overlap_optimizer = True, False
if world == 1:   # (the replay is single-process only: the NCCL exchange is not captured)
    tr.cuda_graph = True
    res["cuda_graph_replay"] = min(timed(), timed())
if rank == 0:
    grad_bytes = sum(s.numel for s in tr.states.values()) * 4
    ms = min(res.values())
    print(json.dumps({"config": "cfg4 fine-tuning step, 8 videos x 8 frames per GPU, both modalities, fwd + bwd + all-reduce + Adam",
                      "n_gpus": world, "ms_per_step": res, "clips_per_s_all_gpus": world * B * F / (ms * 1e-3),
                      "allreduce_bytes_fp32": grad_bytes,
                      "backward_launches": sum(s.last_backward_launches for s in tr.states.values())}))
if world > 1:
    dist.destroy_process_group()
