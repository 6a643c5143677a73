"""Diagnostic: data-parallel equivalence of the fine-tuning step at full model size, per parameter segment.
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist
from secondary import _train_batch
from mraudio_b200.training import QFormerTrainer
from mraudio_b200.xinstructblip import XInstructBLIPQFormers

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
layers = int(os.environ.get("LAYERS", "12"))
torch.manual_seed(0)
model = XInstructBLIPQFormers(modalities=("video", "audio"), num_hidden_layers=layers).to(dev)
tr = QFormerTrainer(model, accum_grad_iters=1, warmup_steps=0, init_lr=1e-5, grad_comm_dtype=torch.float32)
F, T, Bc = 8, 32, 2
parts = [_train_batch(Bc, F, T, 1000 + r, dev, text_seed=77) for r in range(world)]
whole = ({m: torch.cat([p[0][m] for p in parts]) for m in parts[0][0]}, torch.cat([p[1] for p in parts]),
         torch.cat([p[2] for p in parts]), {m: torch.cat([p[3][m] for p in parts]) for m in parts[0][3]})
shard = parts[rank]


def grads(batch, reduce, overlap=True, parallel=True):
    tr.overlap_allreduce, tr.allreduce_enabled, tr.parallel_modalities = overlap, reduce, parallel
    for st in tr.states.values():
        st.zero_grad()
    tr.train_step(batch[0], batch[1], batch[2], surrogate=batch[3], apply_optimizer=False)
    torch.cuda.synchronize()
    return {m: st.grad.clone() for m, st in tr.states.items()}


g_whole = grads(whole, False)
g_whole2 = grads(whole, False)
for name, kw in (("overlapped", dict(overlap=True)), ("flat", dict(overlap=False)), ("overlapped, one stream", dict(overlap=True, parallel=False))):
    g = grads(shard, True, **kw)
    for m, st in tr.states.items():
        ref = g_whole[m]
        worst = []
        for seg, (off, shp) in st.seg.items():
            n = 1
            for d in shp:
                n *= d
            a, b = g[m][off:off + n], ref[off:off + n]
            den = b.abs().max().item()
            err = (a - b).abs().max().item() / den if den > 0 else (a.abs().max().item())
            worst.append((err, seg))
        worst.sort(reverse=True)
        tot = ((g[m] - ref).abs().max() / ref.abs().max()).item()
        rr = ((g_whole2[m] - ref).abs().max() / ref.abs().max()).item()
        if rank == 0:
            print(f"[{name}] {m}: whole-buffer rel err {tot:.3e} (run-to-run of the whole batch: {rr:.3e}); worst segments:",
                  [(f"{e:.2e}", s) for e, s in worst[:6]], flush=True)
dist.destroy_process_group()
