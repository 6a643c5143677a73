"""cfg4 fine-tuning step only (tools/secondary.py::cfg4_finetune) under torchrun: one JSON line on rank 0.  Used to sweep NCCL
settings (NCCL_MAX_CTAS ...) for the overlap of the gradient exchange with the backward."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import secondary

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
out = secondary.cfg4_finetune(secondary.Ctx(world, rank, dev))
if rank == 0:
    out["env"] = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
    out.pop("workload", None)
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
