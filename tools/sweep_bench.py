"""Config 5 of BASELINE.json: moment-retrieval evaluation sweep -- 150 s videos at 0.5 fps (75 clips), 128 videos per step
across the ranks (16 per GPU on 8 GPUs), then GPU mr_eval R1 / mAP scoring of all queries with the fixed-width gather to rank 0.
    python tools/sweep_bench.py [--steps S]                                  # 1 GPU: 16 videos per step
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sweep_bench.py
Per step every rank runs both Q-Formers + projections on its 16 x 75 clips (one (video, query) pair per video and step).
The LLM that turns the projected tokens into text is out of scope: its generations are synthetic strings in the
reference's output format, parsed by mraudio_b200.parsing (evaluate.py:48), scored on this rank's GPU and gathered.
Prints one JSON line on rank 0 (CUDA-event times, max over ranks)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mraudio_b200 import mr_eval
from mraudio_b200.parsing import parse_output
from mraudio_b200.xinstructblip import XInstructBLIPQFormers

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--videos", type=int, default=16)
ap.add_argument("--frames", type=int, default=75)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = XInstructBLIPQFormers(modalities=("video", "audio")).to(dev).eval()
g = torch.Generator().manual_seed(100 + rank)
B, F, T = args.videos, args.frames, 32
feats = {"video": torch.randn(B, F, 257, 1408, generator=g).to(torch.bfloat16).to(dev),
         "audio": torch.randn(B, F, 256, 768, generator=g).to(torch.bfloat16).to(dev)}
ids = torch.randint(1000, 30000, (B, T), generator=g).to(dev)
mask = torch.ones(B, T, dtype=torch.long, device=dev)


def synth_text(rng, n_max):
    """an LLM-style generation: 1..n_max windows on the 2-s grid of a 150-s video, with the occasional format slip"""
    wins = []
    for _ in range(int(rng.integers(1, n_max + 1))):
        a, b = sorted(rng.integers(0, 76, size=2).tolist())
        b = max(b, a + 1)
        wins.append(f"[{2 * a} {2 * b}]" if rng.random() < 0.1 else f"[{2 * a}, {2 * b}]")
    return "junk" if rng.random() < 0.02 else "[" + ", ".join(wins) + "]</s>"


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


with torch.no_grad():
    for _ in range(2):
        model.encode_modalities(feats, ids, mask)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    for _ in range(args.steps):
        out, _ = model.encode_modalities(feats, ids, mask)
e1.record()
barrier()
ms_fwd = max_over_ranks(e0.elapsed_time(e1) / args.steps)

# ---- scoring: every (video, query) pair of this rank -> parse -> score on the GPU -> gather to rank 0
rng = np.random.default_rng(7 + rank)
n_local = B * args.steps * 100          # 100 queries per (video, step): a sweep-sized scoring job (cf. QVH val: 1 550 queries)
base = rank * n_local
t0 = time.perf_counter()
records = [{"qid": base + i, "_order": base + i, "pred_relevant_windows": parse_output(synth_text(rng, 5)),
            "relevant_windows": parse_output(synth_text(rng, 3).replace("junk", "[[0, 2]]"))} for i in range(n_local)]
t_parse = time.perf_counter() - t0
mr_eval.score_records_distributed(records[:64], records[:64])     # warm-up: NCCL gather channels, pinned staging, module load
barrier()
t0 = time.perf_counter()
rec = mr_eval.score_records_distributed(records, records)
torch.cuda.synchronize()
t_score = time.perf_counter() - t0
t_score = max_over_ranks(t_score)
if rank == 0:
    total = n_local * world
    stub = [{"qid": i, "pred_relevant_windows": [[0, 0]], "relevant_windows": [[0, 0]]} for i in range(total)]
    t0 = time.perf_counter()
    res = mr_eval.eval_submission(stub, stub, verbose=False, _records=rec)
    t_reduce = time.perf_counter() - t0
    clips = B * F
    print(json.dumps({"config": "cfg5 sweep: 75-clip videos, %d videos per GPU per step, both Q-Formers + projections, then GPU mr_eval" % B,
                      "n_gpus": world, "forward_ms_per_step": ms_fwd, "clips_per_s_all_gpus": world * clips / (ms_fwd * 1e-3),
                      "videos_per_s_all_gpus": world * B / (ms_fwd * 1e-3),
                      "frac_of_sustained_bf16_peak_per_gpu": clips * 33.95e9 / (ms_fwd * 1e-3) / 1367.2e12,
                      "scored_queries": total, "parse_s_per_rank": t_parse, "score_and_gather_s": t_score, "reduce_s_rank0": t_reduce,
                      "brief": {k: res["brief"][k] for k in ("MR-full-R1@0.5", "MR-full-R1@0.7", "MR-full-mAP", "MR-full-mIoU",
                                                             "MR-full-invalid_pred_num")}}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
