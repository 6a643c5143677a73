#!/usr/bin/env python
"""Headline benchmark: X-InstructBLIP video+audio Q-Former + llm_proj forward, clips/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A *step* is one pass of the hot path over one batch of synthetic encoder features: per rank 32 videos x 8 frames
(config 2 of BASELINE.json) -> 256 video rows [257 x 1408] + 256 audio rows [256 x 768], T = 32 prompt tokens, both
Q-Formers (12 layers) + both llm_proj (768 -> 4096).  1 clip = 1 video row + 1 audio row + both projections
(33.95 GFLOP algorithmic).  Weak scaling: every rank processes its own 32 videos; no data-path collective.

Timed regions (CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks):
  value    inputs resident in HBM (285 MB of features per step > 126 MB L2, so nothing is served from cache)
  e2e      the public host API (XInstructBLIPQFormers.encode_modalities_host): pinned host features -> H2D -> Q-Formers ->
           D2H of inputs_llm into pinned host memory, double-buffered over three streams
  roofline an instrumented repeat of the K steps with CUDA events around every tensor-core GEMM launch
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, I_FF, NQ, D_LLM, LAYERS, HEADS = 768, 3072, 32, 4096, 12, 12
MODAL = {"video": (257, 1408), "audio": (256, 768)}   # Nk, W


def flops_per_row(Nk, W, T, layers=LAYERS, cross_freq=2):
    """2*M*N*K per GEMM incl. QK^T and PV; returns (linear_flops, attention_core_flops, cross_kv_flops)."""
    S = NQ + T
    lc = sum(1 for i in range(layers) if i % cross_freq == 0)
    lin = layers * (3 * 2 * S * H * H + 2 * S * H * H)          # self q,k,v + out
    lin += lc * (2 * 2 * NQ * H * H)                             # cross q + out
    kv = lc * (2 * 2 * Nk * W * H)                               # cross k,v
    lin += kv
    lin += layers * 4 * NQ * H * I_FF + layers * 4 * T * H * I_FF
    lin += 2 * NQ * H * D_LLM
    core = layers * 4 * S * S * H + lc * 4 * NQ * Nk * H
    return float(lin), float(core), float(kv)


def load_ncu_traffic():
    """dram bytes per launch and tensor-pipe activity of the captured launch types (ONE `ncu --set full` capture of a
    config-2 step, profiles/r01g_ncu_full_summary.json, written by tools/ncu_summary.py)"""
    p = os.path.join(ROOT, "profiles", "r01g_ncu_full_summary.json")
    if not os.path.exists(p):
        return None
    rows = json.load(open(p))
    names = ["cross_kv_video (gemm_tc_kernel, 2-CTA MMA)", "cross_kv_audio (gemm_tc_kernel, 2-CTA MMA)", "qkv_grouped (gemm_tc_kernel)",
             "attn_out+LN (gemm_ln_kernel)", "cross_q (gemm_tc_kernel)", "cross_out+LN (gemm_ln_kernel)", "ffn_up GELU (gemm_tc_kernel)",
             "ffn_down+LN (gemm_ln_kernel)"]
    out = {}
    for name, r in zip(names, rows):
        out[name] = {"dram_bytes_per_launch": r.get("dram_bytes"),
                     "tensor_pipe_active_pct": float(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])}
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(pw) if pw else None,
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_reference_run(steps, warmup, videos_per_step, frames, T, threads=None):
    """The reference's CPU PyTorch path (fp32 eager) restated by oracle/qformer_oracle.py -- LAVIS itself is an absent
    dependency of the reference -- timed on this box's host cores.  Returns (clips_per_s, ms_per_step, cores, sample)."""
    import torch
    from oracle import qformer_oracle as qo
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would otherwise time the
    # reference on ONE core)
    if not threads:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    g = torch.Generator().manual_seed(1234)
    state = {}
    for m, (Nk, W) in MODAL.items():
        cfg = qo.QFormerOracleConfig(encoder_width=W)
        state[m] = (cfg, qo.init_qformer_weights(cfg, seed=1234),
                    torch.randn(videos_per_step, frames, Nk, W, generator=g))
    ids = torch.randint(1000, 30000, (videos_per_step, T), generator=g)
    mask = torch.ones(videos_per_step, T, dtype=torch.long)

    def step():
        with torch.no_grad():
            for m in ("video", "audio"):
                cfg, w, feats = state[m]
                qo.xinstructblip_encode(w, cfg, feats, ids, mask)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    clips = videos_per_step * frames
    sample = (f"{videos_per_step} video(s) x {frames} frames per step ({clips} clips), fp32 eager PyTorch, "
              f"{steps} timed steps after {warmup} warm-up")
    return clips / dt, dt * 1e3, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, ms, cores, sample = cpu_reference_run(args.steps, args.warmup, args.ref_videos, args.frames, args.text_len)
    line = {
        "impl": "reference", "metric": "qformer_video_audio_clips_per_sec", "value": v, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, per_step_videos=args.ref_videos, note="CPU reference arm: bounded sample per step"),
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, per_step_videos, note=None):
    c = {"workload": "X-InstructBLIP video+audio Q-Former (12 layers, cross-attn every 2nd) + llm_proj forward, "
                     "BASELINE.json configs[1]",
         "videos_per_gpu_per_step": per_step_videos, "frames": args.frames, "clips_per_gpu_per_step": per_step_videos * args.frames,
         "video_tokens": "257x1408", "audio_tokens": "256x768", "text_len": args.text_len, "queries": NQ, "llm_dim": D_LLM,
         "parallelism": f"dp{args.gpus} (videos sharded across ranks, no data-path collective)",
         "cache": "inputs larger than L2 (285 MB of features + 0.74 GB of weights per step vs 126 MB L2)"}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from mraudio_b200 import _lib
    from mraudio_b200.xinstructblip import XInstructBLIPQFormers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback on the product path)"
    torch.cuda.set_device(local)
    # pin this rank to the CPUs next to its GPU BEFORE any pinned host buffer is allocated (first-touch places the pages
    # on that NUMA node): with 8 ranks pulling 286 MB per step each, un-bound ranks pile their buffers onto one socket
    numa = _lib.bind_to_gpu_cpus(local) if not args.no_numa_bind else None
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torch.distributed.run)"

    B, F, T = args.videos, args.frames, args.text_len
    torch.manual_seed(1234)
    model = XInstructBLIPQFormers(modalities=("video", "audio")).to(dev).eval()
    g = torch.Generator().manual_seed(1234 + rank)
    host_feats = {m: torch.randn(B, F, Nk, W, generator=g).to(torch.bfloat16).pin_memory() for m, (Nk, W) in MODAL.items()}
    ids_h = torch.randint(1000, 30000, (B, T), generator=g).pin_memory()
    mask_h = torch.ones(B, T, dtype=torch.long).pin_memory()
    feats = {m: t.to(dev) for m, t in host_feats.items()}
    ids, mask = ids_h.to(dev), mask_h.to(dev)
    clips = B * F

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

    def set_profile(mode):
        for m in model.modalities:
            getattr(model, f"{m}_Qformer").bert.set_profile_mode(mode)

    def read_profile():
        tot_ms, tot_n = [0.0] * 5, [0] * 5
        for m in model.modalities:
            ms, n = getattr(model, f"{m}_Qformer").bert.read_profile()
            tot_ms = [a + b for a, b in zip(tot_ms, ms)]
            tot_n = [a + b for a, b in zip(tot_n, n)]
        return tot_ms, tot_n

    def device_step():
        with torch.no_grad():
            return model.encode_modalities(feats, ids, mask)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    for _ in range(args.warmup):
        device_step()
    if args.device_only:
        print(json.dumps({"ms_per_step": timed(device_step, args.steps), "launches_per_step": model.last_launches}))
        return
    sampler = ClockSampler(local) if rank == 0 else None
    t_clock0 = time.time()
    # ---- value: device-resident inputs
    ms_dev = timed(device_step, args.steps)
    launches = model.last_launches * args.steps
    # ---- roofline: same K steps with events around every GEMM launch
    set_profile(_lib.PROFILE_DOMINANT)
    ms_instr = timed(device_step, args.steps)
    prof_ms, prof_n = read_profile()
    # ---- e2e: host buffers through the public host API
    pipe = model.host_pipeline(B, F, {m: v[0] for m, v in MODAL.items()}, T, slots=int(os.environ.get("MRA_BENCH_SLOTS", "2")))
    set_profile(_lib.PROFILE_OFF)

    def e2e_steps(n):
        for i in range(n):
            pipe.submit(host_feats, ids_h, mask_h)
        pipe.drain()

    # PCIe probe: pinned host -> device and back, alone on the bus (explains e2e when the step is copy-bound)
    def copy_gbs(src, dst):
        best = 0.0
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dst.copy_(src, non_blocking=True)
            b.record()
            torch.cuda.synchronize()
            best = max(best, src.numel() * src.element_size() / (a.elapsed_time(b) * 1e-3) / 1e9)
        return best
    h2d_gbs = copy_gbs(host_feats["video"], feats["video"])
    slot0 = pipe.slots[0]
    d2h_gbs = copy_gbs(torch.empty_like(slot0.out_host["video"], device=dev), slot0.out_host["video"])

    e2e_steps(max(2, args.warmup))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps(args.steps)          # drain() makes the current stream wait for the last D2H copy
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    t_clock1 = time.time()
    clocks = sampler.stop(t_clock0, t_clock1) if sampler else None
    # ---- per-category breakdown (not part of any headline number)
    set_profile(_lib.PROFILE_ALL)
    for _ in range(2):
        device_step()
    torch.cuda.synchronize()
    all_ms, all_n = read_profile()
    set_profile(_lib.PROFILE_OFF)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    lin = core = kv = 0.0
    for m, (Nk, W) in MODAL.items():
        a, b, c = flops_per_row(Nk, W, T)
        lin, core, kv = lin + a * clips, core + b * clips, kv + c * clips
    # dominant kernel = gemm_tc_kernel (every Linear); per-launch figures are flops-weighted over its launches of a step
    gemm_ms = (prof_ms[0] + prof_ms[1]) / args.steps
    gemm_n = (prof_n[0] + prof_n[1]) // args.steps
    peak = peaks["bf16_tflops_sustained"]
    achieved = lin / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    kv_ms = prof_ms[0] / args.steps
    step_flops = lin + core
    # CPU baseline: rank 0 at N = 1 only (bounded sample of the same workload)
    cpu = None
    if world == 1:
        cpu_v, cpu_ms, cores, sample = cpu_reference_run(6, 1, 4, F, T)   # ~5-10 s of CPU work on a 16-core host
        cpu = {"value": cpu_v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample}
    line = {
        "metric": "qformer_video_audio_clips_per_sec", "value": clips * world / (ms_dev * 1e-3), "unit": "clips/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, per_step_videos=B),
        "videos_per_sec": B * world / (ms_dev * 1e-3),
        "step_tflops_algorithmic": step_flops / 1e12,
        "frac_of_bf16_peak_whole_step": step_flops / (ms_dev * 1e-3) / 1e12 / peak,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": (load_ncu_traffic() or {}).get("cross_kv_video (gemm_tc_kernel, 2-CTA MMA)", {}).get("dram_bytes_per_launch"),
                     "traffic_note": "dram read+write bytes of the largest launch (cross-K/V GEMM, video: 1.42 GB algorithmic) from "
                                     "profiles/r01g_ncu_full_summary.json; other captured launch types in ncu_captures",
                     "ncu_captures": load_ncu_traffic(), "peak_source": peak_src + ", sustained figure (kernel timed inside a long step)",
                     "kernel": "tcgen05 Linear kernels gemm_tc_kernel + gemm_ln_kernel (all launches of a step, flops-weighted)",
                     "launches_per_step": gemm_n, "ms_per_step_in_kernel": gemm_ms, "ms_per_step_instrumented": ms_instr,
                     "algorithmic_tflop_per_step": lin / 1e12,
                     "cross_kv_launch": {"tflop": kv / 1e12, "ms": kv_ms, "achieved": kv / (kv_ms * 1e-3) / 1e12 if kv_ms else 0.0}},
        "cpu_baseline": cpu,
        "e2e": {"value": clips * world / (ms_e2e * 1e-3), "unit": "clips/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                "pcie_probe_gbs": {"h2d": h2d_gbs, "d2h": d2h_gbs},
                "copy_bound_ms_per_step": max(pipe.h2d_bytes / (h2d_gbs * 1e9), pipe.d2h_bytes / (d2h_gbs * 1e9)) * 1e3,
                "host_cpu_binding": numa,
                "api": "XInstructBLIPQFormers.host_pipeline(...).submit(pinned host features) -> pinned host inputs_llm"},
        "gpu_launches": launches,
        "clocks": clocks,
        "breakdown_ms_per_step": {k: v / 2 for k, v in zip(_lib.PROFILE_CATS, all_ms)},
        "breakdown_launches_per_step": {k: v // 2 for k, v in zip(_lib.PROFILE_CATS, all_n)},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--videos", type=int, default=32, help="videos per GPU per step (config 2: 32)")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--text-len", type=int, default=32)
    ap.add_argument("--device-only", action="store_true", help="run warm-up + timed device steps and exit (for ncu)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to the CPUs next to its GPU (A/B)")
    ap.add_argument("--ref-videos", type=int, default=1, help="videos per step of the CPU reference arm (bounded sample)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
