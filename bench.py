#!/usr/bin/env python
"""bench.py — the reference's headline metric on its headline config, measured on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            # this framework (hand-written sm_100a attention kernels)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host CPU cores (oracle)

Metric (BASELINE.json): 512x512, 50-step images/s.  One "step" = ONE IMAGE of BASELINE.json configs[1]:
SD-v1-4 architecture, 512x512 (64x64 latent), 50 PLMS steps, 2-3 objects, alpha inner-opt ON (3 epochs, each a
differentiable 51-evaluation trajectory + VAE decode + CLIP loss + backward + Adam), batch = 1 per GPU.  Weights are
seeded random tensors of the exact architecture and prompts/layouts/text embeddings are synthetic (no checkpoint,
dataset access or layout predictor offline) — `data: synthetic`.  N > 1: one process per GPU (torchrun), prompts
sharded round-robin, no data-path collective (weights broadcast once at start-up) -> `scaling: weak`.

`value`    images/s with each step's inputs already resident in HBM and the result left on the device.
`e2e`      the same through the public API (SpaceTimeAttnPipeline.generate) with HOST inputs: every step copies its
           conditioning + x_T from pinned host memory and reads the decoded image back to the host.
`roofline` the kernel with the largest share of device time inside the timed region, timed per launch with CUDA
           events on the launching stream; achieved = algorithmic FLOPs per launch / mean launch time; peak = the
           driver-measured MEASURED_PEAKS.json figure (sustained: the kernel runs inside a long step).
`cpu_baseline` the CPU oracle (a port of the reference algorithm) timed on this box's host cores on a bounded sample
           (forward and forward+backward UNet evaluations, extrapolated to one alpha-optimised image).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "images_per_sec_512x512_50step"
UNIT = "images/s"
EVALS_PER_IMAGE = 3 * 51  # 3 alpha epochs x (50 PLMS steps + 1 extra evaluation at the first step)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2, help="timed images per GPU (each ~ seconds)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--ddim_steps", type=int, default=50)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--checkpoint_min_tokens", type=int, default=int(os.environ.get("STA_CKPT_MIN_TOKENS", "0")))
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs (block-level checkpointing, as the reference)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"tflops": d.get("bf16_tflops_sustained") or d["bf16_tflops"], "burst_tflops": d["bf16_tflops"],
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)"}
    return {"tflops": 1400.0, "burst_tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------------
# algorithmic work of one launch (no padding, no recompute)
# ------------------------------------------------------------------------------------------------------
def launch_flops(kind, key):
    if kind.startswith("groupnorm"):
        b, hw, c = key
        return (10.0 if kind.endswith("fwd") else 30.0) * b * hw * c  # elementwise: reported, not a roofline quantity
    if kind in ("sattn_fwd", "sattn_bwd"):
        b, n, h, d = key
        f = 4.0 * n * n * h * d * b
        return f if kind == "sattn_fwd" else 2.0 * f
    B, n, h, d, n_obj = key
    return 4.0 * 77 * n * h * d * (2 + n_obj) * B  # fwd; bwd (no dK/dV, no recompute) has the same count


def launch_bytes(kind, key):
    if kind.startswith("groupnorm"):
        b, hw, c = key
        return (3 if kind.endswith("fwd") else 5) * b * hw * c * 2  # fwd: x twice + y; bwd: (x, dy) twice + dx
    if kind in ("sattn_fwd", "sattn_bwd"):
        b, n, h, d = key
        io = 4 if kind == "sattn_fwd" else 8  # q,k,v,o  |  + do,dq,dk,dv
        return io * b * n * h * d * 2
    B, n, h, d, n_obj = key
    C = h * d
    return 2 * (2 * B * n * C) * 2 + (2 + n_obj) * B * 77 * C * 2 * 2 + B * n_obj * n


def standalone_kernel_ms(kind, key, iters=10):
    """Mean device time of ONE launch of a sta_* kernel at the step's exact geometry: `iters` back-to-back launches
    (one CUDA graph) between two CUDA events on the launching stream, after a warm-up, with an L2 flush (256 MiB
    memset) before."""
    import torch

    from diffusion_spacetime_attn_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).half()
    if kind.startswith("groupnorm"):
        b, hw, c = key
        side = int(hw ** 0.5)
        x = rnd(b, c, side, hw // side).contiguous(memory_format=torch.channels_last)
        gam, bet = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        y, stats, xn = ops.groupnorm_fwd(x, gam, bet, 1e-5, True)
        fn = (lambda: ops.groupnorm_fwd(x, gam, bet, 1e-5, True)) if kind.endswith("fwd") else (
            lambda: ops.groupnorm_bwd(xn, y, gam, bet, stats, 1e-5, True))
    elif kind.startswith("sattn"):
        b, n, h, d = key
        q, k, v, do = rnd(b, n, h * d), rnd(b, n, h * d), rnd(b, n, h * d), rnd(b, n, h * d) * 0.1
        out, lse = ops.sattn_fwd(q, k, v, h)
        fn = (lambda: ops.sattn_fwd(q, k, v, h)) if kind == "sattn_fwd" else (lambda: ops.sattn_bwd(q, k, v, out, lse, do, h))
    else:
        B, n, h, d, n_obj = key
        C = h * d
        q, do = rnd(2 * B, n, C), rnd(2 * B, n, C) * 0.1
        kc, vc = rnd(B, 2 + n_obj, 77, C), rnd(B, 2 + n_obj, 77, C)
        mask = coef = None
        if n_obj:
            from diffusion_spacetime_attn_b200.ldm.modules.attention import build_object_masks

            boxes = [[0.25 + 0.5 * (i % 2), 0.25 + 0.5 * ((i // 2) % 2)] for i in range(n_obj)]
            mask = build_object_masks(boxes, n, "cuda").unsqueeze(0).expand(B, -1, -1).contiguous()
            coef = torch.full((B, n_obj), 2.5, device="cuda")
        out, lse = ops.xattn_fwd(q, kc, vc, mask, coef, h)
        fn = (lambda: ops.xattn_fwd(q, kc, vc, mask, coef, h)) if kind == "xattn_fwd" else (
            lambda: ops.xattn_bwd(q, kc, vc, mask, coef, lse, do, h, out=out))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    # the launches are captured into one CUDA graph (as they run inside the step) so that the host launch rate of
    # this Python process is not part of the measurement (it dominated the small streaming kernels)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(iters):
            fn()
    graph.replay()
    flush.zero_()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    graph.replay()
    b_.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b_) / iters


# ------------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------------
def cpu_oracle_images_per_sec(n_evals: int):
    """Time the CPU oracle (fp32, all host threads) at BASELINE configs[1]'s geometry (batch 2 CFG, 64x64 latent, 2 objects)
    on a bounded sample — `n_evals` forward UNet evaluations and `n_evals` forward+backward evaluations (gradients w.r.t.
    the latent and alpha, torch autograd through the out-of-place restatement; the reference itself cannot back-propagate
    on CPU) — and extrapolate one alpha-optimised image = 3 x 51 forwards + 3 x 50 backwards, WITHOUT the
    reference's per-block recompute (the favourable reading for the CPU):  t_image = 153 t_fwd + 150 (t_fwd+bwd - t_fwd).
    Returns (images/s, t_fwd, t_fwd+bwd, cores)."""
    import torch

    from oracle import sta_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.UNetConfig()
    g = torch.Generator().manual_seed(0)
    p = {}
    for k, shp in O.unet_param_shapes(cfg).items():  # fast init (values do not matter for timing)
        t = torch.empty(shp)
        if k.endswith("weight") and len(shp) > 1:
            t.normal_(0, (1.0 / max(1, int(torch.tensor(shp[1:]).prod()))) ** 0.5, generator=g)
        elif k.endswith("bias"):
            t.zero_()
        else:
            t.fill_(1.0)
        p[k] = t
    x = torch.randn(2, 4, 64, 64, generator=g)
    ctx = torch.randn(2, 77, 768, generator=g)
    locs = [torch.randn(2, 77, 768, generator=g) for _ in range(2)]
    bboxes = [[0.3, 0.5], [0.7, 0.5]]
    t_in = torch.full((2,), 501, dtype=torch.long)
    with torch.no_grad():
        O.unet_forward(x, t_in, ctx, torch.tensor([2.5, 2.5]), bboxes, locs, p, cfg)  # warm-up
        t0 = time.perf_counter()
        for _ in range(n_evals):
            O.unet_forward(x, t_in, ctx, torch.tensor([2.5, 2.5]), bboxes, locs, p, cfg)
        t_f = (time.perf_counter() - t0) / n_evals
    t0 = time.perf_counter()
    for _ in range(n_evals):
        xg = x.clone().requires_grad_(True)
        coef = torch.tensor([2.5, 2.5], requires_grad=True)
        y = O.unet_forward(xg, t_in, ctx, coef, bboxes, locs, p, cfg)
        torch.autograd.grad(y, [xg, coef], torch.ones_like(y))
    t_fb = (time.perf_counter() - t0) / n_evals
    t_image = EVALS_PER_IMAGE * t_f + 3 * 50 * max(t_fb - t_f, 0.0)
    return 1.0 / t_image, t_f, t_fb, cores


def _cpu_sample_text(n, t_f, t_fb):
    return (f"{n} forward and {n} forward+backward UNet evaluation(s) of the CPU oracle (fp32, batch 2, 64x64 latent, 2 objects): "
            f"{t_f:.2f} s and {t_fb:.2f} s each; one alpha-optimised image extrapolated as 153 t_fwd + 150 (t_fwd+bwd - t_fwd), "
            "no per-block recompute (the reference has no working CPU backward; this is the oracle port's autograd)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 1  # one forward + one forward/backward UNet evaluation per "step" (bounded sample of an image's 153 + 150)
    n = min(max(1, args.steps), 6) * per_step  # bounded: ~8 s per step on 16 host cores, the run must end within minutes
    ips, t_f, t_fb, cores = cpu_oracle_images_per_sec(n)
    sample = _cpu_sample_text(n, t_f, t_fb)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / ips, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: SD-v1-4 architecture 512x512, %d PLMS steps, 2-3 objects, "
                               "alpha inner-opt on (%d epochs), batch=1 per GPU" % (args.ddim_steps, args.epochs)},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    from diffusion_spacetime_attn_b200 import native, ops, prompts as P
    from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline, broadcast_weights, shard_prompts

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    native.load()

    pipe = SpaceTimeAttnPipeline(device=f"cuda:{local_rank}", seed=0, steps=args.ddim_steps, num_epochs=args.epochs,
                                 use_checkpoint=True, checkpoint_min_tokens=args.checkpoint_min_tokens,
                                 save_images=False, cuda_graphs=not args.eager, half_weights=not args.eager)
    bcast_bytes = broadcast_weights(pipe.model) + (broadcast_weights(pipe.clip_loss) if world > 1 else 0)

    total_images = args.warmup + 2 * args.steps
    records = P.read_gpt(P.SYNTHETIC_GPT)
    while len(records) < total_images * world:
        records = records + records
    items = P.build_work_items(records)
    mine = [items[i] for i in shard_prompts(len(items), rank, world)][:total_images]
    conds = [pipe.encode([it]) for it in mine]  # pinned host tensors

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also primes autocast weight caches, cuDNN heuristics, the kernels' smem attributes) ----
    for i in range(args.warmup):
        pipe.generate([mine[i]], conds[i], to_host=False)
    barrier()

    sampler = ClockSampler(local_rank)
    # ---- timed region 1: inputs resident in HBM ----
    idx = list(range(args.warmup, args.warmup + args.steps))
    dev_conds = [pipe.to_device(conds[i]) for i in idx]
    torch.cuda.reset_peak_memory_stats()
    barrier()
    launches0 = ops.launch_count()
    runner = pipe.model.graph_runner
    if runner is not None:
        runner.reset_counters()
    else:
        ops.profile_start()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i, dc in zip(idx, dev_conds):
        pipe.generate([mine[i]], dc, to_host=False)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    records_k = ops.profile_stop() if runner is None else []
    hist_graph = runner.launch_histogram() if runner is not None else None
    dev_ms = ev0.elapsed_time(ev1)
    launches = ops.launch_count() - launches0
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
    reserved_mem = torch.cuda.memory_reserved() / 2 ** 30  # includes the graphs' private pools (activation slots)

    # ---- timed region 2: end to end through the public API with host buffers ----
    idx2 = list(range(args.warmup + args.steps, args.warmup + 2 * args.steps))
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    d2h = 0
    for i in idx2:
        img = pipe.generate([mine[i]], conds[i], to_host=True)  # H2D of the step's inputs + D2H of the image inside
        d2h = img.numel() * img.element_size()
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    h2d = pipe.h2d_bytes(conds[idx2[0]])

    times = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(times[0]), float(times[1])
    err = native.device_error()

    # ---- per-kernel device time of the sta_* kernels launched inside timed region 1 (rank-local) ----
    # Eager mode: CUDA events around every launch.  CUDA-graph mode (default): launches inside a replay cannot be
    # bracketed by events, so the launch histogram of the replayed graphs is combined with the mean launch time of
    # each (kernel, geometry) measured right here, standalone, on the same stream (back-to-back launches).
    kernels = []
    if hist_graph:
        for (kind, key), n in hist_graph.items():
            if not (kind.startswith("groupnorm") or kind.startswith("sattn") or kind.startswith("xattn")):
                continue  # token-major streaming kernels (LayerNorm / GEGLU / bias adds): counted in gpu_launches only
            ms = standalone_kernel_ms(kind, key)
            kernels.append({"kernel": "sta_" + kind, "geometry": list(key), "launches": n, "total_ms": ms * n,
                            "mean_us": 1000.0 * ms, "tflops": launch_flops(kind, key) / ms / 1e9,
                            "gbs": launch_bytes(kind, key) / ms / 1e6, "timing": "standalone x launches"})
    else:
        for (kind, key), (n, tot) in ops.kernel_time_summary(records_k).items():
            if not (kind.startswith("groupnorm") or kind.startswith("sattn") or kind.startswith("xattn")):
                continue
            kernels.append({"kernel": "sta_" + kind, "geometry": list(key), "launches": n, "total_ms": tot,
                            "mean_us": 1000.0 * tot / n, "tflops": launch_flops(kind, key) / (tot / n) / 1e9,
                            "gbs": launch_bytes(kind, key) / (tot / n) / 1e6, "timing": "events in step"})
    kernels.sort(key=lambda r: -r["total_ms"])
    peaks = measured_peaks()
    traffic = None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    roofline = roofline_xattn = None
    if kernels:
        top = kernels[0]
        if tf.exists():
            traffic = json.loads(tf.read_text()).get(top["kernel"] + ":" + "x".join(map(str, top["geometry"])))
        attn = [k for k in kernels if "attn" in k["kernel"]]
        top = attn[0] if attn else kernels[0]  # the dominant ATTENTION kernel (the GroupNorm kernels are HBM streams)
        roofline = {"kernel": top["kernel"], "geometry": top["geometry"], "bound": "tensor", "achieved": top["tflops"],
                    "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": top["tflops"] / peaks["tflops"],
                    "traffic": traffic, "peak_source": peaks["source"], "launches": top["launches"],
                    "mean_launch_us": top["mean_us"], "share_of_step": top["total_ms"] / dev_ms,
                    "flops_convention": "algorithmic, no recompute (SURVEY.md 8d); sattn_bwd = 2.0 x fwd"}
        # the fused dual cross-attention (north_star's named kernel) is HBM/latency-bound (38.5*(2+n) FLOP/B): reported
        # against the measured copy bandwidth, with its tensor throughput next to it
        xa = [k for k in kernels if k["kernel"] == "sta_xattn_fwd"]
        if xa:
            x0 = xa[0]
            tr = json.loads(tf.read_text()).get("sta_xattn_fwd:" + "x".join(map(str, x0["geometry"]))) if tf.exists() else None
            roofline_xattn = {"kernel": x0["kernel"], "geometry": x0["geometry"], "bound": "hbm", "achieved": x0["gbs"],
                              "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": x0["gbs"] / peaks["hbm_gbs"], "traffic": tr,
                              "mean_launch_us": x0["mean_us"], "tflops": x0["tflops"],
                              "note": "algorithmic bytes: q + out + projected contexts + masks (SURVEY.md 8d)"}

    if rank == 0:
        n_img = args.steps * world
        line = {
            "metric": METRIC, "value": n_img / (dev_ms / 1000.0), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: SD-v1-4 architecture 512x512, %d PLMS steps, 2-3 objects, "
                                   "alpha inner-opt on (%d epochs), batch=1 per GPU" % (args.ddim_steps, args.epochs),
                       "weights": pipe.weights, "prompts": "synthetic_gpt.txt (gpt.txt record format)",
                       "l2": "inputs larger than L2: each step streams ~3.4 GB of weights 300+ times",
                       "execution": ("CUDA graphs per UNet evaluation, fp16 weights; differentiable evaluations keep their "
                                     "activations in HBM slots (fwd-with-grad graph + bwd graph per slot), recompute graph "
                                     "when no slot is free: %s" % json.dumps(runner.slot_summary())
                                     if pipe.cuda_graphs else "eager, block-level gradient checkpointing"),
                       "parallelism": "dp%d (prompt-sharded, weights broadcast once: %d bytes)" % (world, bcast_bytes)},
            "e2e": {"value": n_img / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_xattn": roofline_xattn,
            "kernels": kernels[:12],
            "peak_mem_gib": peak_mem,
            "reserved_mem_gib": reserved_mem,
            "device_error": err,
        }
        if world == 1 and not args.no_cpu_baseline:
            ips, t_f, t_fb, cores = cpu_oracle_images_per_sec(1)
            line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": _cpu_sample_text(1, t_f, t_fb)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
