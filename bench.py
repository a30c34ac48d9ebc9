#!/usr/bin/env python
"""Benchmark of the ConvoFusion sampling hot path (BASELINE.json metric: motion-seconds generated per second).

A "step" is one pass of the hot path over one batch of synthetic clips: conditioning projections (once per clip),
the DDIM loop (50 denoiser evaluations under 7-branch modality guidance + scheduler), VAE decode to joints.
One clip = 128 frames @ 25 fps = 5.12 motion-seconds.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm, BASELINE configs[1] (N>1: under torchrun)
  python bench.py --impl reference [...]                         the reference algorithm on the host CPU cores
  python bench.py --dyadic                                       configs[2]: DnD-shaped conditioning
  python bench.py --windows 115 [--batch 1|64]                   configs[3]: ~5 min of unbounded synthesis per stream
  python bench.py --sweep 4096 [--gpus N]                        configs[4]: 4096 clips, static shard, seeds 1234+clip_id

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import json
import os
import queue
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MOTION_S_PER_CLIP = 128 / 25.0
SCHED_KW = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=True)
MEM_LEN = (32, 161, 32, 8, 1)      # spk text, lsn audio, lsn text, active-passive bits, listener id


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=18)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step (BASELINE.json configs[1])")
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--dyadic", action="store_true", help="configs[2]: DnD-shaped conditioning")
    ap.add_argument("--windows", type=int, default=0,
                    help="configs[3]: unbounded synthesis, this many serial 128-frame windows at 50 %% overlap per stream "
                         "(latent inpainting of the previous window + decode per window; 115 = ~5 min); 0 = bounded clips")
    ap.add_argument("--sweep", type=int, default=0,
                    help="configs[4]: this many clips in total, statically sharded over the ranks in batches of --batch, "
                         "per-clip seeds 1234 + clip_id; a step is one batch of the shard and --steps is ignored")
    ap.add_argument("--in-flight", type=int, default=3,
                    help="independent batches kept in flight on one GPU (one sampler handle + stream each); every step "
                         "is still one full pass over one batch of --batch clips, steps of different handles overlap")
    ap.add_argument("--chains", type=int, default=-1,
                    help="concurrent chains inside every lane's captured step (-1 = SamplerPool default: 3 with two lanes, "
                         "2 with three or more, the library default of 6 with one batch in flight)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the PyTorch-eager-on-the-B200 baseline")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the kernel micro-measurements (profiling runs)")
    return ap.parse_args()


def n_branches(args):
    """Guidance branches our arm evaluates: never the weight-0 full-cond branch; the speaker-only branch only for dyadic
    clips (monadic: it repeats the unconditional branch exactly, convofusion_b200.conditioning.guidance_branches)."""
    return 6 if args.dyadic else 5


# ------------------------------------------------------------------------------------------ flops
def denoiser_flops(n_clips, n_branch, mem_len=MEM_LEN, d=512, ff=1024, L=9, ntok=16, lat=128):
    """FLOPs of one denoiser evaluation (2*m*n*k over every GEMM and attention contraction).
    `executed`: what this implementation issues in the shared-slot plan: per layer the
    full-batch GEMMs in_proj / out_proj / 2 x TimeBlock / shared scores (N = sum of 32-padded memory lengths) /
    shared values (K = sum of 64-padded lengths) / FFN, the conditional row groups (one stream per single-modality
    branch: query projection + fuser block), the per-pair and self attention, plus the per-step memory-side
    pre-projection.  The padding (scores N = 320 for 234 keys, values K = 448) is ~3 % of `executed`.
    `as_written`: the reference's own count for 7 branches."""
    R = n_clips * n_branch * ntok
    n_tot = sum((m + 31) // 32 * 32 for m in mem_len)
    k_tot = sum((m + 63) // 64 * 64 for m in mem_len)
    M = sum(mem_len)
    full = 2.0 * R * (3 * d * d + d * d + d * d + n_tot * d + d * k_tot + d * d + ff * d + d * ff)
    n_cond_rows = (n_branch - 1) * n_clips * ntok                           # every branch but the first: one conditional stream
    cond = 2.0 * n_cond_rows * 2 * d * d
    self_att = 2.0 * 2 * n_clips * n_branch * ntok * ntok * d
    pair_att = 2.0 * 2 * n_clips * ntok * M * d                             # each clip's own memory, once per stream
    per_step = 2.0 * (n_tot * L * d * d + L * d * k_tot * d)                # Z and Y^T for all layers
    ends = 2.0 * (n_clips * ntok * lat * d + R * d * lat)                   # latent_embd (replicated) + latent_proj
    executed = L * (full + cond + self_att + pair_att) + per_step + ends
    as_written = n_clips * 7 * (L * (212.3e6 + 0.0328e6 * M + 1.0486e6 * M) + 5.2e6)
    return {"executed": executed, "as_written": as_written}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU DURING the timed region, every 200 ms.  Read in-process through NVML
    (pynvml): spawning `nvidia-smi` five times a second from each of eight ranks takes driver locks that stall kernel
    submission for everybody (measured: 49.4 vs 46.2 ms per pass at 8 GPUs).  Falls back to nvidia-smi without pynvml."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nvml, self.handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:      # the CUDA ordinal is an index into CUDA_VISIBLE_DEVICES
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
        return [str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits]

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1 if self.nvml is not None else 0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        reasons = sorted({n for r in self.rows for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------ reference algorithm
class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ReferencePass:
    """The reference algorithm (oracle port of Denoiser.forward / 7-branch guidance / DDIM / VAE decode,
    oracle/convofusion_oracle.py, pinned to the reference modules by tests/golden) on the FULL batch of the workload:
    all 7 guidance branches as written (convofusion.py:499-541).  `run(n)` times conditioning + n denoiser evaluations
    of the 7*B batch + one decode; only the NUMBER of DDIM steps is extrapolated (linearly: every step is the same
    work).  device 'cpu' = the host-core baseline; a CUDA device = PyTorch eager on the B200 (TF32 off)."""

    def __init__(self, n_clips, ddim_steps, dyadic, device="cpu", threads=None, autocast=False):
        import torch
        import convofusion_b200 as cf
        from convofusion_b200.synthetic import randomize_, synthetic_clip
        from oracle import convofusion_oracle as O
        self.torch, self.O, self.n_clips, self.ddim_steps, self.autocast = torch, O, n_clips, ddim_steps, autocast
        self.dev = torch.device(device)
        if self.dev.type == "cpu" and threads:
            torch.set_num_threads(threads)
        if self.dev.type == "cuda":
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
        sd = randomize_(cf.ConvoFusionSampler(precision="fp32"), 1234).state_dict()
        self.sd = {k: v.to(self.dev) for k, v in sd.items()}
        syn = synthetic_clip(n_clips, seed=1234, dyadic=dyadic)
        clip = {k: (v.to(self.dev) if torch.is_tensor(v) else v) for k, v in syn["clip"].items()}
        clip["text_lsn_mask"], clip["text_spk_mask"] = ~clip["text_lsn_attn"].bool(), ~clip["text_spk_attn"].bool()
        self.clip, self.U, self.Um = clip, syn["uncond_text"].to(self.dev), (~syn["uncond_text_attn"].bool()).to(self.dev)
        self.lat0 = torch.randn(n_clips, 16, 128, generator=torch.Generator().manual_seed(1)).to(self.dev)
        self.sch = O.DDIMSchedulerOracle(**SCHED_KW)
        self.sch.set_timesteps(ddim_steps)

    def _sync(self):
        if self.dev.type == "cuda":
            self.torch.cuda.synchronize(self.dev)

    def run(self, timed_steps):
        torch, O = self.torch, self.O
        timed_steps = max(1, min(timed_steps, self.ddim_steps))
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if (self.autocast and self.dev.type == "cuda") else _Null()
        with torch.no_grad(), ctx:
            self._sync()
            t0 = time.perf_counter()
            enc, masks = O.assemble_guidance_batch(self.sd, self.clip, self.U, self.Um)
            self._sync()
            t_cond = time.perf_counter() - t0
            lat = self.lat0
            t0 = time.perf_counter()
            for t in self.sch.timesteps[:timed_steps]:
                eps, _ = O.denoiser_forward(self.sd, torch.cat([lat] * 7), t, enc, masks, prefix="denoiser.")
                lat = self.sch.step(O.guidance_combine(eps.float(), 7.5), t, lat, eta=0.0).prev_sample
            self._sync()
            t_step = (time.perf_counter() - t0) / timed_steps
            t0 = time.perf_counter()
            O.vae_decode(self.sd, O.latents_to_vae_input(lat.permute(1, 0, 2)), [128] * self.n_clips, prefix="vae.")
            self._sync()
            t_dec = time.perf_counter() - t0
        total = t_cond + self.ddim_steps * t_step + t_dec
        return {"seconds_per_pass": total, "ms_per_denoiser_step": t_step * 1e3, "decode_ms": t_dec * 1e3,
                "conditioning_ms": t_cond * 1e3, "value": self.n_clips * MOTION_S_PER_CLIP / total}


def cpu_sample_text(B, ddim_steps, timed, cores):
    return (f"all {B} clips of the batch (7x{B} denoiser batch, 7 branches as written), conditioning + {timed} of {ddim_steps} "
            f"DDIM steps (each step is identical work: only the step count is extrapolated) + VAE decode of the {B} "
            f"clips; oracle port of the reference, torch fp32, {cores} threads")


def workload_config(args, batch):
    """Identical in both arms (the driver compares them); how each arm evaluates it is in `execution`."""
    if args.sweep:
        tag = f"BASELINE.json configs[4]: {args.sweep} clips data-parallel, static shard, seeds 1234+clip_id"
    elif args.windows:
        tag = "BASELINE.json configs[3]: unbounded synthesis"
    else:
        tag = "BASELINE.json configs[%d]" % (2 if args.dyadic else 1)
    return {"workload": ("dyadic DnD-shaped" if args.dyadic else "monadic BEAT-shaped") +
            f" config_cf_beatdnd random-init, batch {batch} clips/GPU, {args.ddim_steps} DDIM steps, 7-branch guidance 7.5, "
            f"VAE decode to 128x189 joints ({tag})",
            "clips_per_gpu": batch, "ddim_steps": args.ddim_steps, "guidance_scale": 7.5,
            "unbounded_windows": args.windows, "sweep_clips": args.sweep, "memory_tokens": sum(MEM_LEN),
            "l2": "no flush: one pass streams 186 MB of bf16 weights 50x plus >150 MB of activations, above the 126 MB L2",
            "reference": "the reference arm (--impl reference) runs the same workload through the oracle port of the "
                         "reference's algorithm on the host cores, all 7 guidance branches as written"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ref = ReferencePass(args.batch, args.ddim_steps, args.dyadic, "cpu", cores)
    # every step = one bounded sample: conditioning + `timed` denoiser evaluations of the full 7x64 batch + decode,
    # sized from a first probe so that the whole --steps/--warmup run stays within a few minutes
    probe = ref.run(1)                                   # also pages in the weights / warms the thread pool
    budget_s = 150.0 / max(1, args.steps + args.warmup)
    timed = int(max(1, min(4, budget_s * 1e3 // max(1.0, probe["ms_per_denoiser_step"]))))
    vals = []
    for i in range(args.warmup + args.steps):
        r = ref.run(timed)
        if i >= args.warmup:
            vals.append(r)
    v = sum(r["value"] for r in vals) / len(vals)
    ms = sum(r["seconds_per_pass"] for r in vals) / len(vals) * 1e3
    print(json.dumps({
        "impl": "reference", "metric": "motion_seconds_per_second", "value": v, "unit": "motion-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.sweep else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.batch),
        "execution": {"guidance_branches_evaluated": 7, "batches_in_flight": 1, "device": "host CPU cores"},
        "ms_per_denoiser_step": sum(r["ms_per_denoiser_step"] for r in vals) / len(vals),
        "cpu_baseline": {"value": v, "unit": "motion-s/s", "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(args.batch, args.ddim_steps, timed, cores)},
        "e2e": {"value": v, "unit": "motion-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def load_peaks():
    pk = ROOT / "MEASURED_PEAKS.json"
    return json.loads(pk.read_text()) if pk.exists() else {}


def assemble_line(args, *, world, B, F, n_branch, ms_dev, ms_e2e, launches, clocks, parts, roof, single, h2d_bytes,
                  d2h_bytes, steps=None, clips_total=None, extra=None):
    """The contract's JSON line from the measurements of run_ours (pure host arithmetic: unit-tested on the CPU)."""
    W = args.windows
    steps = steps if steps is not None else args.steps
    if clips_total is None:
        clips_total = B * world * steps
    motion_s = MOTION_S_PER_CLIP if W == 0 else MOTION_S_PER_CLIP * (W + 1) / 2.0   # 50 % overlap between windows
    value = clips_total * motion_s / (ms_dev * 1e-3)
    e2e_value = clips_total * motion_s / (ms_e2e * 1e-3)
    peaks = load_peaks()
    # B200_PROFILING.md: burst peak for a kernel timed alone (the kernel mix runs by itself for ~15 ms at full clocks),
    # sustained peak for work timed inside a long step (the whole denoiser step)
    peak_burst = peaks.get("bf16_tflops", 1590.0)
    peak_tf = peaks.get("bf16_tflops_sustained", peak_burst)
    peak_src = ("MEASURED_PEAKS.json bf16_tflops (burst; kernels timed alone) (of measured)" if peaks
                else "fallback 1.59 PFLOP/s (of fallback)")
    fl = denoiser_flops(B, n_branch)
    ms_den = parts["loop_ms"] / args.ddim_steps          # `parts` times one sample() call = one window
    line = {
        "metric": "motion_seconds_per_second", "value": value, "unit": "motion-s/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / steps, "higher_is_better": True,
        "scaling": "strong" if args.sweep else "weak",
        "vs_baseline": None, "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
        "config": workload_config(args, B),
        "execution": {"guidance_branches_evaluated": n_branch, "batches_in_flight": F, "chains": getattr(args, "chains", -1),
                      "operands": ("16-bit GEMM / attention operands (bf16; LayerNorm outputs as fp16 unless CFB_BF16_ACT_F16=0), "
                                   "fp32 accumulation, residual, LayerNorm, softmax, guidance, scheduler")
                      if args.precision == "bf16" else "fp32 (three-way bf16 split on the tensor cores unless CFB_FP32_TC=0)",
                      "in_flight": ("every step is one full pass over its own batch of %d clips; %d independent batches "
                                    "overlap on the GPU (SamplerPool lanes), see one_batch_in_flight for the latency view"
                                    % (B, F)) if F > 1 else "one batch at a time"},
        "e2e": {"value": e2e_value, "unit": "motion-s/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "one_batch_in_flight": single,
        "ms_per_denoiser_step": ms_den, "pass_split_ms": parts,
        # whole pass (conditioning + decode included) / DDIM steps at the measured throughput: with several batches
        # in flight this is below the single-batch latency figure above
        "ms_per_denoiser_step_at_throughput": ms_dev / steps / args.ddim_steps / max(1, W),
        "denoiser_step_tflops": {"executed": fl["executed"] / (ms_den * 1e-3) / 1e12,
                                 "reference_equivalent": fl["as_written"] / (ms_den * 1e-3) / 1e12,
                                 "executed_gflop_per_step": fl["executed"] / 1e9},
    }
    if W:
        line["windows_per_second"] = clips_total * W / (ms_dev * 1e-3)
    if roof:
        tf = roof["tflops"]
        traffic = None
        for name in ("r02_traffic.json", "r01_traffic.json"):
            if (ROOT / "profiles" / name).exists():
                traffic = json.loads((ROOT / "profiles" / name).read_text())
                break
        line["roofline"] = {"bound": "tensor", "kernel": roof["kernel"],
                            "achieved": tf, "peak": peak_burst, "unit": "TFLOP/s", "frac": tf / peak_burst,
                            "frac_of_sustained_peak": tf / peak_tf, "sustained_peak": peak_tf,
                            "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_detail": traffic,
                            "peak_source": peak_src, "launches_timed": roof["launches"], "ms_timed": roof["ms"],
                            "concurrent_streams": roof.get("concurrent_streams"),
                            "how": "achieved = 2 M N K summed over the timed launches / CUDA-event time of one graph replay of "
                                   "them, issued with the step's own concurrency (one stream per chain and lane); "
                                   "one_stream_back_to_back is the same launch list on a single stream (each launch "
                                   "alone on the GPU: the kernel's latency-bound floor at 28-56 CTAs per launch)",
                            "one_stream_back_to_back": None if not roof.get("serial") else {
                                "achieved": roof["serial"]["tflops"], "frac": roof["serial"]["tflops"] / peak_burst,
                                "ms_timed": roof["serial"]["ms"], "launches_timed": roof["serial"]["launches"]},
                            "shapes": roof.get("shapes"),
                            "whole_step_frac": fl["executed"] / (ms_den * 1e-3) / 1e12 / peak_tf,
                            "whole_step_frac_at_throughput":
                                fl["executed"] * args.ddim_steps * max(1, W) / (ms_dev / steps * 1e-3) / 1e12 / peak_tf}
        if roof.get("mem"):
            line["roofline_mem"] = roof["mem"]
    if extra:
        line.update(extra)
    return line


# ------------------------------------------------------------------------------------------ kernel micro-measurements
def _graph_time(torch, side, one_pass, iters):
    """Device time of `one_pass` (launches on stream `side`) replayed as a CUDA graph: an eager launch through ctypes
    costs ~12 us of host time, more than most of these kernels run, so the events bracket device time only."""
    with torch.cuda.stream(side):
        for _ in range(3):
            one_pass()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            one_pass()
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_roofline(torch, batch, n_branch, dev, chains, iters=20, lanes=1):
    """Device time of the tcgen05 GEMM kernel family over the GEMM shape mix of one denoiser evaluation AT THE SHAPES
    THE STEP LAUNCHES: every chain's row count (n_batch / chains entries x 16 tokens), the eight full-batch operators
    of a layer, the conditional projections (query projection + fuser block of the chain's conditional stream) and the
    per-step memory-side pre-projection; weights of all 9 layers rotate so operands exceed L2.  Chains run back to
    back here (the step overlaps them on forked streams), so this is the kernel's efficiency at production shapes,
    not the step time."""
    from convofusion_b200 import _lib
    d, L = 512, 9
    n_batch = batch * n_branch
    per = ((n_batch + chains - 1) // chains + 7) // 8 * 8
    rows = [min(per, n_batch - c * per) * 16 for c in range((n_batch + per - 1) // per)]
    n_tot = sum((m + 31) // 32 * 32 for m in MEM_LEN)
    k_tot = sum((m + 63) // 64 * 64 for m in MEM_LEN)
    spec = [(3 * d, d, "bf16"), (d, d, "res"), (d, d, "res"), (n_tot, d, "f32"), (d, k_tot, "res"), (d, d, "res"),
            (1024, d, "gelu"), (d, 1024, "res"),
            (d, d, "bf16"), (d, d, "res")]             # conditional rows: query projection, fuser block
    R = max(rows)
    Ws = [[torch.randn(n, k, device=dev).bfloat16() for (n, k, _) in spec] for _ in range(L)]
    As = {k: torch.randn(R, k, device=dev).bfloat16() for k in (d, k_tot, 1024)}
    outs = {"bf16": {n: torch.empty(R, n, device=dev, dtype=torch.bfloat16) for n in (3 * d, 1024, d)},
            "f32": {n: torch.zeros(R, n, device=dev) for n in (d, n_tot)}}
    Am = torch.randn(max(MEM_LEN), d, device=dev).bfloat16()
    Wz = torch.randn(L * d, d, device=dev).bfloat16()
    Oz = torch.empty(max(MEM_LEN), L * d, device=dev, dtype=torch.bfloat16)
    side = torch.cuda.Stream(device=dev)
    st = side.cuda_stream
    lib = _lib.lib()
    acc = {"flops": 0.0, "n": 0}

    def lin(A, W, out, M, n, k, kind):
        obf = kind in ("bf16", "gelu")
        _lib.check(lib.cfb_linear(A.data_ptr(), 1, W.data_ptr(), 0, out.data_ptr(), int(obf), M, n, k,
                                  1 if kind == "gelu" else 0, 0, int(kind == "res"), _lib.GEMM_TCGEN05, st))
        acc["flops"] += 2.0 * M * n * k
        acc["n"] += 1

    def one_pass():
        acc["flops"], acc["n"] = 0.0, 0
        for x_len in MEM_LEN:                           # memory side: Z and Y^T of the unconditional slot, all layers
            lin(Am, Wz, Oz, x_len, L * d, d, "bf16")
            lin(Am, Wz, Oz, x_len, L * d, d, "bf16")
        for M in rows:
            for layer in Ws:
                for (n, k, kind), w in zip(spec, layer):
                    lin(As[k], w, outs["bf16" if kind in ("bf16", "gelu") else "f32"][n], M, n, k, kind)

    ms = _graph_time(torch, side, one_pass, iters)
    serial = {"tflops": acc["flops"] / (ms * 1e-3) / 1e12, "ms": ms, "launches": acc["n"]}

    # The same launches with the concurrency they have inside the step: every chain of every lane on its own stream
    # (forked from / joined into the captured stream).  The chains of a step are dependent sequences of 5-10 us launches
    # of 28-56 CTAs each; what the step gets out of the kernel is its aggregate rate with `lanes x chains` of them
    # side by side, which is what the whole-step figures are built from.
    n_streams = lanes * len(rows)
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    lane_outs = [{"bf16": {n: torch.empty(R, n, device=dev, dtype=torch.bfloat16) for n in (3 * d, 1024, d)},
                  "f32": {n: torch.zeros(R, n, device=dev) for n in (d, n_tot)}} for _ in range(n_streams)]

    def lin_on(stream, A, W, out, M, n, k, kind):
        obf = kind in ("bf16", "gelu")
        _lib.check(lib.cfb_linear(A.data_ptr(), 1, W.data_ptr(), 0, out.data_ptr(), int(obf), M, n, k,
                                  1 if kind == "gelu" else 0, 0, int(kind == "res"), _lib.GEMM_TCGEN05, stream.cuda_stream))
        acc["flops"] += 2.0 * M * n * k
        acc["n"] += 1

    def concurrent_pass():
        acc["flops"], acc["n"] = 0.0, 0
        for i, s in enumerate(streams):
            s.wait_stream(side)
            M, o = rows[i % len(rows)], lane_outs[i]
            for layer in Ws:
                for (n, k, kind), w in zip(spec, layer):
                    lin_on(s, As[k], w, o["bf16" if kind in ("bf16", "gelu") else "f32"][n], M, n, k, kind)
        for s in streams:
            side.wait_stream(s)

    ms_c = _graph_time(torch, side, concurrent_pass, iters)
    return {"tflops": acc["flops"] / (ms_c * 1e-3) / 1e12, "ms": ms_c, "launches": acc["n"], "concurrent_streams": n_streams,
            "serial": serial,
            "kernel": "gemm_tc_tma_kernel family (tcgen05, TMA store / L2 reduce-add epilogue) at the step's per-chain shapes, "
                      f"{n_streams} chains side by side as in the step ({lanes} batches in flight x {len(rows)} chains)",
            "shapes": {"chain_rows": rows, "per_layer_NK": [(n, k) for n, k, _ in spec],
                       "memory_side": f"{len(MEM_LEN)} streams x 2 x [len, {L * d}, {d}] (serial figure only)"}}


def memory_bound_roofline(torch, batch, n_branch, dev, chains):
    """Achieved GB/s of the memory-bound row kernels at the step's shapes, CUDA events over graph replays, against the
    measured HBM copy peak: LayerNorm rows (fp32 in, bf16 out), the per-step memory normalisation, the fused guidance
    combine + scheduler step."""
    from convofusion_b200 import _lib
    lib = _lib.lib()
    peaks = load_peaks()
    peak = peaks.get("hbm_gbs", 6650.0)
    d = 512
    n_batch = batch * n_branch
    per = ((n_batch + chains - 1) // chains + 7) // 8 * 8
    side = torch.cuda.Stream(device=dev)
    st = side.cuda_stream
    out = {}
    g, b = torch.ones(d, device=dev), torch.zeros(d, device=dev)

    def ln_case(rows, note=None):
        x = torch.randn(rows, d, device=dev)
        y = torch.empty(rows, d, device=dev, dtype=torch.bfloat16)
        ms = _graph_time(torch, side, lambda: _lib.check(lib.cfb_layernorm(x.data_ptr(), g.data_ptr(), b.data_ptr(),
                                                                            y.data_ptr(), 1, rows, d, st)), 50)
        by = rows * d * 6
        r = {"rows": rows, "bytes": by, "us": ms * 1e3, "gbs": by / (ms * 1e-3) / 1e9, "frac": by / (ms * 1e-3) / 1e9 / peak}
        if note:
            r["note"] = note
        return r

    out["ln_rows_chain"] = ln_case(per * 16)
    out["ln_rows_full_batch"] = ln_case(n_batch * 16)
    out["mem_hat_rows"] = ln_case((1 + batch) * sum(MEM_LEN),
                                  "timed through cfb_layernorm on the memory rows (same row-kernel body; mem_hat adds one 2 KB time-embedding row)")
    n = 16 * 128
    eps = torch.randn(n_branch, batch, n, device=dev)
    xl = torch.randn(batch, n, device=dev)
    coef = torch.tensor([0.5, 0.8, 0.9, 0.1, 0.0, 0.0, 0.0, 0.0], device=dev)
    ms = _graph_time(torch, side, lambda: _lib.check(lib.cfb_guidance_sched_step(
        eps.data_ptr(), xl.data_ptr(), 0, coef.data_ptr(), n_branch, batch, n, _lib.SCHED_DDIM, 1, 7.5, st)), 50)
    by = (n_branch + 2) * batch * n * 4
    out["guidance_sched_step"] = {"bytes": by, "us": ms * 1e3, "gbs": by / (ms * 1e-3) / 1e9, "frac": by / (ms * 1e-3) / 1e9 / peak,
                                  "note": "launch-latency sized: %d KB per step" % (by // 1024)}
    return {"bound": "hbm", "peak": peak, "unit": "GB/s",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
            "kernels": out}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import convofusion_b200 as cf
    from convofusion_b200 import _lib
    from convofusion_b200.distributed import batches as shard_batches, clip_seeds, shard_range
    from convofusion_b200.synthetic import randomize_, synthetic_clip

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: convofusion_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, n_branch = args.batch, n_branches(args)
    W = args.windows
    F = max(1, args.in_flight)
    if F > 1 and any(k in os.environ for k in ("CUDA_INJECTION64_PATH", "NV_NSIGHT_INJECTION_PORT_BASE")):
        # Nsight Compute serialises every kernel anyway and its injection library did not survive the lane threads'
        # concurrent graph captures (profiles/README.md): profile one lane
        print("bench.py: profiler injection detected, running with one batch in flight", file=sys.stderr)
        F = 1
    # host cores of this rank: lane threads and the gather thread are pinned inside this set, so that the 2-3 host
    # threads of each of the N processes of a node do not migrate over each other (8-GPU tail)
    cores_all = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    per_rank = max(1, len(cores_all) // max(1, world))
    my_cores = cores_all[local * per_rank:(local + 1) * per_rank] or cores_all

    sampler = randomize_(cf.ConvoFusionSampler(precision=args.precision, num_inference_timesteps=args.ddim_steps), 1234)
    sampler = sampler.to(dev).eval()
    lengths = [128] * B

    # ---- workload: the host batches (pinned) this rank walks through
    def make_batch(seeds):
        if isinstance(seeds, int):
            syn = synthetic_clip(B, seed=seeds, dyadic=args.dyadic)
            clip = syn["clip"]
            init = torch.randn(B, 16, 128, generator=torch.Generator().manual_seed(77 + seeds))
        else:                                   # per-clip seeds (config 5): any sharding generates the same clips
            one = [synthetic_clip(1, seed=s, dyadic=args.dyadic)["clip"] for s in seeds]
            clip = {}
            for k, v0 in one[0].items():
                if torch.is_tensor(v0):
                    clip[k] = torch.cat([c[k] for c in one])
                elif isinstance(v0, list):
                    clip[k] = [c[k][0] for c in one]
                else:
                    clip[k] = v0
            init = torch.cat([torch.randn(1, 16, 128, generator=torch.Generator().manual_seed(77 + s)) for s in seeds])
        host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in clip.items()}
        return host, init.pin_memory()

    syn0 = synthetic_clip(B, seed=1234 + rank, dyadic=args.dyadic)   # the unconditional prompt (one per run)
    Ud, Uad = syn0["uncond_text"].to(dev), syn0["uncond_text_attn"].to(dev)
    if args.sweep:
        lo, hi = shard_range(args.sweep, rank, world)
        spans = shard_batches(lo, hi, B)
        if any(b - a != B for a, b in spans):
            raise ValueError("--sweep needs the shard of every rank to be a multiple of --batch")
        work = [make_batch(clip_seeds(a, b)) for a, b in spans]
        if not args.dyadic:     # monadic clips: the speaker text IS the run's unconditional prompt (dataset.py:185-199)
            for h, _ in work:
                h["text_spk"] = syn0["uncond_text"].unsqueeze(0).repeat(B, 1, 1).pin_memory()
                h["text_spk_attn"] = syn0["uncond_text_attn"].unsqueeze(0).repeat(B, 1).pin_memory()
        n_steps = len(work)
    else:
        work = [make_batch(1234 + rank)]
        n_steps = args.steps
    h2d_bytes = sum(v.numel() * v.element_size() for v in work[0][0].values() if torch.is_tensor(v)) + work[0][1].numel() * 4
    d2h_bytes = B * 128 * 189 * 4
    dwork = [({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in h.items()}, x.to(dev)) for h, x in work]
    stream = torch.cuda.Stream(device=dev)
    comm_stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()      # the inputs above were copied on the default stream; every pass runs on other streams

    def one_pass(clip, x):
        if W > 0:   # serial windows of B independent streams (unbounded_synthesis.py:244-512)
            return sampler.synthesize_unbounded([clip] * W, Ud, Uad, [x] * W, use_graph=not args.no_graph)[-1]
        return sampler.generate(clip, Ud, Uad, lengths, x, use_graph=not args.no_graph)["m_rst"]

    def device_pass(i, k=0):
        clip, x = dwork[i % len(dwork)]
        return one_pass(clip, x)

    # e2e: two pinned result buffers per lane; a buffer is waited for only when its turn comes again, so a lane
    # enqueues its next pass while the previous pass's joints are still on their way to the host
    outs_host = [[torch.empty(B, 128, 189).pin_memory() for _ in range(2)] for _ in range(F)]
    landed = [[None, None] for _ in range(F)]
    turn = [0] * F

    def e2e_pass(i, k=0):
        host, hx = work[i % len(work)]
        c = {kk: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for kk, v in host.items()}
        x = hx.to(dev, non_blocking=True)
        out = one_pass(c, x)
        j = turn[k] & 1
        turn[k] += 1
        if landed[k][j] is not None:
            landed[k][j].synchronize()                 # the joints written two passes ago are on the host
        outs_host[k][j].copy_(out, non_blocking=True)
        landed[k][j] = torch.cuda.Event()
        landed[k][j].record()
        return out

    # ---- the only collective: output motions.  Each pass's all_gather is issued on a communication stream as soon as
    # that pass's joints exist (it overlaps the next pass), in step order on every rank, by one gather thread.
    gathered = [torch.empty(B, 128, 189, device=dev) for _ in range(world)] if world > 1 else None

    class Gatherer:
        def __init__(self, n):
            self.n, self.q, self.pending = n, queue.Queue(), {}
            self.t = threading.Thread(target=self.run, daemon=True)
            if world > 1:
                self.t.start()

        def put(self, i, out):
            if world > 1:
                ev = torch.cuda.Event()
                ev.record()                    # on the producing lane's stream
                self.q.put((i, out, ev))

        def run(self):
            torch.cuda.set_device(dev)
            if hasattr(os, "sched_setaffinity") and my_cores:
                try:
                    os.sched_setaffinity(0, {my_cores[-1]})
                except OSError:
                    pass
            nxt = 0
            while nxt < self.n:
                i, out, ev = self.q.get()
                self.pending[i] = (out, ev)
                while nxt in self.pending:
                    o, e = self.pending.pop(nxt)
                    with torch.cuda.stream(comm_stream):
                        comm_stream.wait_event(e)
                        dist.all_gather(gathered, o)
                        o.record_stream(comm_stream)
                    nxt += 1

        def join(self):
            if world > 1:
                self.t.join()
                torch.cuda.current_stream().wait_stream(comm_stream)

    pool = cf.SamplerPool(sampler, lanes=F, affinity=my_cores, chains=args.chains if args.chains >= 0 else None) if F > 1 else None
    if F == 1 and args.chains >= 0:
        sampler.denoiser.step_chains = args.chains
    use_pool = [pool is not None]

    trace_on = bool(os.environ.get("CFB_BENCH_TRACE"))
    trace_ev = []

    def run_steps(fn, n):
        g = Gatherer(n)

        def done(i, out):
            if trace_on:                       # per-pass completion events (diagnosis of run-to-run variance)
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                trace_ev.append((i, time.perf_counter(), e))
            g.put(i, out)

        if use_pool[0]:
            pool.map(lambda i, k: fn(i, k), list(range(n)), on_result=done, keep_results=False)
        else:
            for i in range(n):
                done(i, fn(i, 0))
        g.join()

    def timed(fn, warmup, steps, sample_clocks):
        with torch.cuda.stream(stream):
            run_steps(fn, max(warmup, 1) * (F if use_pool[0] else 1))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            clocks = ClockSampler(local) if (sample_clocks and rank == 0) else None   # the line is rank 0's
            if clocks:
                clocks.start()
            l0 = _lib.lib().cfb_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t_host0 = time.perf_counter()
            del trace_ev[:]
            run_steps(fn, steps)
            e1.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if trace_on and rank == 0:
                rows = sorted((e0.elapsed_time(e), i, (th - t_host0) * 1e3) for i, th, e in trace_ev)
                sys.stderr.write(f"[trace] {fn.__name__}: total {ms:.1f} ms; pass i: device-done ms (host-enqueued ms): " +
                                 " ".join(f"{i}:{d:.0f}({h:.0f})" for d, i, h in rows) + "\n")
            launches = _lib.lib().cfb_launch_count() - l0
            ck = clocks.stop() if clocks else None
        per_rank_ms = None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            allt = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank_ms = [float(v) for v in allt]
            ms = max(per_rank_ms)
        return ms, launches, ck, per_rank_ms

    ms_dev, launches, clocks, per_rank_ms = timed(device_pass, args.warmup, n_steps, True)
    ms_e2e, _, _, _ = timed(e2e_pass, 2, n_steps, False)
    single = None
    if pool is not None and not args.sweep:
        use_pool[0] = False                              # one batch in flight: the latency view
        k1 = max(2, n_steps // 2)
        ms_one, _, _, _ = timed(device_pass, args.warmup, k1, False)
        use_pool[0] = True
        motion = MOTION_S_PER_CLIP if W == 0 else MOTION_S_PER_CLIP * (W + 1) / 2.0
        single = {"value": B * world * k1 * motion / (ms_one * 1e-3), "unit": "motion-s/s", "ms_per_step": ms_one / k1}

    # split of one pass: conditioning / loop / decode (device events; the first repetition warms this call pattern up)
    roof, extra = None, {}
    with torch.cuda.stream(stream):
        clip0, x0 = dwork[0]
        for rep in range(2):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            enc, masks = sampler.encode_conditions(clip0, Ud, Uad)
            ev[1].record()
            z, _, _ = sampler.sample(enc, masks, B, x0, use_graph=not args.no_graph, spk_is_uncond=not args.dyadic)
            ev[2].record()
            sampler.decode(z, lengths)
            ev[3].record()
            torch.cuda.synchronize()
        parts = {"conditioning_ms": ev[0].elapsed_time(ev[1]), "loop_ms": ev[1].elapsed_time(ev[2]),
                 "decode_ms": ev[2].elapsed_time(ev[3])}
        if rank == 0 and args.precision == "bf16" and not args.no_roofline:
            chains = 6 if F <= 1 else 3 if F == 2 else 2
            roof = gemm_roofline(torch, B, n_branch, dev, chains, lanes=max(1, F))
            roof["mem"] = memory_bound_roofline(torch, B, n_branch, dev, chains)
    if per_rank_ms:
        extra["per_rank_ms"] = per_rank_ms
    if rank == 0 and world == 1 and not args.no_gpu_eager and not W and not args.sweep:
        # PyTorch eager on the same B200 (SURVEY 8d / BASELINE.md 4.5): the oracle port with its tensors on the GPU,
        # TF32 off, and once under bf16 autocast; the full 7 x B batch, 5 of the DDIM steps timed
        try:
            eager = {}
            for tag, ac in (("fp32", False), ("bf16_autocast", True)):
                ref = ReferencePass(B, args.ddim_steps, args.dyadic, dev, autocast=ac)
                ref.run(2)
                r = ref.run(5)
                eager[tag] = {"value": r["value"], "unit": "motion-s/s", "ms_per_denoiser_step": r["ms_per_denoiser_step"],
                              "decode_ms": r["decode_ms"]}
                del ref
            eager["what"] = ("oracle port of the reference (plain torch ops, all 7 branches as written) with its tensors on "
                             "this GPU, TF32 off; conditioning + 5 of the DDIM steps timed (step count extrapolated) + decode")
            extra["gpu_eager_baseline"] = eager
            torch.cuda.empty_cache()
        except Exception as exc:      # a baseline leg must never take the benchmark line down
            extra["gpu_eager_baseline"] = {"unavailable": repr(exc)[:300]}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        ref = ReferencePass(B, args.ddim_steps, args.dyadic, "cpu", cores)
        ref.run(1)
        r = ref.run(2)
        extra["cpu_baseline"] = {"value": r["value"], "unit": "motion-s/s", "cores": cores, "kind": "port",
                                 "sample": cpu_sample_text(B, args.ddim_steps, 2, cores),
                                 "ms_per_denoiser_step": r["ms_per_denoiser_step"]}

    if rank == 0:
        line = assemble_line(args, world=world, B=B, F=F, n_branch=n_branch, ms_dev=ms_dev, ms_e2e=ms_e2e, launches=launches,
                             clocks=clocks, parts=parts, roof=roof, single=single, h2d_bytes=h2d_bytes, d2h_bytes=d2h_bytes,
                             steps=n_steps, clips_total=args.sweep if args.sweep else None, extra=extra)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
