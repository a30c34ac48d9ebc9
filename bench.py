#!/usr/bin/env python
"""Benchmark of the ConvoFusion sampling hot path (BASELINE.json metric: motion-seconds generated per second).

A "step" is one pass of the hot path over one batch of synthetic clips: conditioning projections (once per clip),
the DDIM loop (50 denoiser evaluations under 7-branch modality guidance + scheduler), VAE decode to joints.
One clip = 128 frames @ 25 fps = 5.12 motion-seconds.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (N>1: launched under torchrun)
  python bench.py --impl reference [...]                         the reference algorithm on the host CPU

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MOTION_S_PER_CLIP = 128 / 25.0
SCHED_KW = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU per step (BASELINE.json configs[1])")
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--dyadic", action="store_true", help="configs[2]: DnD-shaped conditioning")
    ap.add_argument("--windows", type=int, default=0,
                    help="configs[3]: unbounded synthesis, this many serial 128-frame windows at 50 %% overlap per stream "
                         "(latent inpainting of the previous window + decode per window); 0 = bounded clips")
    ap.add_argument("--in-flight", type=int, default=2,
                    help="independent batches kept in flight on one GPU (one sampler handle + stream each); every step "
                         "is still one full pass over one batch of --batch clips, steps of different handles overlap")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the GEMM micro-measurement (profiling runs)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ flops
def denoiser_flops(n_clips, n_branch, mem_len=(32, 161, 32, 8, 1), d=512, ff=1024, L=9, ntok=16, lat=128):
    """FLOPs of one denoiser evaluation (2*m*n*k over every GEMM and attention contraction).
    `executed`: what this implementation issues in the shared-slot plan: per layer the
    full-batch GEMMs in_proj / out_proj / 2 x TimeBlock / shared scores (N = sum of 32-padded memory lengths) /
    shared values (K = sum of 64-padded lengths) / FFN, the conditional row groups (one stream per single-modality
    branch: query projection + fuser block), the per-pair and self attention, plus the per-step memory-side
    pre-projection.  `as_written`: the reference's own count for 7 branches."""
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
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(n_clips, ddim_steps, timed_denoiser_steps, dyadic, threads):
    """The reference algorithm on host cores: oracle port of Denoiser.forward / guidance / DDIM / VAE decode
    (oracle/convofusion_oracle.py, pinned to the reference modules by tests/golden).  Times `timed_denoiser_steps`
    evaluations of the 7*B batch plus one decode and extrapolates the loop linearly to `ddim_steps`."""
    import torch
    import convofusion_b200 as cf
    from convofusion_b200.synthetic import randomize_, synthetic_clip
    from oracle import convofusion_oracle as O
    torch.set_num_threads(threads)
    sd = {k: v for k, v in randomize_(cf.ConvoFusionSampler(precision="fp32"), 1234).state_dict().items()}
    syn = synthetic_clip(n_clips, seed=1234, dyadic=dyadic)
    clip = dict(syn["clip"])
    clip["text_lsn_mask"], clip["text_spk_mask"] = ~clip["text_lsn_attn"].bool(), ~clip["text_spk_attn"].bool()
    with torch.no_grad():
        t0 = time.perf_counter()
        enc, masks = O.assemble_guidance_batch(sd, clip, syn["uncond_text"], ~syn["uncond_text_attn"].bool())
        t_cond = time.perf_counter() - t0
        sch = O.DDIMSchedulerOracle(**SCHED_KW)
        sch.set_timesteps(ddim_steps)
        lat = torch.randn(n_clips, 16, 128, generator=torch.Generator().manual_seed(1))
        den = lambda x, t: O.denoiser_forward(sd, x, t, enc, masks, prefix="denoiser.")
        den(torch.cat([lat] * 7), sch.timesteps[0])          # warm-up
        t0 = time.perf_counter()
        timed_denoiser_steps = min(timed_denoiser_steps, ddim_steps)
        for t in sch.timesteps[:timed_denoiser_steps]:
            eps, _ = den(torch.cat([lat] * 7), t)
            lat = sch.step(O.guidance_combine(eps, 7.5), t, lat, eta=0.0).prev_sample
        t_step = (time.perf_counter() - t0) / timed_denoiser_steps
        t0 = time.perf_counter()
        O.vae_decode(sd, O.latents_to_vae_input(lat.permute(1, 0, 2)), [128] * n_clips, prefix="vae.")
        t_dec = time.perf_counter() - t0
    total = t_cond + ddim_steps * t_step + t_dec
    return {"seconds_per_pass": total, "ms_per_denoiser_step": t_step * 1e3, "decode_ms": t_dec * 1e3,
            "value": n_clips * MOTION_S_PER_CLIP / total}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_clips, timed = 8, 8          # bounded sample: ~5-10 s of host work per step
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_run(n_clips, args.ddim_steps, timed, args.dyadic, cores)
        if i >= args.warmup:
            vals.append(r)
    v = sum(r["value"] for r in vals) / len(vals)
    ms = sum(r["seconds_per_pass"] for r in vals) / len(vals) * 1e3
    sample = (f"{n_clips} of the {args.batch} clips (7x{n_clips} denoiser batch), {timed} of {args.ddim_steps} DDIM steps timed and "
              f"extrapolated linearly, + conditioning + VAE decode; torch fp32, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": "motion_seconds_per_second", "value": v, "unit": "motion-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload as our arm (BASELINE configs[1]: batches of 64 clips); each step times a bounded sample of it
        # (cpu_baseline.sample).  The reference evaluates all 7 guidance branches as written, ours skips the weight-0 one.
        "config": dict(workload_config(args, args.batch), batches_in_flight=1, reference_branches_evaluated=7),
        "cpu_baseline": {"value": v, "unit": "motion-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "motion-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, batch, branches=None):
    if branches is None:
        branches = 6 if args.dyadic else 5
    return {"workload": ("dyadic DnD-shaped" if args.dyadic else "monadic BEAT-shaped") +
            f" config_cf_beatdnd random-init, batch {batch} clips/GPU, {args.ddim_steps} DDIM steps, 7-branch guidance 7.5, "
            "VAE decode to 128x189 joints (BASELINE.json configs[%d])" % (2 if args.dyadic else 1),
            "clips_per_gpu": batch, "ddim_steps": args.ddim_steps, "guidance_branches_evaluated": branches,
            "unbounded_windows": args.windows,
            "memory_tokens": 234 if args.dyadic else 234,
            "l2": "no flush: one pass streams 186 MB of bf16 weights 50x plus >150 MB of activations, above the 126 MB L2"}


def assemble_line(args, *, world, B, F, n_branch, ms_dev, ms_e2e, launches, clocks, parts, roof, single, h2d_bytes,
                  d2h_bytes):
    """The contract's JSON line from the measurements of run_ours (pure host arithmetic: unit-tested on the CPU)."""
    W = args.windows
    clips_total = B * world * args.steps
    motion_s = MOTION_S_PER_CLIP if W == 0 else MOTION_S_PER_CLIP * (W + 1) / 2.0   # 50 % overlap between windows
    value = clips_total * motion_s / (ms_dev * 1e-3)
    e2e_value = clips_total * motion_s / (ms_e2e * 1e-3)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    # B200_PROFILING.md: burst peak for a kernel timed alone (the GEMM mix runs by itself for ~15 ms at full clocks),
    # sustained peak for work timed inside a long step (the whole denoiser step)
    peak_burst = peaks.get("bf16_tflops", 1590.0)
    peak_tf = peaks.get("bf16_tflops_sustained", peak_burst)
    peak_src = ("MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone) (of measured)" if peaks
                else "fallback 1.59 PFLOP/s (of fallback)")
    fl = denoiser_flops(B, n_branch)
    ms_den = parts["loop_ms"] / args.ddim_steps          # `parts` times one sample() call = one window
    line = {
        "metric": "motion_seconds_per_second", "value": value, "unit": "motion-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
        "config": dict(workload_config(args, B), batches_in_flight=F,
                       in_flight=("every step is one full pass over its own batch of %d clips; %d independent batches "
                                  "overlap on the GPU (SamplerPool lanes), see one_batch_in_flight for the latency view"
                                  % (B, F)) if F > 1 else "one batch at a time"),
        "e2e": {"value": e2e_value, "unit": "motion-s/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "one_batch_in_flight": single,
        "ms_per_denoiser_step": ms_den, "pass_split_ms": parts,
        # whole pass (conditioning + decode included) / DDIM steps at the measured throughput: with several batches
        # in flight this is below the single-batch latency figure above
        "ms_per_denoiser_step_at_throughput": ms_dev / args.steps / args.ddim_steps / max(1, W),
        "denoiser_step_tflops": {"executed": fl["executed"] / (ms_den * 1e-3) / 1e12,
                                 "reference_equivalent": fl["as_written"] / (ms_den * 1e-3) / 1e12,
                                 "executed_gflop_per_step": fl["executed"] / 1e9},
    }
    if roof:
        tf, gemm_ms, n_gemm = roof
        traffic = None
        tfile = ROOT / "profiles" / "r01_traffic.json"
        if tfile.exists():
            traffic = json.loads(tfile.read_text())
        line["roofline"] = {"bound": "tensor", "kernel": "gemm_tc_tma_kernel (tcgen05, TMA store / L2 reduce-add epilogue)",
                            "achieved": tf, "peak": peak_burst, "unit": "TFLOP/s", "frac": tf / peak_burst,
                            "frac_of_sustained_peak": tf / peak_tf, "sustained_peak": peak_tf,
                            "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_detail": traffic,
                            "peak_source": peak_src, "launches_timed": n_gemm, "ms_per_72_gemms": gemm_ms,
                            "whole_step_frac": fl["executed"] / (ms_den * 1e-3) / 1e12 / peak_tf,
                            "whole_step_frac_at_throughput":
                                fl["executed"] * args.ddim_steps * max(1, W) / (ms_dev / args.steps * 1e-3) / 1e12 / peak_tf}
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        r = cpu_reference_run(8, args.ddim_steps, 8, args.dyadic, cores)
        line["cpu_baseline"] = {"value": r["value"], "unit": "motion-s/s", "cores": cores, "kind": "port",
                                "sample": f"8 of the {B} clips (7x8 denoiser batch), 8 of {args.ddim_steps} DDIM steps timed and "
                                          "extrapolated linearly + conditioning + decode; oracle port, torch fp32",
                                "ms_per_denoiser_step": r["ms_per_denoiser_step"]}
    return line


# ------------------------------------------------------------------------------------------ our arm
def gemm_roofline(torch, lib_mod, batch, n_branch, dev, iters=20):
    """Device time of the dominant kernel (gemm_tc_kernel) over exactly the GEMM shape mix of one denoiser
    evaluation, CUDA events on the launching stream; weights of all 9 layers rotate so operands exceed L2."""
    from convofusion_b200 import _lib
    R, d = batch * n_branch * 16, 512
    # (N, K, epilogue) of the eight full-batch GEMMs of one layer in the shared-slot plan: in_proj, self out_proj,
    # time_block1, shared scores, shared values, time_block2, linear1 (GELU), linear2.  "res" = fp32 residual update.
    spec = [(3 * d, d, "bf16"), (d, d, "res"), (d, d, "res"), (320, d, "f32"), (d, 448, "res"), (d, d, "res"),
            (1024, d, "gelu"), (d, 1024, "res")]
    shapes = [(n, k) for n, k, _ in spec]
    Ws = [[torch.randn(n, k, device=dev).bfloat16() for (n, k) in shapes] for _ in range(9)]
    As = {k: torch.randn(R, k, device=dev).bfloat16() for k in (d, 448, 1024)}
    outs = {"bf16": {n: torch.empty(R, n, device=dev, dtype=torch.bfloat16) for n in (3 * d, 1024)},
            "f32": {n: torch.zeros(R, n, device=dev) for n in (d, 320)}}
    side = torch.cuda.Stream(device=dev)
    st = side.cuda_stream

    def one_pass():
        for layer in Ws:
            for (n, k, kind), w in zip(spec, layer):
                obf = kind in ("bf16", "gelu")
                out = outs["bf16" if obf else "f32"][n]
                _lib.check(_lib.lib().cfb_linear(As[k].data_ptr(), 1, w.data_ptr(), 0, out.data_ptr(), int(obf), R, n, k,
                                                 1 if kind == "gelu" else 0, 0, int(kind == "res"), _lib.GEMM_TCGEN05, st))
    # An eager launch through ctypes costs ~12 us of host time, more than most of these kernels run, so the pass is
    # captured into a CUDA graph and replayed: the events below bracket device time only.
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
    ms = e0.elapsed_time(e1) / iters
    flops = 9 * sum(2.0 * R * n * k for (n, k) in shapes)
    return flops / (ms * 1e-3) / 1e12, ms, 9 * len(shapes)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import convofusion_b200 as cf
    from convofusion_b200 import _lib
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
    # guidance branches evaluated: the weight-0 full-cond branch never; the speaker-only branch only for dyadic clips
    # (monadic: it repeats the unconditional branch exactly, convofusion_b200.conditioning.guidance_branches)
    B, n_branch = args.batch, (6 if args.dyadic else 5)

    sampler = randomize_(cf.ConvoFusionSampler(precision=args.precision, num_inference_timesteps=args.ddim_steps), 1234)
    sampler = sampler.to(dev).eval()
    syn = synthetic_clip(B, seed=1234 + rank, dyadic=args.dyadic)     # config 5: seeds 1234 + shard id
    clip, U, Ua = syn["clip"], syn["uncond_text"], syn["uncond_text_attn"]
    init = torch.randn(B, 16, 128, generator=torch.Generator().manual_seed(77 + rank))
    host = {k: v.pin_memory() for k, v in clip.items() if torch.is_tensor(v)}
    host_init = init.pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values()) + host_init.numel() * 4
    out_host = torch.empty(B, 128, 189).pin_memory()
    d2h_bytes = out_host.numel() * 4
    Ud, Uad = U.to(dev), Ua.to(dev)
    dclip = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in clip.items()}
    dinit = init.to(dev)
    lengths = [128] * B
    gathered = [torch.empty(B, 128, 189, device=dev) for _ in range(world)] if world > 1 else None
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()      # the inputs above were copied on the default stream; every pass runs on other streams

    W = args.windows
    inits = [dinit] * W

    def pass_device():
        if W > 0:   # serial windows of B independent streams (unbounded_synthesis.py:244-512)
            out = sampler.synthesize_unbounded([dclip] * W, Ud, Uad, inits, use_graph=not args.no_graph)[-1]
        else:
            out = sampler.generate(dclip, Ud, Uad, lengths, dinit, use_graph=not args.no_graph)["m_rst"]
        if world > 1:
            dist.all_gather(gathered, out)        # the only collective: output motions at the end
        return out

    def pass_e2e():
        c = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        c["lsn_id"], c["spk_is_uncond"] = clip["lsn_id"], clip["spk_is_uncond"]
        x = host_init.to(dev, non_blocking=True)
        if W > 0:
            out = sampler.synthesize_unbounded([c] * W, Ud, Uad, [x] * W, use_graph=not args.no_graph)[-1]
        else:
            out = sampler.generate(c, Ud, Uad, lengths, x, use_graph=not args.no_graph)["m_rst"]
        if world > 1:
            dist.all_gather(gathered, out)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    def timed(fn, warmup, steps, sample_clocks):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            clocks = ClockSampler(local) if sample_clocks else None
            if clocks:
                clocks.start()
            l0 = _lib.lib().cfb_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            launches = _lib.lib().cfb_launch_count() - l0
            ck = clocks.stop() if clocks else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches, ck

    F = max(1, args.in_flight)
    single = None
    if F > 1 and any(k in os.environ for k in ("CUDA_INJECTION64_PATH", "NV_NSIGHT_INJECTION_PORT_BASE")):
        # Nsight Compute serialises every kernel anyway and its injection library did not survive the lane threads'
        # concurrent graph captures (profiles/README.md): profile one lane
        print("bench.py: profiler injection detected, running with one batch in flight", file=sys.stderr)
        F = 1
    if W > 0:
        F = 1      # windows of one stream are serial (preseq + root hand-off)
    if F > 1:
        # Successive steps are independent batches, so F of them are kept in flight (convofusion_b200.SamplerPool:
        # one handle over the same packed weights + one stream + one host thread per lane).  The timed region is
        # still exactly K full passes over K batches of B clips.
        pool = cf.SamplerPool(sampler, lanes=F)
        # e2e: two pinned result buffers per lane; a buffer is waited for only when its turn comes again, so a lane
        # enqueues its next pass while the previous pass's joints are still on their way to the host
        outs_host = [[torch.empty(B, 128, 189).pin_memory() for _ in range(2)] for _ in range(F)]
        landed = [[None, None] for _ in range(F)]
        turn = [0] * F

        def lane_device(_i, k):
            return sampler.generate(dclip, Ud, Uad, lengths, dinit, use_graph=not args.no_graph)["m_rst"]

        def lane_e2e(_i, k):
            c = {kk: v.to(dev, non_blocking=True) for kk, v in host.items()}
            c["lsn_id"], c["spk_is_uncond"] = clip["lsn_id"], clip["spk_is_uncond"]
            x = host_init.to(dev, non_blocking=True)
            out = sampler.generate(c, Ud, Uad, lengths, x, use_graph=not args.no_graph)["m_rst"]
            j = turn[k] & 1
            turn[k] += 1
            if landed[k][j] is not None:
                landed[k][j].synchronize()                 # the joints written two passes ago are on the host
            outs_host[k][j].copy_(out, non_blocking=True)
            landed[k][j] = torch.cuda.Event()
            landed[k][j].record()
            return out

        def run_steps(fn, n):
            outs = pool.map(fn, list(range(n)))
            if world > 1:      # the only collective: output motions, gathered by the main thread once the lanes are
                for o in outs:  # done (a fixed order on every rank; NCCL calls are not issued from the lane threads)
                    dist.all_gather(gathered, o)

        def timed_pool(fn, warmup, steps, sample_clocks):
            with torch.cuda.stream(stream):
                run_steps(fn, max(warmup, 1) * F)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                    torch.cuda.synchronize()
                clocks = ClockSampler(local) if sample_clocks else None
                if clocks:
                    clocks.start()
                l0 = _lib.lib().cfb_launch_count()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run_steps(fn, steps)
                e1.record()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                launches = _lib.lib().cfb_launch_count() - l0
                ck = clocks.stop() if clocks else None
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
            return ms, launches, ck

        ms_dev, launches, clocks = timed_pool(lane_device, args.warmup, args.steps, True)
        ms_e2e, _, _ = timed_pool(lane_e2e, 1, args.steps, False)
        ms_one, _, _ = timed(pass_device, args.warmup, max(2, args.steps // 2), False)   # one batch in flight: the latency view
        single = {"value": B * world * max(2, args.steps // 2) * MOTION_S_PER_CLIP / (ms_one * 1e-3), "unit": "motion-s/s",
                  "ms_per_step": ms_one / max(2, args.steps // 2)}
    else:
        ms_dev, launches, clocks = timed(pass_device, args.warmup, args.steps, True)
        ms_e2e, _, _ = timed(pass_e2e, 1, args.steps, False)

    # split of one pass: conditioning / loop / decode (device events, rank 0 only, untimed extra pass)
    parts = {}
    with torch.cuda.stream(stream):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        enc, masks = sampler.encode_conditions(dclip, Ud, Uad)
        ev[1].record()
        z, _, _ = sampler.sample(enc, masks, B, dinit, use_graph=not args.no_graph)
        ev[2].record()
        sampler.decode(z, lengths)
        ev[3].record()
        torch.cuda.synchronize()
        parts = {"conditioning_ms": ev[0].elapsed_time(ev[1]), "loop_ms": ev[1].elapsed_time(ev[2]),
                 "decode_ms": ev[2].elapsed_time(ev[3])}
        roof = None
        if rank == 0 and args.precision == "bf16" and not args.no_roofline:
            tf, gemm_ms, n_gemm = gemm_roofline(torch, _lib, B, n_branch, dev)
            roof = (tf, gemm_ms, n_gemm)

    if rank == 0:
        line = assemble_line(args, world=world, B=B, F=F, n_branch=n_branch, ms_dev=ms_dev, ms_e2e=ms_e2e, launches=launches,
                             clocks=clocks, parts=parts, roof=roof, single=single, h2d_bytes=h2d_bytes, d2h_bytes=d2h_bytes)
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
