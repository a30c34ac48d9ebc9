#!/usr/bin/env python
"""Device timeline of the captured sampling step (CUPTI through torch.profiler): which kernels overlap, how busy
the machine is, and where a chain waits.  Not a benchmark -- CUPTI adds a little time per kernel.

    python tools/timeline.py [--batch 64] [--ddim-steps 6] [--out gpurun_out/timeline.json]
"""
import argparse
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch.profiler import ProfilerActivity, profile

import convofusion_b200 as cf
from convofusion_b200.synthetic import randomize_, synthetic_clip, to_device


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"cfb::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:60]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ddim-steps", type=int, default=6)
    ap.add_argument("--out", default="")
    ap.add_argument("--generate", action="store_true", help="trace the whole pass (conditioning + sampling + VAE decode)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    s = randomize_(cf.ConvoFusionSampler(precision="bf16", num_inference_timesteps=a.ddim_steps), 1234).to(dev).eval()
    syn = to_device(synthetic_clip(a.batch, seed=1234), dev)
    init = torch.randn(a.batch, 16, 128, generator=torch.Generator().manual_seed(77)).to(dev)
    enc, masks = s.encode_conditions(syn["clip"], syn["uncond_text"], syn["uncond_text_attn"])
    if a.generate:
        run = lambda: s.generate(syn["clip"], syn["uncond_text"], syn["uncond_text_attn"], [128] * a.batch, init)
    else:
        run = lambda: s.sample(enc, masks, a.batch, init)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run()
        torch.cuda.synchronize()
    ev = []
    for e in prof.profiler.kineto_results.events():      # raw CUPTI activity records (ns), stream = resource id
        if e.device_type() == torch.autograd.DeviceType.CUDA and e.duration_ns() > 0:
            ev.append((e.name(), e.start_ns() / 1e3, (e.start_ns() + e.duration_ns()) / 1e3, e.device_resource_id()))
    ev = [e for e in ev if "Memcpy" not in e[0] and "Memset" not in e[0]]
    ev.sort(key=lambda e: e[1])
    t0, t1 = ev[0][1], max(e[2] for e in ev)
    wall = t1 - t0
    print(f"{len(ev)} kernels over {wall / 1e3:.3f} ms  ({a.ddim_steps} steps -> {wall / 1e3 / a.ddim_steps:.3f} ms/step under CUPTI)")
    # sweep line: concurrency histogram
    pts = sorted([(e[1], 1) for e in ev] + [(e[2], -1) for e in ev])
    hist, cur, last = defaultdict(float), 0, t0
    for t, d in pts:
        hist[cur] += t - last
        cur, last = cur + d, t
    print("kernels in flight : share of wall time")
    for k in sorted(hist):
        if hist[k] / wall > 0.005:
            print(f"   {k:3d} : {100 * hist[k] / wall:5.1f}%")
    avg = sum(k * v for k, v in hist.items()) / wall
    print(f"average kernels in flight {avg:.2f}; idle {100 * hist[0] / wall:.1f}%")
    by = defaultdict(lambda: [0, 0.0])
    for n, b, e, _ in ev:
        by[short(n)][0] += 1
        by[short(n)][1] += e - b
    tot = sum(v[1] for v in by.values())
    print(f"summed kernel time {tot / 1e3:.3f} ms = {tot / wall:.2f} x wall")
    print("  share  total_us  count  avg_us  kernel")
    for n, (c, d) in sorted(by.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"  {100 * d / tot:5.1f}% {d:9.1f} {c:6d} {d / c:7.2f}  {n}")
    # per-stream busy share and dependent-launch gaps
    streams = defaultdict(list)
    for n, b, e, st in ev:
        streams[st].append((b, e, n))
    print("stream: kernels, busy share, median gap to the previous kernel on the same stream (us)")
    for st, lst in sorted(streams.items(), key=lambda kv: -len(kv[1]))[:14]:
        busy = sum(e - b for b, e, _ in lst)
        gaps = sorted(max(0.0, lst[i][0] - lst[i - 1][1]) for i in range(1, len(lst)))
        med = gaps[len(gaps) // 2] if gaps else 0.0
        p90 = gaps[int(len(gaps) * 0.9)] if gaps else 0.0
        print(f"   {st}: {len(lst):5d} kernels, busy {100 * busy / wall:5.1f}%, gap median {med:.2f} p90 {p90:.2f}")
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        json.dump([{"name": short(n), "ts": b - t0, "dur": e - b, "stream": st} for n, b, e, st in ev], open(a.out, "w"))


if __name__ == "__main__":
    main()
