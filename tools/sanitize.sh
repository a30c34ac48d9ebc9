#!/bin/bash
# compute-sanitizer over the kernel unit tests and a small sampling run (the reference has no race /
# memory checking; the hand-rolled mbarrier / TMEM / TMA protocols of gemm_tc.cu and rowblock.cu are what is checked).
# Usage (on the GPU box):  bash tools/sanitize.sh <outdir>
out=${1:-gpurun_out/sanitize}
mkdir -p "$out"
SAMPLE='import sys, torch; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import convofusion_b200 as cf
from convofusion_b200 import _lib
from convofusion_b200.synthetic import synthetic_clip, to_device
from helpers import state_dict
mask = int(sys.argv[1])
s = cf.ConvoFusionSampler(precision="bf16", num_inference_timesteps=2); s.load_state_dict(state_dict()); s = s.to("cuda:0").eval()
_lib.check(_lib.lib().cfb_set_rowblock(mask))
syn = to_device(synthetic_clip(8, seed=5, dyadic=True), "cuda:0")
init = torch.randn(8, 16, 128, generator=torch.Generator().manual_seed(6)).cuda()
out = s.generate(syn["clip"], syn["uncond_text"], syn["uncond_text_attn"], [128] * 8, init, use_graph=False)
torch.cuda.synchronize(); print("sample ok", float(out["m_rst"].abs().mean()))'
for tool in memcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py -x -q -k "${SAN_K:-not writer}" > "$out/${tool}_kernels.log" 2>&1
  echo "exit=$?" >> "$out/${tool}_kernels.log"
  for mask in ${SAN_MASKS:-0 7}; do
    timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -c "$SAMPLE" $mask > "$out/${tool}_sample_rowblock${mask}.log" 2>&1
    echo "exit=$?" >> "$out/${tool}_sample_rowblock${mask}.log"
  done
done
grep -H "ERROR SUMMARY\|exit=\|passed\|failed\|sample ok" "$out"/*.log > "$out/summary.txt"
cat "$out/summary.txt"
