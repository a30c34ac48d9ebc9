"""Small bf16 sampling run (8 dyadic clips, 2 DDIM steps, eager launches) for compute-sanitizer:
`compute-sanitizer --tool memcheck python tools/sanitize_sample.py [rowblock_mask]` (see tools/sanitize.sh)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import convofusion_b200 as cf
from convofusion_b200 import _lib
from convofusion_b200.synthetic import synthetic_clip, to_device
from helpers import state_dict

mask = int(sys.argv[1]) if len(sys.argv) > 1 else 0
s = cf.ConvoFusionSampler(precision="bf16", num_inference_timesteps=2)
s.load_state_dict(state_dict())
s = s.to("cuda:0").eval()
_lib.check(_lib.lib().cfb_set_rowblock(mask))
syn = to_device(synthetic_clip(8, seed=5, dyadic=True), "cuda:0")
init = torch.randn(8, 16, 128, generator=torch.Generator().manual_seed(6)).cuda()
out = s.generate(syn["clip"], syn["uncond_text"], syn["uncond_text_attn"], [128] * 8, init, use_graph=False)
torch.cuda.synchronize()
print("sample ok", float(out["m_rst"].abs().mean()))
