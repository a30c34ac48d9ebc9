import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import convofusion_b200 as cf
from convofusion_b200 import _lib
from convofusion_b200.synthetic import synthetic_clip, to_device
from helpers import SCHED_KW, golden, rel_err, state_dict
s = cf.ConvoFusionSampler(precision="fp32"); s.load_state_dict(state_dict()); s = s.to("cuda:0").eval()
s.scheduler = cf.DDIMScheduler(clip_sample=True, **SCHED_KW); s.num_inference_timesteps = 50
g = golden("sample_ddim50_clip.pt")
syn = to_device(synthetic_clip(1, seed=1235, dyadic=False), "cuda:0")
enc, masks = s.encode_conditions(syn["clip"], syn["uncond_text"], syn["uncond_text_attn"])
init = torch.randn(1, 16, 128, generator=torch.Generator().manual_seed(100)).cuda()
_lib.check(_lib.lib().cfb_set_fp32_tensor_cores(4))
_, rec, _ = s.sample(enc, masks, 1, init, record=True)
print(sys.argv[1], " ".join(f"{rel_err(rec[i].cpu(), g['record'][i]):.3f}" for i in (0, 24, 49)))
