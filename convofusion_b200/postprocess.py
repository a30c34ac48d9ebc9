"""Output post-processing of the reference's test / demo writers (convofusion/models/modeltype/base.py:204-209)."""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib

NJOINTS = 63


def keypoints3d(feats: Tensor) -> Tensor:
    """[..., 189] decoded joint features -> [..., 63, 3] key points: / 3, fingers re-attached to their wrist (joints
    43.. to joint 11, joints 23..42 to joint 7), everything re-attached to the root -- the arithmetic and order of
    base.py:204-209 (which then writes pred.npy / gt.npy).  Runs on the device; raises for CPU tensors."""
    if feats.device.type != "cuda":
        raise _lib.CfbError("keypoints3d needs a CUDA tensor: convofusion_b200 has no CPU path")
    if feats.shape[-1] != NJOINTS * 3:
        raise ValueError(f"last dimension must be {NJOINTS * 3}, got {tuple(feats.shape)}")
    x = feats.detach().to(torch.float32).contiguous()
    out = torch.empty(*x.shape[:-1], NJOINTS, 3, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().cfb_keypoints3d(x.data_ptr(), x.numel() // (NJOINTS * 3), out.data_ptr(), _lib.stream_ptr()))
    return out
