"""Output writer of the test / demo drivers: the `.npy` layout `quant_eval` and `scripts/visualize.py` consume
(convofusion/models/modeltype/base.py:128-357, the BEAT/DnD branch).

For every sample `<out>/<keyid>/pred.npy` holds the generated motion as key points [length, 63, 3] float32: features
/ 3, fingers re-attached to their wrist, everything re-attached to the root (base.py:204-209 -- computed on the device
by `cfb_keypoints3d`, bit-exact), cut to the sample's length (base.py:176-177); `gt.npy` / `motion_spk.npy` likewise
when ground truth / the speaker's motion are passed (base.py:211-236); attention maps are written as
`<att_name>/att_<t>.npy` per recorded timestep (base.py:243-259; like the reference, the map of batch entry 0 is
written into every sample's directory).

The device -> host copy is asynchronous: key points are produced on the caller's stream, copied on a dedicated copy
stream into a ring of pinned buffers, and turned into files by a background thread, so the sampler's next batch is
not held up by the transfer or the file system (the reference blocks on `.cpu().numpy()` per tensor).  Wav / png /
wordmap side files are outside the hot path.
"""
from __future__ import annotations

import os
import queue
import threading
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib
from .postprocess import keypoints3d

ATT_NAMES = ("att_spk", "att_alsn", "att_tlsn", "att_apb", "att_lsnemb")     # base.py:165


class MotionWriter:
    def __init__(self, output_dir, ring: int = 3):
        self.output_dir = Path(output_dir)
        self.ring = max(2, int(ring))
        self._pinned: List[Dict[str, Tensor]] = [dict() for _ in range(self.ring)]
        self._free: "queue.Queue[int]" = queue.Queue()
        for i in range(self.ring):
            self._free.put(i)
        self._jobs: "queue.Queue" = queue.Queue()
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._errors: List[BaseException] = []
        self._thread = threading.Thread(target=self._drain, name="cfb-writer", daemon=True)
        self._thread.start()
        self.files_written = 0

    # ------------------------------------------------------------------ device side
    def _pin(self, slot: int, name: str, like: Tensor) -> Tensor:
        buf = self._pinned[slot].get(name)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = torch.empty(like.shape, dtype=like.dtype).pin_memory()
            self._pinned[slot][name] = buf
        return buf

    def submit(self, m_rst: Tensor, lengths: Sequence[int], keyids: Sequence[str], m_ref: Optional[Tensor] = None,
               motion_spk: Optional[Tensor] = None, att_maps: Optional[Dict[int, Sequence[Tensor]]] = None) -> None:
        """Queue one batch: m_rst / m_ref / motion_spk [B, T, 189] CUDA tensors, att_maps {t: 5 x [B, layers, 16, M_x]}
        as returned by `_diffusion_reverse` (or None).  Returns as soon as the copies are enqueued."""
        if self._errors:
            raise self._errors[0]
        if m_rst.device.type != "cuda":
            raise _lib.CfbError("MotionWriter.submit needs CUDA tensors: convofusion_b200 has no CPU path")
        if len(lengths) != m_rst.shape[0] or len(keyids) != m_rst.shape[0]:
            raise ValueError("lengths / keyids must have one entry per sample")
        dev = m_rst.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        slot = self._free.get()            # blocks only when `ring` batches are still on their way to disk
        tensors = {"pred": keypoints3d(m_rst)}
        if m_ref is not None:
            tensors["gt"] = keypoints3d(m_ref)
        if motion_spk is not None:
            tensors["motion_spk"] = keypoints3d(motion_spk)
        if att_maps:
            for t, maps in att_maps.items():
                for name, a in zip(ATT_NAMES, maps):
                    tensors[f"{name}/att_{int(t)}"] = a[0].detach().to(torch.float32).contiguous()   # base.py:251
        produced = torch.cuda.current_stream(dev)
        self._copy_stream.wait_stream(produced)
        host = {}
        with torch.cuda.stream(self._copy_stream):
            for name, t in tensors.items():
                buf = self._pin(slot, name, t)
                buf.copy_(t, non_blocking=True)
                t.record_stream(self._copy_stream)
                host[name] = buf
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        self._jobs.put((slot, done, host, [int(l) for l in lengths], [str(k) for k in keyids]))

    # ------------------------------------------------------------------ host side
    def _drain(self):
        while True:
            job = self._jobs.get()
            if job is None:
                return
            slot, done, host, lengths, keyids = job
            try:
                done.synchronize()
                for i, key in enumerate(keyids):
                    sample_dir = self.output_dir / key                       # base.py:171-174
                    os.makedirs(sample_dir, exist_ok=True)
                    for name, buf in host.items():
                        if "/" in name:                                      # attention map of batch entry 0
                            path = sample_dir / (name + ".npy")
                            os.makedirs(path.parent, exist_ok=True)
                            np.save(path, buf.numpy())
                        else:
                            np.save(sample_dir / (name + ".npy"), buf[i, :lengths[i]].numpy())
                        self.files_written += 1
            except BaseException as exc:          # surfaced by the next submit() / close()
                self._errors.append(exc)
            finally:
                self._free.put(slot)
                self._jobs.task_done()

    def flush(self) -> None:
        self._jobs.join()
        if self._errors:
            raise self._errors[0]

    def close(self) -> None:
        self.flush()
        self._jobs.put(None)
        self._thread.join(timeout=10)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
