"""Several independent batches in flight on one GPU.

One guided sampling pass (`ConvoFusionSampler.generate`) is a chain of ~150 dependent short kernels per step: at the
reference's batch of 64 clips it is bound as much by that chain's latency as by throughput, and the
clips of different batches never interact (clips / streams are the independent units).  `SamplerPool`
therefore runs `lanes` passes side by side: every lane owns a device handle (workspace + captured step graph) over the
SAME packed weights, a CUDA stream and a host thread.  The host thread matters: a thread blocks inside
`cudaGraphLaunch` once its stream's launch queue is full (50 replays of a ~1000-node graph), so a single host thread
would not get the next batch enqueued before the first one has almost drained.

This is the data-parallel path of BASELINE.json configs[4] (4096 clips in batches of 64) inside one GPU; across GPUs
the batches are partitioned by `distributed.shard_range`.
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence

import torch

from .modules import lane


def _record_stream(obj: Any, stream: torch.cuda.Stream) -> None:
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_stream(v, stream)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_stream(v, stream)


class SamplerPool:
    def __init__(self, sampler, lanes: int = 2, chains: Optional[int] = None, affinity: Optional[Sequence[int]] = None):
        """chains: concurrent chains inside each lane's captured step (Denoiser.step_chains) while the pool runs;
        default 3 with two or more lanes (the lanes already supply concurrency: measured 6.18 k vs 6.00 k motion-s/s
        with 6), the library default otherwise.  affinity: host cores the lane threads are pinned to (lane k ->
        affinity[k % len]); several processes per node (one per GPU) then do not migrate over each other."""
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        # (Round 1 refused the opt-in CTA-pair GEMM, CFB_TC_2CTA=1, with several lanes: one such run never finished.  Two
        # pair CTAs per SM can deadlock on the tensor-memory allocation permits; the kernel now keeps one pair CTA per SM
        # -- gemm_tc.cu -- and lanes + pair kernel are covered by tests/test_gpu_zz_cta_pair.py.)
        self.sampler = sampler
        self.lanes = int(lanes)
        # measured on the B200 (profiles/r02_lanes_ab.txt): 3 lanes x 2 chains 7.42 k motion-s/s, 2 x 3 7.17 k, 4 x 1 7.26 k
        self.chains = int(chains) if chains is not None else (0 if lanes <= 1 else 3 if lanes == 2 else 2)
        self.affinity = list(affinity) if affinity else None
        self._streams: Optional[List[torch.cuda.Stream]] = None

    def _device(self) -> torch.device:
        return self.sampler.denoiser._device()      # raises on CPU parameters: there is no CPU path

    def streams(self) -> List[torch.cuda.Stream]:
        if self._streams is None:
            dev = self._device()
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(self.lanes)]
        return self._streams

    def map(self, fn: Callable[[Any, int], Any], items: Sequence[Any],
            on_result: Optional[Callable[[int, Any], None]] = None, keep_results: bool = True) -> List[Any]:
        """Runs fn(item, lane_index) for every item, `lanes` at a time; results in item order.  on_result(i, result)
        is called on the lane's thread (lane stream current) as soon as item i has been enqueued.  With
        `keep_results=False` a result is dropped once on_result has seen it (the returned list holds None): a long
        run that keeps every pass's output alive makes the caching allocator grow, and every cudaMalloc it then issues
        synchronises the device -- measured as ~90 ms stalls of BOTH lanes in bench.py.

        Inside fn the calling thread's current stream is the lane's stream and Denoiser / ConvoFusionVae calls run
        on the lane's handle.  Work already queued on the caller's current stream is visible to every lane, and the
        caller's stream waits for all lanes before `map` returns (no host synchronisation of the device)."""
        dev = self._device()
        streams = self.streams()
        caller = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(caller)
        todo: "queue.SimpleQueue" = queue.SimpleQueue()
        for i, it in enumerate(items):
            todo.put((i, it))
        results: List[Any] = [None] * len(items)
        errors: List[BaseException] = []

        def worker(k: int):
            try:
                torch.cuda.set_device(dev)
                if self.affinity and hasattr(os, "sched_setaffinity"):
                    try:
                        os.sched_setaffinity(0, {self.affinity[k % len(self.affinity)]})   # this thread only
                    except OSError:
                        pass
                with lane(k, chains=self.chains), torch.cuda.stream(streams[k]), torch.no_grad():
                    while not errors:
                        try:
                            i, it = todo.get_nowait()
                        except queue.Empty:
                            return
                        r = fn(it, k)
                        if on_result is not None:
                            on_result(i, r)
                        if keep_results:
                            results[i] = r
                        del r
            except BaseException as exc:      # re-raised on the calling thread
                errors.append(exc)

        n = min(self.lanes, max(1, len(items)))
        if n == 1 and not self.affinity:
            worker(0)
        else:      # (with an affinity list even one lane runs on its own thread: the caller's affinity is left alone)
            threads = [threading.Thread(target=worker, args=(k,), name=f"cfb-lane-{k}") for k in range(n)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        for s in streams:
            caller.wait_stream(s)
        if errors:
            raise errors[0]
        _record_stream(results, caller)      # allocated on a lane's stream, consumed on the caller's
        return results

    def generate_many(self, batches: Iterable[Dict[str, Any]]) -> List[Dict[str, Any]]:
        """`ConvoFusionSampler.generate(**batch)` for every batch (dicts of its keyword arguments)."""
        return self.map(lambda kw, _k: self.sampler.generate(**kw), list(batches))
