"""Host side of unbounded synthesis: cutting a long recording into the overlapping windows the sampler consumes.

Mirrors the window bookkeeping of `process_samples` / `process_text` (unbounded_synthesis.py:244-312, 189-241):
128-frame windows (5.12 s at 25 fps) that advance by half a window, so a recording of n_parts * 128 frames gives
2 * n_parts - 1 windows; window k takes 161 mel frames and 8 active-passive bits starting at k/2 of a part, and the
words of the time-stamped transcript that fall into (or straddle the edges of) its time span.  The device work per
window is `ConvoFusionSampler.synthesize_unbounded`.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple, Union

import torch
from torch import Tensor

MOTION_LEN = 128                  # unbounded_synthesis.py:272
WINDOW_SECONDS = MOTION_LEN / 25  # :273
UNCOND_PROMPT = "-" * 10          # convofusion.py:910 / unbounded_synthesis.py:196

Segment = Tuple[Tuple[float, float], str]      # ((start_s, end_s), word)


def window_spans(n_frames: int) -> List[Tuple[float, float]]:
    """(start_s, end_s) of every window of an n_frames recording (unbounded_synthesis.py:275-276, 291)."""
    n_parts = n_frames // MOTION_LEN
    return [((k / 2) * WINDOW_SECONDS, (k / 2 + 1) * WINDOW_SECONDS) for k in range(2 * n_parts - 1)]


def _word_in_window(s: float, e: float, first: bool, a: float, b: float) -> bool:
    """The seven acceptance rules of process_text (unbounded_synthesis.py:203-230) for a word spanning [s, e] and a
    window [a, b]; `first` = the word is the first of its transcript."""
    span, mid = b - a, (a + b) / 2
    return (
        (s >= a and e <= b)                                                                     # inside the window
        or (mid <= e <= b and ((s < a - span / 2 and not first) or (s < a and first)))          # long word ending in the 2nd half
        or (a - 1 <= s < a and b < e <= b + 1)                                                  # covers the window, <= 1 s slack
        or (a <= s <= mid and b <= e <= b + 1)                                                  # starts in the 1st half, ends just after
        or (a - 1 <= s <= a and mid <= e <= b)                                                  # starts just before, ends in the 2nd half
        or (mid < s <= b - 1 and e <= b + 1)                                                    # starts in the 2nd half
        or (s >= a - 1 and a + 2 <= e < mid)                                                    # ends in the 1st half, >= 2 s in
    )


def window_text(transcript: Union[str, Sequence[Segment]], start_s: float, end_s: float) -> str:
    """Text of one stream for one window; the unconditional prompt passes through (unbounded_synthesis.py:196-198)."""
    if isinstance(transcript, str) and transcript == UNCOND_PROMPT:
        return transcript
    picked = [word for i, ((s, e), word) in enumerate(transcript)
              if _word_in_window(float(s), float(e), i == 0, start_s, end_s)]
    return " ".join(picked)


def slice_windows(batch: Dict[str, object], featurise: Callable[[List[str]], Tuple[Tensor, Tensor]]) -> List[Dict[str, object]]:
    """Per-window conditioning for `ConvoFusionSampler.synthesize_unbounded` from a dataloader batch in the layout
    process_samples reads (:247-268): `motion_lsn` [B, n_parts*128, 189] (only its length is used), `melspec_lsn`
    [B, n_parts*160 (+1), 80], `active_passive_lsn` [B, n_parts*8], `lsn_id`, `seg_lsn` / `seg_spk` (time-stamped
    transcripts per stream).  `featurise(texts) -> (T5 last hidden state [B,Lt,768], attention mask [B,Lt])` stands
    for the frozen T5 body, which is outside the hot path."""
    n_frames = batch["motion_lsn"].shape[1]
    n_parts = n_frames // MOTION_LEN
    mel, apb = batch["melspec_lsn"], batch["active_passive_lsn"]
    mel_len, apb_len = mel.shape[1] // n_parts, apb.shape[1] // n_parts          # :278-279
    out = []
    for k, (t0, t1) in enumerate(window_spans(n_frames)):
        texts_lsn = [window_text(seg, t0, t1) for seg in batch["seg_lsn"]]       # :293
        texts_spk = [window_text(seg, t0, t1) for seg in batch["seg_spk"]]       # :305
        hid_l, attn_l = featurise(texts_lsn)
        hid_s, attn_s = featurise(texts_spk)
        out.append({
            "mel_lsn": mel[:, int(k / 2 * mel_len):int((k / 2 + 1) * mel_len) + 1].contiguous(),    # :302 (one extra frame)
            "apb": apb[:, int(k / 2 * apb_len):int((k / 2 + 1) * apb_len)].contiguous(),            # :308
            "lsn_id": list(batch["lsn_id"]),
            "text_lsn": hid_l, "text_lsn_attn": attn_l, "text_spk": hid_s, "text_spk_attn": attn_s,
            "texts": {"lsn": texts_lsn, "spk": texts_spk},
        })
    return out
