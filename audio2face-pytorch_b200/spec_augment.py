"""SpecAugment time masks for the training forward of the wav2vec2 encoder -- host side.

Replaces `_compute_mask_indices` + the masking branch of ref:src/model/wav2vec.py:25-72,149-162 (call:
`_compute_mask_indices((B, T), config.mask_time_prob, config.mask_time_length, attention_mask=None, min_masks=2)`,
HF `Wav2Vec2Config` defaults mask_time_prob=0.05, mask_time_length=10, mask_feature_prob=0 -> no feature masking).

The reference draws from numpy's GLOBAL generator, on the host, every training step.  To be a drop-in the draw order
is kept identical, so that `np.random.seed(s)` followed by one training forward masks exactly the same frames as the
reference does (pinned by tests/golden/spec_augment.npz, generated from the live reference):

    1. one `np.random.rand()`                        -> number of spans per utterance (probabilistic rounding)
    2. per utterance one `np.random.choice(T - L', n, replace=False)` -> span starts
    3. per utterance whose span union is longer than the shortest union, one
       `np.random.choice(indices, shortest, replace=False)`          -> every row masks the same number of frames

The mask goes to the device as one byte per frame; applying it (and its backward) are the a2f_spec_mask_* kernels.
"""
from __future__ import annotations

import numpy as np

MASK_TIME_PROB = 0.05       # HF Wav2Vec2Config.mask_time_prob
MASK_TIME_LENGTH = 10       # HF Wav2Vec2Config.mask_time_length
MIN_MASKS = 2               # ref:src/model/wav2vec.py:156


def time_mask(batch: int, frames: int, mask_prob: float = MASK_TIME_PROB, span: int = MASK_TIME_LENGTH,
              min_masks: int = MIN_MASKS, rng=np.random) -> np.ndarray:
    """bool [batch, frames]; True = frame replaced by `masked_spec_embed`.  `rng` needs rand() and choice() (numpy's
    global module by default, like the reference)."""
    n_spans = max(min_masks, int(mask_prob * frames / float(span) + rng.rand()))
    span_len = span
    if n_spans == 0:
        # min_masks == 0 and a zero draw; the reference's `lengths[0] = ...` on an empty array raises as well
        raise ValueError("SpecAugment drew zero spans (min_masks must be >= 1)")
    if frames - span_len <= n_spans:        # sequence too short to place the spans without replacement
        span_len = frames - n_spans - 1
    if span_len < 0 or frames - span_len < n_spans:
        raise ValueError(f"sequence of {frames} frames is too short for {n_spans} SpecAugment spans")
    per_row = []
    for _ in range(batch):
        starts = rng.choice(frames - span_len, n_spans, replace=False)
        # spans keep the NOMINAL length even when the start range was shrunk above; frames past the end are dropped
        covered = (np.asarray(starts)[:, None] + np.arange(span)[None, :]).reshape(-1)
        per_row.append(np.unique(covered[covered < frames]))
    keep = min(len(ix) for ix in per_row)
    mask = np.zeros((batch, frames), dtype=bool)
    for b, ix in enumerate(per_row):
        if len(ix) > keep:
            ix = rng.choice(ix, keep, replace=False)
        mask[b, ix] = True
    return mask
