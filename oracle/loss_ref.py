"""CPU restatement of the losses on the aligner hot path.  TEST INFRASTRUCTURE ONLY.

* Cross entropy: ``CrossEntropyLoss(ignore_index=-100)(logits.view(-1, V), labels.view(-1))``
  (reference: thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py:241-246, labels built at :577-581 with pad -> -100).
  Mean over rows with label != -100 of ``logsumexp(z) - z[label]``, computed in fp32 (autocast promotes);
  ``dlogits = (softmax - onehot) / n_valid`` on valid rows, 0 on ignored rows; an all-ignored batch gives NaN, as torch does.
* Masked MSE: NOT in the reference (grep ``mse``: none); it is the loss BASELINE.json's north_star adds.
  Oracle adopted: ``F.mse_loss(y[valid].float(), t[valid].float())``
  = sum_valid (y - t)^2 / (n_valid * D); ``dy = 2 (y - t) / (n_valid * D)`` on valid rows, 0 elsewhere.  Parity unpinned by
  the reference; pinned against torch's F.mse_loss in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

IGNORE_INDEX = -100


def cross_entropy_fwd_bwd(logits: np.ndarray, labels: np.ndarray, grad_scale: float = 1.0):
    """logits [R, V] (any float dtype, promoted to fp32), labels [R] int64.  Returns (loss fp32, dlogits fp32 [R, V], n_valid)."""
    z = logits.astype(np.float32)
    valid = labels != IGNORE_INDEX
    n_valid = int(valid.sum())
    zmax = z.max(axis=1, keepdims=True)
    e = np.exp(z - zmax, dtype=np.float32)
    s = e.sum(axis=1, keepdims=True, dtype=np.float32)
    lse = (np.log(s, dtype=np.float32) + zmax)[:, 0]
    safe = np.where(valid, labels, 0)
    picked = z[np.arange(z.shape[0]), safe]
    per_row = np.where(valid, lse - picked, 0.0).astype(np.float32)
    if n_valid == 0:
        return np.float32(np.nan), np.zeros_like(z), 0
    loss = np.float32(per_row.sum(dtype=np.float64) / n_valid)
    d = e / s
    d[np.arange(z.shape[0]), safe] -= 1.0
    d *= (valid[:, None] * (grad_scale / n_valid)).astype(np.float32)
    return loss, d.astype(np.float32), n_valid


def masked_mse_fwd_bwd(y: np.ndarray, t: np.ndarray, row_valid: np.ndarray | None = None, grad_scale: float = 1.0):
    """y, t [M, D]; row_valid [M] bool (None = all rows).  Returns (loss fp32, dy fp32 [M, D], n_valid)."""
    yf, tf = y.astype(np.float32), t.astype(np.float32)
    M, D = yf.shape
    valid = np.ones(M, dtype=bool) if row_valid is None else row_valid.astype(bool)
    n_valid = int(valid.sum())
    if n_valid == 0:
        return np.float32(np.nan), np.zeros_like(yf), 0
    diff = (yf - tf) * valid[:, None]
    loss = np.float32((diff.astype(np.float64) ** 2).sum() / (n_valid * D))
    dy = (2.0 * grad_scale / (n_valid * D)) * diff
    return loss, dy.astype(np.float32), n_valid
