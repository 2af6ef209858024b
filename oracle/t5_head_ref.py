"""CPU restatement of the frozen T5 decoder's output head + loss and of a frozen bias-free Linear (SURVEY.md section 8 f-1,
first slice).  TEST INFRASTRUCTURE ONLY -- the product path never imports this.

Reference (thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py, T5ForDecoder.forward, under the bf16 autocast of
thinkdiff/tasks/base_task.py:237):
    :231     sequence_output = decoder_outputs[0]
    :234-237 (tie_word_embeddings only -- google/flan-t5-xxl is untied) sequence_output *= model_dim ** -0.5
    :239     lm_logits = self.lm_head(sequence_output)            nn.Linear(d_model, vocab, bias=False): bf16 under autocast
    :243-246 loss = CrossEntropyLoss(ignore_index=-100)(lm_logits.view(-1, V), labels.view(-1))   fp32 (autocast promotes)
Backward as autograd forms it: dlogits fp32 -> rounded to bf16 (the cast's backward) -> dseq = bf16(dlogits . W), no dW (frozen,
:715-717).  The cross-attention K / V projections (transformers T5Attention.k / .v, requirements.txt:14) are the same frozen
bias-free Linear.  Pinned by tests/golden/lm_head_ce_small.npz = this expression executed by torch under CPU bf16 autocast
(oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch

from . import loss_ref


def _bf16(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16)


def frozen_linear_fwd(x: np.ndarray, W: np.ndarray) -> np.ndarray:
    """bf16(x_bf16 . W_bf16^T) with fp32 accumulation; returned as fp32 values that are exactly representable in bf16."""
    y = _bf16(x).float() @ _bf16(W).float().t()
    return y.to(torch.bfloat16).float().numpy()


def frozen_linear_dx(dy: np.ndarray, W: np.ndarray) -> np.ndarray:
    dx = _bf16(dy).float() @ _bf16(W).float()
    return dx.to(torch.bfloat16).float().numpy()


def lm_head_ce_fwd_bwd(seq: np.ndarray, W_lm: np.ndarray, labels: np.ndarray, grad_scale: float = 1.0):
    """Returns (loss fp32, dseq fp32-valued bf16 [R, K], logits fp32-valued bf16 [R, V])."""
    logits = frozen_linear_fwd(seq, W_lm)
    loss, dlogits, _ = loss_ref.cross_entropy_fwd_bwd(logits, labels, grad_scale)
    return loss, frozen_linear_dx(dlogits, W_lm), logits
