"""Load the reference's OWN functions for the hot path by ast-extracting them from /root/reference and exec'ing them.

The reference modules cannot be imported as modules here (omegaconf, qwen_vl_utils, peft, the forked vLLM and a
symbol removed from transformers 5.x are missing -- SURVEY.md section 8c), but the two functions on the hot path
depend only on torch (+ T5LayerNorm from the installed transformers), so their source segments are compiled as-is.
Nothing is copied into the repo: the source is read where it lies, at run time, in the dev container only.
/root/reference does not exist on the GPU box -- callers must check ``available()``.
"""
from __future__ import annotations

import ast
import os
import random
import re
import types

REFERENCE_ROOT = os.environ.get("THINKDIFF_REFERENCE_ROOT", "/root/reference")
_MODEL_FILE = "thinkdiff/models/mllama_vllm_t5_embed_decoder_2.py"
_DATASET_FILE = "thinkdiff/datasets/datasets/llava_instruct_dataset_mllama_embed_2.py"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _MODEL_FILE))


def _extract(path: str, name: str, cls: str | None = None) -> str:
    with open(os.path.join(REFERENCE_ROOT, path)) as f:
        src = f.read()
    tree = ast.parse(src)
    nodes = tree.body
    if cls is not None:
        nodes = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    node = next(n for n in nodes if isinstance(n, ast.FunctionDef) and n.name == name)
    seg = ast.get_source_segment(src, node)
    if cls is not None:  # method: dedent by its own indentation
        import textwrap

        lines = src.splitlines()[node.lineno - 1 : node.end_lineno]
        seg = textwrap.dedent("\n".join(lines))
    return seg


def load_build_vision_projector():
    """The reference's ``build_vision_projector`` (mllama_vllm_t5_embed_decoder_2.py:41-79), exec'd verbatim."""
    import torch
    from torch import nn
    from transformers.models.t5.modeling_t5 import T5LayerNorm

    class IdentityMap(nn.Module):  # thinkdiff/models/model_utils.py:7-16 (only reached for type 'identity')
        def forward(self, x, *args, **kwargs):
            return x

    ns = {"nn": nn, "re": re, "torch": torch, "T5LayerNorm": T5LayerNorm, "IdentityMap": IdentityMap}
    exec(compile(_extract(_MODEL_FILE, "build_vision_projector"), _MODEL_FILE, "exec"), ns)
    return ns["build_vision_projector"]


def build_reference_projector(mm_hidden_size: int, hidden_size: int, projector_type: str = "mlp2x_gelu_t5_norm"):
    cfg = types.SimpleNamespace(
        mm_projector_type=projector_type, mm_hidden_size=mm_hidden_size, hidden_size=hidden_size
    )
    return load_build_vision_projector()(cfg)


def load_collater():
    """The reference collater (llava_instruct_dataset_mllama_embed_2.py:34-185) as ``collater(build_info, samples)``.

    It draws ``random.randint`` from Python's global ``random`` module: seed it with ``random.seed(s)`` before the
    call to make the split points replayable.
    """
    import torch
    from torch.nn import functional as F

    ns = {"torch": torch, "F": F, "random": random}
    exec(compile(_extract(_DATASET_FILE, "collater", cls="LlavaInstructMllamaEmbedDataset_2"), _DATASET_FILE, "exec"), ns)
    fn = ns["collater"]

    def collater(build_info: dict, samples: list):
        return fn(types.SimpleNamespace(build_info=build_info), samples)

    return collater


_OPTIMS_FILE = "thinkdiff/common/optims.py"


def load_lr_schedulers() -> dict:
    """The reference's LR schedulers (thinkdiff/common/optims.py:13-112), exec'd from the source where it lies with a stub for
    the ``registry`` decorator (the only import of that file besides ``math``). Returns {registered name: class}."""
    import math

    registered = {}

    class _Registry:
        @staticmethod
        def register_lr_scheduler(name):
            def deco(cls):
                registered[name] = cls
                return cls

            return deco

    with open(os.path.join(REFERENCE_ROOT, _OPTIMS_FILE)) as f:
        src = f.read()
    src = src.replace("from thinkdiff.common.registry import registry", "")
    ns = {"math": math, "registry": _Registry}
    exec(compile(src, _OPTIMS_FILE, "exec"), ns)
    return registered


_RUNNER_FILE = "thinkdiff/runners/runner_base.py"


def build_reference_optimizer(model, init_lr: float, weight_decay: float, beta2: float | None = None):
    """The reference runner's ``optimizer`` property (thinkdiff/runners/runner_base.py:98-127), exec'd from the source where it
    lies against a stub runner holding ``model`` and the three ``run_cfg`` values it reads. Returns the ``torch.optim.AdamW``."""
    import contextlib
    import io
    import logging

    import torch

    class _RunCfg(dict):
        __getattr__ = dict.__getitem__

    cfg = _RunCfg(init_lr=init_lr, weight_decay=weight_decay)
    if beta2 is not None:
        cfg["beta2"] = beta2
    ns = {"torch": torch, "logging": logging}
    exec(compile(_extract(_RUNNER_FILE, "optimizer", cls="RunnerBase"), _RUNNER_FILE, "exec"), ns)
    runner = types.SimpleNamespace(model=model, config=types.SimpleNamespace(run_cfg=cfg), _optimizer=None)
    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints every parameter name
        return ns["optimizer"](runner)
