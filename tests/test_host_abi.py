"""CPU-side checks: the C-ABI library loads and exports every symbol include/*.h declares, and the host logic
(reference-compatible builder, state-dict contract, collater, optimizer groups, sharding) behaves like the reference.
No compute calls are made (there is no GPU here)."""
import os
import random
import re

import numpy as np
import pytest
import torch

import thinkdiff_mlre_b200 as td
from oracle import pack_ref
from oracle.golden import load_golden
from thinkdiff_mlre_b200 import _lib
from thinkdiff_mlre_b200.sharding import shard_bounds, shard_sizes
from thinkdiff_mlre_b200.train_step import reference_param_groups

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "thinkdiff_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(td_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes

    h = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/thinkdiff_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signatures out of sync with the header"
    assert _lib.lib().td_version() == 100


def test_no_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = td.ThinkDiffAligner(64, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(3, 64))
    assert _lib.lib().td_device_check() != 0
    assert "sm_100" in _lib.last_error() or "CUDA" in _lib.last_error()


def test_builder_contract():
    class Cfg:
        mm_projector_type = "mlp2x_gelu_t5_norm"
        mm_hidden_size = 128
        hidden_size = 256

    m = td.build_vision_projector(Cfg())
    assert isinstance(m, torch.nn.Sequential) and isinstance(m, td.ThinkDiffAligner)
    assert list(m.state_dict().keys()) == ["0.weight", "0.bias", "2.weight", "2.bias", "3.weight"]
    assert m.state_dict()["0.weight"].shape == (256, 128) and m.state_dict()["2.weight"].shape == (256, 256)
    assert all(p.dtype == torch.float32 for p in m.parameters())
    assert [type(c).__name__ for c in m] == ["Linear", "GELU", "Linear", "T5LayerNorm"]
    assert torch.equal(m[3].weight, torch.ones(256))  # T5LayerNorm init
    for t in ("linear", "identity", "mlp2x_gelu", "mlp3x_gelu_t5_norm", "mlp2x_gelu_rms_norm", "mlp2x_gelu_norm"):
        Cfg.mm_projector_type = t
        with pytest.raises(NotImplementedError):
            td.build_vision_projector(Cfg())
    Cfg.mm_projector_type = "conv"
    with pytest.raises(ValueError, match="Unknown projector type"):
        td.build_vision_projector(Cfg())
    with pytest.raises(ValueError):
        td.ThinkDiffAligner(100, 256)


def test_t5_layernorm_isinstance_and_checkpoint_roundtrip(tmp_path):
    T5LayerNorm = pytest.importorskip("transformers.models.t5.modeling_t5").T5LayerNorm
    m = td.ThinkDiffAligner(64, 128)
    norms = [c for c in m.modules() if isinstance(c, T5LayerNorm)]  # the reference's re-init scan (...embed_decoder_2.py:697-700)
    assert len(norms) == 1
    norms[0].load_state_dict(T5LayerNorm(128).state_dict())
    # the reference saves only requires_grad params under their full names (runner_clip_t5.py:262-288)
    sd = {"mm_projector." + k: v for k, v in m.state_dict().items()}
    torch.save({"model": sd}, tmp_path / "ckpt.pth")
    host = torch.nn.Module()
    host.mm_projector = td.ThinkDiffAligner(64, 128)
    msg = host.load_state_dict(torch.load(tmp_path / "ckpt.pth")["model"], strict=False)
    assert not msg.missing_keys and not msg.unexpected_keys


def test_optimizer_groups_match_reference_rule():
    m = td.ThinkDiffAligner(64, 128)
    decay, no_decay = reference_param_groups(m, 0.05)
    names = {id(p): n for n, p in m.named_parameters()}
    assert sorted(names[id(p)] for p in decay["params"]) == ["0.weight", "2.weight"]
    assert sorted(names[id(p)] for p in no_decay["params"]) == ["0.bias", "2.bias", "3.weight"]
    assert decay["weight_decay"] == 0.05 and no_decay["weight_decay"] == 0.0


def _samples_from_golden(g):
    full = [int(v) for v in g["full_lens"]]
    off = np.concatenate([[0], np.cumsum(full)])
    out = []
    for i in range(len(full)):
        e = torch.from_numpy(g["src_bits"][off[i] : off[i + 1]].view(np.int16).copy()).view(torch.bfloat16)
        ids = [int(v) for v in g["src_ids_flat"][off[i] : off[i + 1]]]
        out.append({"json": {"generated_text": f"sample {i}", "output_token_ids": ids},
                    "model.norm.input_embed.pth": e, "model.norm.output_embed.pth": e})
    return out


@pytest.mark.parametrize("name", ["collater_random_split.npz", "collater_fixed_max.npz", "collater_fixed_max_uncapped.npz"])
def test_flat_collater_reproduces_reference_collater(name):
    """FlatCollater (CPU, worker-safe) + the pack oracle == the reference collater's padded batch, mask and ids."""
    g = load_golden(name)
    bi = {k[3:]: int(v) for k, v in g.items() if k.startswith("bi_")}
    random.seed(int(g["seed"]))
    fb = td.FlatCollater(bi, pin_memory=False)(_samples_from_golden(g))
    assert fb.flat.dtype == torch.bfloat16 and fb.src_row_start.dtype == torch.int64 and fb.lens.dtype == torch.int32
    assert fb.l_max == g["out_mask"].shape[1]
    assert fb.lens.tolist() == g["out_mask"].sum(1).tolist()
    bits = fb.flat.view(torch.int16).numpy().view(np.uint16)
    packed, cu = pack_ref.pack_from_flat(bits, fb.src_row_start.tolist(), fb.lens.tolist())
    padded, mask = pack_ref.unpack_padded(packed, cu, fb.l_max)
    np.testing.assert_array_equal(padded, g["out_embed_bits"])
    np.testing.assert_array_equal(mask, g["out_mask"])
    ids = [list(g["ids_flat"][g["ids_off"][i] : g["ids_off"][i + 1]]) for i in range(len(fb.lens))]
    assert [list(map(int, t)) for t in fb.extras["output_token_ids"]] == ids
    assert fb.extras["embed_key"] == "model.norm.output_embed"


@pytest.mark.parametrize("name", ["collater_random_split.npz", "collater_fixed_max.npz"])
def test_flat_collater_truncate_on_host_is_equivalent(name):
    """truncate_on_host=True ships only the kept rows; the packed / padded result is the same reference-collater output."""
    g = load_golden(name)
    bi = {k[3:]: int(v) for k, v in g.items() if k.startswith("bi_")}
    random.seed(int(g["seed"]))
    fb = td.FlatCollater(bi, pin_memory=False, truncate_on_host=True)(_samples_from_golden(g))
    assert fb.flat.shape[0] == int(fb.lens.sum())  # nothing but kept rows crosses PCIe
    assert fb.src_row_start.tolist() == [0] + torch.cumsum(fb.lens.long(), 0)[:-1].tolist()
    bits = fb.flat.view(torch.int16).numpy().view(np.uint16)
    packed, cu = pack_ref.pack_from_flat(bits, fb.src_row_start.tolist(), fb.lens.tolist())
    padded, mask = pack_ref.unpack_padded(packed, cu, fb.l_max)
    np.testing.assert_array_equal(padded, g["out_embed_bits"])
    np.testing.assert_array_equal(mask, g["out_mask"])
    np.testing.assert_array_equal(packed, bits)  # already compact


def test_flat_collater_input_branch_and_errors():
    g = load_golden("collater_input_embed.npz")
    bi = {k[3:]: int(v) for k, v in g.items() if k.startswith("bi_")}
    fb = td.FlatCollater(bi, pin_memory=False, which="input")(_samples_from_golden(g))
    assert fb.l_max == g["in_mask"].shape[1] and fb.lens.tolist() == g["in_mask"].sum(1).tolist()
    with pytest.raises(ValueError, match="No input or output embeds"):
        td.FlatCollater(dict(use_input_embed=0, use_output_embed=0))
    with pytest.raises(ValueError):  # the reference's randint(1, 0) on a length-1 sample
        td.kept_lengths([1], dict(random_split_output_embed=1, output_embed_max_split_len=4))


def test_flat_collater_dual_stream_and_pass_through_keys():
    """use_input_embed and use_output_embed both set: the reference returns both padded tensors from one call, and passes the
    `gpt` / `revised_generated_text` lists through when the samples carry them (...mllama_embed_2.py:38-46, :64-99, :172-181)."""
    g = load_golden("collater_input_embed.npz")
    bi = {k[3:]: int(v) for k, v in g.items() if k.startswith("bi_")}
    assert bi["use_input_embed"] and bi["use_output_embed"]
    samples = _samples_from_golden(g)
    for i, s_ in enumerate(samples):
        s_["json"]["gpt"] = f"answer {i}"
        s_["json"]["revised_generated_text"] = f"revised {i}"
    random.seed(int(g["seed"]))
    fb = td.FlatCollater(bi, pin_memory=False)(samples)
    fin = fb.extras["input_batch"]
    for batch, embed_bits, mask, key in ((fb, g["out_embed_bits"], g["out_mask"], "model.norm.output_embed"),
                                         (fin, g["in_embed_bits"], g["in_mask"], "model.norm.input_embed")):
        bits = batch.flat.view(torch.int16).numpy().view(np.uint16)
        packed, cu = pack_ref.pack_from_flat(bits, batch.src_row_start.tolist(), batch.lens.tolist())
        padded, m = pack_ref.unpack_padded(packed, cu, batch.l_max)
        np.testing.assert_array_equal(padded, embed_bits)
        np.testing.assert_array_equal(m, mask)
        assert batch.extras["embed_key"] == key
    assert fb.extras["mask_key"] == "output_embed_mask" and fin.extras["mask_key"] == "input_embed_mask"
    assert fb.extras["llava_gpts"] == [f"answer {i}" for i in range(len(samples))]
    assert fb.extras["revised_generated_texts"] == [f"revised {i}" for i in range(len(samples))]
    plain = td.FlatCollater(bi, pin_memory=False, which="output")(_samples_from_golden(g))
    assert "llava_gpts" not in plain.extras and "revised_generated_texts" not in plain.extras and "input_batch" not in plain.extras


def test_synthetic_batch_shape_contract():
    b = td.synthetic_lvlm_batch(64, 256, 128, 64, seed=1234, pin=False)
    assert b.flat.shape[1] == 128 and b.extras["flat_target"].shape == (b.flat.shape[0], 64)
    full = torch.diff(torch.cat([b.src_row_start, torch.tensor([b.flat.shape[0]])]))
    assert torch.equal(full, b.lens.long() + 1 + torch.arange(64) % 32)  # L_i = len_i + 1 + (i mod 32)
    assert 1 <= int(b.lens.min()) and int(b.lens.max()) <= 256 and b.l_max == int(b.lens.max())


def test_shard_bounds():
    for n, w in ((512, 8), (1024, 8), (10, 4), (3, 8), (0, 2)):
        sizes = shard_sizes(n, w)
        assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_peer_fold_struct_matches_the_header_layout():
    """ctypes mirror of `td_peer_fold` (include/thinkdiff_b200.h): field order, natural alignment, 40 bytes on LP64."""
    import ctypes as C
    import re

    from thinkdiff_mlre_b200 import _lib as L

    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "thinkdiff_b200.h")).read()
    body = re.search(r"typedef struct td_peer_fold \{(.*?)\} td_peer_fold;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [re.search(r"(\w+)\s*$", decl.strip()).group(1) for decl in body.split(";") if decl.strip()]
    assert names == [f[0] for f in L.PeerFold._fields_]
    assert C.sizeof(L.PeerFold) == 40
    assert L.PeerFold.signal_slot.offset == 8 and L.PeerFold.small_dst.offset == 16 and L.PeerFold.small_numel.offset == 32



HOST_ONLY = {"td_last_error", "td_version", "td_device_check", "td_profile_enable", "td_profile_report", "td_profile_timeline",
             "td_gemm_schedule", "td_scatter_tile_owner", "td_step_ctl_bytes", "td_gemm_workspace_bytes", "td_peer_free", "td_peer_close"}


def test_every_device_entry_point_refuses_cleanly_without_a_gpu():
    """No CPU path behind the C ABI: on a host without an sm_100 device every entry point that would enqueue device work returns
    a negative status and leaves a message in td_last_error() -- with all-zero arguments, i.e. before touching a single pointer."""
    import ctypes as C

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.lib()
    called = 0
    for name, (res, args) in _lib.SIGNATURES.items():
        if name in HOST_ONLY or name.endswith("_bytes"):
            continue
        zero = []
        for a in args:
            if a in (C.c_void_p, C.c_char_p) or (isinstance(a, type) and issubclass(a, C._Pointer)):
                zero.append(None)
            elif a is C.c_float:
                zero.append(0.0)
            else:
                zero.append(0)
        rc = getattr(lib, name)(*zero)
        assert rc < 0, f"{name} returned {rc} on a host without a GPU"
        assert _lib.last_error(), name
        called += 1
    assert called >= 30


def test_workspace_size_functions_are_pure_and_consistent():
    """The *_bytes entry points are host arithmetic (callable anywhere): 256-byte aligned, monotone in M, and ordered the way the
    calls nest (the fused MSE forward carves the plain forward's workspace plus its y buffer; the lm_head call carves a loss
    workspace plus a GEMM workspace)."""
    lib = _lib.lib()
    din, d = 3584, 4096
    prev = None
    for m in (1, 255, 256, 8451, 65536):
        f, b, mse = lib.td_aligner_fwd_workspace_bytes(m, din, d), lib.td_aligner_bwd_workspace_bytes(m, din, d), lib.td_aligner_mse_fwd_workspace_bytes(m, din, d)
        assert f > 0 and b > 0 and f % 256 == 0 and b % 256 == 0
        assert mse >= f + 2 * m * d
        assert lib.td_rmsnorm_bwd_workspace_bytes(m, d) > 0 and lib.td_loss_workspace_bytes(m) > 0
        assert lib.td_lm_head_ce_workspace_bytes(m) >= lib.td_loss_workspace_bytes(m) + lib.td_gemm_workspace_bytes()
        assert lib.td_aligner_norm_partials_bytes(m, d) >= 2 * 4 * d  # at least one fp32 [dg | db2] partial row
        if prev is not None:
            assert f >= prev[0] and b >= prev[1] and mse >= prev[2]
        prev = (f, b, mse)
    assert lib.td_gemm_workspace_bytes() % 256 == 0 and lib.td_step_ctl_bytes() >= 32 and _lib.STEP_CTL_SCALE_OFFSET < lib.td_step_ctl_bytes()


class _ListDataset(torch.utils.data.Dataset):
    def __init__(self, samples):
        self.samples = samples

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        return self.samples[i]


def test_flat_collater_as_collate_fn_of_a_multiprocess_dataloader():
    """The collater's seat in the reference (DataLoader(..., collate_fn=dataset.collater), runner_clip_t5.py:71-79): FlatCollater
    runs inside worker PROCESSES (it never touches CUDA or the library's device calls) and its FlatBatch travels back to the
    trainer intact -- same tensors, lengths and pass-through lists as collating in the main process."""
    g = load_golden("collater_fixed_max.npz")
    bi = {k[3:]: int(v) for k, v in g.items() if k.startswith("bi_")}
    samples = _samples_from_golden(g)
    for i, s in enumerate(samples):
        s["json"]["gpt"] = f"g{i}"
    coll = td.FlatCollater(bi, pin_memory=False, truncate_on_host=True)
    want = [coll(samples[i : i + 3]) for i in range(0, len(samples) - len(samples) % 3, 3)]
    loader = torch.utils.data.DataLoader(_ListDataset(samples), batch_size=3, shuffle=False, drop_last=True, num_workers=2, collate_fn=coll,
                                         multiprocessing_context="spawn", timeout=120)  # spawn: this process is multi-threaded by now
    got = list(loader)
    assert len(got) == len(want) >= 2
    for a, b in zip(want, got):
        assert isinstance(b, td.FlatBatch) and b.l_max == a.l_max and b.lens.tolist() == a.lens.tolist()
        assert torch.equal(a.flat.view(torch.int16), b.flat.view(torch.int16)) and torch.equal(a.src_row_start, b.src_row_start)
        assert a.extras["output_token_ids"] == b.extras["output_token_ids"] and a.extras["llava_gpts"] == b.extras["llava_gpts"]


def test_header_is_plain_c_and_a_c_program_links_the_library(tmp_path):
    """The boundary is a C ABI: include/thinkdiff_b200.h compiles as C11 with -Wall -Werror -pedantic (no C++-isms, no torch or CUDA
    headers), and a C program linked against libthinkdiff_b200.so calls through it -- here the host-only entry points and, without
    a GPU, the refusal of a device entry point with its message."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "consumer.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "thinkdiff_b200.h"
int main(void) {
  int32_t plan[8];
  td_peer_fold fold;
  memset(&fold, 0, sizeof fold);
  if (td_version() != 100) return 1;
  if (td_aligner_fwd_workspace_bytes(8451, 3584, 4096) <= 0) return 2;
  if (td_gemm_schedule(-1, 4096, 4096, 74, 1, plan, 8) >= 0) return 3;            /* bad argument: negative status */
  if (td_scatter_tile_owner(0, 4096, 4096, 8, 0, &plan[0], &plan[1]) != 1) return 4;  /* rank 0 starts at owner 1 */
  printf("device_check=%d msg=%s fold=%d\n", (int)td_device_check(), td_last_error(), (int)sizeof fold);
  return 0;
}
''')
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = [gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-l:libthinkdiff_b200.so", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "fold=" in r.stdout
    if not torch.cuda.is_available():
        assert "device_check=-" in r.stdout and ("sm_100" in r.stdout or "CUDA" in r.stdout)
