"""CPU: the C-ABI library loads and exports every symbol include/diffab_b200.h declares; the host
mirror refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from diffab_pytorch_b200 import _lib, diffusion, so3
from diffab_pytorch_b200.diffab_pytorch import InvariantPointAttentionLayer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "diffab_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"#ifdef DAB_DEBUG_HOOKS.*?#endif", "", text, flags=re.S)     # debug build only (libdiffab_b200_dbg.so)
    return sorted(set(re.findall(r"\b(dab_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTS), "ctypes table and header disagree"
    assert _lib.lib().dab_version() >= 100
    assert isinstance(_lib.lib().dab_last_error(), bytes)


def test_integration_guide_names_every_entry_point():
    """INTEGRATION.md is the binding surface a reference maintainer reads: every exported entry point must appear there
    next to the reference function it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared_symbols() if n not in doc]
    assert not missing, missing


def test_product_library_has_no_debug_hooks():
    """The process-global profiling hooks live in the debug build only (make -C csrc debug)."""
    handle = ctypes.CDLL(_lib.LIB_PATH)
    assert not any(n.startswith("dab_debug") for n in _declared_symbols())
    for name in _lib.DEBUG_EXPORTS:
        assert not hasattr(handle, name), name


def test_workspace_queries_run_without_a_gpu():
    d = _lib.DabIpaDims(32, 128, 128, 64, 8, 32, 8, 8)
    fwd = _lib.lib().dab_ipa_f32_workspace_bytes(ctypes.byref(d), 0)
    bwd = _lib.lib().dab_ipa_f32_workspace_bytes(ctypes.byref(d), 1)
    rows = 32 * 128
    assert fwd >= rows * (1344 + 1024) * 4
    assert bwd >= 2 * fwd + 2 * 32 * 8 * 128 * 128 * 4


def test_cpu_tensors_are_refused():
    with pytest.raises(RuntimeError, match="no CPU path"):
        so3.vector_to_rotation_matrix(torch.randn(4, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        so3.scale_rot(torch.eye(3).expand(2, 5, 3, 3).contiguous(), torch.rand(2))
    layer = InvariantPointAttentionLayer(32, 16, 16, 4, 4, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer(torch.rand(2, 16, 32), torch.rand(2, 16, 16, 16), torch.rand(2, 16, 3, 3), torch.rand(2, 16, 3))
    sd = diffusion.SequenceDiffuser(T=100, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        sd.forward_prob_single_step(torch.zeros(2, 5, dtype=torch.long), torch.ones(2, dtype=torch.long),
                                    torch.ones(2, 5, dtype=torch.bool))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdiffab_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()
