"""SO(3) maps and IGSO(3) sampling on the GPU - mirror of ``diffab_pytorch/so3.py``.

Same public names and argument meaning as the reference module (``so3.py:9-259``); the arithmetic
runs in ``csrc/so3_kernels.cu`` through the C ABI.  Differences, all listed in DESIGN.md:
CUDA tensors only; the IGSO(3) table is computed on the device at construction (no disk cache,
so3.py:18-50); maps accept any leading dims (the reference's ``log_rotmat`` is 4-D only);
``SO3.sample_isotropic_gaussian`` takes an optional ``noise=`` dict so tests can inject the
reference's draws.
"""
import math

import torch

from . import _lib
from ._lib import ptr


def _flat(t, width, name):
    t = _lib.dev(t, torch.float32, name)
    if width == 9:
        if t.shape[-2:] != (3, 3):
            raise ValueError(f"{name}: last two dims must be (3, 3), got {tuple(t.shape)}")
        return t, t.shape[:-2], t.numel() // 9
    if t.shape[-1] != 3:
        raise ValueError(f"{name}: last dim must be 3, got {tuple(t.shape)}")
    return t, t.shape[:-1], t.numel() // 3


def tensor_trace(T):
    """so3.py:142-143 (plain indexing, no kernel needed)."""
    return T.diagonal(offset=0, dim1=-2, dim2=-1).sum(dim=-1)


def small_matmul(a, b):
    """a @ b for (..., n, 3) x (..., 3, 3).  On the GPU a batch of 3 x 3 products goes through a 64 x 256-tile TF32
    GEMM kernel (>100 us for 8192 rotations, and TF32-rounded under torch.set_float32_matmul_precision("high"));
    the broadcast form is one exact-fp32 elementwise kernel.  CPU tensors keep ``@`` (bit-identical to the reference)."""
    if not a.is_cuda:
        return a @ b
    return (a.unsqueeze(-1) * b.unsqueeze(-3)).sum(-2)


def _exp_torch(v):
    """Differentiable restatement used only by autograd backward passes."""
    n = v.norm(dim=-1)[..., None, None]
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    S = torch.stack([torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], -2)
    eye = torch.eye(3, device=v.device, dtype=v.dtype).expand_as(S)
    return eye + S * torch.sin(n) / n + small_matmul(S, S) * (1 - torch.cos(n)) / n**2


class _ExpVec(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v):
        vc, lead, n = _flat(v, 3, "v")
        R = torch.empty(*lead, 3, 3, device=v.device, dtype=torch.float32)
        _lib.check(_lib.lib().dab_so3_exp(ptr(vc), ptr(R), n, _lib.stream_ptr()), "dab_so3_exp")
        ctx.save_for_backward(vc)
        return R

    @staticmethod
    def backward(ctx, gR):
        (v,) = ctx.saved_tensors
        with torch.enable_grad():
            vv = v.detach().requires_grad_(True)
            (gv,) = torch.autograd.grad(_exp_torch(vv), vv, gR)
        return gv


def vector_to_rotation_matrix(v: torch.Tensor) -> torch.Tensor:
    """so3.py:207-216: rotation vector (*, 3) -> rotation matrix (*, 3, 3)."""
    return _ExpVec.apply(v)


def vector_to_skew_symmetric_mat(v: torch.Tensor) -> torch.Tensor:
    """so3.py:185-204 (pure indexing)."""
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack([torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], -2)


def skew_symmetric_mat_to_vector(S):
    """so3.py:165-170 (pure indexing)."""
    return torch.stack([S[..., 2, 1], S[..., 0, 2], S[..., 1, 0]], dim=-1)


def log_rotmat(R):
    """so3.py:146-162: skew-symmetric log of a rotation matrix."""
    Rc, lead, n = _flat(R, 9, "R")
    S = torch.empty_like(Rc)
    _lib.check(_lib.lib().dab_so3_log_skew(ptr(Rc), ptr(S), n, _lib.stream_ptr()), "dab_so3_log_skew")
    return S


def rotation_matrix_to_vector(R: torch.Tensor) -> torch.Tensor:
    """so3.py:173-182."""
    Rc, lead, n = _flat(R, 9, "R")
    v = torch.empty(*lead, 3, device=R.device, dtype=torch.float32)
    _lib.check(_lib.lib().dab_so3_log(ptr(Rc), ptr(v), n, _lib.stream_ptr()), "dab_so3_log")
    return v


def exp_skew_symmetric_mat(S):
    """so3.py:219-237."""
    Sc, lead, n = _flat(S, 9, "S")
    R = torch.empty_like(Sc)
    _lib.check(_lib.lib().dab_so3_exp_skew(ptr(Sc), ptr(R), n, _lib.stream_ptr()), "dab_so3_exp_skew")
    return R


def scale_rot(R: torch.FloatTensor, k: torch.FloatTensor) -> torch.FloatTensor:
    """so3.py:240-259: exp(k log R); ``k`` has the leading dims of ``R`` (right-broadcast)."""
    if k.ndim > R.ndim:
        raise ValueError(f"Dimension of k ({k.ndim}) cannot be larger than that of R ({R.ndim})")
    Rc, lead, n = _flat(R, 9, "R")
    kc = _lib.dev(k, torch.float32, "k")
    if tuple(kc.shape) != tuple(lead[:kc.ndim]):
        raise ValueError(f"k shape {tuple(kc.shape)} is not a prefix of R's leading dims {tuple(lead)}")
    group = 1
    for s in lead[kc.ndim:]:
        group *= s
    out = torch.empty_like(Rc)
    _lib.check(_lib.lib().dab_so3_scale_rot(ptr(Rc), ptr(kc), n, max(group, 1), ptr(out), _lib.stream_ptr()),
               "dab_so3_scale_rot")
    return out


def uniform(*size, device="cuda", generator=None):
    """Uniform rotations (so3.py:129-139 uses scipy on the host; here normalised Gaussian quaternions)."""
    assert len(size) >= 2 and size[-2] == size[-1] == 3, "last two dimensions must be 3"
    from .synth import uniform_rotations
    return uniform_rotations(*size[:-2], generator=generator, device=device)


class SO3:
    """IGSO(3) table + sampler (so3.py:9-126).  ``sigmas_to_consider`` is a 1-D tensor."""

    def __init__(self, sigmas_to_consider, cache_prefix=None, sigma_threshold=0.1, n_bins=8192, num_iters=1024,
                 device="cuda"):
        self.n_bins = n_bins
        self.num_iters = num_iters
        self.sigma_threshold = sigma_threshold
        self.device = torch.device(device)
        self._sigmas_cpu = torch.as_tensor(sigmas_to_consider, dtype=torch.float32).detach().cpu().contiguous()
        self._sigmas = None
        self._histograms = None  # built on the device at first use (the reference builds in the ctor, so3.py:34)

    @property
    def sigmas_to_consider(self):
        if self._sigmas is None:
            if self.device.type != "cuda":
                raise RuntimeError("SO3 tables live on the GPU - diffab_pytorch_b200 has no CPU path")
            self._sigmas = self._sigmas_cpu.to(self.device)
        return self._sigmas

    @property
    def histograms(self):
        if self._histograms is None:
            self._histograms = self._initialize()
        return self._histograms

    def _initialize(self):
        sig = self.sigmas_to_consider
        out = torch.empty(sig.numel(), self.n_bins, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().dab_igso3_table(ptr(sig), sig.numel(), self.n_bins, self.num_iters, ptr(out),
                                                  _lib.stream_ptr()), "dab_igso3_table")
        return out

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("SO3 tables live on the GPU - diffab_pytorch_b200 has no CPU path")
        if device != self.device:
            self.device = device
            self._sigmas = None
            self._histograms = None if self._histograms is None else self._histograms.to(device)
        return self

    def draw_noise(self, n, num_samples, generator=None):
        """The reference's four draws in its order (so3.py:114,78,83,93)."""
        d = self.device
        return {
            "axis": torch.randn(n, num_samples, 3, device=d, generator=generator),
            "hist_exp": torch.empty(n, self.n_bins, device=d).exponential_(generator=generator),
            "jitter": torch.rand(n, num_samples, device=d, generator=generator),
            "gauss": torch.randn(n, num_samples, device=d, generator=generator),
        }

    def sample_isotropic_gaussian(self, sigma_idx: torch.LongTensor, num_samples: int, noise=None,
                                  return_bins=False, out=None) -> torch.FloatTensor:
        """so3.py:98-126: (n,) indices -> (n, num_samples, 3) rotation vectors (written into ``out`` when given)."""
        idx = _lib.dev(sigma_idx, torch.int64, "sigma_idx")
        n = idx.numel()
        if noise is None:
            noise = self.draw_noise(n, num_samples)
        f = lambda k: _lib.dev(noise[k], torch.float32, k)
        if out is None:
            out = torch.empty(n, num_samples, 3, device=self.device, dtype=torch.float32)
        else:
            out = _lib.dev(out, torch.float32, "out")
            if out.numel() != n * num_samples * 3:
                raise ValueError("sample_isotropic_gaussian: out must hold (n, num_samples, 3) floats")
        bins = torch.empty(n, num_samples, device=self.device, dtype=torch.int64) if return_bins else None
        _lib.check(_lib.lib().dab_igso3_sample(
            ptr(self.histograms), ptr(self.sigmas_to_consider), self.sigmas_to_consider.numel(), self.n_bins,
            ptr(idx), n, num_samples, ptr(f("axis")), ptr(f("hist_exp")), ptr(f("jitter")), ptr(f("gauss")),
            float(self.sigma_threshold), ptr(out), ptr(bins), _lib.stream_ptr()), "dab_igso3_sample")
        return (out, bins) if return_bins else out
