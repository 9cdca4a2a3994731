"""CPU oracle: SO(3) maps and the IGSO(3) table/sampler (test infrastructure, see oracle/__init__).

Restates ``/root/reference/diffab_pytorch/so3.py``; each function cites the lines it follows.
All random draws are INJECTED as tensors: the caller draws them with torch in the
reference's call order and passes them in, so that a CUDA kernel fed the same tensors can be
compared bit-for-bit on integer outputs and within tolerance on floats.
"""
import math

import torch

PI = math.pi


# --------------------------------------------------------------------------- maps (so3.py:142-259)
def trace3(R):
    """``tensor_trace`` so3.py:142-143."""
    return R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]


def log_skew(R):
    """``log_rotmat`` so3.py:146-162: theta/(2 sin theta) * (R - R^T), theta = acos((tr-1)/2).

    No guard at theta in {0, pi} (NaN/inf there, as in the reference).  The reference accepts only
    4-D input (so3.py:159); this accepts any leading dims and is identical on 4-D.
    """
    theta = torch.acos((trace3(R) - 1) / 2)
    coef = theta / (2 * torch.sin(theta))
    return coef[..., None, None] * (R - R.transpose(-1, -2))


def vee(S):
    """``skew_symmetric_mat_to_vector`` so3.py:165-170: (S21, S02, S10)."""
    return torch.stack([S[..., 2, 1], S[..., 0, 2], S[..., 1, 0]], dim=-1)


def hat(v):
    """``vector_to_skew_symmetric_mat`` so3.py:185-204."""
    x, y, z = v.unbind(-1)
    o = torch.zeros_like(x)
    return torch.stack(
        [torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], -2
    )


def log_vec(R):
    """``rotation_matrix_to_vector`` so3.py:173-182."""
    return vee(log_skew(R))


def exp_skew(S):
    """``exp_skew_symmetric_mat`` so3.py:219-237: I + S sin(n)/n + S@S (1-cos n)/n^2, n=|vee(S)|."""
    n = vee(S).norm(dim=-1)[..., None, None]
    eye = torch.eye(3, dtype=S.dtype).expand_as(S)
    return eye + S * torch.sin(n) / n + (S @ S) * (1 - torch.cos(n)) / n**2


def exp_vec(v):
    """``vector_to_rotation_matrix`` so3.py:207-216."""
    return exp_skew(hat(v))


def scale_rot(R, k):
    """``scale_rot`` so3.py:240-259: exp(k * log R); k broadcast by right-unsqueeze."""
    if k.ndim > R.ndim:
        raise ValueError("k has more dims than R")
    while k.ndim < R.ndim:
        k = k.unsqueeze(-1)
    return exp_skew(k * log_skew(R))


def uniform_rotations(*lead, generator=None, dtype=torch.float32):
    """Uniform SO(3) from normalised Gaussian quaternions (SURVEY §8d synthetic inputs).

    Stands in for ``so3.uniform`` (so3.py:129-139, scipy-based, tests only).
    """
    q = torch.randn(*lead, 4, generator=generator, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack(
        [
            torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
            torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
            torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1),
        ],
        -2,
    )
    return R.to(dtype)


# ------------------------------------------------------------------ IGSO(3) table (so3.py:52-72)
def igso3_pdf_row(sigma, n_bins=8192, n_terms=1024):
    """One histogram row: ``_precompute_histogram`` + ``_angular_pdf`` so3.py:52-72 (fp32).

    Bin centres (k + 1/2) pi / n_bins; pdf = (1-cos th)/pi * sum_l (2l+1) exp(-l(l+1) s^2)
    sin((l+1/2) th) / sin(th/2); then nan_to_num and clamp at 0.  Rows are NOT normalised.
    """
    sigma = torch.as_tensor(sigma, dtype=torch.float32)
    binsize = PI / n_bins
    theta = torch.arange(0, PI, binsize) + binsize / 2.0  # fp32
    l = torch.arange(n_terms).view(-1, 1)
    a = (1 - torch.cos(theta)) / PI
    b = (2 * l + 1) * torch.exp(-l * (l + 1) * sigma**2)
    c = torch.sin((l + 0.5) * theta) / torch.sin(theta / 2.0)
    return torch.nan_to_num((a * b * c).sum(dim=0)).clamp_min(0.0)


def igso3_table(sigmas, n_bins=8192, n_terms=1024):
    """``SO3._initialize`` so3.py:37-50 without the disk cache: (n_sigma, n_bins) fp32."""
    return torch.stack([igso3_pdf_row(s, n_bins, n_terms) for s in sigmas])


# ------------------------------------------------------------- IGSO(3) sampler (so3.py:74-126)
def multinomial_from_exponential(p, q, num_samples):
    """What ``torch.multinomial(p, n)`` (no replacement) does given its Exp(1) draw ``q``:
    indices of the top-n of p / q in descending order (n == 1: argmax).  SURVEY §7 H4 probe."""
    key = p / q
    if num_samples == 1:
        return key.argmax(dim=-1, keepdim=True)
    return key.topk(num_samples, dim=-1).indices


def igso3_sample(hist, sigmas, sigma_idx, n_samples, axis_noise, exp_noise, jitter, gauss_noise,
                 sigma_threshold=0.1, return_bins=False):
    """``SO3.sample_isotropic_gaussian`` so3.py:98-126 with its four draws injected.

    axis_noise  (n, s, 3) ~ randn          so3.py:114
    exp_noise   (n, n_bins) ~ Exp(1)       the draw inside torch.multinomial, so3.py:78
    jitter      (n, s) ~ U[0,1)            so3.py:83
    gauss_noise (n, s) ~ randn             so3.py:93
    Both angle branches are always evaluated; ``where(sigma < thr, hist, gauss)`` picks (so3.py:122-125).
    """
    n_bins = hist.shape[-1]
    u = torch.nn.functional.normalize(axis_noise, dim=-1)
    bins = multinomial_from_exponential(hist[sigma_idx], exp_noise, n_samples)
    binsize = PI / n_bins
    starts = torch.arange(0, PI, binsize)
    theta_h = starts[bins] + binsize * jitter
    sig = sigmas[sigma_idx]
    theta_g = (2.0 * sig[:, None] + sig[:, None] * gauss_noise) % PI
    theta = torch.where((sig < sigma_threshold)[:, None], theta_h, theta_g)
    out = u * theta[..., None]
    return (out, bins) if return_bins else out
