"""CPU oracle: variance schedule and the three diffusers (test infrastructure, see oracle/__init__).

Restates ``/root/reference/diffab_pytorch/diffusion.py`` with every random draw injected as a
tensor (SURVEY §3.1 draws #2-#7, §7 H4).  ``V = 21`` classes is hard-coded in the reference
(``diffusion.py:47``) and therefore here.
"""
import math

import torch

from . import so3 as oso3

V = 21


def cosine_schedule(T, s=8e-3, beta_max=0.999):
    """``cosine_variance_schedule`` diffusion.py:11-35 (T+1 entries, beta_0 = 0, alpha_bar NOT
    recomputed from the clipped beta)."""
    t = torch.arange(T + 1)
    f = torch.cos((t / T + s) / (1 + s) * math.pi / 2.0).square()
    alpha_bar = f / f[0]
    beta = torch.cat([torch.zeros(1), (1 - alpha_bar[1:] / alpha_bar[:-1]).clip(1e-5, beta_max)])
    return {
        "alpha": 1 - beta,
        "alpha_bar": alpha_bar,
        "alpha_bar_sqrt": alpha_bar.sqrt(),
        "one_minus_alpha_bar_sqrt": (1 - alpha_bar).sqrt(),
        "beta": beta,
    }


def _mix_with_uniform(seq_idx, w_keep, w_noise, generation_mask):
    """Shared body of forward_prob_single_step / forward_prob_from_t0 (diffusion.py:68-79,124-135):
    p = w_keep * onehot + w_noise * (1/21), evaluated as fl(fl(w_keep*oh) + fl(w_noise*u)) in fp32
    with u = fl32(1/21); residues outside the generation mask keep the exact one-hot."""
    onehot = torch.nn.functional.one_hot(seq_idx, num_classes=V)
    unif = torch.ones_like(onehot) / V
    p = w_keep[:, None, None] * onehot + w_noise[:, None, None] * unif
    return torch.where(generation_mask[..., None].expand_as(onehot), p, onehot)


def seq_prob_single_step(sched, seq_idx, t, generation_mask):
    """``SequenceDiffuser.forward_prob_single_step`` diffusion.py:49-79."""
    beta = sched["beta"][t]
    return _mix_with_uniform(seq_idx, 1 - beta, beta, generation_mask)


def seq_prob_from_t0(sched, seq_idx_t0, t, generation_mask):
    """``SequenceDiffuser.forward_prob_from_t0`` diffusion.py:105-135."""
    ab = sched["alpha_bar"][t]
    return _mix_with_uniform(seq_idx_t0, ab, 1 - ab, generation_mask)


def seq_posterior(sched, seq_idx_t, seq_idx_t0, t, generation_mask):
    """``SequenceDiffuser.posterior_single_step`` diffusion.py:168-192."""
    p = seq_prob_single_step(sched, seq_idx_t, t, generation_mask) * seq_prob_from_t0(
        sched, seq_idx_t0, t - 1, generation_mask
    )
    return p / p.sum(dim=-1, keepdim=True)


def seq_diffuse_from_t0(sched, seq_idx_t0, t, generation_mask, exp_noise):
    """``SequenceDiffuser.diffuse_from_t0`` diffusion.py:137-166; ``exp_noise`` (b*L, 21) is the
    Exp(1) draw inside ``torch.multinomial`` (argmax of p / q)."""
    p = seq_prob_from_t0(sched, seq_idx_t0, t, generation_mask)
    s_t = oso3.multinomial_from_exponential(p.view(-1, V), exp_noise.view(-1, V), 1).view(p.shape[:-1])
    return s_t, seq_posterior(sched, s_t, seq_idx_t0, t, generation_mask)


def coord_diffuse_from_t0(sched, x0, t, generation_mask, eps):
    """``CoordinateDiffuser.diffuse_from_t0`` diffusion.py:199-236 (eps is returned unmasked)."""
    a = sched["alpha_bar_sqrt"][t][:, None, None]
    b = sched["one_minus_alpha_bar_sqrt"][t][:, None, None]
    x_t = a * x0 + b * eps
    return torch.where(generation_mask[..., None], x_t, x0), eps


def orient_diffuse_from_t0(sched, hist, O0, generation_mask, t, axis_noise, exp_noise, jitter,
                           gauss_noise):
    """``OrientationDiffuser.diffuse_from_t0`` diffusion.py:262-294:
    O_t = scale_rot(O_0, sqrt(abar_t)) @ exp(IGSO3 sample at sigma_t = sqrt(1 - abar_t))."""
    mean = oso3.scale_rot(O0, sched["alpha_bar_sqrt"][t])
    rotvec = oso3.igso3_sample(hist, sched["one_minus_alpha_bar_sqrt"], t, O0.shape[1], axis_noise,
                               exp_noise, jitter, gauss_noise)
    O_t = mean @ oso3.exp_vec(rotvec)
    return torch.where(generation_mask[..., None, None], O_t, O0)


def draw_add_noise_tensors(bsz, L, n_bins=8192, generator=None):
    """Draw the six noise tensors of ``DiffAb._add_noise`` in the reference's order
    (validated bit-exact against the reference in tools/make_goldens.py)."""
    g = generator
    return {
        "seq_exp": torch.empty(bsz * L, V).exponential_(generator=g),
        "eps": torch.randn(bsz, L, 3, generator=g),
        "axis": torch.randn(bsz, L, 3, generator=g),
        "hist_exp": torch.empty(bsz, n_bins).exponential_(generator=g),
        "jitter": torch.rand(bsz, L, generator=g),
        "gauss": torch.randn(bsz, L, generator=g),
    }


def add_noise(sched, hist, seq_idx_t0, x0, O0, generation_mask, t, noise):
    """``DiffAb._add_noise`` diffab_pytorch.py:778-806."""
    s_t, post = seq_diffuse_from_t0(sched, seq_idx_t0, t, generation_mask, noise["seq_exp"])
    x_t, eps = coord_diffuse_from_t0(sched, x0, t, generation_mask, noise["eps"])
    O_t = orient_diffuse_from_t0(sched, hist, O0, generation_mask, t, noise["axis"],
                                 noise["hist_exp"], noise["jitter"], noise["gauss"])
    return {"seq_idx_t": s_t, "seq_posterior": post, "translations_t": x_t,
            "translations_eps": eps, "orientations_t": O_t}
