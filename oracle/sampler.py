"""CPU oracle: reverse-diffusion step and sampling loop (test infrastructure).

PARITY UNPINNED BY THE REFERENCE.  ``DiffAb.sample`` is an empty stub in the reference
(``diffab_pytorch/diffab_pytorch.py:770-776``) and no reverse step exists anywhere in it, so there
is nothing to pin this file against.  The composition below is the only sampler consistent with
what each head of the reference's epsilon network is trained to predict, and is
built exclusively from primitives that ARE pinned (schedule, exp map, IGSO(3) sampler,
multinomial-by-exponential).  It is frozen by our own goldens (tests/golden/reverse_step.pt).

One step, for a batch that shares nothing across patches (t may differ per patch):

  s_{t-1} ~ Multinomial(seq_posterior)          net is trained against q(s_{t-1}|s_t,s_0) (:857-859)
  x_{t-1} = (x_t - beta_t/sqrt(1-abar_t) eps_theta)/sqrt(alpha_t) + sqrt(beta_t) z ;  z = 0 at t = 1
                                                net is trained against eps (:860-862)
  O_{t-1} = O0_theta @ exp(IGSO3 rotvec at sigma = sqrt(beta_t)) ; no noise at t = 1
                                                net is trained against O_0 (:863-865)
  every output is where(generation_mask, new, old)   (convention of diffusion.py:79,231,292)

Noise tensors per step, drawn in this fixed order (same shapes/order as ``_add_noise``):
seq_exp (B*L,21) Exp(1); z (B,L,3) randn; axis (B,L,3) randn; hist_exp (B,n_bins) Exp(1);
jitter (B,L) U[0,1); gauss (B,L) randn.
"""
import torch

from . import diffusion as odiff
from . import ipa as oipa
from . import so3 as oso3


def reverse_sigmas(sched):
    """sigma table of the reverse-step IGSO(3) noise: sqrt(beta_t), t = 0..T."""
    return sched["beta"].sqrt()


def reverse_step(sched, hist_rev, s_t, x_t, O_t, eps_theta, O0_theta, seq_post, generation_mask,
                 t, noise, return_bins=False):
    """One reverse step; ``t`` is (B,) int64 in [1, T]; ``hist_rev`` = igso3_table(sqrt(beta))."""
    B, L = s_t.shape
    m = generation_mask
    # sequence
    s_new = oso3.multinomial_from_exponential(seq_post.reshape(-1, odiff.V),
                                              noise["seq_exp"].view(-1, odiff.V), 1).view(B, L)
    s_prev = torch.where(m, s_new, s_t)
    # positions
    beta = sched["beta"][t][:, None, None]
    c_eps = beta / sched["one_minus_alpha_bar_sqrt"][t][:, None, None]
    inv_sqrt_alpha = 1.0 / sched["alpha"][t].sqrt()[:, None, None]
    noisy = (t > 1)
    sig = (beta.sqrt() * noisy[:, None, None]).to(x_t.dtype)
    x_new = (x_t - c_eps * eps_theta) * inv_sqrt_alpha + sig * noise["z"]
    x_prev = torch.where(m[..., None], x_new, x_t)
    # orientations
    rotvec, bins = oso3.igso3_sample(hist_rev, reverse_sigmas(sched), t, L, noise["axis"],
                                     noise["hist_exp"], noise["jitter"], noise["gauss"],
                                     return_bins=True)
    O_noised = O0_theta @ oso3.exp_vec(rotvec)
    O_new = torch.where(noisy[:, None, None, None], O_noised, O0_theta)
    O_prev = torch.where(m[..., None, None], O_new, O_t)
    out = {"seq_idx": s_prev, "translations": x_prev, "orientations": O_prev}
    if return_bins:
        out["bins"] = bins
    return out


def draw_step_noise(B, L, n_bins=8192, generator=None, device="cpu"):
    g = generator
    return {
        "seq_exp": torch.empty(B * L, odiff.V, device=device).exponential_(generator=g),
        "z": torch.randn(B, L, 3, generator=g, device=device),
        "axis": torch.randn(B, L, 3, generator=g, device=device),
        "hist_exp": torch.empty(B, n_bins, device=device).exponential_(generator=g),
        "jitter": torch.rand(B, L, generator=g, device=device),
        "gauss": torch.randn(B, L, generator=g, device=device),
    }


def draw_initial_state(seq_idx, x, O, generation_mask, generator=None):
    """t = T prior on generated residues: s ~ U{0..20}, x ~ N(0, I), O ~ uniform SO(3)."""
    B, L = seq_idx.shape
    s = torch.randint(0, odiff.V, (B, L), generator=generator)
    xT = torch.randn(B, L, 3, generator=generator)
    OT = oso3.uniform_rotations(B, L, generator=generator)
    m = generation_mask
    return (torch.where(m, s, seq_idx), torch.where(m[..., None], xT, x),
            torch.where(m[..., None, None], OT, O))


def sample_loop(state, sched, hist_rev, s, x, O, res_ctx, pair_ctx, generation_mask, n_layers,
                n_head, noises, t_start=None, t_stop=1):
    """Run steps t_start..t_stop (inclusive, descending); ``noises[t]`` is that step's noise dict."""
    T = sched["beta"].numel() - 1
    t_start = T if t_start is None else t_start
    B = s.shape[0]
    for step in range(t_start, t_stop - 1, -1):
        t = torch.full((B,), step, dtype=torch.long)
        beta = sched["beta"][t]
        out = oipa.denoiser_forward(state, s, x, O, res_ctx, pair_ctx, beta, n_layers, n_head)
        nxt = reverse_step(sched, hist_rev, s, x, O, out["translations_eps"],
                           out["orientations_t0"], out["seq_posterior"], generation_mask, t,
                           noises[step])
        s, x, O = nxt["seq_idx"], nxt["translations"], nxt["orientations"]
    return {"seq_idx": s, "translations": x, "orientations": O}
