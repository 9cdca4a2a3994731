"""CPU oracle: invariant point attention layer and the epsilon network (test infrastructure).

Index-explicit restatement of ``/root/reference/diffab_pytorch/diffab_pytorch.py:315-607`` that
works on a plain ``{state-dict key: tensor}`` mapping (so fp64 arbitration is a ``.double()`` of the
inputs) and is differentiable by autograd (it is the checker for the backward kernels).

Quirks preserved on purpose: ``gamma`` is used raw (no softplus, ``:373,429``);
frames act on ROW vectors, ``p_glob = p_loc @ R + t`` (``:324``) and ``p_loc = (p_glob - t) @ R^T``
(``:336``); there is no attention mask; no residual / LayerNorm between layers (``:494-498``).
"""
import torch

from . import so3 as oso3

IPA_KEYS = ("gamma", "to_q_scalar.weight", "to_k_scalar.weight", "to_v_scalar.weight",
            "to_pair_bias.weight", "to_q_point.weight", "to_k_point.weight", "to_v_point.weight",
            "to_out.weight", "to_out.bias")


def ipa_layer(w, x, e, R, t, n_head, return_attn=False, use_pair_bias=True):
    """``InvariantPointAttentionLayer.forward`` diffab_pytorch.py:389-465 (``use_pair_bias=False``: no bias term, no pair
    features, two independent logits, ``:374-387,438-462``).

    x (B,L,D)  e (B,L,L,C)  R (B,L,3,3)  t (B,L,3)  ->  (B,L,D)
    Feature order of the projections is (h d) for scalars and (h p c) for points (``:395-408``);
    the concatenation fed to ``to_out`` is [scalar (h d) | pair (h c) | point (h p c) | norm (h p)]
    (``:456-462``).
    """
    B, L, _ = x.shape
    H = n_head
    lin = lambda name: x @ w[name].transpose(0, 1)
    qs = lin("to_q_scalar.weight").view(B, L, H, -1)
    ks = lin("to_k_scalar.weight").view(B, L, H, -1)
    vs = lin("to_v_scalar.weight").view(B, L, H, -1)
    ds = qs.shape[-1]

    def points(name):
        p = lin(name).view(B, L, H, -1, 3)                       # local frame, (h p c)
        return torch.einsum("blhpk,blkc->blhpc", p, R) + t[:, :, None, None, :]   # :315-324

    qp, kp, vp = points("to_q_point.weight"), points("to_k_point.weight"), points("to_v_point.weight")
    Pq = qp.shape[3]

    scale_scalar = ds ** -0.5                                     # :359
    scale_point = (4.5 * Pq) ** -0.5                              # :372
    scale_total = (3 if use_pair_bias else 2) ** -0.5             # :385-387
    logit_scalar = torch.einsum("bihd,bjhd->bhij", qs, ks) * scale_scalar          # :416-419
    bias = torch.einsum("bijc,hc->bhij", e, w["to_pair_bias.weight"]) if use_pair_bias else 0.0   # :423
    diff = qp[:, :, None] - kp[:, None, :]                        # (B,i,j,H,P,3)   :426-428
    d2 = diff.pow(2).sum(-1).sum(-1).permute(0, 3, 1, 2)          # (B,H,i,j)       :435
    logit_point = -0.5 * scale_point * w["gamma"].view(1, H, 1, 1) * d2            # :431-436
    logit = scale_total * (logit_scalar + bias + logit_point)     # :439
    attn = logit.softmax(dim=-1)                                  # :443

    o_scalar = torch.einsum("bhij,bjhd->bihd", attn, vs).reshape(B, L, -1)         # :445-446
    o_pair = torch.einsum("bhij,bijc->bihc", attn, e).reshape(B, L, -1)            # :449-450
    og = torch.einsum("bhij,bjhpc->bihpc", attn, vp)                               # :452
    ol = torch.einsum("bihpk,bick->bihpc", og - t[:, :, None, None, :], R)         # :327-336,453
    nrm = ol.norm(dim=-1)                                                           # :454
    parts = [o_scalar, o_pair, ol.reshape(B, L, -1), nrm.reshape(B, L, -1)]
    cat = torch.cat(parts if use_pair_bias else parts[:1] + parts[2:], dim=-1)                  # :459-462
    y = cat @ w["to_out.weight"].transpose(0, 1) + w["to_out.bias"]                # :464
    return (y, attn, logit) if return_attn else y


def layer_weights(state, prefix):
    """Pick one layer's ten tensors out of a (reference-keyed) state dict."""
    return {k: state[prefix + k] for k in IPA_KEYS}


def ipa_module(state, x, e, R, t, n_layers, n_head, prefix="denoiser.ipa.layers."):
    """``InvariantPointAttentionModule.forward`` diffab_pytorch.py:494-498: plain chain."""
    for i in range(n_layers):
        x = ipa_layer(layer_weights(state, f"{prefix}{i}."), x, e, R, t, n_head)
    return x


def _mlp(state, prefix, idxs, h):
    for n, i in enumerate(idxs):
        h = h @ state[f"{prefix}.{i}.weight"].transpose(0, 1) + state[f"{prefix}.{i}.bias"]
        if n + 1 < len(idxs):
            h = torch.relu(h)
    return h


def denoiser_forward(state, seq_idx_t, x_t, O_t, res_ctx, pair_ctx, beta, n_layers, n_head,
                     prefix="denoiser."):
    """``Denoiser.forward`` diffab_pytorch.py:558-607 (the two mask arguments are unused there)."""
    s_emb = state[prefix + "sequence_embedding.weight"][seq_idx_t]                 # :572
    h = _mlp(state, prefix + "to_res_emb", (0, 2), torch.cat([res_ctx, s_emb], -1))  # :573-574
    h = ipa_module(state, h, pair_ctx, O_t, x_t, n_layers, n_head, prefix + "ipa.layers.")
    t_emb = torch.stack([beta, torch.sin(beta), torch.cos(beta)], -1)              # :584
    h = torch.cat([h, t_emb[:, None, :].expand(-1, h.shape[1], -1)], -1)           # :585-588
    eps = _mlp(state, prefix + "coordinate_denoising", (0, 2, 4), h)               # :591
    v = _mlp(state, prefix + "orientation_denoising", (0, 2, 4), h)                # :594
    O0 = O_t @ oso3.exp_vec(v)                                                     # :595-596
    post = _mlp(state, prefix + "sequence_denoising", (0, 2, 4), h).softmax(-1)    # :599
    return {"translations_eps": eps, "orientations_t0": O0, "seq_posterior": post,
            "rotvec": v, "res_emb": h}


def orientation_loss(pred, target):
    """``OrientationLoss`` diffab_pytorch.py:610-625, reduction='none': (R_pred^T R_true - I)^2."""
    d = torch.einsum("blij,blik->bljk", pred, target)
    return (d - torch.eye(3, dtype=d.dtype)).pow(2)


def losses(denoised, noised, orientations_t0, generation_mask, residue_mask):
    """Loss tail of ``DiffAb._shared_step`` diffab_pytorch.py:856-880."""
    m = generation_mask & residue_mask
    denom = m.sum()
    target = noised["seq_posterior"]
    kl = torch.nn.functional.kl_div(denoised["seq_posterior"].log(), target, reduction="none")
    mse = (denoised["translations_eps"] - noised["translations_eps"]).pow(2)
    rot = orientation_loss(denoised["orientations_t0"], orientations_t0)
    return ((kl * m[..., None]).sum() / denom, (mse * m[..., None]).sum() / denom,
            (rot * m[..., None, None]).sum() / denom)
