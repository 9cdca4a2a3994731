"""CPU baseline runner: the reference's own modules driven through a reverse-diffusion pass (test infrastructure).

Used only by ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg).  When the unmodified reference is
importable (``ref_shim``: ``$DIFFAB_REFERENCE_ROOT``, ``baseline/_ref`` or ``/root/reference``) the epsilon network and the
context encoders that run are the REFERENCE's (``DiffAb.encode_context`` ``diffab_pytorch.py:680-724`` and
``DiffAb.denoise`` ``:726-768``, stock code path, fp32, all host threads); only the reverse update itself comes from
``oracle/sampler.py`` because the reference has none (``DiffAb.sample`` is a stub, ``:770-776``).  Without the reference
the same pass runs on the oracle port (``oracle/ipa.py``), which then also skips the context encoders (it has no port of
them) - the result says which (``kind``: "reference" | "port").
"""
import os
import time

import torch

from . import diffusion as odiff
from . import ref_shim
from . import sampler as osamp
from . import so3 as oso3

TRAIN_CFG = (128, 64, 6, 32, 8, 8, 8)   # train.py:62-70
_MODEL = {}


def reference_model(state):
    """The reference's ``DiffAb`` (train.py configuration) with ``state`` loaded, or None.  Cached per process:
    its constructor builds the IGSO(3) table on the CPU (6-17 s) and writes ``./.cache/so3_histograms``."""
    if "m" not in _MODEL:
        pkg = ref_shim.load_reference()
        if pkg is None:
            _MODEL["m"] = None
        else:
            from diffab_pytorch.diffab_pytorch import DiffAb
            m = DiffAb(*TRAIN_CFG).eval()
            m.load_state_dict(state)
            _MODEL["m"] = m
    return _MODEL["m"]


def _hist_rows(sched, steps, n_bins=8192):
    """Reverse-step IGSO(3) rows (sigma = sqrt(beta_t)) for the steps that take the histogram branch; the others keep a
    row of ones (never selected: so3.py:122-125 picks the Gaussian branch at sigma >= 0.1)."""
    key = ("hist", tuple(sorted(steps)))
    if key not in _MODEL:
        sig = osamp.reverse_sigmas(sched)
        hist = torch.ones(sig.numel(), n_bins)
        for t in steps:
            if float(sig[t]) < 0.1:
                hist[t] = oso3.igso3_pdf_row(sig[t], n_bins)
        _MODEL[key] = hist
    return _MODEL[key]


def timed_pass(state, batch, steps, threads, with_context=True, seed=0):
    """One bounded sample of the workload on the host: context encoding of ``batch`` + the reverse steps ``steps``
    (descending list of t).  Returns dict(kind, t_context_s, t_steps_s, n_steps, n_patches)."""
    torch.set_num_threads(threads)
    model = reference_model(state)
    B, L = batch["seq_idx"].shape
    g = torch.Generator().manual_seed(seed)
    sched = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    hist_rev = _hist_rows(sched, steps)
    s, x, O = osamp.draw_initial_state(batch["seq_idx"], batch["xyz"][:, :, 1], batch["orientations"],
                                       batch["generation_mask"], generator=g)
    noises = {t: osamp.draw_step_noise(B, L, generator=g) for t in steps}
    m = batch["generation_mask"]
    with torch.no_grad():
        t0 = time.perf_counter()
        if model is not None and with_context:
            from diffab_pytorch_b200 import synth
            distmat = batch["distmat"] if "distmat" in batch else synth.pairwise_atom_distances(batch["xyz"])
            res_ctx, pair_ctx = model.encode_context(
                batch["seq_idx"], batch["xyz"], batch["orientations"], batch["backbone_dihedrals"], distmat,
                batch["pairwise_dihedrals"], batch["atom_mask"], batch["chain_idx"], batch["residue_idx"], m,
                batch["residue_mask"])
        else:
            res_ctx = torch.randn(B, L, TRAIN_CFG[0], generator=g)
            pair_ctx = torch.randn(B, L, L, TRAIN_CFG[1], generator=g)
        t_ctx = time.perf_counter() - t0
        t0 = time.perf_counter()
        if model is not None:
            for step in steps:
                t = torch.full((B,), step, dtype=torch.long)
                out = model.denoise(s, x, O, res_ctx, pair_ctx, sched["beta"][t], m, batch["residue_mask"])
                nxt = osamp.reverse_step(sched, hist_rev, s, x, O, out["translations_eps"], out["orientations_t0"],
                                         out["seq_posterior"], m, t, noises[step])
                s, x, O = nxt["seq_idx"], nxt["translations"], nxt["orientations"]
        else:
            from . import ipa as oipa
            for step in steps:
                t = torch.full((B,), step, dtype=torch.long)
                out = oipa.denoiser_forward(state, s, x, O, res_ctx, pair_ctx, sched["beta"][t], TRAIN_CFG[2], TRAIN_CFG[6])
                nxt = osamp.reverse_step(sched, hist_rev, s, x, O, out["translations_eps"], out["orientations_t0"],
                                         out["seq_posterior"], m, t, noises[step])
                s, x, O = nxt["seq_idx"], nxt["translations"], nxt["orientations"]
        t_steps = time.perf_counter() - t0
    ok = bool(torch.isfinite(x).all() and torch.isfinite(O).all())
    return {"kind": "reference" if model is not None else "port", "with_context": bool(model is not None and with_context),
            "t_context_s": t_ctx, "t_steps_s": t_steps, "n_steps": len(steps), "n_patches": B, "finite": ok,
            "root": ref_shim.REFERENCE_ROOT if model is not None else None, "threads": threads,
            "cores": os.cpu_count() or 1}


def patches_per_s(r, T=100):
    """Whole-pass rate implied by a bounded sample: context once + T steps at the sample's mean step time."""
    per_step = r["t_steps_s"] / max(r["n_steps"], 1)
    return r["n_patches"] / (r["t_context_s"] + T * per_step)
