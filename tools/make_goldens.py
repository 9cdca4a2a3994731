#!/usr/bin/env python
"""Generate tests/golden/*.pt by running the UNMODIFIED reference (build container only).

Imports ``/root/reference`` through ``oracle/ref_shim.py`` (nothing in the reference is edited or
copied), runs every hot-path function of SURVEY §8a on seeded synthetic inputs, and stores the
outputs as small fixtures.  Inputs are NOT stored when they can be rebuilt from a seed with
``diffab_pytorch_b200.synth`` (same torch version on the GPU box); a checksum of each rebuilt
input is stored so a mismatch fails loudly.  Large outputs are stored as a strided subsample plus
sum / abs-sum.  While generating, the oracle restatement is checked against the reference and the
max deviations are printed (the same checks run from the fixtures in tests/test_oracle_vs_golden.py).

Run from a scratch CWD-independent location:  python tools/make_goldens.py
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import diffab_pytorch_b200  # noqa: E402,F401
from diffab_pytorch_b200 import synth  # noqa: E402
from oracle import diffusion as odiff  # noqa: E402
from oracle import ipa as oipa  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle import sampler as osamp  # noqa: E402
from oracle import so3 as oso3  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TRAIN_CFG = dict(d_residue_emb=128, d_pair_emb=64, n_ipa_layers=6, d_scalar_per_head=32,
                 n_query_point_per_head=8, n_value_point_per_head=8, n_head=8)  # train.py:62-70
TINY_CFG = dict(D=32, C=16, ds=16, Pq=4, Pv=4, H=8)  # tests/test_modules.py:143-149


def checksum(t):
    t = t.detach().double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64) % 97 + 1
    return float((t * w).sum())


def summarize(t, stride):
    """Subsample + sums for a large tensor."""
    idx = tuple(slice(None, None, s) for s in stride)
    return {"sub": t[idx].clone(), "stride": stride, "sum": float(t.double().sum()),
            "abssum": float(t.double().abs().sum()), "shape": tuple(t.shape)}


def maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())


def main():
    os.makedirs(GOLD, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="diffab_gold_")
    os.chdir(scratch)  # the reference writes ./.cache/so3_histograms (so3.py:13,19)
    ref = ref_shim.load_reference()
    assert ref is not None, "reference tree not found"
    from diffab_pytorch import diffab_pytorch as rmod
    from diffab_pytorch import diffusion as rdiff
    from diffab_pytorch import so3 as rso3

    torch.set_num_threads(4)

    # ------------------------------------------------------------------ schedule (A10)
    sched = rdiff.cosine_variance_schedule(100, s=0.01, beta_max=0.999)
    osched = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    for k in sched:
        assert torch.equal(sched[k], osched[k]), k
    torch.save({k: v.clone() for k, v in sched.items()}, os.path.join(GOLD, "schedule.pt"))
    print("schedule: oracle bit-equal")

    # ------------------------------------------------------------------ SO(3) maps (A5-A7)
    g = torch.Generator().manual_seed(11)
    R = synth.uniform_rotations(8, 100, generator=g)
    k = torch.rand(8, generator=g)
    v = torch.randn(8, 100, 3, generator=g) * 1.2
    gold = {
        "R": R, "k": k, "v": v,
        "log_skew": rso3.log_rotmat(R),
        "log_vec": rso3.rotation_matrix_to_vector(R),
        "exp_vec": rso3.vector_to_rotation_matrix(v),
        "exp_log": rso3.exp_skew_symmetric_mat(rso3.log_rotmat(R)),
        "scale_rot": rso3.scale_rot(R, k),
    }
    print("so3: oracle max diffs",
          maxdiff(oso3.log_skew(R), gold["log_skew"]), maxdiff(oso3.log_vec(R), gold["log_vec"]),
          maxdiff(oso3.exp_vec(v), gold["exp_vec"]), maxdiff(oso3.scale_rot(R, k), gold["scale_rot"]))
    torch.save(gold, os.path.join(GOLD, "so3_maps.pt"))

    # ------------------------------------------------------------------ model (ctor builds the IGSO3 table, A8)
    torch.manual_seed(0)
    model = rmod.DiffAb(**TRAIN_CFG)
    model.eval()
    shapes = {k2: tuple(v2.shape) for k2, v2 in model.state_dict().items()}
    torch.save(shapes, os.path.join(GOLD, "state_shapes.pt"))
    print("state dict:", len(shapes), "tensors,", sum(v2.numel() for v2 in model.state_dict().values()), "params")
    state = synth.synthetic_state(shapes, seed=0)
    model.load_state_dict(state)

    hist = model.orientation_diffuser.so3.histograms           # (101, 8192) sigma = sqrt(1-abar)
    rows = [0, 1, 2, 5, 6, 50, 100]
    so3_rev = rso3.SO3(sigmas_to_consider=sched["beta"].sqrt())  # reverse-step table, sigma = sqrt(beta)
    hist_rev = so3_rev.histograms
    rows_rev = [1, 3, 20, 50, 100]
    torch.save({"rows": {r: hist[r].clone() for r in rows},
                "rows_rev": {r: hist_rev[r].clone() for r in rows_rev}},
               os.path.join(GOLD, "igso3_table.pt"))
    for r in (1, 6, 100):
        o = oso3.igso3_pdf_row(sched["one_minus_alpha_bar_sqrt"][r])
        print(f"igso3 row {r}: oracle max diff {maxdiff(o, hist[r]):.3e} (row max {float(hist[r].max()):.3e})")

    # ------------------------------------------------------------------ _add_noise (A9, A11-A16)
    B, L = 4, 128
    batch = synth.make_patches(B, L, seed=3, with_distmat=False)
    t = torch.tensor([1, 5, 6, 100])
    torch.manual_seed(123)
    ref_out = model._add_noise(batch["seq_idx"], batch["xyz"][:, :, 1], batch["orientations"],
                               batch["generation_mask"], t)
    torch.manual_seed(123)
    noise = odiff.draw_add_noise_tensors(B, L)
    full_hist = hist
    ora = odiff.add_noise(osched, full_hist, batch["seq_idx"], batch["xyz"][:, :, 1],
                          batch["orientations"], batch["generation_mask"], t, noise)
    for key in ref_out:
        same = torch.equal(ref_out[key], ora[key])
        print(f"_add_noise {key}: bit-equal={same} maxdiff={maxdiff(ref_out[key], ora[key]):.3e}")
        assert same or ref_out[key].dtype.is_floating_point
    assert torch.equal(ref_out["seq_idx_t"], ora["seq_idx_t"])
    # a second case with every residue generated (exercises the mask-free branch everywhere)
    all_mask = torch.ones(B, L, dtype=torch.bool)
    torch.manual_seed(124)
    ref_out2 = model._add_noise(batch["seq_idx"], batch["xyz"][:, :, 1], batch["orientations"], all_mask, t)
    torch.save({"seed_patches": 3, "t": t, "seed_noise": 123, "out": ref_out,
                "seed_noise_all": 124, "out_all": ref_out2,
                "hist_rows": {int(r): hist[int(r)].clone() for r in t},
                "chk": {kk: checksum(batch[kk]) for kk in ("seq_idx", "xyz", "orientations")}},
               os.path.join(GOLD, "add_noise.pt"))

    # sequence-diffuser probabilities (A11, A13) on a random mask
    g = torch.Generator().manual_seed(5)
    seq = torch.randint(0, 20, (4, 100), generator=g)
    gm = torch.randint(0, 2, (4, 100), generator=g).bool()
    tt = torch.tensor([1, 10, 90, 100])
    sd = model.seq_diffuser
    seq_t = torch.where(gm, torch.randint(0, 21, (4, 100), generator=g), seq)  # context residues keep s_0
    gold = {"seq": seq, "mask": gm, "t": tt, "seq_t": seq_t,
            "p_single": sd.forward_prob_single_step(seq, tt, gm),
            "p_from_t0": sd.forward_prob_from_t0(seq, tt, gm),
            "posterior": sd.posterior_single_step(seq_t, seq, tt, gm)}
    assert torch.equal(gold["p_single"], odiff.seq_prob_single_step(osched, seq, tt, gm))
    assert torch.equal(gold["p_from_t0"], odiff.seq_prob_from_t0(osched, seq, tt, gm))
    print("seq posterior oracle maxdiff", maxdiff(gold["posterior"], odiff.seq_posterior(osched, seq_t, seq, tt, gm)))
    torch.save(gold, os.path.join(GOLD, "seq_probs.pt"))

    # ------------------------------------------------------------------ IPA layer (A1/A2), fwd + grads
    def ipa_case(name, B, L, D, C, H, ds, Pq, Pv, seed, stride_e):
        shp = synth.ipa_layer_shapes(D, C, H, ds, Pq, Pv)
        w = synth.synthetic_state(shp, seed=seed)
        layer = rmod.InvariantPointAttentionLayer(D, C, ds, Pq, Pv, H)
        layer.load_state_dict(w)
        x, e, R, t3 = synth.make_ipa_inputs(B, L, D, C, seed=seed + 100)
        gy = torch.randn(B, L, D, generator=torch.Generator().manual_seed(seed + 200))
        out = {}
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            lay = layer.to(dt)
            xi = x.detach().to(dt).clone().requires_grad_(True)
            ei = e.detach().to(dt).clone().requires_grad_(True)
            y = lay(xi, ei, R.to(dt), t3.to(dt))
            (y * gy.to(dt)).sum().backward()
            out[tag] = {"y": y.detach().clone(), "dx": xi.grad.clone(),
                        "de": summarize(ei.grad, stride_e),
                        "dw": {n: p.grad.clone() for n, p in lay.named_parameters()}}
            for p in lay.parameters():
                p.grad = None
        # oracle check (fp64)
        w64 = {kk: vv.double() for kk, vv in w.items()}
        yo = oipa.ipa_layer(w64, x.double(), e.double(), R.double(), t3.double(), H)
        print(f"ipa[{name}]: oracle-vs-ref fp64 maxdiff {maxdiff(yo, out['f64']['y']):.3e}; "
              f"ref fp32-vs-fp64 {maxdiff(out['f32']['y'], out['f64']['y']):.3e} "
              f"(|y|max {float(out['f64']['y'].abs().max()):.3f})")
        # keep fp64 arbiter small: y, dx, de-summary, small dw only
        out["f64"]["dw"] = {n: (g2 if g2.numel() <= 4096 else summarize(g2, (4, 4)))
                            for n, g2 in out["f64"]["dw"].items()}
        out["f32"]["dw"] = {n: (g2 if g2.numel() <= 4096 else summarize(g2, (4, 4)))
                            for n, g2 in out["f32"]["dw"].items()}
        out["cfg"] = dict(B=B, L=L, D=D, C=C, H=H, ds=ds, Pq=Pq, Pv=Pv, seed=seed)
        out["chk"] = {"x": checksum(x), "e": checksum(e), "w": checksum(w["to_out.weight"])}
        torch.save(out, os.path.join(GOLD, f"ipa_{name}.pt"))

    ipa_case("train", 1, 128, 128, 64, 8, 32, 8, 8, seed=0, stride_e=(1, 8, 8, 1))
    ipa_case("tiny", 4, 16, seed=1, stride_e=(1, 1, 1, 1), **TINY_CFG)
    ipa_case("ragged", 2, 37, 48, 24, 4, 12, 3, 5, seed=2, stride_e=(1, 1, 1, 1))

    # ------------------------------------------------------------------ encode_context / Denoiser (A4, A17) / losses
    B = 2
    batch = synth.make_patches(B, 128, seed=7)
    with torch.no_grad():
        res_ctx, pair_ctx = model.encode_context(
            batch["seq_idx"], batch["xyz"], batch["orientations"], batch["backbone_dihedrals"],
            batch["distmat"], batch["pairwise_dihedrals"], batch["atom_mask"], batch["chain_idx"],
            batch["residue_idx"], batch["generation_mask"], batch["residue_mask"])
        torch.manual_seed(321)
        t = torch.tensor([37, 88])
        noised = model._add_noise(batch["seq_idx"], batch["xyz"][:, :, 1], batch["orientations"],
                                  batch["generation_mask"], t)
        beta = sched["beta"][t]
        den = model.denoise(noised["seq_idx_t"], noised["translations_t"], noised["orientations_t"],
                            res_ctx, pair_ctx, beta, batch["generation_mask"], batch["residue_mask"])
        oden = oipa.denoiser_forward(state, noised["seq_idx_t"], noised["translations_t"],
                                     noised["orientations_t"], res_ctx, pair_ctx, beta, 6, 8)
        for kk in den:
            print(f"denoiser {kk}: oracle maxdiff {maxdiff(den[kk], oden[kk]):.3e}")
        # losses via the reference's own criteria
        seq_loss = model.aa_loss(den["seq_posterior"].log(), noised["seq_posterior"])
        pos_loss = model.coordinate_loss(den["translations_eps"], noised["translations_eps"])
        rot_loss = model.orientation_loss(den["orientations_t0"], batch["orientations"])
        m = batch["generation_mask"] & batch["residue_mask"]
        denom = m.sum()
        ref_losses = torch.stack([(seq_loss * m[..., None]).sum() / denom,
                                  (pos_loss * m[..., None]).sum() / denom,
                                  (rot_loss * m[..., None, None]).sum() / denom])
        ol = torch.stack(oipa.losses(oden, noised, batch["orientations"], batch["generation_mask"],
                                     batch["residue_mask"]))
        print("losses ref", ref_losses.tolist(), "oracle maxdiff", maxdiff(ref_losses, ol))
    torch.save({"seed_patches": 7, "seed_state": 0, "t": t, "seed_noise": 321,
                "res_ctx": res_ctx, "pair_ctx": summarize(pair_ctx, (1, 8, 8, 1)),
                "noised": noised, "denoised": den, "losses": ref_losses,
                "chk": {kk: checksum(batch[kk]) for kk in ("seq_idx", "xyz", "orientations", "distmat")}},
               os.path.join(GOLD, "denoiser.pt"))

    # full _shared_step forward under a global seed (draw order #1-#7)
    with torch.no_grad():
        torch.manual_seed(99)
        sl = model._shared_step(batch, 0)
    torch.save({"seed_patches": 7, "seed_state": 0, "seed_step": 99, "losses": torch.stack(list(sl))},
               os.path.join(GOLD, "shared_step.pt"))
    print("_shared_step losses", [float(v2) for v2 in sl])

    # ------------------------------------------------------------------ reverse step (O3; OUR composition, unpinned by the reference)
    g = torch.Generator().manual_seed(77)
    tt = torch.tensor([50, 3])
    noise = osamp.draw_step_noise(B, 128, generator=g)
    rev = osamp.reverse_step(osched, hist_rev, noised["seq_idx_t"], noised["translations_t"],
                             noised["orientations_t"], den["translations_eps"], den["orientations_t0"],
                             den["seq_posterior"], batch["generation_mask"], tt, noise, return_bins=True)
    tt1 = torch.tensor([1, 100])
    rev1 = osamp.reverse_step(osched, hist_rev, noised["seq_idx_t"], noised["translations_t"],
                              noised["orientations_t"], den["translations_eps"], den["orientations_t0"],
                              den["seq_posterior"], batch["generation_mask"], tt1, noise, return_bins=True)
    torch.save({"t": tt, "t1": tt1, "seed_noise": 77, "out": rev, "out1": rev1,
                "hist_rev_rows": {int(r): hist_rev[int(r)].clone() for r in (50, 3, 1, 100)}},
               os.path.join(GOLD, "reverse_step.pt"))
    total = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print(f"wrote {GOLD}: {total / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
