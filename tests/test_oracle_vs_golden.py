"""CPU: the oracle restatement against fixtures produced by the UNMODIFIED reference
(tools/make_goldens.py).  This is what pins the oracle."""
import pytest
import torch

from conftest import checksum, load_golden
from diffab_pytorch_b200 import synth
from oracle import diffusion as odiff
from oracle import ipa as oipa
from oracle import sampler as osamp
from oracle import so3 as oso3


def _hist_from_rows(rows, n=101, n_bins=8192):
    h = torch.zeros(n, n_bins)
    for r, v in rows.items():
        h[int(r)] = v
    return h


def test_schedule_bit_exact():
    g = load_golden("schedule.pt")
    s = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    for k in g:
        assert torch.equal(s[k], g[k]), k
    assert s["beta"][0] == 0 and float(s["alpha_bar"][100]) < 1e-12  # alpha_bar not recomputed from clipped beta


def test_so3_maps():
    g = load_golden("so3_maps.pt")
    assert torch.equal(oso3.log_skew(g["R"]), g["log_skew"])
    assert torch.equal(oso3.log_vec(g["R"]), g["log_vec"])
    assert torch.equal(oso3.exp_vec(g["v"]), g["exp_vec"])
    assert torch.equal(oso3.scale_rot(g["R"], g["k"]), g["scale_rot"])
    # reference property tests (tests/test_so3.py:24-31, 79-93)
    S = oso3.log_skew(g["R"])
    assert torch.allclose(S, -S.transpose(-1, -2))
    Rs = oso3.scale_rot(g["R"], g["k"])
    assert torch.allclose(Rs.transpose(-1, -2) @ Rs, torch.eye(3).expand_as(Rs), rtol=1e-5, atol=1e-5)


def test_igso3_rows():
    g = load_golden("igso3_table.pt")
    s = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    for r in (1, 6):
        assert torch.equal(oso3.igso3_pdf_row(s["one_minus_alpha_bar_sqrt"][r]), g["rows"][r])
    assert torch.equal(oso3.igso3_pdf_row(s["beta"].sqrt()[50]), g["rows_rev"][50])


def test_seq_probs_bit_exact():
    g = load_golden("seq_probs.pt")
    s = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    assert torch.equal(odiff.seq_prob_single_step(s, g["seq"], g["t"], g["mask"]), g["p_single"])
    assert torch.equal(odiff.seq_prob_from_t0(s, g["seq"], g["t"], g["mask"]), g["p_from_t0"])
    assert torch.equal(odiff.seq_posterior(s, g["seq_t"], g["seq"], g["t"], g["mask"]), g["posterior"])


def test_add_noise_bit_exact():
    g = load_golden("add_noise.pt")
    s = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    batch = synth.make_patches(4, 128, seed=g["seed_patches"], with_distmat=False)
    for k, v in g["chk"].items():
        assert checksum(batch[k]) == pytest.approx(v, rel=1e-12), f"synthetic input {k} differs from the generator's"
    hist = _hist_from_rows(g["hist_rows"])
    for seed, key, mask in ((g["seed_noise"], "out", batch["generation_mask"]),
                            (g["seed_noise_all"], "out_all", torch.ones(4, 128, dtype=torch.bool))):
        torch.manual_seed(seed)
        noise = odiff.draw_add_noise_tensors(4, 128)
        out = odiff.add_noise(s, hist, batch["seq_idx"], batch["xyz"][:, :, 1], batch["orientations"],
                              mask, g["t"], noise)
        for name, ref in g[key].items():
            assert torch.equal(out[name], ref), (key, name)


@pytest.mark.parametrize("name", ["train", "tiny", "ragged"])
def test_ipa_layer(name):
    g = load_golden(f"ipa_{name}.pt")
    c = g["cfg"]
    w = synth.synthetic_state(synth.ipa_layer_shapes(c["D"], c["C"], c["H"], c["ds"], c["Pq"], c["Pv"]), seed=c["seed"])
    x, e, R, t = synth.make_ipa_inputs(c["B"], c["L"], c["D"], c["C"], seed=c["seed"] + 100)
    assert checksum(x) == pytest.approx(g["chk"]["x"], rel=1e-12)
    assert checksum(e) == pytest.approx(g["chk"]["e"], rel=1e-12)
    gy = torch.randn(c["B"], c["L"], c["D"], generator=torch.Generator().manual_seed(c["seed"] + 200))
    w64 = {k: v.double().requires_grad_(True) for k, v in w.items()}
    x64, e64 = x.double().requires_grad_(True), e.double().requires_grad_(True)
    y = oipa.ipa_layer(w64, x64, e64, R.double(), t.double(), c["H"])
    (y * gy.double()).sum().backward()
    ref = g["f64"]
    assert (y - ref["y"]).abs().max() < 1e-12
    assert (x64.grad - ref["dx"]).abs().max() < 1e-11
    idx = tuple(slice(None, None, s) for s in ref["de"]["stride"])
    assert (e64.grad[idx] - ref["de"]["sub"]).abs().max() < 1e-11
    assert float(e64.grad.sum()) == pytest.approx(ref["de"]["sum"], rel=1e-9, abs=1e-9)
    for n, gr in ref["dw"].items():
        mine = w64[n].grad
        if isinstance(gr, dict):
            idx = tuple(slice(None, None, s) for s in gr["stride"])
            assert (mine[idx] - gr["sub"]).abs().max() < 1e-10, n
            assert float(mine.abs().sum()) == pytest.approx(gr["abssum"], rel=1e-9)
        else:
            assert (mine - gr).abs().max() < 1e-10, n
    # the reference's own fp32 result sits ~1e-6 from its fp64 one: the 1e-4 budget is comfortable
    assert (g["f32"]["y"].double() - ref["y"]).abs().max() < 1e-4 * ref["y"].abs().max()


def test_denoiser_and_losses():
    g = load_golden("denoiser.pt")
    shapes = load_golden("state_shapes.pt")
    assert len(shapes) == 106 and sum(torch.Size(v).numel() for v in shapes.values()) == 2538468
    state = synth.synthetic_state(shapes, seed=g["seed_state"])
    batch = synth.make_patches(2, 128, seed=g["seed_patches"], with_distmat=False)
    sched = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    n = g["noised"]
    with torch.no_grad():
        # pair_ctx is only stored subsampled; the oracle denoiser is checked on the IPA-free part
        # via the full reference outputs below only when the full pair tensor is available, so here
        # we check the loss tail and the reverse-step fixture instead.
        ol = torch.stack(oipa.losses(g["denoised"], n, batch["orientations"], batch["generation_mask"],
                                     batch["residue_mask"]))
    assert torch.allclose(ol, g["losses"], rtol=1e-6, atol=0)
    # reverse step fixture (our composition) reproduces from the stored reference tensors
    r = load_golden("reverse_step.pt")
    hist_rev = _hist_from_rows(r["hist_rev_rows"])
    gen = torch.Generator().manual_seed(r["seed_noise"])
    noise = osamp.draw_step_noise(2, 128, generator=gen)
    d = g["denoised"]
    for tkey, okey in (("t", "out"), ("t1", "out1")):
        out = osamp.reverse_step(sched, hist_rev, n["seq_idx_t"], n["translations_t"], n["orientations_t"],
                                 d["translations_eps"], d["orientations_t0"], d["seq_posterior"],
                                 batch["generation_mask"], r[tkey], noise, return_bins=True)
        for k in ("seq_idx", "bins"):
            assert torch.equal(out[k], r[okey][k]), k
        for k in ("translations", "orientations"):
            assert torch.equal(out[k], r[okey][k]), k
    # context residues are untouched, generated ones moved
    m = batch["generation_mask"]
    assert torch.equal(out["translations"][~m], n["translations_t"][~m])


def test_ipa_layer_without_pair_bias():
    """use_pair_bias=False (diffab_pytorch.py:374-387,438-462): oracle variant vs the reference's fp64 forward + gradients."""
    g = load_golden("ipa_nopb.pt")
    c = g["cfg"]
    x, e, R, t = synth.make_ipa_inputs(c["B"], c["L"], c["D"], c["C"], seed=c["seed"] + 100)
    gy = torch.randn(c["B"], c["L"], c["D"], generator=torch.Generator().manual_seed(c["seed"] + 200))
    w = {k: v.clone().requires_grad_(True) for k, v in g["state"].items()}
    assert "to_pair_bias.weight" not in w
    x64 = x.double().requires_grad_(True)
    y = oipa.ipa_layer(w, x64, e.double(), R.double(), t.double(), c["H"], use_pair_bias=False)
    (y * gy.double()).sum().backward()
    assert (y - g["y"]).abs().max() < 1e-12
    assert (x64.grad - g["dx"]).abs().max() < 1e-11
    for n, gr in g["dw"].items():
        assert (w[n].grad - gr).abs().max() < 1e-10, n
