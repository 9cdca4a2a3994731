"""GPU parity: forward noising, sequence probabilities and the reverse step vs the reference goldens
and the oracle, through the C ABI.  Integer outputs (sequences) bit-exact; floats <= 1e-4 relative."""
import pytest
import torch

from conftest import checksum, load_golden
from diffab_pytorch_b200 import _lib, diffusion, so3, synth
from oracle import diffusion as odiff
from oracle import sampler as osamp

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _hist(rows, device):
    h = torch.zeros(101, 8192)
    for r, v in rows.items():
        h[int(r)] = v
    return h.to(device)


def _cuda(d):
    return {k: v.to(DEV) for k, v in d.items()}


def test_seq_probs_bit_exact():
    g = load_golden("seq_probs.pt")
    sd = diffusion.SequenceDiffuser(T=100, s=0.01, beta_max=0.999, device=DEV)
    seq, t, m, st = g["seq"].to(DEV), g["t"].to(DEV), g["mask"].to(DEV), g["seq_t"].to(DEV)
    assert torch.equal(sd.forward_prob_single_step(seq, t, m).cpu(), g["p_single"])
    assert torch.equal(sd.forward_prob_from_t0(seq, t, m).cpu(), g["p_from_t0"])
    post = sd.posterior_single_step(st, seq, t, m).cpu()
    assert (post - g["posterior"]).abs().max() < 1e-6


def test_reference_sequence_diffuser_properties():
    # tests/test_diffusion.py:16-103 of the reference
    sd = diffusion.SequenceDiffuser(T=100, s=0.01, beta_max=0.999, device=DEV)
    bsz, L = 32, 100
    seq = torch.randint(0, 20, (bsz, L), device=DEV)
    all_m = torch.ones(bsz, L, dtype=torch.bool, device=DEV)
    gm = torch.randint(0, 2, (bsz, L), device=DEV).bool()
    full = lambda v: torch.full((bsz,), v, device=DEV, dtype=torch.long)
    for fn in (sd.forward_prob_single_step, sd.forward_prob_from_t0):
        p1, p90 = fn(seq, full(1), all_m), fn(seq, full(90), all_m)
        assert p1.shape == p90.shape == (bsz, L, 21)
        assert (p1.gather(-1, seq[..., None]) > p90.gather(-1, seq[..., None])).all()
    p10 = sd.forward_prob_from_t0(seq, full(10), gm)
    sampled = torch.multinomial(p10.view(-1, 21), 1).view(bsz, L)
    post = sd.posterior_single_step(sampled, seq, full(10), gm)
    assert (post.gather(-1, seq[..., None]) > 1 / 20.0).all()
    s2, post2 = sd.diffuse_from_t0(seq, full(2), all_m, return_posterior=True)
    s99, post99 = sd.diffuse_from_t0(seq, full(99), all_m, return_posterior=True)
    assert s2.shape == s99.shape == (bsz, L) and post2.shape == post99.shape == (bsz, L, 21)
    assert (s2 != seq).sum() < (s99 != seq).sum()
    cd = diffusion.CoordinateDiffuser(T=100, device=DEV)
    xyz = torch.randn(bsz, L, 3, device=DEV)
    x_t, eps = cd.diffuse_from_t0(xyz, torch.randint(0, 100, (bsz,), device=DEV), gm, return_eps=True)
    assert x_t.shape == eps.shape == (bsz, L, 3)
    od = diffusion.OrientationDiffuser(T=100, device=DEV)
    O_t = od.diffuse_from_t0(so3.uniform(bsz, L, 3, 3, device=DEV), gm, full(50))
    assert O_t.shape == (bsz, L, 3, 3)


def test_add_noise_against_reference_golden():
    g = load_golden("add_noise.pt")
    batch = synth.make_patches(4, 128, seed=g["seed_patches"], with_distmat=False)
    for k, v in g["chk"].items():
        assert checksum(batch[k]) == pytest.approx(v, rel=1e-12)
    sched = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    dsched = _lib.Schedule(sched, DEV)
    table = so3.SO3(sched["one_minus_alpha_bar_sqrt"], device=DEV)
    table._histograms = _hist(g["hist_rows"], DEV)
    for seed, key, mask in ((g["seed_noise"], "out", batch["generation_mask"]),
                            (g["seed_noise_all"], "out_all", torch.ones(4, 128, dtype=torch.bool))):
        torch.manual_seed(seed)
        noise = odiff.draw_add_noise_tensors(4, 128)         # CPU generator, reference's draw order
        out = diffusion.fused_add_noise(dsched, table, batch["seq_idx"].to(DEV), batch["xyz"][:, :, 1].contiguous().to(DEV),
                                        batch["orientations"].to(DEV), mask.to(DEV), g["t"].to(DEV), _cuda(noise))
        ref = g[key]
        assert torch.equal(out["seq_idx_t"].cpu(), ref["seq_idx_t"])               # bit-exact integers
        assert (out["seq_posterior"].cpu() - ref["seq_posterior"]).abs().max() < 1e-6
        assert torch.equal(out["translations_eps"].cpu(), ref["translations_eps"])
        assert (out["translations_t"].cpu() - ref["translations_t"]).abs().max() < 1e-5
        d = (out["orientations_t"].cpu() - ref["orientations_t"]).abs()
        assert d.max() < 1e-4, float(d.max())
        # context residues are copied bit-for-bit
        keep = ~mask
        assert torch.equal(out["orientations_t"].cpu()[keep], batch["orientations"][keep])
        assert torch.equal(out["translations_t"].cpu()[keep], batch["xyz"][:, :, 1][keep])


def test_add_noise_full_table_on_device_matches_golden_sequences():
    """Same as above but with the IGSO(3) table computed by our own kernel (tolerance-only table):
    sequences stay bit-exact, orientations within tolerance."""
    g = load_golden("add_noise.pt")
    batch = synth.make_patches(4, 128, seed=g["seed_patches"], with_distmat=False)
    sched = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    od = diffusion.OrientationDiffuser(T=100, s=0.01, beta_max=0.999, device=DEV)
    torch.manual_seed(g["seed_noise"])
    noise = odiff.draw_add_noise_tensors(4, 128)
    out = diffusion.fused_add_noise(_lib.Schedule(sched, DEV), od.so3, batch["seq_idx"].to(DEV),
                                    batch["xyz"][:, :, 1].contiguous().to(DEV), batch["orientations"].to(DEV),
                                    batch["generation_mask"].to(DEV), g["t"].to(DEV), _cuda(noise))
    assert torch.equal(out["seq_idx_t"].cpu(), g["out"]["seq_idx_t"])
    assert (out["orientations_t"].cpu() - g["out"]["orientations_t"]).abs().max() < 1e-3


def test_reverse_step_against_oracle_golden():
    g = load_golden("denoiser.pt")
    r = load_golden("reverse_step.pt")
    batch = synth.make_patches(2, 128, seed=g["seed_patches"], with_distmat=False)
    sched = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    dsched = _lib.Schedule(sched, DEV)
    table = so3.SO3(sched["beta"].sqrt(), device=DEV)
    table._histograms = _hist(r["hist_rev_rows"], DEV)
    noise = osamp.draw_step_noise(2, 128, generator=torch.Generator().manual_seed(r["seed_noise"]))
    n, d = g["noised"], g["denoised"]
    # the kernel takes the predicted rotation vector; recover it from O0 = O_t @ exp(v)
    from oracle import so3 as oso3
    v_theta = oso3.log_vec(n["orientations_t"].transpose(-1, -2) @ d["orientations_t0"])
    for tkey, okey in (("t", "out"), ("t1", "out1")):
        out = diffusion.fused_reverse_step(dsched, table, n["seq_idx_t"].to(DEV), n["translations_t"].to(DEV),
                                           n["orientations_t"].to(DEV), d["translations_eps"].to(DEV),
                                           v_theta.to(DEV), d["seq_posterior"].to(DEV),
                                           batch["generation_mask"].to(DEV), r[tkey].to(DEV), _cuda(noise),
                                           return_O0=True)
        ref = r[okey]
        assert torch.equal(out["seq_idx"].cpu(), ref["seq_idx"])                   # bit-exact integers
        assert (out["translations"].cpu() - ref["translations"]).abs().max() < 1e-5
        assert (out["orientations"].cpu() - ref["orientations"]).abs().max() < 2e-4
        assert (out["orientations_t0"].cpu() - d["orientations_t0"]).abs().max() < 2e-4
        keep = ~batch["generation_mask"]
        assert torch.equal(out["translations"].cpu()[keep], n["translations_t"][keep])
        assert torch.equal(out["orientations"].cpu()[keep], n["orientations_t"][keep])


def test_edge_sizes():
    sched = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    sd = diffusion.SequenceDiffuser(T=100, device=DEV)
    out = sd.forward_prob_from_t0(torch.zeros(0, 5, dtype=torch.long, device=DEV),
                                  torch.zeros(0, dtype=torch.long, device=DEV),
                                  torch.zeros(0, 5, dtype=torch.bool, device=DEV))
    assert out.shape == (0, 5, 21)
    # t = T (alpha_bar ~ 1e-15): probabilities are uniform on generated residues
    p = sd.forward_prob_from_t0(torch.zeros(1, 3, dtype=torch.long, device=DEV),
                                torch.full((1,), 100, dtype=torch.long, device=DEV),
                                torch.ones(1, 3, dtype=torch.bool, device=DEV))
    assert (p - 1 / 21).abs().max() < 1e-6
