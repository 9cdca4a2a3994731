"""GPU parity: SO(3) maps and IGSO(3) table/sampler vs the oracle and the reference goldens,
through the C ABI (ctypes).  Tolerances: fp32 path <= 1e-4 relative (north_star); bin indices
bit-exact."""
import math

import pytest
import torch

from conftest import load_golden
from diffab_pytorch_b200 import so3, synth
from oracle import diffusion as odiff
from oracle import so3 as oso3

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_maps_match_reference_goldens():
    g = load_golden("so3_maps.pt")
    R, k, v = g["R"].to(DEV), g["k"].to(DEV), g["v"].to(DEV)
    assert (so3.log_rotmat(R).cpu() - g["log_skew"]).abs().max() < 2e-5
    assert (so3.rotation_matrix_to_vector(R).cpu() - g["log_vec"]).abs().max() < 2e-5
    assert (so3.vector_to_rotation_matrix(v).cpu() - g["exp_vec"]).abs().max() < 2e-6
    assert (so3.exp_skew_symmetric_mat(g["log_skew"].to(DEV)).cpu() - g["exp_log"]).abs().max() < 2e-5
    # scale_rot: compare as rotations (angle of the relative rotation), away from theta ~ pi
    got = so3.scale_rot(R, k).cpu()
    rel = torch.einsum("blij,blkj->blik", got, g["scale_rot"])
    ang = torch.acos(((oso3.trace3(rel) - 1) / 2).clamp(-1, 1))
    far = ((oso3.trace3(g["R"]) - 1) / 2 + 1).abs() > 1e-2
    assert ang[far].max() < 1e-3


def test_reference_property_tests():
    # tests/test_so3.py:24-31 (skew symmetry), :44-62 (exp(log R) round trip), :79-93 (scale_rot orthogonal)
    bsz, L = 32, 100
    R = so3.uniform(bsz, L, 3, 3, device=DEV)
    S = so3.log_rotmat(R)
    assert S.shape == (bsz, L, 3, 3)
    assert torch.allclose(S, -S.transpose(2, 3))
    assert so3.skew_symmetric_mat_to_vector(S).shape == (bsz, L, 3)
    assert so3.tensor_trace(R).shape == (bsz, L)
    R_recon = so3.exp_skew_symmetric_mat(S)
    cos_theta = (so3.tensor_trace(R) - 1) / 2
    ok = ((cos_theta - 1).abs() >= 1e-2) & ((cos_theta + 1).abs() >= 1e-2)
    assert ((R - R_recon).abs().sum(dim=(-1, -2))[ok] < 1e-4).all()
    k = torch.rand(bsz, device=DEV)
    Rs = so3.scale_rot(R, k)
    prod = Rs.transpose(2, 3) @ Rs
    assert torch.allclose(prod[ok], torch.eye(3, device=DEV).expand_as(prod)[ok], rtol=1e-5, atol=1e-5)
    # any leading dims; empty input
    assert so3.rotation_matrix_to_vector(R[0, 0]).shape == (3,)
    assert so3.vector_to_rotation_matrix(torch.empty(0, 3, device=DEV)).shape == (0, 3, 3)


def test_large_flat_round_trip():
    # full-size property (H7): 2^22 rotations, exp(log R) == R
    n = 1 << 22
    R = synth.uniform_rotations(n, device=DEV)
    v = so3.rotation_matrix_to_vector(R)
    cos_theta = (so3.tensor_trace(R) - 1) / 2     # same exclusion as tests/test_so3.py:56-59 of the reference
    ok = ((cos_theta - 1).abs() >= 1e-2) & ((cos_theta + 1).abs() >= 1e-2)
    back = so3.vector_to_rotation_matrix(v)
    assert (back - R).abs().amax(dim=(-1, -2))[ok].max() < 5e-5


def test_exp_backward_matches_autograd_of_oracle():
    v = torch.randn(5, 7, 3)
    gR = torch.randn(5, 7, 3, 3)
    v64 = v.double().requires_grad_(True)
    (oso3.exp_vec(v64) * gR.double()).sum().backward()
    vg = v.to(DEV).requires_grad_(True)
    (so3.vector_to_rotation_matrix(vg) * gR.to(DEV)).sum().backward()
    assert (vg.grad.cpu() - v64.grad).abs().max() < 1e-4 * v64.grad.abs().max()


def test_igso3_table_rows_match_reference():
    g = load_golden("igso3_table.pt")
    sched = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    table = so3.SO3(sched["one_minus_alpha_bar_sqrt"], device=DEV).histograms.cpu()
    assert table.shape == (101, 8192)
    for r, ref in g["rows"].items():
        if r == 0:
            continue  # sigma = 0: all 1024 terms of an alternating series at full weight; never sampled (t >= 1)
        err = (table[r] - ref).abs().max() / ref.max()
        # tolerance-only parity: 1024-term fp32 series with sin arguments up to 3215 rad; the reference's own
        # table carries ~1e-4 of summation-order noise at the smallest sigmas
        assert err < 5e-4, (r, float(err))
    rev = so3.SO3(sched["beta"].sqrt(), device=DEV).histograms.cpu()
    for r, ref in g["rows_rev"].items():
        assert (rev[r] - ref).abs().max() / ref.max() < 5e-4, r
    assert (table >= 0).all() and torch.isfinite(table).all()


def _hist(rows, device):
    h = torch.zeros(101, 8192)
    for r, v in rows.items():
        h[int(r)] = v
    return h.to(device)


def test_igso3_sampler_bins_bit_exact_and_angles():
    g = load_golden("add_noise.pt")
    sched = odiff.cosine_schedule(100, s=0.01, beta_max=0.999)
    sampler = so3.SO3(sched["one_minus_alpha_bar_sqrt"], device=DEV)
    sampler._histograms = _hist(g["hist_rows"], DEV)          # the reference's own table rows
    t = g["t"]
    torch.manual_seed(g["seed_noise"])
    noise = odiff.draw_add_noise_tensors(4, 128)
    ref, ref_bins = oso3.igso3_sample(_hist(g["hist_rows"], "cpu"), sched["one_minus_alpha_bar_sqrt"], t, 128,
                                      noise["axis"], noise["hist_exp"], noise["jitter"], noise["gauss"],
                                      return_bins=True)
    got, bins = sampler.sample_isotropic_gaussian(t.to(DEV), 128, noise={k: v.to(DEV) for k, v in noise.items()},
                                                  return_bins=True)
    assert torch.equal(bins.cpu(), ref_bins)                  # integer work: bit-exact, both sigma branches
    assert (got.cpu() - ref).abs().max() < 1e-5
    # the sampling loop's form: the draw written into a caller's buffer (made on a side stream beside the epsilon network)
    buf = torch.empty(4, 128, 3, device=DEV)
    ret = sampler.sample_isotropic_gaussian(t.to(DEV), 128, noise={k: v.to(DEV) for k, v in noise.items()}, out=buf)
    assert ret.data_ptr() == buf.data_ptr() and torch.equal(buf, got)
    with pytest.raises(ValueError):
        sampler.sample_isotropic_gaussian(t.to(DEV), 128, noise={k: v.to(DEV) for k, v in noise.items()},
                                          out=torch.empty(4, 64, 3, device=DEV))
    # ragged / edge sizes: L = 1, L = 37, B = 1
    for B, L in ((1, 1), (3, 37)):
        tt = torch.tensor([1, 5, 100][:B])
        gen = torch.Generator().manual_seed(B * 100 + L)
        nz = {"axis": torch.randn(B, L, 3, generator=gen), "hist_exp": torch.empty(B, 8192).exponential_(generator=gen),
              "jitter": torch.rand(B, L, generator=gen), "gauss": torch.randn(B, L, generator=gen)}
        r, rb = oso3.igso3_sample(_hist(g["hist_rows"], "cpu"), sched["one_minus_alpha_bar_sqrt"], tt, L, nz["axis"],
                                  nz["hist_exp"], nz["jitter"], nz["gauss"], return_bins=True)
        o, ob = sampler.sample_isotropic_gaussian(tt.to(DEV), L, noise={k: v.to(DEV) for k, v in nz.items()},
                                                  return_bins=True)
        assert torch.equal(ob.cpu(), rb) and (o.cpu() - r).abs().max() < 1e-5
