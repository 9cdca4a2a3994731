"""dab_pair_mlp_bwd_layer_sm100 (one layer of PairEmbedding's MLPs backward in one pass) against the same gradients
composed from PyTorch fp32 ops on the same bf16 inputs."""
import pytest
import torch

from diffab_pytorch_b200 import _lib
from diffab_pytorch_b200._lib import ptr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(B, L, seed, masked):
    g_ = torch.Generator(device=DEV).manual_seed(seed)
    P = B * L * L
    g = torch.randn(P, 64, device=DEV, generator=g_).bfloat16()
    a = torch.relu(torch.randn(P, 64, device=DEV, generator=g_)).bfloat16()
    W = (0.2 * torch.randn(64, 64, device=DEV, generator=g_)).bfloat16()
    m = torch.ones(B, L, device=DEV, dtype=torch.uint8)
    if masked:
        m[0, 1] = 0
        m[B - 1, L - 3] = 0
    return g, a, W, m


@pytest.mark.parametrize("B,L,masked,with_prev", [(2, 128, True, True), (3, 128, False, False), (1, 20, True, True),
                                                  (2, 100, True, True)])
def test_layer_backward_matches_composed_gradients(B, L, masked, with_prev):
    g, a, W, m = _case(B, L, 3 * B + L, masked)
    P = B * L * L
    pm = (m[:, :, None] & m[:, None, :]).reshape(P, 1).float()
    a = (a.float() * pm).bfloat16()                       # rows of the layer input are zero for masked pairs
    dW = torch.zeros(64, 64, device=DEV)
    db = torch.zeros(64, device=DEV)
    dbp = torch.zeros(64, device=DEV)
    g_prev = torch.empty(P, 64, device=DEV, dtype=torch.bfloat16)
    lib = _lib.lib()
    for rep in range(2):                                    # the outputs are accumulated: the second call doubles them
        _lib.check(lib.dab_pair_mlp_bwd_layer_sm100(ptr(g), ptr(a), ptr(W), ptr(m) if masked else None, B, L, ptr(g_prev),
                                                    ptr(dW), ptr(db), ptr(dbp) if with_prev else None, _lib.stream_ptr()),
                   "dab_pair_mlp_bwd_layer_sm100")
    torch.cuda.synchronize()
    gf, af, Wf = g.double(), a.double(), W.double()
    ref_dW = 2 * gf.t() @ af
    ref_db = 2 * (gf * pm.double()).sum(0)
    ref_prev = ((gf @ Wf) * (af > 0)).float()
    assert (dW.double() - ref_dW).abs().max() <= 1e-5 * ref_dW.abs().max()
    assert (db.double() - ref_db).abs().max() <= 1e-5 * ref_db.abs().max() + 1e-3
    # bf16 rounding of an fp32 accumulation against the rounding of the fp64 value: one bf16 ulp
    err = (g_prev.float() - ref_prev).abs()
    assert float((err / (ref_prev.abs() + 1e-2)).max()) < 1e-2
    assert torch.equal(g_prev == 0, ref_prev.bfloat16() == 0) or float(((g_prev == 0) != (ref_prev.bfloat16() == 0)).float().mean()) < 1e-5
    if with_prev:
        ref_dbp = 2 * g_prev.double().sum(0)
        assert (dbp.double() - ref_dbp).abs().max() <= 1e-5 * ref_dbp.abs().max() + 1e-3


def test_layer_backward_without_data_gradient():
    g, a, W, m = _case(1, 128, 5, False)
    dW = torch.zeros(64, 64, device=DEV)
    db = torch.zeros(64, device=DEV)
    lib = _lib.lib()
    _lib.check(lib.dab_pair_mlp_bwd_layer_sm100(ptr(g), ptr(a), ptr(W), None, 1, 128, None, ptr(dW), ptr(db), None,
                                                _lib.stream_ptr()), "dab_pair_mlp_bwd_layer_sm100")
    torch.cuda.synchronize()
    ref = g.double().t() @ a.double()
    assert (dW.double() - ref).abs().max() <= 1e-5 * ref.abs().max()
    assert (db.double() - g.double().sum(0)).abs().max() <= 1e-3


@pytest.mark.parametrize("B", [1, 3])
def test_table_gradients_on_tensor_cores_match_index_add(B):
    """dab_pair_table_grad_sm100 (one-hot GEMMs) against torch.index_add_ on the same bf16 gradient, and against the
    CUDA-core kernel dab_pair_table_grad; several chains, repeated residue indices, offsets beyond the clamp."""
    L, D, md = 128, 64, 32
    g_ = torch.Generator(device=DEV).manual_seed(B)
    g1 = torch.randn(B * L * L, D, device=DEV, generator=g_).bfloat16()
    seq = torch.randint(0, 21, (B, L), device=DEV, generator=g_)
    ridx = torch.cumsum(torch.randint(0, 3, (B, L), device=DEV, generator=g_), 1)          # repeats and gaps
    chain = torch.randint(0, 4, (B, L), device=DEV, generator=g_)                           # 0 = padding: weight zero
    lib = _lib.lib()
    s_type = torch.zeros(441, D, device=DEV)
    s_rel = torch.zeros(2 * md + 1, D, device=DEV)
    _lib.check(lib.dab_pair_table_grad_sm100(ptr(g1), ptr(seq), ptr(ridx), ptr(chain), B, L, md, ptr(s_type), ptr(s_rel),
                                             _lib.stream_ptr()), "dab_pair_table_grad_sm100")
    t_old = torch.zeros(441, D, device=DEV)
    r_old = torch.zeros(2 * md + 1, D, device=DEV)
    ws = torch.empty(lib.dab_pair_table_grad_workspace_bytes(B, L, md), device=DEV, dtype=torch.uint8)
    _lib.check(lib.dab_pair_table_grad(ptr(g1), ptr(seq), ptr(ridx), ptr(chain), B, L, md, ptr(t_old), ptr(r_old), ptr(ws),
                                       ws.numel(), _lib.stream_ptr()), "dab_pair_table_grad")
    torch.cuda.synchronize()
    pt = (seq[:, :, None] * 21 + seq[:, None, :]).reshape(-1)
    rel = ((ridx[:, :, None] - ridx[:, None, :]).clamp(-md, md) + md).reshape(-1)
    cp = (chain[:, :, None] * chain[:, None, :]).reshape(-1, 1).double()
    ref_t = torch.zeros(441, D, device=DEV, dtype=torch.float64).index_add_(0, pt, g1.double())
    ref_r = torch.zeros(2 * md + 1, D, device=DEV, dtype=torch.float64).index_add_(0, rel, g1.double() * cp)
    assert (s_type.double() - ref_t).abs().max() <= 1e-5 * ref_t.abs().max() + 1e-4
    assert (s_rel.double() - ref_r).abs().max() <= 1e-5 * ref_r.abs().max() + 1e-4
    assert (s_type - t_old).abs().max() <= 1e-4 * t_old.abs().max() + 1e-4
    assert (s_rel - r_old).abs().max() <= 1e-4 * r_old.abs().max() + 1e-4


def test_mixed_precision_mlp_run_matches_fp32_autograd():
    """_MixedMlpFunction (dab_linear_bf16 / dab_bias_grad / dab_gemm_bf16_tn) against the same three layers in fp32 PyTorch:
    output and every gradient within bf16 tolerance; the run is picked up by _run_mlp only inside _mixed_glue."""
    import torch.nn as nn
    from diffab_pytorch_b200.diffab_pytorch import _mixed_glue, _mlp, _run_mlp
    torch.manual_seed(0)
    mlp = _mlp([256, 128, 128, 64]).to(DEV)                # Linear, ReLU, Linear, ReLU, Linear
    x = torch.randn(4, 128, 256, device=DEV, requires_grad=True)
    gy = torch.randn(4, 128, 64, device=DEV)
    ref = mlp(x)
    ref.backward(gy)
    want = [x.grad.clone()] + [p.grad.clone() for p in mlp.parameters()]
    x.grad = None
    mlp.zero_grad()
    with _mixed_glue(True):
        got_y = _run_mlp(mlp, x)
    assert got_y.dtype == torch.float32 and got_y.shape == ref.shape
    got_y.backward(gy)
    got = [x.grad.clone()] + [p.grad.clone() for p in mlp.parameters()]
    assert float((got_y.detach() - ref.detach()).abs().max() / ref.detach().abs().max()) < 2e-2
    for a, b in zip(got, want):
        assert a.shape == b.shape
        # gradients behind a ReLU: a bf16 pre-activation on the other side of zero flips a whole term (random inputs are
        # centred on zero, the worst case); measured 5 % Frobenius for dx behind two ReLUs, < 1 % for the last layer
        assert float((a - b).norm() / b.norm()) < 0.1
        assert float((a.flatten() @ b.flatten()) / (a.norm() * b.norm())) > 0.995
    # outside the context the layers stay on the fp32 path (bit-identical forward)
    assert torch.equal(_run_mlp(mlp, x), ref)
