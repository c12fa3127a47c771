"""tcgen05 tensor-core primitives (TF32 x 3 split) against fp64 torch: descriptors, TMEM round trip, both operand majors."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-30))


@pytest.mark.parametrize("M,N,K", [(1, 64, 64), (100, 64, 36), (128, 192, 192), (300, 256, 256), (1117, 192, 192), (3001, 256, 256), (129, 16, 8)])
def test_gemm_tf32x3(M, N, K):
    from seggroup_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g) * 0.3
    C = ops.gemm_tf32x3(A.cuda(), B.cuda())
    ref = A.double() @ B.double().t()
    # fp32 SIMT-level accuracy: the split keeps ~21 mantissa bits per product
    assert _rel(C, ref) < 1e-5, _rel(C, ref)
    # and clearly better than plain TF32 would be (1e-3): sanity check that all three partial products are present
    assert _rel(C, ref) < 1e-4


def test_gemm_tf32x3_split_k():
    """tall contraction (the GCN weight gradient dW = dZ^T X with K = thousands of clusters): split-K slabs summed in order"""
    from seggroup_b200 import ops
    g = torch.Generator().manual_seed(5)
    for M, N, K in [(192, 192, 8600), (256, 256, 5368), (256, 256, 1024), (192, 192, 1028)]:
        A = torch.randn(M, K, generator=g)
        B = torch.randn(N, K, generator=g) * 0.3
        C = ops.gemm_tf32x3(A.cuda(), B.cuda())
        assert _rel(C, A.double() @ B.double().t()) < 1e-5


@pytest.mark.parametrize("S,C", [(351, 192), (1047, 256), (66, 192), (2, 256), (8599, 192)])
def test_gcn_linear_relu_forward_backward(S, C):
    """relu(fc(x)) of the GCN layer (model.py:146-151) on tcgen05 with the ReLU fused into the epilogue, and its backward (two
    more GEMMs), against torch fp64."""
    from seggroup_b200.pipeline import LinearReluFn
    g = torch.Generator().manual_seed(S + C)
    x = torch.randn(S, C, generator=g)
    W = torch.randn(C, C, generator=g) * (1.0 / C ** 0.5)
    go = torch.randn(S, C, generator=g)
    xd, Wd = x.cuda().requires_grad_(True), W.cuda().requires_grad_(True)
    y = LinearReluFn.apply(xd, Wd)
    (y * go.cuda()).sum().backward()
    x64, W64 = x.double().requires_grad_(True), W.double().requires_grad_(True)
    z = x64 @ W64.t()
    # common active set (the derivative of ReLU jumps at 0): the GPU's
    mask = (y.detach() > 0).cpu()
    flipped = (z.detach() > 0) != mask
    assert float(z.detach()[flipped].abs().max() if flipped.any() else 0.0) < 1e-5 * float(z.detach().abs().max())
    ref = torch.relu(z)
    (z * mask * go.double()).sum().backward()
    assert _rel(y.detach(), ref.detach()) < 1e-5
    assert _rel(xd.grad, x64.grad) < 1e-5 and _rel(Wd.grad, W64.grad) < 1e-5


@pytest.mark.parametrize("n_points,neg_gamma", [(8000, False), (8000, True), (333, False), (20011, False)])
def test_edgeconv_tensor_core_path(n_points, neg_gamma):
    """MLP3 forward with the second layer on tcgen05 (inference path) against the oracle and against the SIMT path."""
    import numpy as np
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    g = torch.Generator().manual_seed(n_points)
    x9 = torch.randn(n_points, 9, generator=g)
    x9[:, 3:6] = torch.rand(n_points, 3, generator=g) * 2 - 1
    knn = torch.randint(0, n_points, (n_points, 20), generator=g)
    knn[:, 0] = torch.arange(n_points)
    p = O.init_params(1)
    torch.manual_seed(11)
    p["mlp_3.bn1.weight"] = 0.5 + torch.rand(64); p["mlp_3.bn1.bias"] = 0.2 * torch.randn(64)
    p["mlp_3.bn2.weight"] = 0.5 + torch.rand(64); p["mlp_3.bn2.bias"] = 0.2 * torch.randn(64)
    if neg_gamma:
        p["mlp_3.bn2.weight"][::3] *= -1.0           # exercises the min-candidate branch
        p["mlp_3.bn2.weight"][5] = 0.0
    ref = O.mlp3_forward(p, x9, knn).detach()
    c = lambda k: p[k].detach().cuda()
    args = (x9.cuda(), knn.to(torch.int32).cuda(), c("mlp_3.conv1.0.weight"), c("mlp_3.bn1.weight"), c("mlp_3.bn1.bias"),
            c("mlp_3.conv2.0.weight"), c("mlp_3.bn2.weight"), c("mlp_3.bn2.bias"))
    train = ops.edgeconv_fwd(*args, want_backward=True)        # GRAM variant: also the hidden-layer second moments
    tc = ops.edgeconv_fwd(*args, want_argk=True, want_backward=False)
    torch.cuda.synchronize()
    for o in (tc, train):
        assert _rel(o["out"], ref) < 1e-4, _rel(o["out"], ref)
    assert _rel(tc["out"], train["out"]) < 1e-5
    assert _rel(tc["stats2"], train["stats2"]) < 1e-5
    assert _rel(tc["var2"], train["var2"]) < 1e-5
    same = (tc["argk"] == train["argk"]).float().mean().item()
    assert same > 0.999, same
    # hidden-layer moments against an fp64 evaluation of model.py:131-133 (conv1 -> BN(train) -> LeakyReLU)
    xi = x9.double().unsqueeze(1).expand(-1, 20, -1)
    e = torch.cat([x9.double()[knn] - xi, xi], -1).reshape(-1, 18)
    y = e @ p["mlp_3.conv1.0.weight"].detach().double().reshape(64, 18).t()
    y = (y - y.mean(0)) / torch.sqrt(y.var(0, unbiased=False) + 1e-5) * p["mlp_3.bn1.weight"].double() + p["mlp_3.bn1.bias"].double()
    h = torch.nn.functional.leaky_relu(y, 0.2)
    m2 = train["mom2"].cpu()
    G = m2[:4096].view(64, 64)
    assert _rel(0.5 * (G + G.t()), h.t() @ h) < 2e-5, _rel(0.5 * (G + G.t()), h.t() @ h)
    assert _rel(m2[4096:], h.sum(0)) < 2e-5
