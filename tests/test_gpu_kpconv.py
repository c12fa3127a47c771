"""Parity of the KPConv operator set (grid subsampling, radius neighbours, KPConv) on the GPU against the oracle
(oracle/kpconv_oracle.py, itself pinned against the compiled reference cores in tests/test_kpconv_oracle.py) and,
when oracle/_ref is present, against those cores directly."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def cloud(seed, n, batches=1):
    from seggroup_b200 import synth
    return synth.make_cloud(seed, n, batches=batches)


def cu(a, dt=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    return (t if dt is None else t.to(dt)).cuda()


@pytest.mark.parametrize("dl", [0.04, 0.1, 0.5])
def test_grid_subsampling_compute(dl):
    """B2: numpy in / out, barycentres + feature means bit-exact, labels with the canonical tie rule."""
    from oracle import kpconv_oracle as K
    from seggroup_b200.kpconv_ops import grid_subsampling
    pts, _ = cloud(1, 20000)
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((len(pts), 4)).astype(np.float32)
    cls = rng.integers(0, 5, (len(pts), 2)).astype(np.int32)
    sub, subf, subc = grid_subsampling.compute(pts, features=feats, classes=cls, sampleDl=dl)
    rs, rf, rc, _ = K.grid_subsampling(pts, feats, cls, dl)
    assert sub.dtype == np.float32 and subc.dtype == np.int32 and subc.shape[1] == 2
    assert np.array_equal(sub, rs) and np.array_equal(subf, rf) and np.array_equal(subc, rc)
    only = grid_subsampling.compute(pts, sampleDl=dl)
    assert isinstance(only, np.ndarray) and np.array_equal(only, rs)
    p2, c2 = grid_subsampling.compute(pts, classes=cls[:, 0], sampleDl=dl)
    assert c2.shape == (len(rs), 1) and np.array_equal(c2[:, 0], rc[:, 0])


def test_grid_subsampling_errors():
    from seggroup_b200.kpconv_ops import grid_subsampling
    pts, _ = cloud(1, 8000)
    with pytest.raises(RuntimeError, match="points.shape is not"):
        grid_subsampling.compute(pts[:, :2])
    with pytest.raises(RuntimeError, match="features.shape is not"):
        grid_subsampling.compute(pts, features=np.zeros((5, 2), np.float32))
    with pytest.raises(RuntimeError, match="classes.shape is not"):
        grid_subsampling.compute(pts, classes=np.zeros((5,), np.int32))
    with pytest.raises(RuntimeError, match="Error parsing method"):
        grid_subsampling.compute(pts, method="median")


def test_grid_subsampling_vs_compiled_reference():
    from oracle import kpconv_oracle as K
    from oracle import kpconv_ref
    if not kpconv_ref.available():
        pytest.skip("oracle/_ref not built")
    from seggroup_b200.kpconv_ops import batch_grid_subsampling
    pts, lens = cloud(2, 12000, batches=3)
    sub, sb = batch_grid_subsampling(cu(pts), cu(lens, torch.int32), 0.05)
    rp, rb = kpconv_ref.batch_grid_subsampling(pts, lens, 0.05)
    assert np.array_equal(sb.cpu().numpy(), rb)
    sub = sub.cpu().numpy()
    s = so = 0
    for b, m in zip(lens, rb):
        perm = K.reference_to_canonical(rp[so:so + m], pts[s:s + b], 0.05)
        assert np.array_equal(sub[so:so + m], rp[so:so + m][perm])
        s += b; so += m


def test_big_voxels_keep_input_order():
    """a voxel holding thousands of points exercises the big-bucket sort; the fp32 sum must still run in input order"""
    from oracle import kpconv_oracle as K
    from seggroup_b200.kpconv_ops import grid_subsampling_op
    rng = np.random.default_rng(3)
    pts = (rng.random((6000, 3)) * np.array([2.0, 2.0, 0.5])).astype(np.float32)
    sub = grid_subsampling_op(cu(pts), 1.0).cpu().numpy()
    ref, _, _, _ = K.grid_subsampling(pts, dl=1.0)
    assert len(ref) <= 8 and np.array_equal(sub, ref)


@pytest.mark.parametrize("radius", [0.08, 0.2])
def test_batch_neighbors(radius):
    from oracle import kpconv_oracle as K
    from seggroup_b200.kpconv_ops import batch_grid_subsampling, batch_ordered_neighbors, ordered_neighbors
    pts, lens = cloud(3, 15000, batches=2)
    P, Lb = cu(pts), cu(lens, torch.int32)
    sub, sb = batch_grid_subsampling(P, Lb, 0.04)
    q, qb = batch_grid_subsampling(P, Lb, 0.08)
    nb = batch_ordered_neighbors(q, sub, qb, sb, radius).cpu().numpy()
    ref = K.batch_neighbors(q.cpu().numpy(), sub.cpu().numpy(), qb.cpu().numpy(), sb.cpu().numpy(), radius)
    assert nb.shape == ref.shape and nb.dtype == np.int32
    assert np.array_equal(nb, ref)
    # single cloud signature; queries == supports puts every point first in its own row
    one = sub[: int(sb[0])].contiguous()
    nb1 = ordered_neighbors(one, one, radius).cpu().numpy()
    assert np.array_equal(nb1[:, 0], np.arange(len(one)))


def test_batch_neighbors_vs_compiled_reference():
    from oracle import kpconv_oracle as K
    from oracle import kpconv_ref
    if not kpconv_ref.available():
        pytest.skip("oracle/_ref not built")
    from seggroup_b200.kpconv_ops import batch_ordered_neighbors
    pts, lens = cloud(4, 6000, batches=2)
    sub, sb = K.batch_grid_subsampling(pts, lens, 0.05)
    nb = batch_ordered_neighbors(cu(sub), cu(sub), cu(sb, torch.int32), cu(sb, torch.int32), 0.12).cpu().numpy()
    ref = kpconv_ref.batch_neighbors(sub, sub, sb, sb, 0.12, nanoflann=True)
    assert np.array_equal(nb, K.canonical_rows(ref, sub, sub))


@pytest.mark.parametrize("cin,cout", [(5, 64), (64, 64), (128, 128), (64, 32), (16, 16), (4, 32), (32, 32)])
@pytest.mark.parametrize("influence,mode", [("linear", "sum"), ("gaussian", "sum"), ("linear", "closest"), ("constant", "sum")])
def test_kpconv_forward_backward(cin, cout, influence, mode):
    from oracle import kpconv_oracle as K
    from seggroup_b200.kpconv_ops import KPConv_ops, batch_ordered_neighbors
    pts, lens = cloud(5, 6000)
    sub, _ = K.batch_grid_subsampling(pts, lens, 0.06)
    qs, _ = K.batch_grid_subsampling(pts, lens, 0.12)
    S, Q = cu(sub), cu(qs)
    radius, extent = 0.15, 0.06
    nb = batch_ordered_neighbors(Q, S, None, None, radius)
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    kp = (torch.rand(15, 3, generator=g) * 2 - 1) * 0.09
    kp[0] = 0
    feats = torch.randn(len(sub), cin, generator=g)
    kv = torch.randn(15, cin, cout, generator=g) * (1.0 / np.sqrt(cin * 15))
    fd = feats.cuda().requires_grad_(True); kd = kv.cuda().requires_grad_(True)
    out = KPConv_ops(Q, S, nb, fd, kp.cuda(), kd, extent, influence, mode)
    f64 = feats.double().requires_grad_(True); k64 = kv.double().requires_grad_(True)
    ref = K.kpconv_ops(Q.cpu(), S.cpu(), nb.cpu(), f64, kp, k64, extent, influence, mode, dtype=torch.float64)
    scale = float(ref.abs().max())
    assert float((out.detach().cpu().double() - ref.detach()).abs().max()) < 1e-4 * scale       # fp32 SIMT path: 1e-4 relative
    go = torch.randn(out.shape, generator=g)
    (out * go.cuda()).sum().backward()
    (ref * go.double()).sum().backward()
    assert float((fd.grad.cpu().double() - f64.grad).abs().max()) < 1e-4 * float(f64.grad.abs().max())
    assert float((kd.grad.cpu().double() - k64.grad).abs().max()) < 1e-4 * float(k64.grad.abs().max())


@pytest.mark.parametrize("cin,cout,wcap", [(64, 64, 64), (64, 64, 20), (32, 32, 40), (128, 64, 64), (64, 128, 48), (128, 128, 64), (32, 256, 33), (192, 16, 64)])
@pytest.mark.parametrize("influence,mode", [("linear", "sum"), ("gaussian", "sum"), ("linear", "closest"), ("constant", "sum")])
def test_kpconv_tensor_core_contraction(cin, cout, wcap, influence, mode):
    """sgb_kpconv_fwd_tc (tcgen05 contraction, TF32 x 3) against the fp64 restatement of convolution_ops.py:161-249 and
    against the fp32 SIMT kernel: 1e-4 relative to the former (north_star bar), 5e-5 to the latter (two fp32 summation orders)."""
    from oracle import kpconv_oracle as K
    from seggroup_b200 import _lib
    from seggroup_b200.kpconv_ops import KPConv_ops, batch_ordered_neighbors
    pts, lens = cloud(7, 5000)
    sub, _ = K.batch_grid_subsampling(pts, lens, 0.05)
    qs, _ = K.batch_grid_subsampling(pts, lens, 0.07)          # a query count that is not a multiple of the 128-row tile
    S, Q = cu(sub), cu(qs)
    radius, extent = 0.14, 0.055
    nb = batch_ordered_neighbors(Q, S, None, None, radius)[:, :wcap].contiguous()      # rows are distance sorted: keep the nearest
    assert nb.shape[1] <= 64 and (nb == len(sub)).any(), "the case must hold shadow neighbours"
    assert _lib.call("sgb_kpconv_tc_supported", nb.shape[1], cin, cout, 15, len(sub)) == 1
    g = torch.Generator().manual_seed(cin * 1000 + cout + wcap)
    kp = (torch.rand(15, 3, generator=g) * 2 - 1) * 0.08
    kp[0] = 0
    feats = torch.randn(len(sub), cin, generator=g)
    kv = torch.randn(15, cin, cout, generator=g) * (1.0 / np.sqrt(cin * 15))
    launches = _lib.launch_count()
    out = KPConv_ops(Q, S, nb, feats.cuda(), kp.cuda(), kv.cuda(), extent, influence, mode)
    assert _lib.launch_count() - launches == 2                  # prep + the tcgen05 kernel: the tensor-core path really ran
    simt = KPConv_ops(Q, S, nb, feats.cuda(), kp.cuda(), kv.cuda(), extent, influence, mode, tensor_cores=False)
    ref = K.kpconv_ops(Q.cpu(), S.cpu(), nb.cpu(), feats.double(), kp, kv.double(), extent, influence, mode, dtype=torch.float64)
    scale = float(ref.abs().max())
    assert float((out.cpu().double() - ref).abs().max()) < 1e-4 * scale
    assert float((out - simt).abs().max()) < 5e-5 * scale


@pytest.mark.parametrize("cin,cout,wcap", [(64, 64, 64), (64, 64, 20), (32, 32, 40), (128, 64, 64), (64, 128, 48), (128, 128, 64), (32, 96, 33)])
@pytest.mark.parametrize("influence,mode", [("linear", "sum"), ("gaussian", "sum"), ("linear", "closest"), ("constant", "sum")])
def test_kpconv_tensor_core_backward(cin, cout, wcap, influence, mode):
    """sgb_kpconv_bwd_tc (dK = WF^T g and GW = g K^T on tcgen05, TF32 x 3; gradient of convolution_ops.py:240-247) against the
    fp64 restatement (1e-4 relative, north_star bar) and the fp32 SIMT backward; dK must be bit-reproducible."""
    from oracle import kpconv_oracle as K
    from seggroup_b200 import _lib
    from seggroup_b200.kpconv_ops import KPConv_ops, batch_ordered_neighbors
    pts, lens = cloud(7, 5000)
    sub, _ = K.batch_grid_subsampling(pts, lens, 0.05)
    qs, _ = K.batch_grid_subsampling(pts, lens, 0.07)          # a query count that is neither a multiple of 64 nor of 128
    S, Q = cu(sub), cu(qs)
    radius, extent = 0.14, 0.055
    nb = batch_ordered_neighbors(Q, S, None, None, radius)[:, :wcap].contiguous()
    assert (nb == len(sub)).any(), "the case must hold shadow neighbours"
    assert _lib.call("sgb_kpconv_bwd_tc_supported", len(qs), nb.shape[1], cin, cout, 15, len(sub)) == 1
    g = torch.Generator().manual_seed(cin * 1000 + cout + wcap)
    kp = (torch.rand(15, 3, generator=g) * 2 - 1) * 0.08
    kp[0] = 0
    feats = torch.randn(len(sub), cin, generator=g)
    kv = torch.randn(15, cin, cout, generator=g) * (1.0 / np.sqrt(cin * 15))
    go = torch.randn(len(qs), cout, generator=g)

    def grads(tc):
        fd = feats.cuda().requires_grad_(True); kd = kv.cuda().requires_grad_(True)
        out = KPConv_ops(Q, S, nb, fd, kp.cuda(), kd, extent, influence, mode, tensor_cores=tc)
        launches = _lib.launch_count()
        (out * go.cuda()).sum().backward()
        return fd.grad, kd.grad, _lib.launch_count() - launches
    gf, gk, nl = grads(True)
    assert nl == 4                                              # dK kernel + ordered reduce + K_values image + dfeat kernel: the tcgen05 path ran
    gf2, gk2, _ = grads(True)
    assert torch.equal(gk, gk2)                                 # deterministic weight gradient
    sf, sk, _ = grads(False)
    f64 = feats.double().requires_grad_(True); k64 = kv.double().requires_grad_(True)
    ref = K.kpconv_ops(Q.cpu(), S.cpu(), nb.cpu(), f64, kp, k64, extent, influence, mode, dtype=torch.float64)
    (ref * go.double()).sum().backward()
    fs, ks = float(f64.grad.abs().max()), float(k64.grad.abs().max())
    assert float((gf.cpu().double() - f64.grad).abs().max()) < 1e-4 * fs
    assert float((gk.cpu().double() - k64.grad).abs().max()) < 1e-4 * ks
    assert float((gf - sf).abs().max()) < 5e-5 * fs and float((gk - sk).abs().max()) < 5e-5 * ks
    assert float((gf - gf2).abs().max()) < 1e-5 * fs            # scatter-add order only


def test_kpconv_tensor_core_backward_small_and_unsupported():
    from oracle import kpconv_oracle as K
    from seggroup_b200 import _lib
    from seggroup_b200.kpconv_ops import KPConv_ops
    assert _lib.call("sgb_kpconv_bwd_tc_supported", 1000, 65, 64, 64, 15, 1000) == 0     # rows wider than 64 neighbours
    assert _lib.call("sgb_kpconv_bwd_tc_supported", 1000, 40, 5, 64, 15, 1000) == 0      # first-layer Cin
    assert _lib.call("sgb_kpconv_bwd_tc_supported", 1000, 40, 64, 256, 15, 1000) == 0    # Cout > 128: SIMT backward
    g = torch.Generator().manual_seed(5)
    for n, n0, W, K_ in [(1, 7, 3, 15), (37, 50, 9, 15), (129, 300, 32, 13), (5, 40, 0, 15), (300, 200, 40, 1)]:
        q = torch.rand(n, 3, generator=g) * 0.2
        s = torch.rand(n0, 3, generator=g) * 0.2
        nb = torch.randint(0, n0 + 3, (n, W), generator=g).clamp(max=n0).to(torch.int32)   # id n0 = shadow neighbour
        kp = (torch.rand(K_, 3, generator=g) * 2 - 1) * 0.06
        feats = torch.randn(n0, 64, generator=g)
        kv = torch.randn(K_, 64, 32, generator=g) * 0.05
        go = torch.randn(n, 32, generator=g)
        fd = feats.cuda().requires_grad_(True); kd = kv.cuda().requires_grad_(True)
        out = KPConv_ops(q.cuda(), s.cuda(), nb.cuda(), fd, kp.cuda(), kd, 0.08, "linear", "sum")
        (out * go.cuda()).sum().backward()
        f64 = feats.double().requires_grad_(True); k64 = kv.double().requires_grad_(True)
        ref = K.kpconv_ops(q, s, nb, f64, kp, k64, 0.08, "linear", "sum", dtype=torch.float64)
        (ref * go.double()).sum().backward()
        assert float((fd.grad.cpu().double() - f64.grad).abs().max()) <= 1e-4 * max(float(f64.grad.abs().max()), 1e-6)
        assert float((kd.grad.cpu().double() - k64.grad).abs().max()) <= 1e-4 * max(float(k64.grad.abs().max()), 1e-6)


def test_kpconv_tensor_core_small_and_unsupported():
    from oracle import kpconv_oracle as K
    from seggroup_b200 import _lib
    from seggroup_b200.kpconv_ops import KPConv_ops
    assert _lib.call("sgb_kpconv_tc_supported", 65, 64, 64, 15, 1000) == 0       # rows wider than 64 neighbours
    assert _lib.call("sgb_kpconv_tc_supported", 40, 5, 64, 15, 1000) == 0        # first-layer Cin
    assert _lib.call("sgb_kpconv_tc_supported", 40, 64, 512, 15, 1000) == 0
    g = torch.Generator().manual_seed(3)
    for n, n0, W in [(1, 7, 3), (37, 50, 9), (129, 300, 32), (5, 40, 0)]:
        q = torch.rand(n, 3, generator=g) * 0.2
        s = torch.rand(n0, 3, generator=g) * 0.2
        nb = torch.randint(0, n0 + 3, (n, W), generator=g).clamp(max=n0).to(torch.int32)   # id n0 = shadow neighbour
        kp = (torch.rand(15, 3, generator=g) * 2 - 1) * 0.06
        feats = torch.randn(n0, 64, generator=g)
        kv = torch.randn(15, 64, 32, generator=g) * 0.05
        out = KPConv_ops(q.cuda(), s.cuda(), nb.cuda(), feats.cuda(), kp.cuda(), kv.cuda(), 0.08, "linear", "sum")
        ref = K.kpconv_ops(q, s, nb, feats.double(), kp, kv.double(), 0.08, "linear", "sum", dtype=torch.float64)
        assert out.shape == (n, 32)
        assert float((out.cpu().double() - ref).abs().max()) <= 1e-4 * max(float(ref.abs().max()), 1e-6)


@pytest.mark.parametrize("d,W", [(64, 33), (128, 17), (6, 5), (256, 40), (32, 1)])
def test_ind_max_pool_and_closest_pool(d, W):
    """a21 (network_blocks.py:49-81): forward bit-exact (max / copy are exact), gradients to 1e-6 (tie shares)."""
    from oracle import kpconv_oracle as K
    from seggroup_b200.kpconv_ops import closest_pool, ind_max_pool
    rng = np.random.default_rng(d * 100 + W)
    n1, n2 = 5000, 3000
    x = rng.standard_normal((n1, d)).astype(np.float32)
    x[rng.integers(0, n1, 200)] = np.round(x[rng.integers(0, n1, 200)], 1)      # ties between rows
    inds = rng.integers(0, n1, (n2, W)).astype(np.int32)
    fill = rng.integers(1, W + 1, n2)                                             # neighbour counts, the rest is shadow (= n1)
    inds[np.arange(W)[None, :] >= fill[:, None]] = n1
    inds[:7] = n1                                                                  # cells with only shadow entries
    inds[7:20, 1:] = inds[7:20, :1]                                                # duplicated listings (ties by construction)
    g = rng.standard_normal((n2, d)).astype(np.float32)
    for fn, ofn in ((ind_max_pool, K.ind_max_pool), (closest_pool, K.closest_pool)):
        xc = torch.tensor(x, requires_grad=True)
        ref = ofn(xc, torch.tensor(inds))
        ref.backward(torch.tensor(g))
        xg = cu(x).requires_grad_(True)
        out = fn(xg, cu(inds))
        out.backward(cu(g))
        assert np.array_equal(out.detach().cpu().numpy(), ref.detach().numpy()), fn.__name__
        err = (xg.grad.cpu() - xc.grad).abs().max().item()
        assert err <= 1e-5 * max(1.0, xc.grad.abs().max().item()), (fn.__name__, err)


def test_ind_max_pool_int64_indices_and_empty():
    from seggroup_b200.kpconv_ops import closest_pool, ind_max_pool
    x = torch.randn(100, 8, device="cuda")
    inds = torch.randint(0, 101, (50, 4), device="cuda")                          # int64 as TF's gather would accept
    a = ind_max_pool(x, inds)
    xe = torch.cat([x, x.amin(0, keepdim=True)])
    assert torch.equal(a, xe[inds].amax(1))
    assert closest_pool(x, inds[:0]).shape == (0, 8)
    assert ind_max_pool(x, inds[:0]).shape == (0, 8)


def test_kpfcnn_rigid_blocks_forward_backward():
    """a two-level encoder / decoder built from the rigid blocks of network_blocks.py (simple, resnetb,
    resnetb_strided, nearest_upsample, unary) on the library operators, against the fp64 restatement of the same blocks:
    the chained forward within 1e-4 relative; per block, on the same input, output within 1e-4 and input / parameter
    gradients within 2e-4 of their largest entry (for the GPU's LeakyReLU active set)."""
    from types import SimpleNamespace
    from oracle import kpconv_oracle as K
    from seggroup_b200 import kpconv_blocks as B
    from seggroup_b200.kpconv_ops import batch_ordered_neighbors
    pts, lens = cloud(11, 4000)
    p0, _ = K.batch_grid_subsampling(pts, lens, 0.06)
    p1, _ = K.batch_grid_subsampling(p0, np.array([len(p0)], np.int32), 0.12)
    P0, P1 = cu(p0), cu(p1)
    r0, r1 = 0.15, 0.30
    inputs = {"points": [P0, P1], "neighbors": [batch_ordered_neighbors(P0, P0, None, None, r0)[:, :40].contiguous(),
                                                 batch_ordered_neighbors(P1, P1, None, None, r1)[:, :40].contiguous()],
              "pools": [batch_ordered_neighbors(P1, P0, None, None, r0)[:, :40].contiguous()],
              "upsamples": [batch_ordered_neighbors(P0, P1, None, None, r1)[:, :1].contiguous()]}
    g = torch.Generator().manual_seed(3)
    kp = torch.rand(15, 3, generator=g) * 2 - 1
    kp = kp / kp.norm(dim=1, keepdim=True) * torch.rand(15, 1, generator=g) ** (1 / 3)
    kp[0] = 0
    cfg = SimpleNamespace(KP_extent=1.0, density_parameter=5.0, KP_influence="linear", convolution_mode="sum", num_kernel_points=15,
                          use_batch_norm=True, batch_norm_momentum=0.99, fixed_kernel_points="center", K_points=kp)
    torch.manual_seed(5)
    arch = [("simple", 0, 4, 32, r0), ("resnetb", 0, 32, 32, r0), ("resnetb_strided", 0, 64, 64, r0), ("resnetb", 1, 128, 64, r1),
            ("nearest_upsample", 1, None, None, r1), ("unary", 0, 128, 32, r0)]
    blocks = [B.get_block_ops(n)(cin, fd, cfg).cuda() for n, _, cin, fd, _ in arch]
    for b in blocks:                                        # non-trivial BN affine parameters
        for nme, prm in b.named_parameters():
            if nme.endswith("bn.weight"):
                prm.data = 0.5 + torch.rand_like(prm)
            if nme.endswith("bn.bias"):
                prm.data = 0.2 * torch.randn_like(prm)
    feats = torch.randn(len(p0), 4, generator=g)

    # chained forward (the whole encoder / decoder) against the fp64 restatement
    x = feats.cuda()
    with torch.no_grad():
        for b, (n, li, _, _, rad) in zip(blocks, arch):
            x = b(li, inputs, x, rad, cfg, True)
    cpu_inputs = {k: [t.cpu() for t in v] for k, v in inputs.items()}
    y = feats.double()
    stage_in = []
    with torch.no_grad():
        for b, (n, li, _, _, rad) in zip(blocks, arch):
            stage_in.append(y)
            Pd = {k: v.detach().cpu().double() for k, v in b.named_parameters()}
            y = K.block_forward(n, Pd, li, cpu_inputs, y, rad, cfg)
    assert float((x.cpu().double() - y).abs().max()) < 1e-4 * float(y.abs().max())

    # every block on the SAME input (the fp64 features of that stage): output, input gradient and parameter gradients.
    # LeakyReLU's derivative jumps at 0, so gradients are compared for a common active set: the restatement's backward uses
    # the sign pattern the GPU saw, and the two patterns may differ only where the fp64 pre-activation is zero to rounding.
    import torch.nn.functional as F
    for b, (n, li, _, _, rad), fin in zip(blocks, arch, stage_in):
        masks = []
        orig_fwd = B.BatchNorm.forward

        def bn_fwd(self, x_, training=True, slope=1.0, residual=None):       # LeakyReLU is fused into the BN kernel: y > 0 <=> pre-activation > 0
            y_ = orig_fwd(self, x_, training, slope, residual)
            if slope != 1.0:
                masks.append((y_.detach() > 0).cpu())
            return y_
        B.BatchNorm.forward = bn_fwd
        try:
            fd = fin.float().cuda().requires_grad_(True)
            out = b(li, inputs, fd, rad, cfg, True)
        finally:
            B.BatchNorm.forward = orig_fwd
        go = torch.randn(out.shape, generator=g)
        (out * go.cuda()).sum().backward()
        f64 = fin.float().double().requires_grad_(True)
        Pd = {k: v.detach().cpu().double().requires_grad_(True) for k, v in b.named_parameters()}
        pre = []
        ref = K.block_forward(n, Pd, li, cpu_inputs, f64, rad, cfg, pre_activations=pre, lrelu_masks=iter(masks))
        assert len(pre) == len(masks), n
        for xr, m in zip(pre, masks):
            flipped = (xr > 0) != m
            assert float(xr[flipped].abs().max() if flipped.any() else 0.0) < 1e-4 * float(xr.abs().max()), n
        (ref * go.double()).sum().backward()
        assert float((out.detach().cpu().double() - ref.detach()).abs().max()) < 1e-4 * float(ref.detach().abs().max()), n
        assert float((fd.grad.cpu().double() - f64.grad).abs().max()) < 2e-4 * float(f64.grad.abs().max()) + 1e-12, n
        for k, v in b.named_parameters():
            gref = Pd[k].grad
            assert gref is not None, (n, k)
            assert float((v.grad.cpu().double() - gref).abs().max()) < 2e-4 * float(gref.abs().max()) + 1e-9, (n, k)
            v.grad = None


def test_kpconv_input_pipeline_matches_restatement():
    """tf_segmentation_inputs / big_neighborhood_filter / tf_stack_batch_inds / calibrate_neighbors
    (kpconv/datasets/common.py:377-384, 432-475, 551-652, 1021-1158) on the CUDA operators against the numpy restatement built
    on the oracle's pinned subsampling / neighbour functions: every index tensor of the flat input list bit-identical."""
    from types import SimpleNamespace
    from oracle import kpconv_oracle as K
    from seggroup_b200 import kpconv_inputs as KI
    cfg = SimpleNamespace(architecture=["simple", "resnetb", "resnetb_strided", "resnetb", "resnetb_deformable_strided", "resnetb_deformable",
                                        "nearest_upsample", "unary", "nearest_upsample", "unary"],
                          first_subsampling_dl=0.04, KP_extent=1.0, density_parameter=5.0, num_layers=3)
    batches = []
    for seed in (3, 4):
        pts, lens = cloud(seed, 4000, batches=3)
        sub, sl = K.batch_grid_subsampling(pts, lens, cfg.first_subsampling_dl)
        feats = np.ones((len(sub), 1), np.float32)
        labels = np.arange(len(sub), dtype=np.int32) % 20
        binds = np.repeat(np.arange(len(sl)), sl).astype(np.int32)
        batches.append((sub, feats, labels, sl, binds))
    to_dev = lambda b: (cu(b[0]), cu(b[1]), cu(b[2]), cu(b[3], torch.int32), cu(b[4], torch.int32))
    lim_ref = K.calibrate_neighbors(batches, cfg, keep_ratio=0.8, samples_threshold=10 ** 9)
    lim = KI.calibrate_neighbors([to_dev(b) for b in batches], cfg, keep_ratio=0.8, samples_threshold=10 ** 9)
    assert np.array_equal(lim, lim_ref) and len(lim) == 3 and (lim > 0).all()
    ref = K.segmentation_inputs(cfg, *batches[0], neighborhood_limits=lim_ref)
    got = KI.segmentation_inputs(cfg, *to_dev(batches[0]), neighborhood_limits=lim)
    assert len(got) == len(ref) == 4 * 3 + 5
    for i, (a, b) in enumerate(zip(got, ref)):
        a = a.cpu().numpy()
        assert a.shape == b.shape, (i, a.shape, b.shape)
        if a.dtype.kind == "f":
            assert np.array_equal(a, b.astype(a.dtype)), i                   # subsampled points / weights: bit-identical fp32
        else:
            assert np.array_equal(a, b), i
    assert torch.equal(KI.get_batch_inds(cu(batches[0][3], torch.int32)).cpu(), torch.as_tensor(batches[0][4]))
    eq = KI.stack_batch_inds(torch.tensor([4, 4, 4], dtype=torch.int32, device="cuda")).cpu().numpy()
    assert np.array_equal(eq, K.stack_batch_inds([4, 4, 4])) and eq.shape == (3, 5)


# ---- golden vectors minted by executing the unmodified reference on the TensorFlow stand-in (oracle/make_golden_kpconv.py) ----------
def _golden():
    import os
    from oracle import make_golden_kpconv as M
    g = M.geometry()
    gd = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return M, g, M.ops_inputs(g), np.load(os.path.join(gd, "kpconv_ref_ops.npz")), np.load(os.path.join(gd, "kpconv_ref_blocks.npz"))


def _close(a, b, tol):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-12)


def test_kpconv_ops_and_pools_match_reference_vectors():
    """a20 / a21 against the reference itself (fixtures from the unmodified convolution_ops.py / network_blocks.py): KPConv_ops in the six
    influence x aggregation modes (SIMT fp32 and tcgen05 paths), ind_max_pool, closest_pool — outputs and gradients, 1e-4 relative."""
    from seggroup_b200.kpconv_ops import KPConv_ops, closest_pool, ind_max_pool
    M, g, x, gold, _ = _golden()
    for infl, mode in M.OPS_MODES:
        for tc in (False, True):
            f, kv = cu(x["feats"]).requires_grad_(True), cu(x["kv"]).requires_grad_(True)
            y = KPConv_ops(cu(x["q"]), cu(x["s"]), cu(x["idx"]), f, cu(x["kp"]), kv, float(x["extent"]), infl, mode, tensor_cores=tc)
            (y * cu(x["go"])).sum().backward()
            tag = "ops/%s_%s/" % (infl, mode)
            assert _close(y, gold[tag + "out"], 1e-4), (tag, tc)
            assert _close(f.grad, gold[tag + "dfeats"], 1e-4) and _close(kv.grad, gold[tag + "dkv"], 1e-4), (tag, tc)
    for name, fn, idx, src in (("ind_max_pool", ind_max_pool, x["idx"], x["feats"]), ("closest_pool", closest_pool, g["up0"], x["go"][:, :16])):
        f = cu(np.ascontiguousarray(src)).requires_grad_(True)
        y = fn(f, cu(idx))
        (y * cu(gold[name + "/go"].astype(np.float32))).sum().backward()
        assert _close(y, gold[name + "/out"], 1e-6) and _close(f.grad, gold[name + "/dx"], 1e-5), name


def test_deformable_kpconv_matches_reference_vectors():
    """N3: KPConv_deform_ops (explicit offsets, +- modulations; linear / gaussian / constant / closest) and KPConv_deformable (offset
    convolution -> deformed convolution) against the fixtures minted from the unmodified convolution_ops.py:252-493: outputs and the
    gradients w.r.t. features, K_values, offsets, modulations, offset-convolution weights and bias."""
    from seggroup_b200.kpconv_ops import KPConv_deform_ops, KPConv_deformable
    M, g, x, gold, _ = _golden()
    for infl, mode, modulated in [("linear", "sum", False), ("linear", "sum", True), ("gaussian", "sum", False), ("constant", "sum", False),
                                  ("linear", "closest", False)]:
        f, kv = cu(x["feats"]).requires_grad_(True), cu(x["kv"]).requires_grad_(True)
        off = cu(x["offsets"]).requires_grad_(True)
        mod = cu(x["modulations"]).requires_grad_(True) if modulated else None
        y = KPConv_deform_ops(cu(x["q"]), cu(x["s"]), cu(x["idx"]), f, cu(x["kp"]), off, mod, kv, float(x["extent"]), infl, mode)
        (y * cu(x["go"])).sum().backward()
        tag = "deform_ops/%s_%s_%d/" % (infl, mode, int(modulated))
        assert _close(y, gold[tag + "out"], 1e-4), tag
        assert _close(f.grad, gold[tag + "dfeats"], 1e-4) and _close(kv.grad, gold[tag + "dkv"], 1e-4), tag
        assert _close(off.grad, gold[tag + "doffsets"], 2e-4) or np.abs(gold[tag + "doffsets"]).max() == 0, tag
        if np.abs(gold[tag + "doffsets"]).max() == 0:
            assert float(off.grad.abs().max()) == 0.0, tag
        if modulated:
            assert _close(mod.grad, gold[tag + "dmod"], 1e-4), tag
    for modulated in (False, True):
        f, kv = cu(x["feats"]).requires_grad_(True), cu(x["kv"]).requires_grad_(True)
        kv0 = cu(x["kv0m" if modulated else "kv0"]).requires_grad_(True)
        b0 = cu(x["b0m" if modulated else "b0"]).requires_grad_(True)
        y, _ = KPConv_deformable(cu(x["q"]), cu(x["s"]), cu(x["idx"]), f, kv, kv0, b0, KP_extent=float(x["extent"]), KP_influence="linear",
                                 aggregation_mode="sum", modulated=modulated, K_points=cu(x["kp"]))
        (y * cu(x["go"])).sum().backward()
        tag = "deformable/%d/" % int(modulated)
        assert _close(y, gold[tag + "out"], 1e-4), tag
        for a, k in ((f.grad, "dfeats"), (kv.grad, "dkv"), (kv0.grad, "dkv0"), (b0.grad, "db0")):
            assert _close(a, gold[tag + k], 3e-4), (tag, k)


def test_kpfcnn_blocks_match_reference_vectors():
    """N1 / N3: every block of network_blocks.py that the ScanNet architecture uses (training_Scannet.py:78-98: simple, resnetb,
    resnetb_strided, resnetb_deformable, resnetb_deformable_strided, nearest_upsample, unary; + simple_strided, max_pool) in training
    mode against the fixtures minted from the unmodified reference blocks: output 1e-4, gradients 1e-3 of their largest entry."""
    from test_kpconv_reference_pin import _restatement_params
    from seggroup_b200 import kpconv_blocks as B
    M, g, x, _, gold = _golden()
    cfg = M.config(g["kp_unit"])
    inputs = {"points": [cu(g["p0"]), cu(g["p1"])], "neighbors": [cu(g["nb0"]), cu(g["nb1"])], "pools": [cu(g["pool0"])], "upsamples": [cu(g["up0"])]}
    for name in M.BLOCKS:
        li, fdim, radius, feats, V = M.block_case(name, g)
        blk = B.get_block_ops(name)(feats.shape[1], fdim, cfg).cuda()
        P = _restatement_params(name, V)
        named = dict(blk.named_parameters())
        assert set(P) == set(named), (name, sorted(P), sorted(named))
        with torch.no_grad():
            for k, v in P.items():
                named[k].copy_(cu(v))
        f = cu(feats).requires_grad_(True)
        y = blk(li, inputs, f, radius, cfg, True)
        tag = "block/%s/" % name
        (y * cu(gold[tag + "go"].astype(np.float32))).sum().backward()
        assert _close(y, gold[tag + "out"], 1e-4), name
        assert _close(f.grad, gold[tag + "dfeats"], 1e-3), name
        inv = {v2: k2 for k2, v2 in zip(V.keys(), _restatement_params(name, {k: k for k in V}).keys())}
        for k, prm in named.items():
            assert prm.grad is not None, (name, k)
            assert _close(prm.grad, gold[tag + "d/" + inv[k]], 1e-3), (name, k)


def test_deformable_offsets_loss():
    """KPFCNN_model.py:218-286 ('fitting' and 'permissive' offset regularisers of one layer) against a brute-force torch fp64 evaluation."""
    from seggroup_b200.kpconv_ops import deformable_offsets_loss
    M, g, x, _, _ = _golden()
    q, s, idx, kp = x["q"], x["s"], x["idx"], x["kp"]
    ext = float(x["extent"])
    off = cu(x["offsets"]).requires_grad_(True)
    loss = deformable_offsets_loss(cu(q), cu(s), cu(idx), cu(kp), off, ext, "fitting")
    loss.backward()
    o64 = torch.tensor(x["offsets"], dtype=torch.float64, requires_grad=True)
    s_ext = torch.cat([torch.tensor(s, dtype=torch.float64), torch.full((1, 3), 1000.0, dtype=torch.float64)])
    nbr = s_ext[torch.tensor(idx).long()] - torch.tensor(q, dtype=torch.float64).unsqueeze(1)                 # [n,W,3]
    dkp = o64 + torch.tensor(kp, dtype=torch.float64)
    d2 = ((nbr.unsqueeze(2) - dkp.unsqueeze(1)) ** 2).sum(3)                                                    # [n,W,K]
    fit = (d2.min(1)[0] / ext ** 2).mean()
    locs = dkp / ext
    rep = 0
    for i in range(15):
        other = torch.cat([locs[:, :i], locs[:, i + 1:]], 1).detach()
        dist = torch.sqrt(((other - locs[:, i:i + 1]) ** 2).sum(2))
        rep = rep + (torch.clamp(1.5 - dist, min=0) ** 2).sum(1).mean()
    ref = fit + rep
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref))
    assert _close(off.grad, o64.grad.numpy(), 1e-3)
    perm = deformable_offsets_loss(cu(q), cu(s), cu(idx), cu(kp), off.detach(), ext, "permissive")
    assert abs(float(perm) - float(torch.clamp(torch.linalg.norm(locs.detach(), dim=2) - 1, min=0).mean())) < 1e-5


@pytest.mark.parametrize("n,d", [(1, 64), (300, 32), (5000, 128), (1237, 72)])
@pytest.mark.parametrize("slope,use_res", [(0.2, False), (1.0, False), (0.2, True)])
def test_bn_act_matches_torch(n, d, slope, use_res):
    """sgb_bn_act_fwd/_bwd (N1: batch_norm + leaky_relu of network_blocks.py:147-173, residual join of 337 / 581) against
    F.batch_norm + F.leaky_relu evaluated in fp64: outputs, running statistics, and the gradients w.r.t. x, the residual, gamma, beta;
    training and evaluation mode."""
    import torch.nn.functional as F
    from seggroup_b200 import kpconv_ops as KO
    g = torch.Generator().manual_seed(n * 7 + d)
    x = (torch.randn(n, d, generator=g) * 1.7 + 0.8).cuda().requires_grad_(True)
    res = torch.randn(n, d, generator=g).cuda().requires_grad_(True) if use_res else None
    gamma = (torch.rand(d, generator=g) + 0.5).cuda().requires_grad_(True)
    beta = torch.randn(d, generator=g).cuda().requires_grad_(True)
    w = torch.randn(n, d, generator=g).cuda()
    for training in ((True, False) if n > 1 else (False,)):
        rm, rv = torch.randn(d, generator=g).cuda() * 0.1, (torch.rand(d, generator=g) + 0.5).cuda()
        rm_ref, rv_ref = rm.double().clone(), rv.double().clone()
        for t in (x, gamma, beta) + ((res,) if use_res else ()):
            t.grad = None
        y = KO.bn_act(x, gamma, beta, rm, rv, residual=res, eps=1e-6, slope=slope, training=training, momentum=0.01)
        (y * w).sum().backward()
        got = [y.detach().clone(), x.grad.clone(), gamma.grad.clone(), beta.grad.clone()] + ([res.grad.clone()] if use_res else [])
        xd, gd, bd = x.detach().double().requires_grad_(True), gamma.detach().double().requires_grad_(True), beta.detach().double().requires_grad_(True)
        rd = res.detach().double().requires_grad_(True) if use_res else None
        z = F.batch_norm(xd, rm_ref, rv_ref, gd, bd, training, 0.01, 1e-6)
        if use_res:
            z = z + rd
        yr = F.leaky_relu(z, slope) if slope != 1.0 else z
        (yr * w.double()).sum().backward()
        ref = [yr.detach(), xd.grad, gd.grad, bd.grad] + ([rd.grad] if use_res else [])
        # the derivative of LeakyReLU jumps at 0: exclude nothing, fp32 pre-activations within 1e-6 of zero do not occur with these draws
        for a, b in zip(got, ref):
            assert torch.allclose(a.double(), b, rtol=2e-4, atol=2e-4 * float(b.abs().max())), (n, d, slope, use_res, training)
        assert torch.allclose(rm.double(), rm_ref, rtol=1e-5, atol=1e-6) and torch.allclose(rv.double(), rv_ref, rtol=1e-5, atol=1e-6)
