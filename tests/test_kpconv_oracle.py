"""Pins the KPConv-operator oracle (oracle/kpconv_oracle.py) against the compiled, unmodified reference C++ cores
(oracle/_ref/libkpconv_ref.so, built by oracle/build_ref.py).  CPU only."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref():
    from oracle import kpconv_ref
    if not kpconv_ref.available():
        pytest.skip("oracle/_ref is not built and the reference tree is absent")
    return kpconv_ref


def cloud(seed, n, batches=1):
    from seggroup_b200 import synth
    return synth.make_cloud(seed, n, batches=batches)


def test_grid_subsampling_points_features_labels(ref):
    from oracle import kpconv_oracle as K
    pts, _ = cloud(1, 8000)
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((len(pts), 4)).astype(np.float32)
    cls = rng.integers(0, 5, (len(pts), 2)).astype(np.int32)
    for dl in (0.04, 0.1):
        rp, rf, rc = ref.grid_subsampling(pts, feats, cls, dl)
        sub, subf, subc, _ = K.grid_subsampling(pts, feats, cls, dl)
        perm = K.reference_to_canonical(rp, pts, dl)
        assert len(sub) == len(rp)
        assert np.array_equal(sub, rp[perm])                 # barycentres bit-exact
        assert np.array_equal(subf, rf[perm])                # feature means bit-exact
        # labels: equal wherever the vote is not tied (the reference's tie rule is hash-map order)
        diff = (subc != rc[perm])
        assert diff.mean() < 0.2


def test_batch_grid_subsampling(ref):
    from oracle import kpconv_oracle as K
    pts, lens = cloud(2, 6000, batches=3)
    rp, rb = ref.batch_grid_subsampling(pts, lens, 0.05)
    sub, sb = K.batch_grid_subsampling(pts, lens, 0.05)
    assert np.array_equal(rb, sb)
    s = so = 0
    for b, m in zip(lens, sb):
        perm = K.reference_to_canonical(rp[so:so + m], pts[s:s + b], 0.05)
        assert np.array_equal(sub[so:so + m], rp[so:so + m][perm])
        s += b; so += m


@pytest.mark.parametrize("radius", [0.08, 0.15])
def test_batch_neighbors(ref, radius):
    from oracle import kpconv_oracle as K
    pts, lens = cloud(3, 5000, batches=2)
    sub, sb = K.batch_grid_subsampling(pts, lens, 0.04)
    q, qb = K.batch_grid_subsampling(pts, lens, 0.08)
    for nanoflann in (True, False):
        nb = ref.batch_neighbors(q, sub, qb, sb, radius, nanoflann=nanoflann)
        mine = K.batch_neighbors(q, sub, qb, sb, radius)
        assert nb.shape == mine.shape
        assert np.array_equal(K.canonical_rows(nb, q, sub), mine)


def test_kpconv_ops_shapes_and_modes():
    import torch
    from oracle import kpconv_oracle as K
    g = torch.Generator().manual_seed(0)
    n, n0, W, Kp, ci, co = 50, 80, 12, 15, 8, 16
    q = torch.rand(n, 3, generator=g); s = torch.rand(n0, 3, generator=g)
    idx = torch.randint(0, n0 + 1, (n, W), generator=g)
    f = torch.randn(n0, ci, generator=g); kp = torch.rand(Kp, 3, generator=g) * 0.2 - 0.1
    kv = torch.randn(Kp, ci, co, generator=g)
    for infl in ("linear", "constant", "gaussian"):
        for mode in ("sum", "closest"):
            out = K.kpconv_ops(q, s, idx, f, kp, kv, 0.3, infl, mode)
            assert out.shape == (n, co) and torch.isfinite(out).all()
    # a row made only of shadow neighbours contributes nothing for 'linear'
    idx[0] = n0
    assert float(K.kpconv_ops(q, s, idx, f, kp, kv, 0.3, "linear", "sum")[0].abs().max()) == 0.0


def test_index_pool_oracle_matches_loop_restatement():
    """a21 oracle vs a plain numpy loop written from network_blocks.py:49-81."""
    import torch
    from oracle import kpconv_oracle as K
    rng = np.random.default_rng(3)
    x = rng.standard_normal((40, 5)).astype(np.float32)
    inds = rng.integers(0, 41, (25, 6)).astype(np.int32)
    xe_max = np.concatenate([x, x.min(0, keepdims=True)])
    xe_zero = np.concatenate([x, np.zeros((1, 5), np.float32)])
    want_max = np.stack([xe_max[r].max(0) for r in inds])
    want_closest = np.stack([xe_zero[r[0]] for r in inds])
    assert np.array_equal(K.ind_max_pool(torch.tensor(x), torch.tensor(inds)).numpy(), want_max)
    assert np.array_equal(K.closest_pool(torch.tensor(x), torch.tensor(inds)).numpy(), want_closest)
