"""Parity of every C-ABI op against the oracle (CPU restatement of the reference), on seeded inputs.
Index outputs are compared bit-exactly, fp32 outputs within 1e-4 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def dev(a, dt=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dt is not None:
        t = t.to(dt)
    return t.cuda()


def rel_err(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def levels_of(scene):
    from oracle import seggroup_oracle as O
    uf = O.SegUnionFind(scene.seg_offsets, scene.seg_members, np.asarray(scene.weak_label))
    return uf, O.Level(uf)


def merged_uf(scene, n_merge, seed=0):
    """Random admissible unions on the oracle union-find (to get big, multi-segment clusters)."""
    from oracle import seggroup_oracle as O
    uf = O.SegUnionFind(scene.seg_offsets, scene.seg_members, np.asarray(scene.weak_label))
    rng = np.random.default_rng(seed)
    for _ in range(n_merge):
        a, b = rng.integers(0, uf.S, 2)
        uf.union(uf.cid[a], uf.cid[b])
    return uf


def uf_to_device(uf):
    """oracle SegUnionFind -> device int[6][S1] state (parent, next, tail, pnum, ins, sem)."""
    S = uf.S
    parent = uf.cid.copy()
    nxt = np.full(S, -1, np.int64); tail = np.arange(S)
    for r in range(S):
        m = uf.members[r]
        if m:
            assert m[0] == r
            for a, b in zip(m[:-1], m[1:]):
                nxt[a] = b
            tail[r] = m[-1]
    return dev(np.stack([parent, nxt, tail, uf.pnum.astype(np.int64), uf.ins, uf.sem]), torch.int32)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 32768, 32769, 1_000_003])
def test_scan(n):
    from seggroup_b200 import ops
    x = np.random.default_rng(n).integers(0, 7, n).astype(np.int32)
    out = ops.exclusive_scan(dev(x)).cpu().numpy()
    ref = np.concatenate([[0], np.cumsum(x)]).astype(np.int32)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("C", [64, 128, 192, 256, 3])
@pytest.mark.parametrize("use_members", [True, False])
def test_segment_pool_max(C, use_members):
    from seggroup_b200 import ops
    rng = np.random.default_rng(C)
    N, S = 30000, 120
    feat = rng.integers(-4, 5, (N, C)).astype(np.float32) * 0.25        # many exact ties
    cuts = np.sort(rng.choice(np.arange(1, N), S - 1, replace=False))
    cuts[0] = 1                                                          # a single-row segment
    cuts = np.unique(cuts)
    off = np.concatenate([[0], cuts, [N]]).astype(np.int32)
    S = len(off) - 1
    members = rng.permutation(N).astype(np.int32) if use_members else None
    out, arg = ops.segment_pool_max(dev(feat), dev(off), dev(members) if use_members else None)
    out, arg = out.cpu().numpy(), arg.cpu().numpy()
    for s in range(S):
        rows = members[off[s]:off[s + 1]] if use_members else np.arange(off[s], off[s + 1])
        f = feat[rows]
        assert np.array_equal(out[s], f.max(0)), s
        assert np.array_equal(arg[s], rows[f.argmax(0)]), s              # numpy argmax = first maximum
    # backward: gradient lands on the argmax rows only
    g = rng.standard_normal((S, C)).astype(np.float32)
    gf = ops.segment_pool_max_bwd(dev(g), dev(arg), N).cpu().numpy()
    ref = np.zeros((N, C), np.float32)
    for s in range(S):
        ref[arg[s], np.arange(C)] += g[s]
    assert np.array_equal(gf, ref)


@pytest.mark.parametrize("C", [64, 128, 192])
@pytest.mark.parametrize("N", [4096, 20011, 131072 + 7])
def test_segment_pool_staged_edge_cases(C, N):
    """The shared-memory staged kernel (point-sized pools, C = 64 / 128; C = 192 takes the generic kernel): ragged tail chunk, one-row and two-row segments
    next to a segment of thousands of rows, whole segments of -inf, NaN rows (torch.max: the FIRST NaN wins), ties."""
    from seggroup_b200 import ops
    rng = np.random.default_rng(N + C)
    feat = rng.integers(-6, 7, (N, C)).astype(np.float32) * 0.125
    sizes = [1, 2, 1, 31, 32, 33, 64, 3000, 1, 95, 97]
    while sum(sizes) < N - 600:
        sizes.append(int(rng.integers(1, 500)))
    sizes.append(N - sum(sizes))
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    S = len(sizes)
    members = rng.permutation(N).astype(np.int32)
    seg_rows = lambda s: members[off[s]:off[s + 1]]
    feat[seg_rows(5)] = -np.inf                                          # a whole segment of -inf: first member wins
    feat[seg_rows(7)[100:2000:7], 3] = np.nan                            # NaNs inside the big segment (several chunks)
    feat[seg_rows(7)[1500], 5] = np.nan
    feat[seg_rows(9), 0] = np.nan                                        # all-NaN column of a one-row segment
    feat[seg_rows(S - 1)[-1], 1] = np.nan                                # NaN in the ragged tail chunk
    out, arg = ops.segment_pool_max(dev(feat), dev(off), dev(members))
    out, arg = out.cpu().numpy(), arg.cpu().numpy()
    ref_o, ref_a = torch.empty(S, C), torch.empty(S, C, dtype=torch.long)
    for s in range(S):
        v, i = torch.from_numpy(feat[seg_rows(s)]).max(0)                # torch semantics incl. NaN; ties checked below
        ref_o[s], ref_a[s] = v, i
    assert np.array_equal(out, ref_o.numpy(), equal_nan=True)
    for s in range(S):
        rows = seg_rows(s)
        f = feat[rows]
        isn = np.isnan(f)
        first = np.where(isn.any(0), isn.argmax(0), (f == np.nanmax(np.where(isn, -np.inf, f), 0)).argmax(0))
        assert np.array_equal(arg[s], rows[first]), s


def test_cluster_knn(scene20k):
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    sc = scene20k
    for uf in (merged_uf(sc, 0), merged_uf(sc, 400, 1)):
        L = O.Level(uf)
        order = np.concatenate(L.points).astype(np.int32)
        off = np.concatenate([[0], np.cumsum([len(p) for p in L.points])]).astype(np.int32)
        data = dev(sc.data)
        knn = ops.cluster_knn(data, dev(order), dev(off), 20).cpu().numpy()
        ref = O.cluster_knn(torch.from_numpy(sc.data[:, :3].copy()), L.points, 20, tie="canonical").numpy()
        bad = (knn != ref).any(1)
        assert bad.sum() == 0, "kNN rows differ: %d of %d (first %s)" % (bad.sum(), len(bad), np.nonzero(bad)[0][:5])


@pytest.mark.parametrize("case", ["plane_far", "wall", "duplicates", "line"])
def test_cluster_knn_exact_on_hard_geometry(case):
    """The sorted sweep must return exactly the brute-force ranking of the reference's fp32 score, also where that
    score is dominated by cancellation noise (far from the origin), where the sweep axis is degenerate and with ties."""
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    rng = np.random.default_rng(len(case))
    n_big = 6000
    if case == "plane_far":
        big = np.stack([rng.uniform(8, 16, n_big), rng.uniform(-14, -9, n_big), 2.5 + 0.002 * rng.standard_normal(n_big)], 1)
    elif case == "wall":
        big = np.stack([np.full(n_big, 3.25), rng.uniform(0, 6, n_big), rng.uniform(0, 2.5, n_big)], 1)
    elif case == "duplicates":
        base = np.stack([rng.uniform(0, 4, n_big // 4), rng.uniform(0, 4, n_big // 4), np.zeros(n_big // 4)], 1)
        big = np.tile(base, (4, 1))[rng.permutation(n_big)]
    else:
        big = np.stack([rng.uniform(0, 10, n_big), np.full(n_big, 1.0), np.full(n_big, 1.0)], 1)
    small = rng.uniform(0, 1, (300, 3))
    xyz = np.concatenate([big, small]).astype(np.float32)
    N = len(xyz)
    perm = rng.permutation(N)
    members = [perm[np.isin(perm, np.arange(n_big))], perm[~np.isin(perm, np.arange(n_big))]]   # member order = random
    order = np.concatenate(members).astype(np.int32)
    off = np.array([0, len(members[0]), N], np.int32)
    knn = ops.cluster_knn(dev(xyz), dev(order), dev(off), 20).cpu().numpy()
    ref = O.cluster_knn(torch.from_numpy(xyz), members, 20, tie="canonical").numpy()
    bad = (knn != ref).any(1)
    assert bad.sum() == 0, "kNN rows differ: %d of %d (first %s)" % (bad.sum(), len(bad), np.nonzero(bad)[0][:5])


def test_cluster_knn_small_clusters():
    """clusters with n <= k: first n columns = members, rest 0 (model.py:513-518)"""
    from seggroup_b200 import ops
    rng = np.random.default_rng(0)
    sizes = [1, 5, 20, 21, 3]
    N = sum(sizes)
    xyz = rng.standard_normal((N, 3)).astype(np.float32)
    order = rng.permutation(N).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    knn = ops.cluster_knn(dev(xyz), dev(order), dev(off), 20).cpu().numpy()
    for c, n in enumerate(sizes):
        m = order[off[c]:off[c + 1]]
        if n <= 20:
            for p in m:
                assert np.array_equal(knn[p, :n], m) and (knn[p, n:] == 0).all()
        else:
            for p in m:
                assert set(knn[p]) <= set(m) and len(set(knn[p])) == 20


@pytest.mark.parametrize("P", [64, 1024])
def test_cluster_cloud_indices(scene20k, P):
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    sc = scene20k
    uf = merged_uf(sc, 0 if P == 64 else 300, 2)
    L = O.Level(uf)
    order = np.concatenate(L.points).astype(np.int32)
    off = np.concatenate([[0], np.cumsum([len(p) for p in L.points])]).astype(np.int32)
    idx, status = ops.cluster_cloud_indices(dev(sc.data), dev(order), dev(off), P)
    idx = idx.cpu().numpy()
    assert int(status.item()) == 0
    for c, m in enumerate(L.points):
        li = O.cluster_cloud_indices(len(m), sc.data[m, :3], P)
        assert np.array_equal(idx[c], m[li]), c


def test_cluster_cloud_indices_mostly_coincident_members():
    """model.py:407-412 on clusters whose members are mostly coincident: the farthest-point picks run out of distinct points,
    the trailing picks equal member 0 and are overwritten by the leading ones — `choice[-invalid:] = choice[:invalid]` with
    OVERLAPPING slices when invalid > rem / 2 (numpy copies through a temporary).  Picks must equal the oracle's."""
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    rng = np.random.default_rng(4)
    clouds, order, off = [], [], [0]
    for n, distinct in ((40, 3), (50, 2), (37, 5), (63, 4), (33, 2)):          # P = 64 -> rem = 24, 14, 27, 1, 31
        base = rng.random((distinct, 3)).astype(np.float32)
        pts = base[rng.integers(0, distinct, n)]
        pts[0] = base[0]
        clouds.append(pts)
        order += list(range(off[-1], off[-1] + n))
        off.append(off[-1] + n)
    xyz = np.concatenate(clouds).astype(np.float32)
    data = np.concatenate([xyz, np.zeros_like(xyz)], 1)
    idx, status = ops.cluster_cloud_indices(dev(data), dev(np.array(order, np.int32)), dev(np.array(off, np.int32)), 64)
    idx = idx.cpu().numpy()
    hit_overlap = False
    for c, pts in enumerate(clouds):
        n = len(pts)
        li = O.cluster_cloud_indices(n, pts, 64)
        assert np.array_equal(idx[c] - off[c], li), (c, idx[c] - off[c], li)
        rem = 64 % n
        tail = li[64 - rem:] if rem else li[:0]
        hit_overlap |= rem > 0 and (O.fps_indices(pts, rem) == 0).sum() > rem / 2
    assert hit_overlap, "no case exercised the overlapping repair"
    assert int(status.item()) == 0


def test_cloud_transform_and_mlp1(scene20k):
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    sc = scene20k
    uf, L = levels_of(sc)
    cloud_idx = np.stack([m[O.cluster_cloud_indices(len(m), sc.data[m, :3], 64)] for m in L.points]).astype(np.int32)
    data = torch.from_numpy(sc.data.copy())
    clouds_ref = []
    for ci in cloud_idx:
        c = data[torch.as_tensor(ci, dtype=torch.long)].clone()
        c[:, :3] -= c[:, :3].mean(0)
        c[:, :3] /= torch.abs(c[:, :3]).max()
        clouds_ref.append(c)
    clouds_ref = torch.stack(clouds_ref)
    clouds = ops.cluster_cloud_transform(dev(sc.data), dev(cloud_idx))
    # bit-exact: the kernel reproduces torch-CPU's summation order for the [64,3] mean (cluster_cloud.cu)
    assert torch.equal(clouds.cpu(), clouds_ref), "transformed clouds differ in %d entries" % int((clouds.cpu() != clouds_ref).sum())
    p = O.init_params(1, 4.0)
    # feed the oracle's clouds to both sides so the kNN indices are comparable bit-exactly
    o = ops.mlp1_fwd(dev(clouds_ref), p["mlp_1.conv1.0.weight"].cuda(), p["mlp_1.bn1.weight"].cuda(), p["mlp_1.bn1.bias"].cuda())
    for k in O.TRAINABLE:
        p[k].requires_grad_(True)
    feat_ref, knn_ref = O.mlp1_forward(p, clouds_ref, tie="canonical")
    assert np.array_equal(o["knn"].cpu().numpy(), knn_ref.numpy())
    assert rel_err(o["feat"], feat_ref.detach()) < RTOL
    g = torch.randn(feat_ref.shape, generator=torch.Generator().manual_seed(0))
    (feat_ref * g).sum().backward()
    gW, gg, gb = ops.mlp1_bwd(g.cuda(), dev(clouds_ref), o["knn"], o["arg_pt"], p["mlp_1.conv1.0.weight"].detach().cuda(), o["stats"], o["mom"])
    assert rel_err(gW.view(-1), p["mlp_1.conv1.0.weight"].grad.view(-1)) < RTOL
    assert rel_err(gg, p["mlp_1.bn1.weight"].grad) < RTOL
    assert rel_err(gb, p["mlp_1.bn1.bias"].grad) < RTOL


@pytest.mark.parametrize("two", [False, True])
def test_edgeconv_fwd_bwd(scene8k, two):
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    sc = scene8k
    uf = merged_uf(sc, 20, 3)
    L = O.Level(uf)
    data = torch.from_numpy(sc.data.copy())
    knn_ref = O.cluster_knn(data[:, :3], L.points, 20, tie="canonical")
    x9_ref = O.centralized(data, L.points)
    order = np.concatenate(L.points).astype(np.int32)
    off = np.concatenate([[0], np.cumsum([len(p) for p in L.points])]).astype(np.int32)
    x9 = ops.centralize(dev(sc.data), dev(order), dev(off))
    assert rel_err(x9, x9_ref) < 1e-5
    p = O.init_params(1)
    torch.manual_seed(7)
    pre = "mlp_3" if two else "mlp_2"
    for k in list(p):
        if k.startswith(pre) and (k.endswith("bn1.weight") or k.endswith("bn2.weight")):
            p[k] = 0.5 + torch.rand(64)
        if k.startswith(pre) and (k.endswith("bn1.bias") or k.endswith("bn2.bias")):
            p[k] = 0.2 * torch.randn(64)
    for k in O.TRAINABLE:
        p[k].requires_grad_(True)
    out_ref = (O.mlp3_forward if two else O.mlp2_forward)(p, x9_ref, knn_ref)
    c = lambda k: p[k].detach().cuda()
    knn_d = dev(knn_ref.numpy(), torch.int32)
    x9_d = dev(x9_ref.numpy())
    if two:
        o = ops.edgeconv_fwd(x9_d, knn_d, c(pre + ".conv1.0.weight"), c(pre + ".bn1.weight"), c(pre + ".bn1.bias"),
                             c(pre + ".conv2.0.weight"), c(pre + ".bn2.weight"), c(pre + ".bn2.bias"))
    else:
        o = ops.edgeconv_fwd(x9_d, knn_d, c(pre + ".conv1.0.weight"), c(pre + ".bn1.weight"), c(pre + ".bn1.bias"))
    assert rel_err(o["out"], out_ref.detach()) < RTOL
    # pooled + backward
    pooled, arg = ops.segment_pool_max(o["out"], dev(off), dev(order))
    pooled_ref = O.segment_max(out_ref, L.points)
    assert rel_err(pooled, pooled_ref.detach()) < RTOL
    g = torch.randn(pooled_ref.shape, generator=torch.Generator().manual_seed(1))
    (pooled_ref * g).sum().backward()
    if two:
        r = ops.edgeconv_bwd(g.cuda(), arg, o["argk"], x9_d, knn_d, c(pre + ".conv1.0.weight"), o["stats1"], o["mom1"], o["ctr"],
                             c(pre + ".conv2.0.weight"), o["stats2"], o["mom2"])
    else:
        r = ops.edgeconv_bwd(g.cuda(), arg, o["argk"], x9_d, knn_d, c(pre + ".conv1.0.weight"), o["stats1"], o["mom1"], o["ctr"])
    errs = {"gW1": rel_err(r["gW1"].view(-1), p[pre + ".conv1.0.weight"].grad.view(-1)),
            "gg1": rel_err(r["gg1"], p[pre + ".bn1.weight"].grad), "gb1": rel_err(r["gb1"], p[pre + ".bn1.bias"].grad)}
    if two:
        errs.update(gW2=rel_err(r["gW2"].view(-1), p[pre + ".conv2.0.weight"].grad.view(-1)),
                    gg2=rel_err(r["gg2"], p[pre + ".bn2.weight"].grad), gb2=rel_err(r["gb2"], p[pre + ".bn2.bias"].grad))
    assert max(errs.values()) < 5e-4, errs     # gradients: sums of ~1e5 fp32 terms on both sides


# ------------------------------------------------------------------------------------------------
def test_scene_init_level_build_update_adj(scene20k):
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    sc = scene20k
    seg_off, seg_mem = dev(sc.seg_offsets, torch.int32), dev(sc.seg_members, torch.int32)
    sop, sos, uf_d = ops.scene_init(seg_off, seg_mem, dev(sc.weak_label, torch.int32))
    uf = merged_uf(sc, 0)
    assert np.array_equal(uf_d.cpu().numpy(), uf_to_device(uf).cpu().numpy())
    seg_of_point = np.empty(sc.n_points, np.int64)
    for s in range(uf.S):
        seg_of_point[sc.seg_members[sc.seg_offsets[s]:sc.seg_offsets[s + 1]]] = s
    assert np.array_equal(sop.cpu().numpy(), seg_of_point)
    adj1 = ops.update_adj(dev(sc.adj, torch.int32), sop, uf.S).cpu().numpy()
    assert np.array_equal(adj1, O.update_adj(sc.adj, seg_of_point))
    # a merged state: build the level on the device from the oracle's union-find
    L_old = O.Level(uf)
    uf2 = merged_uf(sc, 200, 4)
    L_ref = O.Level(uf2)
    L = ops.level_build(uf_to_device(uf2), seg_off, seg_mem, sos)
    assert L.S == L_ref.S
    assert np.array_equal(L.roots.cpu().numpy(), L_ref.roots)
    assert np.array_equal(L.seg2cl.cpu().numpy(), L_ref.seg2cluster)
    assert np.array_equal(L.order.cpu().numpy(), np.concatenate(L_ref.points))
    assert np.array_equal(L.cl_pt_off.cpu().numpy(), np.concatenate([[0], np.cumsum([len(p) for p in L_ref.points])]))
    assert np.array_equal(L.cl_seg_list.cpu().numpy(), np.concatenate(L_ref.seg_lists))
    assert np.array_equal(L.cl_ins.cpu().numpy(), L_ref.ins) and np.array_equal(L.cl_sem.cpu().numpy(), L_ref.sem)
    assert np.array_equal(L.cl_rootpt.cpu().numpy(), L_ref.root_point)
    L0 = ops.level_build(uf_d, seg_off, seg_mem, sos)
    o2n, ch_off, ch_list = ops.level_children(L0, L)
    groups = O.children_groups(L_ref, L_old)
    assert np.array_equal(ch_list.cpu().numpy(), np.concatenate(groups))
    assert np.array_equal(ch_off.cpu().numpy(), np.concatenate([[0], np.cumsum([len(g) for g in groups])]))
    adj2 = ops.update_adj(dev(adj1, torch.int32), o2n, L.S).cpu().numpy()
    assert np.array_equal(adj2, O.update_adj(adj1, L_ref.seg2cluster[L_old.roots]))
    # label export
    seg, ins, sem = ops.export_labels(dev(sc.unmap), sop, L)
    rs, ri, rm = L_ref.point_labels(sc.n_points, sc.unmap)
    assert np.array_equal(seg.cpu().numpy(), rs) and np.array_equal(ins.cpu().numpy(), ri) and np.array_equal(sem.cpu().numpy(), rm)


def test_group_nearby_and_unlabeled(scene20k):
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    sc = scene20k
    seg_off, seg_mem = dev(sc.seg_offsets, torch.int32), dev(sc.seg_members, torch.int32)
    sop, sos, uf_d = ops.scene_init(seg_off, seg_mem, dev(sc.weak_label, torch.int32))
    uf = merged_uf(sc, 0)
    L1 = O.Level(uf)
    seg_of_point = sop.cpu().numpy().astype(np.int64)
    adj1 = O.update_adj(sc.adj, seg_of_point)
    rng = np.random.default_rng(0)
    dist = rng.uniform(0, 10, len(adj1)).astype(np.float32)
    O.group_nearby(uf, dist, adj1, L1.roots, 3.0)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.group_nearby(dev(adj1, torch.int32), dev(L1.roots, torch.int32), dev(dist), 3.0, uf_d, status)
    assert int(status.item()) == 0
    L2_ref = O.Level(uf)
    L2 = ops.level_build(uf_d, seg_off, seg_mem, sos)
    assert L2.S == L2_ref.S and L2.S < L1.S
    assert np.array_equal(L2.order.cpu().numpy(), np.concatenate(L2_ref.points))
    assert np.array_equal(L2.cl_ins.cpu().numpy(), L2_ref.ins)
    # phase-A step of group_unlabeled on the new level
    adj2 = O.update_adj(adj1, L2_ref.seg2cluster[L1.roots])
    d2 = rng.uniform(0, 5, len(adj2)).astype(np.float32)
    S = L2_ref.S
    dm = torch.ones(S, S) * 1000
    a = torch.as_tensor(adj2)
    dm[a[:, 0], a[:, 1]] = torch.as_tensor(d2); dm[a[:, 1], a[:, 0]] = torch.as_tensor(d2)
    amin_ref = dm.min(-1)[1].numpy()
    for i in range(S):
        c1 = uf.cid[L2_ref.roots[i]]
        if uf.ins[c1] != -1:
            continue
        uf.union(c1, uf.cid[L2_ref.roots[amin_ref[i]]])
    csr = ops.sym_csr(dev(adj2, torch.int32), S)
    amin = ops.group_unlabeled_step(dev(d2), csr, S, L2.roots, uf_d)
    assert np.array_equal(amin.cpu().numpy(), amin_ref)
    L3_ref = O.Level(uf)
    L3 = ops.level_build(uf_d, seg_off, seg_mem, sos)
    assert L3.S == L3_ref.S
    assert np.array_equal(L3.order.cpu().numpy(), np.concatenate(L3_ref.points))


@pytest.mark.parametrize("C", [192, 256])
def test_edge_dist_and_gcn(C):
    from oracle import seggroup_oracle as O
    from seggroup_b200 import ops
    from seggroup_b200.pipeline import EdgeDistFn, GcnAggFn
    rng = np.random.default_rng(C)
    S = 300
    e = rng.integers(0, S, (1500, 2))
    e = e[e[:, 0] != e[:, 1]]
    adj = np.unique(np.sort(e, 1), axis=0)
    X = torch.randn(S, C, generator=torch.Generator().manual_seed(C))
    W = torch.randn(C, C, generator=torch.Generator().manual_seed(C + 1)) * 0.1
    Xr = X.clone().requires_grad_(True); Wr = W.clone().requires_grad_(True)
    sims_r = torch.exp(-O.edge_distance(Xr, adj) * (1 / 8))
    out_r = O.gcn_forward(Wr, Xr, adj, sims_r)
    g = torch.randn(out_r.shape, generator=torch.Generator().manual_seed(3))
    (out_r * g).sum().backward()
    Xd = X.cuda().requires_grad_(True); Wd = W.cuda().requires_grad_(True)
    adj_d = dev(adj, torch.int32)
    csr = ops.sym_csr(adj_d, S)
    d = EdgeDistFn.apply(Xd, adj_d, *csr)
    assert rel_err(d.detach(), O.edge_distance(X, adj)) < 1e-5
    sims = torch.exp(-d * (1 / 8))
    out = torch.relu(torch.nn.functional.linear(GcnAggFn.apply(Xd, sims, adj_d, *csr), Wd))
    assert rel_err(out.detach(), out_r.detach()) < RTOL
    (out * g.cuda()).sum().backward()
    assert rel_err(Xd.grad, Xr.grad) < RTOL
    assert rel_err(Wd.grad, Wr.grad) < RTOL


@pytest.mark.parametrize("S,E", [(1, 0), (37, 60), (300, 1500), (1117, 3400), (2999, 20000)])
def test_sym_csr_exact(S, E):
    """Rows list (neighbour, edge id) by ascending neighbour; hub rows (one cluster adjacent to most others) included."""
    from seggroup_b200 import ops
    rng = np.random.default_rng(S)
    e = rng.integers(0, S, (E, 2))
    if S > 2:
        hub = np.stack([np.full(S // 2, S // 3), rng.permutation(S)[:S // 2]], 1)     # a hub in the middle of the id range
        e = np.concatenate([e, hub])
    e = e[e[:, 0] != e[:, 1]] if len(e) else e.reshape(0, 2)
    adj = np.unique(np.sort(e, 1), axis=0).astype(np.int32).reshape(-1, 2)
    row_off, nbr, eid = ops.sym_csr(dev(adj, torch.int32), S)
    A = len(adj)
    src = np.concatenate([adj[:, 0], adj[:, 1]]); dst = np.concatenate([adj[:, 1], adj[:, 0]]); ids = np.concatenate([np.arange(A), np.arange(A)])
    order = np.lexsort((dst, src))
    ref_off = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=S))])
    assert np.array_equal(row_off.cpu().numpy(), ref_off)
    assert np.array_equal(nbr.cpu().numpy()[:2 * A], dst[order])
    assert np.array_equal(eid.cpu().numpy()[:2 * A], ids[order])


@pytest.mark.parametrize("case", ["scene_like", "many_ids", "nothing_valid"])
def test_evaluate_matches_oracle(case):
    """sgb_evaluate (a17) against the restatement of model.py:608-655 on random labels: counts exact, accuracies to fp32."""
    from oracle import seggroup_oracle as O
    from seggroup_b200 import pipeline
    rng = np.random.default_rng(len(case))
    n = 60000
    n_ids = 40 if case == "scene_like" else 3000                       # > 1024 ids: the global-memory histogram path
    sem_true = rng.integers(0, 41, n)
    ins_true = rng.integers(1, n_ids + 1, n)
    if case == "nothing_valid":
        sem_true[:] = 0
    sem_of_id = rng.integers(1, 41, n_ids + 1)
    ins_pred = np.where(rng.random(n) < 0.7, ins_true, rng.integers(-1, n_ids + 1, n))
    sem_pred = np.where(ins_pred >= 0, sem_of_id[np.maximum(ins_pred, 0)], -1)
    sem_pred = np.where(rng.random(n) < 0.05, rng.integers(1, 41, n), sem_pred)     # instances whose first vertex disagrees
    real = np.stack([sem_true, ins_true], 1).astype(np.int64)
    ref = O.evaluate(real, sem_pred.astype(np.int64), ins_pred.astype(np.int64))
    out = pipeline.evaluate(dev(real), dev(sem_pred, torch.int32), dev(ins_pred, torch.int32))
    for a, b in zip(out, ref):
        a = a.cpu().numpy().reshape(b.shape)
        assert np.allclose(a, b, rtol=1e-6, atol=0, equal_nan=True), (case, a, b)


@pytest.mark.parametrize("sizes,n_labels", [([37], 9), ([40, 3, 1500, 212], 30), ([5, 5], 1), ([2500, 1200, 7], 400)])
def test_classifier_groups_match_numpy_unique(sizes, n_labels):
    """sgb_classifier_groups (a15 grouping) against the reference formulation model.py:902-916 per scene: np.unique of the
    clusters' weak instance labels, np.where order inside a group, semantic label of the group's first cluster.  Bit-exact."""
    from seggroup_b200 import ops
    rng = np.random.default_rng(sum(sizes) + n_labels)
    ins = np.concatenate([rng.integers(-1, n_labels, n) for n in sizes]).astype(np.int32)
    sem = rng.integers(-1, 40, ins.size).astype(np.int32)
    cl_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    B = len(sizes)
    order, off, gold, g_off, G, gmin = ops.classifier_groups(dev(ins), dev(sem), dev(cl_off) if B > 1 else None, B)
    r_order, r_off, r_gold, r_goff = [], [0], [], [0]
    for b in range(B):
        lo, hi = cl_off[b], cl_off[b + 1]
        ins_list = ins[lo:hi]
        for v in np.unique(ins_list):
            idx = np.where(ins_list == v)[0]
            r_order += list(lo + idx)
            r_off.append(len(r_order))
            r_gold.append(sem[lo + idx[0]])
        r_goff.append(len(r_gold))
    assert G == len(r_gold) and gmin == min(b - a for a, b in zip(r_goff[:-1], r_goff[1:]))
    assert np.array_equal(order.cpu().numpy(), np.asarray(r_order))
    assert np.array_equal(off.cpu().numpy(), np.asarray(r_off))
    assert np.array_equal(gold.cpu().numpy(), np.asarray(r_gold))
    assert np.array_equal(g_off.cpu().numpy(), np.asarray(r_goff))


def test_evaluate_scenes_equals_per_scene_calls():
    """sgb_evaluate_scenes (one library call per batch) returns, per scene, exactly what sgb_evaluate returns on the scene's slice."""
    from seggroup_b200 import ops, pipeline
    rng = np.random.default_rng(5)
    sizes = [30000, 1, 45001, 777]
    off = [0] + list(np.cumsum(sizes))
    n = off[-1]
    real = np.stack([rng.integers(0, 41, n), rng.integers(1, 60, n)], 1).astype(np.int64)
    ins_pred = rng.integers(-1, 60, n).astype(np.int32)
    sem_pred = np.where(ins_pred >= 0, 1 + ins_pred % 40, -1).astype(np.int32)
    d_real, d_sem, d_ins = dev(real), dev(sem_pred), dev(ins_pred)
    out = ops.evaluate_scenes(d_real, d_sem, d_ins, off, pipeline.SEM_VALID, pipeline.INS_VALID)
    for b, (lo, hi) in enumerate(zip(off[:-1], off[1:])):
        one = ops.evaluate(d_real[lo:hi].contiguous(), d_sem[lo:hi].contiguous(), d_ins[lo:hi].contiguous(), pipeline.SEM_VALID, pipeline.INS_VALID)
        assert np.array_equal(out[b].cpu().numpy(), one.cpu().numpy(), equal_nan=True), b


@pytest.mark.parametrize("counts,drop", [([41, 17, 2, 33], True), ([9], False), ([64, 64], True)])
def test_classifier_head_forward_backward(counts, drop):
    """a15: the fused classifier head + label-smoothed cross entropy (model.py:154-166, 902-932, util.py:12-29), one CTA per
    scene, against torch fp64 evaluated per scene: loss (sum, count), logits, BatchNorm batch statistics and every gradient."""
    import torch.nn.functional as F
    from seggroup_b200.pipeline import ClassifierHeadFn
    g = torch.Generator().manual_seed(sum(counts))
    G = sum(counts)
    feat = torch.randn(G, 256, generator=g)
    W1 = torch.randn(128, 256, generator=g) / 16
    gamma, beta = 0.5 + torch.rand(128, generator=g), 0.2 * torch.randn(128, generator=g)
    W2, b2 = torch.randn(40, 128, generator=g) / 11, 0.1 * torch.randn(40, generator=g)
    gold = torch.randint(0, 40, (G,), generator=g)
    mask = (torch.rand(G, 128, generator=g) > 0.5).float() if drop else None
    w = torch.randn(len(counts), generator=g)                         # upstream gradient of every scene's loss sum
    off = [0] + list(np.cumsum(counts))
    dv = [t.cuda().requires_grad_(True) for t in (feat, W1, gamma, beta, W2, b2)]
    loss_raw, logits, stats = ClassifierHeadFn.apply(*dv, torch.tensor(off, dtype=torch.int32, device="cuda"), gold.int().cuda(),
                                                     mask.cuda() if drop else None, 2.0 if drop else 1.0)
    (loss_raw[:, 0] * w.cuda()).sum().backward()
    rv = [t.double().requires_grad_(True) for t in (feat, W1, gamma, beta, W2, b2)]
    total = 0
    for b, (lo, hi) in enumerate(zip(off[:-1], off[1:])):
        h = F.linear(rv[0][lo:hi], rv[1])
        assert torch.allclose(stats[b, :128].cpu().double(), h.mean(0).detach(), atol=1e-5)
        assert torch.allclose(stats[b, 128:].cpu().double(), h.var(0, unbiased=False).detach(), rtol=1e-4, atol=1e-6)
        h = F.leaky_relu(F.batch_norm(h, None, None, rv[2], rv[3], True, 0.1, 1e-5), 0.2)
        if drop:
            h = h * mask[lo:hi].double() * 2.0
        lg = F.linear(h, rv[4], rv[5])
        one_hot = torch.zeros_like(lg).scatter(1, gold[lo:hi].view(-1, 1), 1)
        one_hot = one_hot * 0.8 + (1 - one_hot) * 0.2 / 39
        ls = -(one_hot * F.log_softmax(lg, dim=1)).sum()
        assert abs(float(loss_raw[b, 0]) - float(ls)) < 1e-5 * abs(float(ls)) and int(loss_raw[b, 1]) == hi - lo
        assert float((logits[lo:hi].cpu().double() - lg.detach()).abs().max()) < 1e-4
        total = total + ls * w[b].double()
    total.backward()
    for a, r, name in zip(dv, rv, ("feat", "W1", "gamma", "beta", "W2", "b2")):
        err = float((a.grad.cpu().double() - r.grad).abs().max() / (r.grad.abs().max() + 1e-30))
        assert err < 2e-5, (name, err)


@pytest.mark.parametrize("C", [64, 192, 256])
def test_aggregate_cluster_feature_use_avg(C):
    """use_avg variant of aggregate_cluster_feature (model.py:278-288): (max, mean) per new cluster, forward and backward against
    the reference's own loop over clusters evaluated with torch on the same device."""
    from seggroup_b200 import pipeline
    g = torch.Generator().manual_seed(3)
    R, S = 700, 90
    feat = torch.randn(R, C, generator=g).cuda().requires_grad_(True)
    perm = torch.randperm(R, generator=g)
    cuts = torch.sort(torch.randperm(R - 1, generator=g)[:S - 1] + 1).values
    offsets = torch.cat([torch.zeros(1, dtype=torch.long), cuts, torch.tensor([R])]).to(torch.int32).cuda()
    members = perm.to(torch.int32).cuda()
    out = pipeline.aggregate_cluster_feature(feat, offsets, members, use_avg=True)
    assert out.shape == (S, 2 * C)
    w = torch.randn(S, 2 * C, generator=g).cuda()
    (out * w).sum().backward()
    got_grad = feat.grad.clone()
    ref_feat = feat.detach().clone().requires_grad_(True)
    rows = []
    off = offsets.cpu().tolist()
    for s in range(S):                                   # model.py:279-285
        idx = members[off[s]:off[s + 1]].long()
        f1 = torch.max(ref_feat[idx], dim=0, keepdim=True)[0]
        f2 = torch.mean(ref_feat[idx], dim=0, keepdim=True)[0].unsqueeze(0)
        rows.append(torch.cat([f1, f2], dim=-1))
    ref = torch.cat(rows, dim=0)
    (ref * w).sum().backward()
    assert torch.equal(out[:, :C], ref[:, :C])
    assert torch.allclose(out[:, C:], ref[:, C:], rtol=1e-5, atol=1e-6)
    assert torch.allclose(got_grad, ref_feat.grad, rtol=1e-5, atol=1e-6)
    assert torch.equal(pipeline.aggregate_cluster_feature(feat.detach(), offsets, members), ref[:, :C].detach())
