"""Tensor-level wrappers over the C-ABI (`include/seggroup_b200.h`).

PyTorch is used here for device memory and streams only: every function allocates its outputs and
workspaces as CUDA tensors, hands raw pointers + the current stream to libseggroup_b200.so, and returns
the tensors.  No function has a CPU or eager-PyTorch fallback: inputs must be CUDA tensors.
"""
from __future__ import annotations

import torch

from . import _lib

I32 = torch.int32
F32 = torch.float32


def _stream():
    # raw cudaStream_t of the calling thread's current stream (torch.cuda.current_stream() costs ~15 us per call)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise _lib.SgbError("%s must be a CUDA tensor (seggroup_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
def exclusive_scan(x):
    _chk(x, I32, "x")
    n = x.numel()
    out = torch.empty(n + 1, dtype=I32, device=x.device)
    ws = _ws(_lib.call("sgb_scan_ws_bytes", n), x.device)
    _lib.call("sgb_exclusive_scan_i32", x, out, n, ws, ws.numel(), _stream())
    return out


def segment_pool_max(feat, offsets, members=None, want_argmax=True):
    """feat [R,C] f32; offsets [S+1] i32; members [M] i32 row ids (None = rows already segment-major).
    Returns (out [S,C], argmax [S,C] i32 row ids)."""
    _chk(feat, F32, "feat"); _chk(offsets, I32, "offsets")
    R, C = feat.shape
    S = offsets.numel() - 1
    M = R if members is None else members.numel()
    if members is not None:
        _chk(members, I32, "members")
    out = torch.empty(S, C, dtype=F32, device=feat.device)
    arg = torch.empty(S, C, dtype=I32, device=feat.device) if want_argmax else None
    ws = _ws(_lib.call("sgb_segment_pool_ws_bytes", S, C), feat.device)
    _lib.call("sgb_segment_pool_max_fwd", feat, R, C, members, M, offsets, S, out, arg, ws, ws.numel(), _stream())
    return out, arg


def segment_pool_max_bwd(grad_out, argmax, n_rows):
    _chk(grad_out, F32, "grad_out"); _chk(argmax, I32, "argmax")
    S, C = grad_out.shape
    g = torch.zeros(n_rows, C, dtype=F32, device=grad_out.device)
    _lib.call("sgb_segment_pool_max_bwd", grad_out, argmax, S, C, g, _stream())
    return g


def segment_pool_mean(feat, offsets, members=None):
    """use_avg variant of aggregate_cluster_feature (model.py:282-284): feat [R,C] f32 -> [S,C] row means of the segments."""
    _chk(feat, F32, "feat"); _chk(offsets, I32, "offsets")
    R, C = feat.shape
    S = offsets.numel() - 1
    M = R if members is None else _chk(members, I32, "members").numel()
    out = torch.empty(S, C, dtype=F32, device=feat.device)
    _lib.call("sgb_segment_pool_mean_fwd", feat, R, C, members, M, offsets, S, out, _stream())
    return out


def segment_pool_mean_bwd(grad_out, offsets, members, n_rows):
    _chk(grad_out, F32, "grad_out"); _chk(offsets, I32, "offsets")
    S, C = grad_out.shape
    g = torch.zeros(n_rows, C, dtype=F32, device=grad_out.device)
    M = n_rows if members is None else members.numel()
    _lib.call("sgb_segment_pool_mean_bwd", grad_out, S, C, members, M, offsets, g, _stream())
    return g


def cluster_knn(xyz, order, cl_off, k=20, scene_pt_off=None):
    """xyz [N,>=3] f32 (row stride = xyz.stride(0)); order [N] i32; cl_off [S+1] i32 -> knn [N,k] i32.
    scene_pt_off [B+1] i32 (device, scene batch): ids are written relative to the first point of the query's scene."""
    _chk(order, I32, "order"); _chk(cl_off, I32, "cl_off")
    if not xyz.is_cuda or xyz.dtype != F32 or xyz.stride(1) != 1:
        raise ValueError("xyz must be a CUDA float32 tensor with unit inner stride")
    N = order.numel()
    knn = torch.empty(N, k, dtype=I32, device=xyz.device)
    S = cl_off.numel() - 1
    ws = _ws(_lib.call("sgb_cluster_knn_ws_bytes", N, S), xyz.device)
    if scene_pt_off is None:
        _lib.call("sgb_cluster_knn", xyz, xyz.stride(0), N, order, cl_off, S, k, knn, ws, ws.numel(), _stream())
    else:
        _lib.call("sgb_cluster_knn_scenes", xyz, xyz.stride(0), N, order, cl_off, S, k, knn, _chk(scene_pt_off, I32, "scene_pt_off"),
                  scene_pt_off.numel() - 1, ws, ws.numel(), _stream())
    return knn


def cluster_cloud_indices(xyz, order, cl_off, P, status=None):
    """-> (cloud_idx [S,P] i32 point ids, status [1] i32; pass `status` to OR the flag into an existing status word)."""
    _chk(order, I32, "order"); _chk(cl_off, I32, "cl_off")
    if not xyz.is_cuda or xyz.dtype != F32 or xyz.stride(1) != 1:
        raise ValueError("xyz must be a CUDA float32 tensor with unit inner stride")
    S = cl_off.numel() - 1
    N = order.numel()
    idx = torch.empty(S, P, dtype=I32, device=xyz.device)
    if status is None:
        status = torch.zeros(1, dtype=I32, device=xyz.device)
    ws = _ws(_lib.call("sgb_cluster_cloud_ws_bytes", N), xyz.device)
    _lib.call("sgb_cluster_cloud_indices", xyz, xyz.stride(0), N, order, cl_off, S, P, idx, status, ws, ws.numel(), _stream())
    return idx, status


def cluster_cloud_transform(data6, cloud_idx):
    _chk(data6, F32, "data6"); _chk(cloud_idx, I32, "cloud_idx")
    S, P = cloud_idx.shape
    clouds = torch.empty(S, P, 6, dtype=F32, device=data6.device)
    _lib.call("sgb_cluster_cloud_transform", data6, cloud_idx, S, P, clouds, _stream())
    return clouds


def centralize(data6, order, cl_off):
    _chk(data6, F32, "data6"); _chk(order, I32, "order"); _chk(cl_off, I32, "cl_off")
    N = data6.shape[0]
    S = cl_off.numel() - 1
    x9 = torch.empty(N, 9, dtype=F32, device=data6.device)
    nb = _lib.call("sgb_centralize_ws_bytes", N, S)                        # fixed-point coordinate sums + cluster id per point
    _lib.call("sgb_centralize", data6, N, order, cl_off, S, x9, _ws(nb, data6.device), nb, _stream())
    return x9


def mlp1_fwd(clouds, W, gamma, beta, feat=None, knn=None, arg_pt=None):
    """clouds [S,64,6] -> dict(feat [S,128], knn [S,64,10], arg_pt [S,64], stats [4,64], var [64], mom [27] f64).
    feat / knn / arg_pt: optional preallocated outputs (row slices of a scene batch's arrays)."""
    _chk(clouds, F32, "clouds")
    S = clouds.shape[0]
    dev = clouds.device
    W = _chk(W.reshape(64, 6), F32, "W")
    out = dict(feat=torch.empty(S, 128, dtype=F32, device=dev) if feat is None else _chk(feat, F32, "feat"),
               knn=torch.empty(S, 64, 10, dtype=I32, device=dev) if knn is None else _chk(knn, I32, "knn"),
               arg_pt=torch.empty(S, 64, dtype=I32, device=dev) if arg_pt is None else _chk(arg_pt, I32, "arg_pt"),
               stats=torch.empty(4, 64, dtype=F32, device=dev),
               var=torch.empty(64, dtype=F32, device=dev), mom=torch.empty(27, dtype=torch.float64, device=dev))
    ws = _ws(_lib.call("sgb_mlp1_ws_bytes", S), dev)
    _lib.call("sgb_mlp1_fwd", clouds, S, W, gamma, beta, out["feat"], out["knn"], out["arg_pt"], out["stats"], out["var"],
              out["mom"], ws, ws.numel(), _stream())
    return out


def mlp1_bwd(g, clouds, knn, arg_pt, W, stats, mom):
    """-> (gW [64,6], ggamma [64], gbeta [64])"""
    S = clouds.shape[0]
    dev = clouds.device
    W = _chk(W.reshape(64, 6).contiguous(), F32, "W")
    gW = torch.empty(64, 6, dtype=F32, device=dev)
    gg = torch.empty(64, dtype=F32, device=dev)
    gb = torch.empty(64, dtype=F32, device=dev)
    ws = _ws(_lib.call("sgb_mlp1_bwd_ws_bytes", S), dev)
    _lib.call("sgb_mlp1_bwd", _chk(g, F32, "g"), clouds, knn, arg_pt, S, W, stats, mom, gW, gg, gb, ws, ws.numel(), _stream())
    return gW, gg, gb


def edgeconv_bwd(g, arg, argk, x9, knn, W1, stats1, mom1, e0, W2=None, stats2=None, mom2=None):
    """g [S,64] -> dict(gW1, gg1, gb1 [, gW2, gg2, gb2])."""
    S = g.shape[0]
    N = x9.shape[0]
    dev = g.device
    two = W2 is not None
    W1 = _chk(W1.reshape(64, 18).contiguous(), F32, "W1")
    r = dict(gW1=torch.empty(64, 18, dtype=F32, device=dev), gg1=torch.empty(64, dtype=F32, device=dev),
             gb1=torch.empty(64, dtype=F32, device=dev))
    if two:
        W2 = _chk(W2.reshape(64, 64).contiguous(), F32, "W2")
        r.update(gW2=torch.empty(64, 64, dtype=F32, device=dev), gg2=torch.empty(64, dtype=F32, device=dev),
                 gb2=torch.empty(64, dtype=F32, device=dev))
    ws = _ws(_lib.call("sgb_edgeconv_bwd_ws_bytes", N, S, int(two)), dev)
    _lib.call("sgb_edgeconv_bwd", _chk(g, F32, "g"), arg, argk, S, x9, knn, N, int(two), W1, stats1, mom1, e0, W2, stats2, mom2,
              r["gW1"], r["gg1"], r["gb1"], r.get("gW2"), r.get("gg2"), r.get("gb2"), ws, ws.numel(), _stream())
    return r


def edgeconv_fwd(x9, knn, W1, gamma1, beta1, W2=None, gamma2=None, beta2=None, want_argk=True, want_backward=True, out=None, argk=None):
    """x9 [N,9], knn [N,20] -> dict(out [N,64], argk [N,64] u8, stats1, var1, mom1, ctr [, stats2, var2, mom2]).
    want_backward=False (inference): no moments are kept and the second layer of MLP3 runs on the tcgen05 tensor cores.
    out / argk: optional preallocated outputs (row slices of a scene batch's arrays)."""
    _chk(x9, F32, "x9"); _chk(knn, I32, "knn")
    N = x9.shape[0]
    dev = x9.device
    two = W2 is not None
    W1 = _chk(W1.reshape(64, 18), F32, "W1")
    if two:
        W2 = _chk(W2.reshape(64, 64), F32, "W2")
    o = dict(out=torch.empty(N, 64, dtype=F32, device=dev) if out is None else _chk(out, F32, "out"), stats1=torch.empty(4, 64, dtype=F32, device=dev),
             var1=torch.empty(64, dtype=F32, device=dev),
             mom1=torch.empty(189, dtype=torch.float64, device=dev) if want_backward else None,
             ctr=torch.empty(18, dtype=F32, device=dev),
             argk=(torch.empty(N, 64, dtype=torch.uint8, device=dev) if argk is None else _chk(argk, torch.uint8, "argk"))
             if (want_argk or want_backward) else None)
    if two:
        o.update(stats2=torch.empty(4, 64, dtype=F32, device=dev), var2=torch.empty(64, dtype=F32, device=dev),
                 mom2=torch.empty(4160, dtype=torch.float64, device=dev) if want_backward else None)
    ws = _ws(_lib.call("sgb_edgeconv_ws_bytes", N, int(two)), dev)
    _lib.call("sgb_edgeconv_fwd", x9, knn, N, int(two), W1, gamma1, beta1, W2, gamma2, beta2, o["out"], o["argk"], o["stats1"], o["var1"],
              o["mom1"], o.get("stats2"), o.get("var2"), o.get("mom2"), o["ctr"], ws, ws.numel(), _stream())
    return o


# ------------------------------------------------------------------------------------------------
# segment graph
# ------------------------------------------------------------------------------------------------
def scene_init(seg_off, seg_members, weak_label):
    """-> (seg_of_point [N], seg_of_pos [N], uf [6,S1])."""
    _chk(seg_off, I32, "seg_off"); _chk(seg_members, I32, "seg_members"); _chk(weak_label, I32, "weak_label")
    N = seg_members.numel()
    S = seg_off.numel() - 1
    dev = seg_off.device
    sop = torch.empty(N, dtype=I32, device=dev)
    sos = torch.empty(N, dtype=I32, device=dev)
    uf = torch.empty(6, S, dtype=I32, device=dev)
    _lib.call("sgb_scene_init", seg_off, seg_members, weak_label, N, S, sop, sos, uf, _stream())
    return sop, sos, uf


class Level:
    """Device arrays of one clustering level (all sized for S1; `S` is read back once, after build).  Levels made by
    `level_step` also carry their adjacency (`adj` [A,2], `csr`), the map from the previous level (`o2n`, `ch_off`,
    `ch_list`), the number of unlabeled clusters and the status word read back with the counts."""
    __slots__ = ("S", "roots", "seg2cl", "cl_seg_off", "cl_seg_list", "cl_pt_off", "order", "cl_ins", "cl_sem", "cl_rootpt",
                 "counts", "adj", "csr", "o2n", "ch_off", "ch_list", "n_unlabeled", "status",
                 "scene_cl_off", "d_scene_cl_off")       # scene batch: cluster range of every scene (host list / device [B+1])


class SceneSplit:
    """Scene partition of a batch (several scenes concatenated block-diagonally): host offsets + their device copies."""
    __slots__ = ("n", "pt_off", "seg_off", "raw_off", "d_pt_off", "d_seg_off", "max_segs")

    def __init__(self, pt_off, seg_off, raw_off, device):
        self.n = len(pt_off) - 1
        self.pt_off, self.seg_off, self.raw_off = [int(v) for v in pt_off], [int(v) for v in seg_off], [int(v) for v in raw_off]
        self.d_pt_off = torch.tensor(self.pt_off, dtype=I32, device=device)
        self.d_seg_off = torch.tensor(self.seg_off, dtype=I32, device=device)
        self.max_segs = max(b - a for a, b in zip(self.seg_off[:-1], self.seg_off[1:]))


def level_step(mode, uf, seg_off, seg_members, seg_of_pos, status, old=None, dist=None, th=0.0, edges=None, mapping=None,
               sweep_cap=64, split=None):
    """One clustering level in one library call (sgb_level_step): mode 0 group_nearby(old.adj, dist, th), 1 one phase-A
    iteration of group_unlabeled (dist, old.csr), 2 no grouping.  edges/mapping: edge list to re-map (default: old.adj
    through the old->new cluster map of this step).  Returns the new Level with .adj, .csr, .o2n, .ch_off, .ch_list."""
    import numpy as np
    S1 = uf.shape[1]
    N = seg_members.numel()
    dev = uf.device
    S_old = old.S if old is not None else 0
    if edges is None:
        edges = old.adj
    E = edges.shape[0]
    B = split.n if split is not None and split.n > 1 else 1
    # counts (4 ints) directly followed by the scenes' new cluster offsets: the library reads both back with ONE copy
    sizes = [S1, S1, S1 + 1, S1, S1 + 1, N, S1, S1, S1, max(S_old, 1), S1 + 1, max(S_old, 1), 2 * max(E, 1), S1 + 1,
             2 * max(E, 1), 2 * max(E, 1), 4, B + 1]
    buf = torch.empty(sum(sizes), dtype=I32, device=dev)
    parts, o = [], 0
    for n in sizes:
        parts.append(buf[o:o + n]); o += n
    (roots, seg2cl, cl_seg_off, cl_seg_list, cl_pt_off, order, cl_ins, cl_sem, cl_rootpt, o2n, ch_off, ch_list, adj_new, csr_off,
     csr_nbr, csr_eid, counts, scl) = parts
    host = _pinned_counts(4 + B + 1)           # page-locked: the two read-backs per level are plain DMA + one stream sync each
    ws = _ws(_lib.call("sgb_level_step_ws_bytes", S1, S_old), dev)
    oc = old.csr if (old is not None and mode == 1) else (None, None, None)
    common = (int(mode), old.adj if old is not None else None, old.adj.shape[0] if old is not None else 0,
              old.roots if old is not None else None, S_old, dist, float(th), int(sweep_cap), oc[0], oc[1], oc[2],
              edges, E, mapping, uf, S1, N, seg_off, seg_members, seg_of_pos,
              roots, seg2cl, cl_seg_off, cl_seg_list, cl_pt_off, order, cl_ins, cl_sem, cl_rootpt, o2n, ch_off, ch_list,
              adj_new, csr_off, csr_nbr, csr_eid, status, counts, host)
    if B == 1:
        _lib.call("sgb_level_step", *common, ws, ws.numel(), _stream())
    else:
        max_cl_old = max(b - a for a, b in zip(old.scene_cl_off[:-1], old.scene_cl_off[1:])) if old is not None else 0
        _lib.call("sgb_level_step_scenes", *common, split.d_seg_off, old.d_scene_cl_off if old is not None else None, scl, B,
                  split.max_segs, max_cl_old, ws, ws.numel(), _stream())
    hv = host.tolist()
    S, A = hv[0], hv[1]
    L = Level()
    L.scene_cl_off = hv[4:4 + B + 1] if B > 1 else [0, S]
    L.d_scene_cl_off = scl if B > 1 else None
    L.S, L.counts = S, counts
    L.roots, L.cl_seg_off, L.cl_pt_off = roots[:S], cl_seg_off[:S + 1], cl_pt_off[:S + 1]
    L.seg2cl, L.cl_seg_list, L.order = seg2cl, cl_seg_list, order
    L.cl_ins, L.cl_sem, L.cl_rootpt = cl_ins[:S], cl_sem[:S], cl_rootpt[:S]
    L.adj = adj_new[:2 * A].view(A, 2)
    L.csr = (csr_off[:S + 1], csr_nbr[:max(2 * A, 1)], csr_eid[:max(2 * A, 1)])
    L.o2n, L.ch_off, L.ch_list = (o2n[:S_old], ch_off[:S + 1], ch_list[:S_old]) if old is not None else (None, None, None)
    L.n_unlabeled, L.status = hv[2], hv[3]
    return L


_pinned = {}


def _pinned_counts(n):
    """Per-thread page-locked int32 scratch for the level counters (level_step consumes it before it returns)."""
    import threading
    key = (threading.get_ident(), n)
    t = _pinned.get(key)
    if t is None:
        t = _pinned[key] = torch.zeros(n, dtype=I32).pin_memory()
    return t


def level_build(uf, seg_off, seg_members, seg_of_pos):
    S1 = uf.shape[1]
    N = seg_members.numel()
    dev = uf.device
    L = Level()
    e = lambda n: torch.empty(n, dtype=I32, device=dev)
    L.roots, L.seg2cl, L.cl_seg_off, L.cl_seg_list = e(S1), e(S1), e(S1 + 1), e(S1)
    L.cl_pt_off, L.order, L.cl_ins, L.cl_sem, L.cl_rootpt = e(S1 + 1), e(N), e(S1), e(S1), e(S1)
    L.counts = torch.zeros(4, dtype=I32, device=dev)
    ws = _ws(_lib.call("sgb_level_ws_bytes", S1), dev)
    _lib.call("sgb_level_build", uf, S1, N, seg_off, seg_members, seg_of_pos, L.roots, L.seg2cl, L.cl_seg_off, L.cl_seg_list,
              L.cl_pt_off, L.order, L.cl_ins, L.cl_sem, L.cl_rootpt, L.counts, ws, ws.numel(), _stream())
    L.S = int(L.counts[0].item())          # the one host sync per level (shapes of everything downstream)
    # trim views to the live cluster count
    S = L.S
    L.roots, L.cl_seg_off, L.cl_pt_off = L.roots[:S], L.cl_seg_off[:S + 1], L.cl_pt_off[:S + 1]
    L.cl_ins, L.cl_sem, L.cl_rootpt = L.cl_ins[:S], L.cl_sem[:S], L.cl_rootpt[:S]
    return L


def level_children(old: Level, new: Level):
    """-> (old2new [S_old], child_off [S_new+1], child_list [S_old])."""
    dev = old.roots.device
    o2n = torch.empty(old.S, dtype=I32, device=dev)
    off = torch.empty(new.S + 1, dtype=I32, device=dev)
    lst = torch.empty(old.S, dtype=I32, device=dev)
    ws = _ws(_lib.call("sgb_children_ws_bytes", new.S), dev)
    _lib.call("sgb_level_children", old.roots, old.S, new.seg2cl, new.S, o2n, off, lst, ws, ws.numel(), _stream())
    return o2n, off, lst


def update_adj(edges, mapping, S_new):
    """edges [E,2] i32, mapping [n_old] i32 -> adj [A,2] i32 (unique, row-sorted, lexicographic)."""
    _chk(edges, I32, "edges"); _chk(mapping, I32, "mapping")
    E = edges.shape[0]
    dev = mapping.device
    cap = max(1, min(E, S_new * (S_new - 1) // 2))
    out = torch.empty(cap, 2, dtype=I32, device=dev)
    counts = torch.zeros(4, dtype=I32, device=dev)
    ws = _ws(_lib.call("sgb_update_adj_ws_bytes", S_new), dev)
    _lib.call("sgb_update_adj", edges, E, mapping, S_new, out, counts, ws, ws.numel(), _stream())
    A = int(counts[1].item())
    return out[:A]


def sym_csr(adj, S):
    _chk(adj, I32, "adj")
    A = adj.shape[0]
    dev = adj.device
    row_off = torch.empty(S + 1, dtype=I32, device=dev)
    nbr = torch.empty(max(2 * A, 1), dtype=I32, device=dev)
    eid = torch.empty(max(2 * A, 1), dtype=I32, device=dev)
    ws = _ws(_lib.call("sgb_sym_csr_ws_bytes", S), dev)
    _lib.call("sgb_sym_csr", adj, A, S, row_off, nbr, eid, ws, ws.numel(), _stream())
    return row_off, nbr, eid


def edge_dist(feat, adj):
    _chk(feat, F32, "feat"); _chk(adj, I32, "adj")
    A = adj.shape[0]
    d = torch.empty(A, dtype=F32, device=feat.device)
    _lib.call("sgb_edge_dist_fwd", feat, feat.shape[1], adj, A, d, _stream())
    return d


def edge_dist_bwd(feat, adj, dist, gdist, csr, gfeat):
    row_off, nbr, eid = csr
    _lib.call("sgb_edge_dist_bwd", feat, feat.shape[0], feat.shape[1], adj, adj.shape[0], dist, gdist, row_off, eid, gfeat, _stream())
    return gfeat


def gcn_agg(X, sims, csr):
    row_off, nbr, eid = csr
    S, C = X.shape
    AX = torch.empty(S, C, dtype=F32, device=X.device)
    rs = torch.empty(S, dtype=F32, device=X.device)
    _lib.call("sgb_gcn_agg_fwd", X, S, C, sims, row_off, nbr, eid, AX, rs, _stream())
    return AX, rs


def gcn_agg_bwd(dAX, X, AX, sims, rs, adj, csr):
    row_off, nbr, eid = csr
    S, C = X.shape
    A = adj.shape[0]
    dX = torch.empty(S, C, dtype=F32, device=X.device)
    dsims = torch.zeros(max(A, 1), dtype=F32, device=X.device)[:A]
    _lib.call("sgb_gcn_agg_bwd", dAX, X, AX, S, C, sims, rs, adj, A, row_off, nbr, eid, dX, dsims, _stream())
    return dX, dsims


def group_nearby(adj, roots_cur, dist, th, uf, status, sweep_cap=64):
    _lib.call("sgb_group_nearby", adj, adj.shape[0], roots_cur, dist, float(th), uf, uf.shape[1], sweep_cap, status, _stream())


def group_unlabeled_step(dist, csr, S, roots_cur, uf):
    row_off, nbr, eid = csr
    amin = torch.empty(2 * S, dtype=I32, device=uf.device)      # arg-min row + scratch list
    _lib.call("sgb_group_unlabeled_step", dist, row_off, nbr, eid, S, roots_cur, uf, uf.shape[1], amin, _stream())
    return amin[:S]


def group_unlabeled_phase_b(unl, cand, roots_cur, uf):
    _chk(unl, I32, "unl"); _chk(cand, I32, "cand")
    _lib.call("sgb_group_unlabeled_phase_b", unl, unl.numel(), cand, cand.shape[1], roots_cur, uf, uf.shape[1], _stream())


def phase_b_rank(xyz, cloud_idx, unl, level: Level, width, split=None):
    """-> cand [n_unl,width] i32: per unlabeled cluster the clusters of its scene by sampled-cloud distance, -1 padded."""
    _chk(cloud_idx, I32, "cloud_idx"); _chk(unl, I32, "unl")
    cand = torch.empty(unl.numel(), width, dtype=I32, device=unl.device)
    B = split.n if split is not None else 1
    _lib.call("sgb_phase_b_rank", xyz, xyz.stride(0), cloud_idx, cloud_idx.shape[1], unl, unl.numel(),
              level.d_scene_cl_off if B > 1 else None, B, level.S, cand, width, _stream())
    return cand


def export_labels(unmap, seg_of_point, level: Level, want_seg=True, split=None):
    """unmap [N_raw] i64 (or None) -> (seg, ins, sem) int32 [N_raw].  split (scene batch): segment labels are point ids inside the scene."""
    dev = seg_of_point.device
    n_raw = unmap.numel() if unmap is not None else seg_of_point.numel()
    seg = torch.empty(n_raw, dtype=I32, device=dev) if want_seg else None
    ins = torch.empty(n_raw, dtype=I32, device=dev)
    sem = torch.empty(n_raw, dtype=I32, device=dev)
    if split is None or split.n == 1:
        _lib.call("sgb_export_labels", unmap, n_raw, seg_of_point, level.seg2cl, level.cl_rootpt, level.cl_ins, level.cl_sem,
                  seg, ins, sem, _stream())
    else:
        _lib.call("sgb_export_labels_scenes", unmap, n_raw, seg_of_point, level.seg2cl, level.cl_rootpt, level.cl_ins, level.cl_sem,
                  seg, ins, sem, split.d_pt_off, split.n, _stream())
    return seg, ins, sem


def classifier_groups(cl_ins, cl_sem, d_scene_cl_off=None, n_scenes=1):
    """Per-instance grouping of the final clusters (model.py:902-916) in one launch + one 8-byte read-back.
    -> (order [S] i32, off [G+1] i32, gold [G] i32, g_off [n_scenes+1] i32, G, min groups per scene)"""
    _chk(cl_ins, I32, "cl_ins"); _chk(cl_sem, I32, "cl_sem")
    S = cl_ins.numel()
    buf = torch.empty(3 * S + 1 + n_scenes + 1 + 2, dtype=I32, device=cl_ins.device)
    order, off, gold = buf[:S], buf[S:2 * S + 1], buf[2 * S + 1:3 * S + 1]
    g_off, counts = buf[3 * S + 1:3 * S + 2 + n_scenes], buf[3 * S + 2 + n_scenes:]
    _lib.call("sgb_classifier_groups", cl_ins, cl_sem, d_scene_cl_off if n_scenes > 1 else None, int(n_scenes), S, order, off, gold, g_off,
              counts, _stream())
    G, gmin = counts.tolist()                           # the one host sync: G sizes the classifier's tensors
    return order, off[:G + 1], gold[:G], g_off, G, gmin


def classifier_head_fwd(feat, g_off, gold, W1, gamma, beta, W2, b2, mask, drop_scale):
    """-> dict(hpre [G,128], stats [B,256], logits [G,40], loss_raw [B,2])"""
    _chk(feat, F32, "feat"); _chk(g_off, I32, "g_off"); _chk(gold, I32, "gold")
    G, B = feat.shape[0], g_off.numel() - 1
    dev = feat.device
    o = dict(hpre=torch.empty(G, 128, dtype=F32, device=dev), stats=torch.empty(B, 256, dtype=F32, device=dev),
             logits=torch.empty(G, 40, dtype=F32, device=dev), loss_raw=torch.empty(B, 2, dtype=F32, device=dev))
    _lib.call("sgb_classifier_head_fwd", feat, G, g_off, B, gold, _chk(W1, F32, "W1"), _chk(gamma, F32, "gamma"), _chk(beta, F32, "beta"),
              _chk(W2, F32, "W2"), _chk(b2, F32, "b2"), mask, float(drop_scale), o["hpre"], o["stats"], o["logits"], o["loss_raw"], _stream())
    return o


def classifier_head_bwd(feat, g_off, gold, W1, gamma, beta, W2, mask, drop_scale, hpre, stats, logits, grad_loss_sum):
    """-> (dfeat [G,256], dW1 [128,256], dgamma [128], dbeta [128], dW2 [40,128], db2 [40]); per-scene partials summed in scene order"""
    G, B = feat.shape[0], g_off.numel() - 1
    dev = feat.device
    e = lambda *s: torch.empty(*s, dtype=F32, device=dev)
    scratch, dfeat = e(G, 128), e(G, 256)
    dW1p, dgp, dbp, dW2p, db2p = e(B, 128, 256), e(B, 128), e(B, 128), e(B, 40, 128), e(B, 40)
    _lib.call("sgb_classifier_head_bwd", feat, G, g_off, B, gold, W1, gamma, beta, W2, mask, float(drop_scale), hpre, stats, logits,
              _chk(grad_loss_sum, F32, "grad_loss_sum"), scratch, dfeat, dW1p, dgp, dbp, dW2p, db2p, _stream())
    if B == 1:
        return dfeat, dW1p[0], dgp[0], dbp[0], dW2p[0], db2p[0]
    return dfeat, dW1p.sum(0), dgp.sum(0), dbp.sum(0), dW2p.sum(0), db2p.sum(0)


_VALID_IDS = {}


def evaluate(real_label, sem_pred, ins_pred, sem_valid, ins_valid, status=None):
    """model.py:608-655 on the device: real_label [n,2] i64, sem_pred / ins_pred [n] i32 -> out [164] f32
    (IoU_sem [2,40], IoU_ins [2,40], acc [4])."""
    import numpy as np
    _chk(real_label, torch.int64, "real_label"); _chk(sem_pred, I32, "sem_pred"); _chk(ins_pred, I32, "ins_pred")
    n = sem_pred.numel()
    key = (tuple(sem_valid), tuple(ins_valid))
    ids = _VALID_IDS.get(key)
    if ids is None:
        ids = _VALID_IDS[key] = (np.ascontiguousarray(sem_valid, np.int32), np.ascontiguousarray(ins_valid, np.int32))
    out = torch.empty(164, dtype=F32, device=sem_pred.device)
    ws = _ws(_lib.call("sgb_evaluate_ws_bytes"), sem_pred.device)
    _lib.call("sgb_evaluate", real_label, sem_pred, ins_pred, n, ids[0], len(ids[0]), ids[1], len(ids[1]), out, status, ws, ws.numel(),
              _stream())
    return out


def evaluate_scenes(real_label, sem_pred, ins_pred, raw_off, sem_valid, ins_valid, status=None):
    """sgb_evaluate for every scene of a batch in ONE library call; raw_off: host list of n_scenes + 1 raw-vertex offsets.
    -> out [n_scenes,164] f32"""
    import numpy as np
    _chk(real_label, torch.int64, "real_label"); _chk(sem_pred, I32, "sem_pred"); _chk(ins_pred, I32, "ins_pred")
    key = (tuple(sem_valid), tuple(ins_valid))
    ids = _VALID_IDS.get(key)
    if ids is None:
        ids = _VALID_IDS[key] = (np.ascontiguousarray(sem_valid, np.int32), np.ascontiguousarray(ins_valid, np.int32))
    B = len(raw_off) - 1
    off = np.ascontiguousarray(raw_off, np.int32)
    if off[-1] > sem_pred.numel() or off[-1] > real_label.shape[0]:
        raise ValueError("raw_off exceeds the label vectors")
    out = torch.empty(B, 164, dtype=F32, device=sem_pred.device)
    ws = _ws(_lib.call("sgb_evaluate_ws_bytes"), sem_pred.device)
    _lib.call("sgb_evaluate_scenes", real_label, sem_pred, ins_pred, off, B, ids[0], len(ids[0]), ids[1], len(ids[1]), out, status, ws,
              ws.numel(), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# tensor-core primitives (tcgen05, TF32 x 3)
# ------------------------------------------------------------------------------------------------
def gemm_tf32x3(A, B, relu=False):
    """C [M,N] = [relu](A [M,K] @ B [N,K]^T) on the tcgen05 tensor cores with the TF32 x 3 split (fp32-level accuracy).
    N multiple of 16 and <= 256, K multiple of 4."""
    _chk(A, F32, "A"); _chk(B, F32, "B")
    M, K = A.shape
    N = B.shape[0]
    if B.shape[1] != K:
        raise ValueError("A [M,K] and B [N,K] must share K")
    if not relu and M <= 512 and K >= 1024:
        # tall contraction (weight gradient): one CTA per 128 rows would walk K alone -> split K over the grid, sum the slabs in order
        kps = max(128, ((K // 96) + 31) // 32 * 32)
        ns = (K + kps - 1) // kps
        part = torch.empty(ns, M, N, dtype=F32, device=A.device)
        _lib.call("sgb_gemm_tf32x3_splitk", A, B, part, M, N, K, kps, _stream())
        return part.sum(0)
    C = torch.empty(M, N, dtype=F32, device=A.device)
    _lib.call("sgb_linear_tf32x3", A, B, C, M, N, K, int(bool(relu)), _stream())
    return C


def linear_tf32x3(x, Wt):
    """y [n,Cout] = x [n,Cin] @ Wt [Cout,Cin]^T for any Cout / Cin: the tcgen05 GEMM takes N <= 256 (multiple of 16) and K % 4 == 0,
    so wide outputs are produced in column panels of 256 and odd inner / outer sizes are zero-padded."""
    n, cin = x.shape
    cout = Wt.shape[0]
    kp = (-cin) % 4
    if kp:
        x = torch.nn.functional.pad(x, (0, kp)); Wt = torch.nn.functional.pad(Wt, (0, kp))
    x = x.contiguous()
    outs = []
    for c0 in range(0, cout, 256):
        w = Wt[c0:c0 + 256]
        npad = (-w.shape[0]) % 16
        if npad:
            w = torch.nn.functional.pad(w, (0, 0, 0, npad))
        y = gemm_tf32x3(x, w.contiguous())
        outs.append(y[:, :y.shape[1] - npad] if npad else y)
    return outs[0] if len(outs) == 1 else torch.cat(outs, 1)
