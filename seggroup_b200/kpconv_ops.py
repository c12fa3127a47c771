"""KPConv operator set on the CUDA kernels, behind the reference's call signatures (boundaries B2-B4).

  B2  grid_subsampling.compute(points, features=None, classes=None, sampleDl=0.1, method='barycenters', verbose=0)
      numpy in / numpy out, same RuntimeError messages as kpconv/cpp_wrappers/cpp_subsampling/wrapper.cpp:58-286
  B3  batch_grid_subsampling(points, batches, dl), grid_subsampling_op(points, dl),
      batch_ordered_neighbors(queries, supports, q_batches, s_batches, radius), ordered_neighbors(queries, supports, radius)
      CUDA tensors in / out, positional order and dtypes of kpconv/tf_custom_ops/tf_*/tf_*.cpp
  B4  KPConv_ops(query_points, support_points, neighbors_indices, features, K_points, K_values, KP_extent,
                 KP_influence, aggregation_mode)    kpconv/kernels/convolution_ops.py:161-169

Output ORDER of the subsampling is canonical (voxels in order of first occurrence per batch element, label ties ->
smallest label) and neighbour rows are sorted by (distance, index): the reference's orders are libstdc++ hash-map
iteration order and std::sort tie order.  No CPU fallback anywhere.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .ops import F32, I32, _chk, _stream, _ws

_INFLUENCE = {"linear": 0, "constant": 1, "gaussian": 2}


def _status_check(status, what):
    s = int(status.item())
    if s & 4:
        raise RuntimeError("%s: batch lengths do not sum to the number of points" % what)
    if s & 8:
        raise RuntimeError("%s: voxel grid exceeds 2^44 cells (sampleDl too small for this extent)" % what)
    if s & 16:
        raise RuntimeError("%s: more than 2^17 grid cells along an axis (radius too small for this extent)" % what)
    if s & 32:
        raise RuntimeError("%s: a query has more than 1024 neighbours" % what)


def _grid_subsample(points, features, classes, batches, dl):
    dev = points.device
    N = points.shape[0]
    B = 1 if batches is None else batches.numel()
    fdim = 0 if features is None else features.shape[1]
    ldim = 0 if classes is None else classes.shape[1]
    out_xyz = torch.empty(N, 3, dtype=F32, device=dev)
    out_f = torch.empty(N, fdim, dtype=F32, device=dev) if fdim else None
    out_c = torch.empty(N, ldim, dtype=I32, device=dev) if ldim else None
    out_b = torch.empty(B, dtype=I32, device=dev)
    counts = torch.zeros(4, dtype=I32, device=dev)
    status = torch.zeros(1, dtype=I32, device=dev)
    ws = _ws(_lib.call("sgb_grid_subsample_ws_bytes", N, B), dev)
    _lib.call("sgb_grid_subsample", points, features, classes, N, fdim, ldim, batches, B, float(dl), out_xyz, out_f, out_c, None, out_b,
              counts, status, ws, ws.numel(), _stream())
    M = int(counts[0].item())
    _status_check(status, "grid_subsampling")
    return out_xyz[:M], (out_f[:M] if fdim else None), (out_c[:M] if ldim else None), out_b


# ------------------------------------------------------------------------------------------------ B3
def batch_grid_subsampling(points, batches, dl):
    """tf_batch_subsampling.cpp:8-20: points f32 [N,3], batches i32 [B], dl -> (sub_points f32 [M,3], sub_batches i32 [B])."""
    _chk(points, F32, "points"); _chk(batches, I32, "batches")
    if points.dim() != 2 or points.shape[1] != 3:
        raise ValueError("BatchGridSubsampling expects points with shape [None, 3]")
    sub, _, _, sb = _grid_subsample(points, None, None, batches, dl)
    return sub, sb


def grid_subsampling_op(points, dl):
    """tf_subsampling.cpp:8-17: points f32 [N,3], dl -> sub_points f32 [M,3]."""
    _chk(points, F32, "points")
    return _grid_subsample(points, None, None, None, dl)[0]


def batch_ordered_neighbors(queries, supports, q_batches, s_batches, radius):
    """tf_batch_neighbors.cpp:8-30: -> neighbors i32 [Nq, W]; padding / shadow index = supports.shape[0]."""
    _chk(queries, F32, "queries"); _chk(supports, F32, "supports")
    if q_batches is not None:
        _chk(q_batches, I32, "q_batches"); _chk(s_batches, I32, "s_batches")
    dev = queries.device
    Nq, Ns = queries.shape[0], supports.shape[0]
    B = 1 if q_batches is None else q_batches.numel()
    status = torch.zeros(1, dtype=I32, device=dev)
    maxc = torch.zeros(1, dtype=I32, device=dev)
    ws = _ws(_lib.call("sgb_radius_neighbors_ws_bytes", Nq, Ns, B), dev)
    _lib.call("sgb_radius_neighbors_count", queries, Nq, supports, Ns, q_batches, s_batches, B, float(radius), maxc, status, ws, ws.numel(), _stream())
    W = int(maxc.item())                          # the output width is data dependent (neighbors.cpp:296-304)
    _status_check(status, "batch_ordered_neighbors")
    nb = torch.empty(Nq, W, dtype=I32, device=dev)
    _lib.call("sgb_radius_neighbors_fill", queries, Nq, supports, Ns, B, float(radius), W, nb, status, ws, ws.numel(), _stream())
    _status_check(status, "batch_ordered_neighbors")
    return nb


def ordered_neighbors(queries, supports, radius):
    """tf_neighbors.cpp:8-17 (single cloud)."""
    return batch_ordered_neighbors(queries, supports, None, None, radius)


# ------------------------------------------------------------------------------------------------ unary convolution
class _UnaryConvFn(torch.autograd.Function):
    """`conv_ops.unary_convolution(features, w)` = features @ w (kpconv/kernels/convolution_ops.py:58-66, used by every unary /
    resnetb block, network_blocks.py:176-188, 290-337) on the tcgen05 tensor cores (TF32 x 3: fp32-level accuracy), forward and
    both backward products."""
    @staticmethod
    def forward(ctx, x, w):
        from . import ops
        ctx.save_for_backward(x, w)
        return ops.linear_tf32x3(x, w.t())

    @staticmethod
    def backward(ctx, g):
        from . import ops
        x, w = ctx.saved_tensors
        g = g.contiguous()
        dx = ops.linear_tf32x3(g, w)                                   # [n,Cout] x [Cin,Cout]^T
        n = x.shape[0]
        pad = (-n) % 4
        xt, gt = x.t(), g.t()
        if pad:
            xt, gt = torch.nn.functional.pad(xt, (0, pad)), torch.nn.functional.pad(gt, (0, pad))
        dw = ops.linear_tf32x3(xt.contiguous(), gt.contiguous())      # [Cin,n] x [Cout,n]^T, split over n inside
        return dx, dw


def unary_convolution(features, w):
    _chk(features, F32, "features")
    return _UnaryConvFn.apply(features, w)


# ------------------------------------------------------------------------------------------------ N1: batch norm + LeakyReLU
class _BnActFn(torch.autograd.Function):
    """`leaky_relu(batch_norm(x) [+ residual])` of the KPFCNN blocks (kpconv/models/network_blocks.py:147-173, 337, 581) as library
    kernels (`sgb_bn_act_fwd/_bwd`): column statistics in fp32 partials + fp64 reduction, one streaming apply pass; the backward
    returns dx, d(residual), dgamma, dbeta."""
    @staticmethod
    def forward(ctx, x, gamma, beta, residual, running_mean, running_var, eps, slope, training, momentum):
        from . import ops
        x = _chk(x.contiguous(), F32, "x")
        n, d = x.shape
        y = torch.empty_like(x)
        stat = torch.empty(3, d, dtype=F32, device=x.device)
        res = residual.contiguous() if residual is not None else None
        ws = ops._ws(_lib.call("sgb_bn_act_ws_bytes", n, d), x.device)
        _lib.call("sgb_bn_act_fwd", x, n, d, gamma, beta, res, float(eps), float(slope), int(bool(training)), float(momentum),
                  running_mean, running_var, y, stat, ws, ws.numel(), ops._stream())
        ctx.save_for_backward(x, y, stat, gamma if gamma is not None else x.new_empty(0))
        ctx.cfg = (float(slope), int(bool(training)), gamma is not None, beta is not None, residual is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        x, y, stat, gamma = ctx.saved_tensors
        slope, training, has_g, has_b, has_r = ctx.cfg
        n, d = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if has_r else None
        dg = torch.empty(d, dtype=F32, device=x.device) if has_g else None
        db = torch.empty(d, dtype=F32, device=x.device) if has_b else None
        ws = ops._ws(_lib.call("sgb_bn_act_ws_bytes", n, d), x.device)
        _lib.call("sgb_bn_act_bwd", dy, x, y, n, d, gamma if has_g else None, stat, slope, training, dx, dres, dg, db,
                  ws, ws.numel(), ops._stream())
        return dx, dg, db, dres, None, None, None, None, None, None


def bn_act(x, gamma, beta, running_mean=None, running_var=None, residual=None, eps=1e-6, slope=0.2, training=True, momentum=0.01):
    """act(batch_norm(x) + residual): x [n,d] f32 CUDA; slope = 1.0 for no activation.  In training the running statistics (if
    given) are updated in place with torch's momentum rule; in evaluation they are required."""
    return _BnActFn.apply(x, gamma, beta, residual, running_mean, running_var, eps, slope, training, momentum)


# ------------------------------------------------------------------------------------------------ B4
class _KPConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, s, idx, feats, kpts, kvals, extent, influence, closest, use_tc=True):
        n, n0 = q.shape[0], s.shape[0]
        W = idx.shape[1]
        K, Cin, Cout = kvals.shape
        out = torch.empty(n, Cout, dtype=F32, device=q.device)
        if use_tc and _lib.call("sgb_kpconv_tc_supported", W, Cin, Cout, K, n0):
            # contraction on tcgen05 (TF32 x 3 split, fp32-level accuracy); shapes outside its limits take the SIMT kernel
            ws = _ws(_lib.call("sgb_kpconv_tc_ws_bytes", Cin, Cout, K), q.device)
            _lib.call("sgb_kpconv_fwd_tc", q, s, idx, feats, kpts, kvals, n, n0, W, Cin, Cout, K, float(extent), influence, closest, out,
                      ws, ws.numel(), _stream())
        else:
            _lib.call("sgb_kpconv_fwd", q, s, idx, feats, kpts, kvals, n, n0, W, Cin, Cout, K, float(extent), influence, closest, out, _stream())
        ctx.save_for_backward(q, s, idx, feats, kpts, kvals)
        ctx.cfg = (float(extent), influence, closest)
        ctx.use_tc = bool(use_tc)
        return out

    @staticmethod
    def backward(ctx, g):
        q, s, idx, feats, kpts, kvals = ctx.saved_tensors
        extent, influence, closest = ctx.cfg
        n, n0 = q.shape[0], s.shape[0]
        K, Cin, Cout = kvals.shape
        gf = torch.zeros(n0, Cin, dtype=F32, device=q.device)
        gk = torch.zeros(K, Cin, Cout, dtype=F32, device=q.device)
        W = idx.shape[1]
        if ctx.use_tc and n > 0 and _lib.call("sgb_kpconv_bwd_tc_supported", n, W, Cin, Cout, K, n0):
            # both contractions (dK = WF^T g, GW = g K^T) on tcgen05, TF32 x 3; other shapes take the SIMT kernel
            ws = _ws(_lib.call("sgb_kpconv_bwd_tc_ws_bytes", n, Cin, Cout, K), q.device)
            _lib.call("sgb_kpconv_bwd_tc", g.contiguous(), q, s, idx, feats, kpts, kvals, n, n0, W, Cin, Cout, K, extent, influence, closest,
                      gf, gk, ws, ws.numel(), _stream())
        else:
            ws = _ws(_lib.call("sgb_kpconv_bwd_ws_bytes", n, Cin, Cout, K), q.device)
            _lib.call("sgb_kpconv_bwd", g.contiguous(), q, s, idx, feats, kpts, kvals, n, n0, W, Cin, Cout, K, extent, influence, closest,
                      gf, gk, ws, ws.numel(), _stream())
        return None, None, None, gf, None, gk, None, None, None, None


def KPConv_ops(query_points, support_points, neighbors_indices, features, K_points, K_values, KP_extent, KP_influence, aggregation_mode,
               tensor_cores=True):
    """kpconv/kernels/convolution_ops.py:161-249, same argument order; differentiable w.r.t. features and K_values.
    `tensor_cores=False` forces the fp32 SIMT kernel (tests compare the two)."""
    if KP_influence not in _INFLUENCE:
        raise ValueError('Unknown influence function type (config.KP_influence)')
    if aggregation_mode not in ("sum", "closest"):
        raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
    q = _chk(query_points.contiguous(), F32, "query_points"); s = _chk(support_points.contiguous(), F32, "support_points")
    idx = neighbors_indices
    if idx.dtype != I32:
        idx = idx.to(I32)
    idx = _chk(idx.contiguous(), I32, "neighbors_indices")
    return _KPConvFn.apply(q, s, idx, features.contiguous(), _chk(K_points.contiguous(), F32, "K_points"), K_values.contiguous(),
                           KP_extent, _INFLUENCE[KP_influence], int(aggregation_mode == "closest"), bool(tensor_cores))


# ------------------------------------------------------------------------------------------------ N3 deformable KPConv
class _KPConvDeformFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, s, idx, feats, kpts, offsets, mods, kvals, extent, influence, closest):
        n, n0 = q.shape[0], s.shape[0]
        K, Cin, Cout = kvals.shape
        out = torch.empty(n, Cout, dtype=F32, device=q.device)
        _lib.call("sgb_kpconv_deform_fwd", q, s, idx, feats, kpts, offsets, mods, kvals, n, n0, idx.shape[1], Cin, Cout, K, float(extent),
                  influence, closest, out, _stream())
        ctx.save_for_backward(q, s, idx, feats, kpts, offsets, kvals, mods if mods is not None else q.new_empty(0))
        ctx.cfg = (float(extent), influence, closest, mods is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        q, s, idx, feats, kpts, offsets, kvals, mods = ctx.saved_tensors
        extent, influence, closest, has_mod = ctx.cfg
        n, n0 = q.shape[0], s.shape[0]
        K, Cin, Cout = kvals.shape
        dev = q.device
        gf = torch.zeros(n0, Cin, dtype=F32, device=dev)
        gk = torch.zeros(K, Cin, Cout, dtype=F32, device=dev)
        go = torch.zeros(n, K, 3, dtype=F32, device=dev)
        gm = torch.zeros(n, K, dtype=F32, device=dev) if has_mod else None
        ws = _ws(_lib.call("sgb_kpconv_bwd_ws_bytes", n, Cin, Cout, K), dev)
        _lib.call("sgb_kpconv_deform_bwd", g.contiguous(), q, s, idx, feats, kpts, offsets, mods if has_mod else None, kvals, n, n0, idx.shape[1],
                  Cin, Cout, K, extent, influence, closest, gf, gk, go, gm, ws, ws.numel(), _stream())
        return None, None, None, gf, None, go, gm, gk, None, None, None


def KPConv_deform_ops(query_points, support_points, neighbors_indices, features, K_points, offsets, modulations, K_values, KP_extent,
                      KP_influence, mode):
    """kpconv/kernels/convolution_ops.py:371-493, same argument order; differentiable w.r.t. features, offsets, modulations, K_values."""
    if KP_influence not in _INFLUENCE:
        raise ValueError('Unknown influence function type (config.KP_influence)')
    if mode not in ("sum", "closest"):
        raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
    q = _chk(query_points.contiguous(), F32, "query_points"); s = _chk(support_points.contiguous(), F32, "support_points")
    idx = neighbors_indices
    if idx.dtype != I32:
        idx = idx.to(I32)
    idx = _chk(idx.contiguous(), I32, "neighbors_indices")
    return _KPConvDeformFn.apply(q, s, idx, features.contiguous(), _chk(K_points.contiguous(), F32, "K_points"), offsets.contiguous(),
                                 modulations.contiguous() if modulations is not None else None, K_values.contiguous(), KP_extent,
                                 _INFLUENCE[KP_influence], int(mode == "closest"))


def KPConv_deformable(query_points, support_points, neighbors_indices, features, K_values, K_values0, b0, fixed='center', KP_extent=1.0,
                      KP_influence='linear', aggregation_mode='sum', modulated=False, K_points=None):
    """kpconv/kernels/convolution_ops.py:252-368: a rigid KPConv with its own weights K_values0 [K,Cin,3K (+K)] and bias b0 produces
    the kernel-point offsets (in units of KP_extent) and, if `modulated`, 2*sigmoid modulations; the deformed convolution uses K_values.
    K_values0 / b0 are the reference's `offset_conv_weights` / `offset_conv_bias` variables (created as zeros there, :321-322);
    K_points is an explicit input as in `KPConv`.  Returns (features, offsets) — the offsets feed the 'fitting' regulariser."""
    if K_points is None:
        raise ValueError("K_points must be given: kernel-point dispositions are an input of the op (SURVEY.md 8c)")
    n_kp = K_values.shape[0]
    odim = K_values0.shape[2]
    pad = (-odim) % 4                      # the kernels take output widths that are multiples of 4: 3K = 45 -> 48 zero columns, sliced off again
    kv0 = torch.nn.functional.pad(K_values0, (0, pad)) if pad else K_values0
    f0 = KPConv_ops(query_points, support_points, neighbors_indices, features, K_points, kv0, KP_extent, KP_influence, aggregation_mode)
    f0 = (f0[:, :odim] if pad else f0) + b0
    if modulated:
        offsets = f0[:, :3 * n_kp].reshape(-1, n_kp, 3)
        modulations = 2 * torch.sigmoid(f0[:, 3 * n_kp:])
    else:
        offsets = f0.reshape(-1, n_kp, 3)
        modulations = None
    offsets = offsets * KP_extent
    out = KPConv_deform_ops(query_points, support_points, neighbors_indices, features, K_points, offsets, modulations, K_values, KP_extent,
                            KP_influence, aggregation_mode)
    return out, offsets


def deformable_offsets_loss(query_points, support_points, neighbors_indices, K_points, offsets, KP_extent, loss_type="fitting"):
    """The offset regulariser of ONE deformable layer, kpconv/models/KPFCNN_model.py:218-286 (before `offsets_decay`):
    'fitting'   : mean over (query, kernel point) of the squared distance from the deformed kernel point to its closest neighbour,
                  in units of KP_extent^2 (:249-266), plus the repulsive terms sum_i mean_n sum_{j != i} max(0, 1.5 - |kp_i - kp_j| / extent)^2
                  with the other points held constant (:268-286);
    'permissive': mean of max(0, |kp| / extent - 1) (:226-239).
    The arg-min over the neighbour row is a kernel (sgb_deform_closest_neighbor); the differentiable part is elementwise torch."""
    q = _chk(query_points.contiguous(), F32, "query_points"); s = _chk(support_points.contiguous(), F32, "support_points")
    idx = neighbors_indices.to(I32).contiguous()
    kp = _chk(K_points.contiguous(), F32, "K_points")
    n, K = offsets.shape[0], kp.shape[0]
    dkp = offsets + kp                                                                       # deformed_KP, convolution_ops.py:418
    if loss_type == "permissive":
        return torch.clamp(torch.linalg.norm(dkp / KP_extent, dim=2) - 1.0, min=0.0).mean()
    if loss_type != "fitting":
        raise ValueError('Unknown offset loss')
    arg = torch.empty(n, K, dtype=I32, device=q.device)
    _lib.call("sgb_deform_closest_neighbor", q, s, idx, kp, offsets.detach().contiguous(), n, s.shape[0], idx.shape[1], K, arg, _stream())
    s_ext = torch.cat([s, torch.full((1, 3), 1000.0, dtype=F32, device=q.device)])
    nb = torch.gather(idx.long().clamp(min=0, max=s.shape[0]), 1, arg.long())               # [n,K] support ids (shadow = n0)
    rel = s_ext[nb] - q.unsqueeze(1)                                                         # [n,K,3]
    fit = (((rel - dkp) ** 2).sum(2) / (KP_extent ** 2)).mean()
    locs = dkp / KP_extent
    dist = torch.sqrt(((locs.detach().unsqueeze(1) - locs.unsqueeze(2)) ** 2).sum(3) + torch.eye(K, device=q.device) * 1e30)   # [n,i,j], j held constant
    rep = (torch.clamp(1.5 - dist, min=0.0) ** 2).sum(2).mean(0).sum()
    return fit + rep


def KPConv(query_points, support_points, neighbors_indices, features, K_values, fixed='center', KP_extent=1.0,
           KP_influence='linear', aggregation_mode='sum', K_points=None):
    """kpconv/kernels/convolution_ops.py:102-158.  The reference regenerates the kernel-point disposition inside this
    call (random optimisation + rotation, kernel_points.py:182-278, unseeded); here the disposition is an explicit
    input (`K_points`, [K,3], already scaled by the caller's K_radius) — SURVEY.md 8c."""
    if K_points is None:
        raise ValueError("K_points must be given: kernel-point dispositions are an input of the op (SURVEY.md 8c)")
    return KPConv_ops(query_points, support_points, neighbors_indices, features, K_points, K_values, KP_extent, KP_influence, aggregation_mode)


# ------------------------------------------------------------------------------------------------ a21
def _inds(inds):
    if inds.dtype != I32:
        inds = inds.to(I32)
    inds = _chk(inds.contiguous(), I32, "inds")
    if inds.dim() != 2 or inds.shape[1] < 1:
        raise ValueError("inds must have shape [n2, max_num]")
    return inds


class _IndMaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inds):
        n1, d = x.shape
        n2, W = inds.shape
        out = torch.empty(n2, d, dtype=F32, device=x.device)
        ws = _ws(_lib.call("sgb_ind_max_pool_ws_bytes", d), x.device)
        _lib.call("sgb_ind_max_pool_fwd", x, n1, d, inds, n2, W, out, ws, ws.numel(), _stream())
        ctx.save_for_backward(x, inds, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, inds, out = ctx.saved_tensors
        n1, d = x.shape
        n2, W = inds.shape
        gx = torch.empty_like(x)
        ws = _ws(_lib.call("sgb_ind_max_pool_ws_bytes", d), x.device)
        _lib.call("sgb_ind_max_pool_bwd", g.contiguous(), x, n1, d, inds, n2, W, out, gx, ws, ws.numel(), _stream())
        return gx, None


class _ClosestPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inds):
        n1, d = x.shape
        n2, W = inds.shape
        out = torch.empty(n2, d, dtype=F32, device=x.device)
        _lib.call("sgb_closest_pool_fwd", x, n1, d, inds, n2, W, out, _stream())
        ctx.save_for_backward(inds)
        ctx.shape = (n1, d)
        return out

    @staticmethod
    def backward(ctx, g):
        (inds,) = ctx.saved_tensors
        n1, d = ctx.shape
        n2, W = inds.shape
        gx = torch.empty(n1, d, dtype=F32, device=g.device)
        _lib.call("sgb_closest_pool_bwd", g.contiguous(), n1, d, inds, n2, W, gx, _stream())
        return gx, None


def ind_max_pool(x, inds):
    """kpconv/models/network_blocks.py:49-66: x [n1,d], inds [n2,max_num] (index n1 = shadow = column minimum of x)
    -> [n2,d]; differentiable w.r.t. x (tf.reduce_max gradient: ties share equally)."""
    x = _chk(x.contiguous(), F32, "x")
    if x.dim() != 2:
        raise ValueError("x must have shape [n1, d]")
    return _IndMaxPoolFn.apply(x, _inds(inds))


def closest_pool(x, inds):
    """kpconv/models/network_blocks.py:69-81: x [n1,d], inds [n2,max_num] (only column 0 is used; index n1 = shadow =
    zeros) -> [n2,d]; differentiable w.r.t. x."""
    x = _chk(x.contiguous(), F32, "x")
    if x.dim() != 2:
        raise ValueError("x must have shape [n1, d]")
    return _ClosestPoolFn.apply(x, _inds(inds))


# ------------------------------------------------------------------------------------------------ B2
class _GridSubsamplingModule:
    """Stands in for the numpy C-extension `grid_subsampling` (wrapper.cpp): `compute(...)`."""

    @staticmethod
    def compute(points, *, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0):
        if method not in ("barycenters", "voxelcenters"):
            raise RuntimeError('Error parsing method. Valid method names are "barycenters" and "voxelcenters" ')
        try:
            p = np.ascontiguousarray(points, dtype=np.float32)
        except Exception:
            raise RuntimeError("Error converting input points to numpy arrays of type float32")
        f = c = None
        if features is not None:
            try:
                f = np.ascontiguousarray(features, dtype=np.float32)
            except Exception:
                raise RuntimeError("Error converting input features to numpy arrays of type float32")
        if classes is not None:
            try:
                c = np.ascontiguousarray(classes, dtype=np.int32)
            except Exception:
                raise RuntimeError("Error converting input classes to numpy arrays of type int32")
        if p.ndim != 2 or p.shape[1] != 3:
            raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
        if f is not None and (f.ndim != 2 or f.shape[0] != p.shape[0]):
            raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
        if c is not None:
            if c.ndim > 2 or c.shape[0] != p.shape[0]:
                raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
            if c.ndim == 1:
                c = c[:, None]
        if not torch.cuda.is_available():
            raise _lib.SgbError("grid_subsampling.compute needs a CUDA device (seggroup_b200 has no CPU path)")
        dev = torch.device("cuda")
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        sub, sf, sc, _ = _grid_subsample(t(p), t(f), t(c), None, float(sampleDl))
        res = [sub.cpu().numpy()]
        if f is not None:
            res.append(sf.cpu().numpy())
        if c is not None:
            res.append(sc.cpu().numpy())
        return res[0] if len(res) == 1 else tuple(res)


grid_subsampling = _GridSubsamplingModule()
