"""TEST INFRASTRUCTURE — CPU restatement (oracle) of the KPConv operator set: grid subsampling, batch radius
neighbours and the rigid KPConv operator.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import it.

Pinning: grid subsampling and neighbours are checked against the compiled, unmodified reference cores
(oracle/_ref, tests/test_kpconv_oracle.py) through the canonical forms defined here — the reference's output
ORDER is libstdc++ hash-map iteration order / std::sort tie order, which no other implementation can follow
(SURVEY.md 7.3 #4, #5).  `kpconv_ops`, `kpconv_deform_ops`, `kpconv_deformable`, the index
pools and `block_forward` restate kpconv/kernels/convolution_ops.py:161-493 and kpconv/models/network_blocks.py; they are
pinned by golden vectors minted by EXECUTING those unmodified reference files on a torch-backed stand-in for the TensorFlow
primitives they call (oracle/tf_shim.py, oracle/make_golden_kpconv.py, tests/test_kpconv_reference_pin.py; TensorFlow
itself is not installable here): outputs and gradients agree to 2e-6 in float64.

Canonical forms:
  * subsampled voxels are listed per batch element in order of FIRST OCCURRENCE (the point with the smallest
    input index of every voxel, ascending) — i.e. the order in which the reference inserts them into its map;
  * majority label ties -> smallest label;
  * neighbour rows are sorted by (squared distance, index).
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------------------------
# grid subsampling  (kpconv/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106,
#                    kpconv/tf_custom_ops/tf_subsampling/grid_subsampling/grid_subsampling.cpp:5-150)
# ------------------------------------------------------------------------------------------------
def voxel_keys(points, dl):
    """fp32 arithmetic of grid_subsampling.cpp:24-56 -> (uint64 keys, origin, NX, NY)."""
    p = np.ascontiguousarray(points, np.float32)
    dl = np.float32(dl)
    mn, mx = p.min(0), p.max(0)
    inv = np.float32(1) / dl                                 # `1/sampleDl` (int / float -> float)
    origin = (np.floor(mn * inv) * dl).astype(np.float32)
    nx = np.uint64(np.floor((mx[0] - origin[0]) / dl)) + np.uint64(1)
    ny = np.uint64(np.floor((mx[1] - origin[1]) / dl)) + np.uint64(1)
    idx = np.floor((p - origin[None, :]) / dl).astype(np.uint64)
    return idx[:, 0] + nx * idx[:, 1] + nx * ny * idx[:, 2], origin, nx, ny


def grid_subsampling(points, features=None, classes=None, dl=0.1):
    """Single cloud.  Returns (sub_points [M,3], sub_features or None, sub_classes [M,ld] or None, first_index [M])
    in canonical (first occurrence) order.  Sums run in input order in fp32, barycentre = sum * (float)(1.0/count),
    feature mean = sum / (float)count (grid_subsampling.cpp:85-95)."""
    p = np.ascontiguousarray(points, np.float32)
    keys, _, _, _ = voxel_keys(p, dl)
    _, first, inv, cnt = np.unique(keys, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")                  # voxels by first occurrence
    rank = np.empty_like(order); rank[order] = np.arange(len(order))
    vox = rank[inv]                                           # voxel id of every point
    M = len(order)
    sums = np.zeros((M, 3), np.float32)
    for i in range(len(p)):                                   # input order, fp32 (np.add.at is not order-safe for floats)
        sums[vox[i]] += p[i]
    counts = cnt[order]
    scale = (1.0 / counts.astype(np.float64)).astype(np.float32)
    sub = sums * scale[:, None]
    subf = None
    if features is not None:
        f = np.ascontiguousarray(features, np.float32)
        fs = np.zeros((M, f.shape[1]), np.float32)
        for i in range(len(p)):
            fs[vox[i]] += f[i]
        subf = fs / counts.astype(np.float32)[:, None]
    subc = None
    if classes is not None:
        c = np.ascontiguousarray(classes, np.int32)
        if c.ndim == 1:
            c = c[:, None]
        subc = np.empty((M, c.shape[1]), np.int32)
        by_vox = np.argsort(vox, kind="stable")
        starts = np.concatenate([[0], np.cumsum(counts)])
        for v in range(M):
            rows = c[by_vox[starts[v]:starts[v + 1]]]
            for d in range(c.shape[1]):
                vals, n = np.unique(rows[:, d], return_counts=True)
                subc[v, d] = vals[np.argmax(n)]               # ties -> smallest label
    return sub, subf, subc, first[order]


def batch_grid_subsampling(points, batches, dl):
    """tf variant (points only): per batch element its own origin; -> (sub_points, sub_batches)."""
    out, lens, s = [], [], 0
    for b in batches:
        sub, _, _, _ = grid_subsampling(points[s:s + b], dl=dl)
        out.append(sub); lens.append(len(sub)); s += b
    return np.concatenate(out, 0), np.array(lens, np.int32)


def reference_to_canonical(ref_points, points, dl):
    """Permutation that brings the reference's output (libstdc++ hash-map iteration order) into canonical order:
    every barycentre is assigned to its voxel with the ORIGINAL cloud's origin / NX / NY, then voxels are ordered
    by first occurrence.  Returns perm with ref_points[perm] canonical."""
    p = np.ascontiguousarray(points, np.float32)
    keys, origin, nx, ny = voxel_keys(p, dl)
    uk, first = np.unique(keys, return_index=True)
    idx = np.floor((np.asarray(ref_points, np.float32) - origin[None, :]) / np.float32(dl)).astype(np.uint64)
    rk = idx[:, 0] + nx * idx[:, 1] + nx * ny * idx[:, 2]
    pos = np.searchsorted(uk, rk)
    if len(np.unique(rk)) != len(rk) or not np.array_equal(uk[np.clip(pos, 0, len(uk) - 1)], rk):
        raise ValueError("a reference barycentre does not fall inside its own voxel (fp32 edge case): pick another seed")
    return np.argsort(first[pos], kind="stable")


# ------------------------------------------------------------------------------------------------
# radius neighbours  (kpconv/tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp:125-332)
# ------------------------------------------------------------------------------------------------
def canonical_rows(nb, queries, supports):
    """Sort every row of a reference neighbour matrix by (d2, index); padding (index == Ns) stays at the end."""
    q = np.asarray(queries, np.float32); s = np.asarray(supports, np.float32)
    Ns = len(s)
    out = nb.copy()
    for i in range(len(nb)):
        row = nb[i]
        v = row[row < Ns]
        d = q[i][None, :] - s[v]
        d2 = ((d[:, 0] * d[:, 0]).astype(np.float32) + (d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32) + (d[:, 2] * d[:, 2]).astype(np.float32)
        o = np.lexsort((v, d2.astype(np.float32)))
        out[i, :len(v)] = v[o]
    return out


def batch_neighbors(queries, supports, q_batches, s_batches, radius, chunk=2048):
    """Brute force, fp32 d2 = dx*dx + dy*dy + dz*dz (left to right), strict d2 < r*r, rows sorted by (d2, index),
    padded with Ns to the global maximum count."""
    q = np.asarray(queries, np.float32); s = np.asarray(supports, np.float32)
    r2 = np.float32(radius) * np.float32(radius)
    rows = []
    qs = ss = 0
    for qb, sb in zip(q_batches, s_batches):
        S = s[ss:ss + sb]
        for c0 in range(qs, qs + qb, chunk):
            Q = q[c0:min(c0 + chunk, qs + qb)]
            d = Q[:, None, :] - S[None, :, :]
            d2 = ((d[..., 0] * d[..., 0]) + (d[..., 1] * d[..., 1])).astype(np.float32) + (d[..., 2] * d[..., 2])
            d2 = d2.astype(np.float32)
            for i in range(len(Q)):
                v = np.nonzero(d2[i] < r2)[0]
                o = np.lexsort((v, d2[i][v]))
                rows.append(v[o] + ss)
        qs += qb; ss += sb
    W = max((len(r) for r in rows), default=0)
    out = np.full((len(q), W), len(s), np.int32)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


# ------------------------------------------------------------------------------------------------
# KPConv  (kpconv/kernels/convolution_ops.py:161-249)
# ------------------------------------------------------------------------------------------------
def kpconv_ops(query_points, support_points, neighbors_indices, features, K_points, K_values, KP_extent,
               KP_influence="linear", aggregation_mode="sum", dtype=None):
    """torch restatement; tensors in, tensor out (differentiable w.r.t. features and K_values)."""
    import torch
    dt = dtype or features.dtype
    q = query_points.to(dt); s = support_points.to(dt); f = features.to(dt); Kp = K_points.to(dt); Kv = K_values.to(dt)
    idx = neighbors_indices.long()
    n_kp = Kp.shape[0]
    shadow = torch.ones_like(s[:1, :]) * 1e6                                   # :190
    s = torch.cat([s, shadow], dim=0)                                          # :191
    nb = s[idx]                                                                # :194  [n, W, 3]
    nb = nb - q.unsqueeze(1)                                                   # :197
    diff = nb.unsqueeze(2) - Kp.view(1, 1, n_kp, 3)                            # :200-203  [n, W, K, 3]
    sq = (diff ** 2).sum(dim=3)                                                # :205
    if KP_influence == "constant":                                             # :208-212
        w = torch.ones_like(sq).transpose(1, 2)
    elif KP_influence == "linear":                                             # :214-217
        w = torch.clamp(1 - torch.sqrt(sq) / KP_extent, min=0.0).transpose(1, 2)
    elif KP_influence == "gaussian":                                           # :219-222, radius_gaussian :48-55
        sigma = KP_extent * 0.3
        w = torch.exp(-sq / (2 * sigma ** 2 + 1e-9)).transpose(1, 2)
    else:
        raise ValueError("Unknown influence function type (config.KP_influence)")
    if aggregation_mode == "closest":                                          # :227-229
        nn_idx = torch.argmin(sq, dim=2)
        w = w * torch.nn.functional.one_hot(nn_idx, n_kp).to(dt).transpose(1, 2)
    elif aggregation_mode != "sum":
        raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
    f = torch.cat([f, torch.zeros_like(f[:1, :])], dim=0)                      # :234
    nf = f[idx]                                                                # :237  [n, W, Cin]
    wf = torch.matmul(w, nf)                                                   # :240  [n, K, Cin]
    wf = wf.permute(1, 0, 2)                                                   # :243
    out = torch.matmul(wf, Kv)                                                 # :244  [K, n, Cout]
    return out.sum(dim=0)                                                      # :247


# ------------------------------------------------------------------------------------------------
# deformable KPConv  (kpconv/kernels/convolution_ops.py:252-493)
# ------------------------------------------------------------------------------------------------
def kpconv_deform_ops(query_points, support_points, neighbors_indices, features, K_points, offsets, modulations, K_values,
                      KP_extent, KP_influence="linear", mode="sum"):
    """convolution_ops.py:371-493 in torch (differentiable w.r.t. features, K_values, offsets, modulations).

    The reference compacts every neighbour row to the neighbours that lie within KP_extent of at least one DEFORMED kernel
    point (:429-445: in_range -> top_k -> batch_gather, dropped entries re-pointed at the shadow row, whose feature is zero).
    Dropping a neighbour == multiplying all its K influence weights by zero, which is how it is written here: no data-dependent
    shapes, same result (for 'linear' the mask is implied by the clamp; for 'gaussian' / 'constant' / 'closest' it is not)."""
    import torch
    dt = features.dtype
    q = query_points.to(dt); s = support_points.to(dt); Kp = K_points.to(dt)
    idx = neighbors_indices.long()
    n_kp = Kp.shape[0]
    s = torch.cat([s, torch.ones_like(s[:1, :]) * 1000], dim=0)                # :405-406 (shadow point at 1000, not 1e6)
    nb = s[idx] - q.unsqueeze(1)                                               # :409-412  [n, W, 3]
    dkp = offsets.to(dt) + Kp                                                  # :415      [n, K, 3]
    sq = ((nb.unsqueeze(2) - dkp.unsqueeze(1)) ** 2).sum(dim=3)                # :418-423  [n, W, K]
    in_range = (sq < KP_extent ** 2).any(dim=2)                                # :426      [n, W]
    if KP_influence == "constant":                                             # :450-453
        w = (sq < KP_extent ** 2).to(dt)
    elif KP_influence == "linear":                                             # :455-458
        w = torch.clamp(1 - torch.sqrt(sq) / KP_extent, min=0.0)
    elif KP_influence == "gaussian":                                           # :460-464
        w = torch.exp(-sq / (2 * (KP_extent * 0.3) ** 2 + 1e-9))
    else:
        raise ValueError("Unknown influence function type (config.KP_influence)")
    if mode == "closest":                                                      # :469-471
        w = w * torch.nn.functional.one_hot(torch.argmin(sq, dim=2), n_kp).to(dt)
    elif mode != "sum":
        raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
    w = w * in_range.unsqueeze(2).to(dt)                                       # :429-445 (see above)
    f = torch.cat([features, torch.zeros_like(features[:1, :])], dim=0)        # :476
    wf = torch.matmul(w.transpose(1, 2), f[idx])                               # :479-482  [n, K, Cin]
    if modulations is not None:                                                # :485-486
        wf = wf * modulations.to(dt).unsqueeze(2)
    return torch.matmul(wf.permute(1, 0, 2), K_values.to(dt)).sum(dim=0)        # :489-493


def kpconv_deformable(query_points, support_points, neighbors_indices, features, K_points, K_values, K_values0, b0, KP_extent,
                      KP_influence="linear", aggregation_mode="sum", modulated=False):
    """convolution_ops.py:252-368 with K_points explicit: a rigid KPConv with its own weights K_values0 [K, Cin, 3K (+K)]
    and bias b0 produces per-query kernel-point offsets (in units of KP_extent) and, if `modulated`, 2*sigmoid modulations;
    the deformed convolution then uses K_values."""
    import torch
    n_kp = K_points.shape[0]
    f0 = kpconv_ops(query_points, support_points, neighbors_indices, features, K_points, K_values0, KP_extent, KP_influence,
                    aggregation_mode, dtype=features.dtype) + b0                # :324-332
    if modulated:                                                              # :334-342
        offsets = f0[:, :3 * n_kp].reshape(-1, n_kp, 3)
        modulations = 2 * torch.sigmoid(f0[:, 3 * n_kp:])
    else:                                                                      # :344-350
        offsets = f0.reshape(-1, n_kp, 3)
        modulations = None
    offsets = offsets * KP_extent                                              # :353
    return kpconv_deform_ops(query_points, support_points, neighbors_indices, features, K_points, offsets, modulations, K_values,
                             KP_extent, KP_influence, aggregation_mode)


# ------------------------------------------------------------------------------------------------
# index pooling beside KPConv  (kpconv/models/network_blocks.py:49-81)
# ------------------------------------------------------------------------------------------------
def ind_max_pool(x, inds):
    """network_blocks.py:49-66 in torch (CPU): shadow row = column minimum (:58), gather (:61), max over the listed
    rows (:64).  torch.amax / torch.amin split the gradient equally between ties, as tf.reduce_max / reduce_min do."""
    import torch
    xe = torch.cat([x, torch.amin(x, dim=0, keepdim=True)], dim=0)
    return torch.amax(xe[inds.long()], dim=1)


def closest_pool(x, inds):
    """network_blocks.py:69-81: shadow row = zeros (:77), rows of the first listed index (:80)."""
    import torch
    xe = torch.cat([x, torch.zeros(1, x.shape[1], dtype=x.dtype)], dim=0)
    return xe[inds[:, 0].long()]


# ------------------------------------------------------------------------------------------------
# rigid KPFCNN blocks  (kpconv/models/network_blocks.py:147-337, 530-581, 824-948), torch-CPU restatement
# ------------------------------------------------------------------------------------------------
def batch_norm_train(x, gamma, beta, eps=1e-6):
    """network_blocks.py:147-158 with training=True: tf.layers.batch_normalization = batch mean / biased variance."""
    mean = x.mean(dim=0, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=0, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * gamma + beta


def block_forward(name, P, layer_ind, inputs, features, radius, config, pre_activations=None, lrelu_masks=None):
    """One block of network_blocks.py in training mode.  P: dict of the block's variables (fp64 tensors) named as in
    seggroup_b200.kpconv_blocks; inputs: dict of per-layer lists `points`, `neighbors`, `pools`, `upsamples`.
    pre_activations: optional list that receives every LeakyReLU input.
    lrelu_masks: optional iterator of boolean tensors (x > 0 as another implementation saw it), consumed in call order: the
    BACKWARD of each LeakyReLU then uses that active set instead of its own.  At the kink the derivative jumps from 0.2 to 1,
    so two implementations whose pre-activations differ in the last bits disagree there by construction; gradient parity is
    therefore defined for a common active set (the tests check that the two sets differ only within rounding of zero)."""
    import torch
    dt = torch.float64

    class _LreluMasked(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, mask):
            ctx.save_for_backward(mask)
            return torch.nn.functional.leaky_relu(x, 0.2)

        @staticmethod
        def backward(ctx, g):
            (mask,) = ctx.saved_tensors
            return g * torch.where(mask, torch.ones_like(g), torch.full_like(g, 0.2)), None

    def _lrelu(x):                                                             # :166-167
        if pre_activations is not None:
            pre_activations.append(x.detach())
        if lrelu_masks is not None:
            return _LreluMasked.apply(x, next(lrelu_masks))
        return torch.nn.functional.leaky_relu(x, 0.2)

    def bn(x, pre):
        return batch_norm_train(x, P[pre + ".bn.weight"], P[pre + ".bn.bias"])

    def kp(q, s, idx, f, w):                                                   # :84-101
        extent = config.KP_extent * radius / config.density_parameter
        return kpconv_ops(q, s, idx, f, config.K_points.to(dt) * (1.5 * extent), w, extent, config.KP_influence, config.convolution_mode, dtype=dt)

    pts, nbs, pools = inputs["points"], inputs["neighbors"], inputs["pools"]
    if name == "unary":                                                        # :176-188
        return _lrelu(bn(features @ P["w"], "bn"))
    if name in ("simple", "simple_strided"):                                   # :191-238
        st = name.endswith("strided")
        q, s, idx = (pts[layer_ind + 1], pts[layer_ind], pools[layer_ind]) if st else (pts[layer_ind], pts[layer_ind], nbs[layer_ind])
        return _lrelu(bn(kp(q, s, idx, features, P["w"]), "bn"))
    if name in ("resnetb", "resnetb_strided", "resnetb_deformable", "resnetb_deformable_strided"):   # :290-337, 393-440, 530-581, 641-692
        st = name.endswith("strided")
        x = _lrelu(bn(features @ P["conv1_w"], "conv1_bn"))
        q, s, idx = (pts[layer_ind + 1], pts[layer_ind], pools[layer_ind]) if st else (pts[layer_ind], pts[layer_ind], nbs[layer_ind])
        if "deformable" in name:                                               # :104-122 -> convolution_ops.py:252-368
            extent = config.KP_extent * radius / config.density_parameter
            x = kpconv_deformable(q.to(dt), s.to(dt), idx, x, config.K_points.to(dt) * (1.5 * extent), P["conv2_w"], P["conv2_offset_w"],
                                  P["conv2_offset_b"], extent, config.KP_influence, config.convolution_mode, getattr(config, "modulated", False))
            x = _lrelu(bn(x, "conv2_bn"))
        else:
            x = _lrelu(bn(kp(q, s, idx, x, P["conv2_w"]), "conv2_bn"))
        x = bn(x @ P["conv3_w"], "conv3_bn")
        shortcut = ind_max_pool(features, pools[layer_ind]) if st else features
        if "shortcut_w" in P:
            shortcut = bn(shortcut @ P["shortcut_w"], "shortcut_bn")
        return _lrelu(x + shortcut)
    if name == "max_pool":                                                     # :824-832
        return ind_max_pool(features, pools[layer_ind])
    if name == "nearest_upsample":                                             # :940-948
        return closest_pool(features, inputs["upsamples"][layer_ind - 1])
    raise ValueError("Unknown block name in the architecture definition : " + name)


# ------------------------------------------------------------------------------------------------
# KPConv input pipeline  (kpconv/datasets/common.py:377-384, 432-475, 551-652, 1021-1158), numpy restatement
# ------------------------------------------------------------------------------------------------
def stack_batch_inds(stacks_len):
    """common.py:432-475"""
    lens = np.asarray(stacks_len, np.int64)
    num_points, max_points = int(lens.sum()), int(lens.max())
    rows, p = [], 0
    for n in lens:
        rows.append(np.concatenate([np.arange(p, p + n), np.full(max_points - n, num_points)]))
        p += n
    out = np.stack(rows)
    if num_points == max_points * len(lens):
        out = np.concatenate([out, np.full((len(lens), 1), num_points)], 1)
    return out.astype(np.int32)


def segmentation_inputs(config, stacked_points, stacked_features, point_labels, stacks_lengths, batch_inds, neighborhood_limits):
    """common.py:1021-1143 with the oracle's canonical subsampling / neighbour functions."""
    pts = np.asarray(stacked_points, np.float32); lens = np.asarray(stacks_lengths, np.int32)
    weights = (lens.min().astype(np.float32) / lens.astype(np.float32))[np.asarray(batch_inds)]
    r_normal = config.first_subsampling_dl * config.KP_extent * 2.5
    P, NB, PL, UP, BL = [], [], [], [], []
    layer_blocks = []
    arch = config.architecture
    for block_i, block in enumerate(arch):
        if "global" in block or "upsample" in block:
            break
        if not ("pool" in block or "strided" in block):
            layer_blocks += [block]
            if block_i < len(arch) - 1 and not ("upsample" in arch[block_i + 1]):
                continue
        if layer_blocks:
            deform = any("deformable" in b for b in layer_blocks[:-1])
            r = r_normal * config.density_parameter / (config.KP_extent * 2.5) if deform else r_normal
            conv_i = batch_neighbors(pts, pts, lens, lens, r)
        else:
            conv_i = np.zeros((0, 1), np.int32)
        if "pool" in block or "strided" in block:
            dl = 2 * r_normal / (config.KP_extent * 2.5)
            pool_p, pool_b = batch_grid_subsampling(pts, lens, dl)
            r = r_normal * config.density_parameter / (config.KP_extent * 2.5) if "deformable" in block else r_normal
            pool_i = batch_neighbors(pool_p, pts, pool_b, lens, r)
            up_i = batch_neighbors(pts, pool_p, lens, pool_b, 2 * r)
        else:
            pool_i = np.zeros((0, 1), np.int32); up_i = np.zeros((0, 1), np.int32)
            pool_p = np.zeros((0, 3), np.float32); pool_b = np.zeros((0,), np.int32)
        if neighborhood_limits is not None:
            lim = int(neighborhood_limits[len(P)])
            conv_i, pool_i, up_i = conv_i[:, :lim], pool_i[:, :lim], up_i[:, :lim]
        P.append(pts); NB.append(conv_i); PL.append(pool_i); UP.append(up_i); BL.append(lens)
        pts, lens = pool_p, pool_b
        r_normal *= 2
        layer_blocks = []
    return P + NB + PL + UP + [np.asarray(stacked_features), weights, stack_batch_inds(BL[0]), stack_batch_inds(BL[-1]), np.asarray(point_labels)]


def calibrate_neighbors(batches, config, keep_ratio=0.8, samples_threshold=10000):
    """common.py:595-647"""
    hist_n = int(np.ceil(4 / 3 * np.pi * (config.density_parameter + 1) ** 3))
    hists = None
    for b in batches:
        li = segmentation_inputs(config, *b, neighborhood_limits=None)
        L = (len(li) - 5) // 4
        hs = []
        for nb in li[L:2 * L]:
            counts = np.sum(nb < nb.shape[0], axis=1) if nb.shape[0] else np.zeros(0, np.int64)
            hs.append(np.bincount(counts, minlength=hist_n)[:hist_n])
        hs = np.vstack(hs)
        hists = hs if hists is None else hists + hs
        if np.min(np.sum(hists, axis=1)) >= samples_threshold:
            break
    cumsum = np.cumsum(hists.T, axis=0)
    return np.sum(cumsum < (keep_ratio * cumsum[hist_n - 1, :]), axis=0)
