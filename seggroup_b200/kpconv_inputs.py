"""KPConv input pipeline on the GPU: the per-layer index construction the reference runs on the CPU inside
`tf.data.map` — `tf_segmentation_inputs` (kpconv/datasets/common.py:1021-1158), `big_neighborhood_filter` (:377-384),
`tf_get_batch_inds` / `tf_stack_batch_inds` (:386-475) and the neighbour-limit calibration `calibrate_neighbors` (:551-652).

Every neighbour matrix / subsampled cloud comes from the CUDA operators of `kpconv_ops` (`batch_ordered_neighbors`,
`batch_grid_subsampling`), i.e. the caller of family 1 of the north_star runs on the same device as the network: a stacked batch of
input spheres goes in, the flat input list of the reference's network comes out, all CUDA tensors.

    li = segmentation_inputs(config, stacked_points, stacked_features, point_labels, stacks_lengths, batch_inds, neighborhood_limits)
    li == input_points [L] + input_neighbors [L] + input_pools [L] + input_upsamples [L]
          + [stacked_features, stacked_weights, stacked_batch_inds_0, stacked_batch_inds_1, point_labels]     (common.py:1139-1143)

`config` needs: architecture (list of block names), first_subsampling_dl, KP_extent, density_parameter.
"""
from __future__ import annotations

import numpy as np
import torch

from .kpconv_ops import batch_grid_subsampling, batch_ordered_neighbors

I32 = torch.int32


def get_batch_inds(stacks_len):
    """common.py:386-430: [3, 2, 5] -> [0, 0, 0, 1, 1, 2, 2, 2, 2, 2] (int32, device of stacks_len)."""
    # element e belongs to the batch entry whose end is the first one > e: a parallel binary search (torch.repeat_interleave builds the
    # same vector with one thread block — 300 us per million points, profiles/r03z_launches_bench.md)
    ends = torch.cumsum(stacks_len.long(), 0)
    total = int(ends[-1]) if ends.numel() else 0
    return torch.bucketize(torch.arange(total, device=stacks_len.device), ends, right=True).to(I32)


def stack_batch_inds(stacks_len):
    """common.py:432-475: [B, max_len (+1)] point indices of every batch element, padded with num_points; a column of shadow
    indices is appended when no row has padding (all elements equally long)."""
    dev = stacks_len.device
    lens = stacks_len.long()
    num_points, max_points = int(lens.sum()), int(lens.max())
    start = torch.cumsum(lens, 0) - lens
    col = torch.arange(max_points, device=dev).view(1, -1)
    inds = torch.where(col < lens.view(-1, 1), start.view(-1, 1) + col, torch.full((1, 1), num_points, device=dev))
    if num_points == max_points * lens.numel():
        inds = torch.cat([inds, torch.full((lens.numel(), 1), num_points, device=dev)], 1)
    return inds.to(I32)


def _layers(architecture):
    """The layer structure `tf_segmentation_inputs` derives from the block list (common.py:1051-1062, 1070-1076, 1084-1094):
    yields (conv_blocks, closing_block) per layer until a global / upsample block is met."""
    layer_blocks = []
    for block_i, block in enumerate(architecture):
        if "global" in block or "upsample" in block:
            break
        if not ("pool" in block or "strided" in block):
            layer_blocks = layer_blocks + [block]
            if block_i < len(architecture) - 1 and not ("upsample" in architecture[block_i + 1]):
                continue
        yield layer_blocks, block
        layer_blocks = []


def segmentation_inputs(config, stacked_points, stacked_features, point_labels, stacks_lengths, batch_inds, neighborhood_limits,
                        object_labels=None):
    """common.py:1021-1158 on CUDA tensors.  neighborhood_limits: per-layer column caps (calibrate_neighbors), or None = no crop."""
    dev = stacked_points.device
    min_len = stacks_lengths.min()
    batch_weights = min_len.float() / stacks_lengths.float()                                  # :1030-1032
    stacked_weights = batch_weights[batch_inds.long()]
    r_normal = config.first_subsampling_dl * config.KP_extent * 2.5                           # :1035
    input_points, input_neighbors, input_pools, input_upsamples, input_batches_len = [], [], [], [], []
    empty_i = lambda: torch.zeros(0, 1, dtype=I32, device=dev)
    for layer_blocks, block in _layers(config.architecture):
        if layer_blocks:                                                                      # :1067-1079
            if any("deformable" in b for b in layer_blocks[:-1]):
                r = r_normal * config.density_parameter / (config.KP_extent * 2.5)
            else:
                r = r_normal
            conv_i = batch_ordered_neighbors(stacked_points, stacked_points, stacks_lengths, stacks_lengths, r)
        else:
            conv_i = empty_i()
        if "pool" in block or "strided" in block:                                             # :1084-1101
            dl = 2 * r_normal / (config.KP_extent * 2.5)
            pool_p, pool_b = batch_grid_subsampling(stacked_points, stacks_lengths, dl)
            r = r_normal * config.density_parameter / (config.KP_extent * 2.5) if "deformable" in block else r_normal
            pool_i = batch_ordered_neighbors(pool_p, stacked_points, pool_b, stacks_lengths, r)
            up_i = batch_ordered_neighbors(stacked_points, pool_p, stacks_lengths, pool_b, 2 * r)
        else:                                                                                 # :1103-1108
            pool_i, up_i = empty_i(), empty_i()
            pool_p = torch.zeros(0, 3, dtype=torch.float32, device=dev)
            pool_b = torch.zeros(0, dtype=I32, device=dev)
        layer = len(input_points)
        if neighborhood_limits is not None:                                                   # big_neighborhood_filter, :377-384
            lim = int(neighborhood_limits[layer])
            conv_i, pool_i, up_i = conv_i[:, :lim].contiguous(), pool_i[:, :lim].contiguous(), up_i[:, :lim].contiguous()
        input_points.append(stacked_points); input_neighbors.append(conv_i); input_pools.append(pool_i)
        input_upsamples.append(up_i); input_batches_len.append(stacks_lengths)
        stacked_points, stacks_lengths = pool_p, pool_b                                       # :1123-1124
        r_normal *= 2
    li = input_points + input_neighbors + input_pools + input_upsamples
    li += [stacked_features, stacked_weights, stack_batch_inds(input_batches_len[0]), stack_batch_inds(input_batches_len[-1])]
    li += [point_labels]
    if object_labels is not None:                                                             # :1147-1156
        li += [object_labels[batch_inds.long()]]
    return li


def neighbor_histograms(flat_inputs, num_layers, hist_n):
    """One calibration step (common.py:617-620): histogram of the neighbourhood sizes of every layer's conv neighbours."""
    hists = torch.zeros(num_layers, hist_n, dtype=torch.int64, device=flat_inputs[0].device)
    for l, nb in enumerate(flat_inputs[num_layers:2 * num_layers]):
        if nb.shape[0] == 0:
            continue
        counts = (nb < nb.shape[0]).sum(1)                       # shadow index = number of support points = rows (same cloud)
        hists[l] = torch.bincount(counts, minlength=hist_n)[:hist_n]
    return hists


def calibrate_neighbors(batches, config, keep_ratio=0.8, samples_threshold=10000):
    """common.py:551-652: per-layer column cap that keeps `keep_ratio` of the neighbourhoods untouched.  `batches` yields
    (stacked_points, stacked_features, point_labels, stacks_lengths, batch_inds) CUDA tuples (one epoch at most is consumed)."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (config.density_parameter + 1) ** 3))                # :595
    num_layers = sum(1 for _ in _layers(config.architecture))
    total = None
    for b in batches:
        li = segmentation_inputs(config, *b, neighborhood_limits=None)
        h = neighbor_histograms(li, num_layers, hist_n)
        total = h if total is None else total + h
        if int(total.sum(1).min()) >= samples_threshold:                                      # :610
            break
    cumsum = torch.cumsum(total.t(), 0)                                                       # :645-646
    return (cumsum < keep_ratio * cumsum[hist_n - 1, :].double()).sum(0).cpu().numpy()
