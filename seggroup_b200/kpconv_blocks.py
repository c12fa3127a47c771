"""Rigid KPFCNN blocks (kpconv/models/network_blocks.py:147-337, 530-581, 824-948) on the seggroup_b200 operators —
SURVEY.md 8f row N1: the strided / upsampling blocks that sit around `KPConv_ops`, `ind_max_pool` and `closest_pool`.

The reference builds these as TensorFlow-1 graph functions `block(layer_ind, inputs, features, radius, fdim, config,
training)` that create their variables on the fly.  Here every block is an `nn.Module` that owns the same variables (same
shapes, reference names in the docstrings) and whose `forward(layer_ind, inputs, features, radius, config, training)`
takes the same `inputs` dictionary (`points`, `neighbors`, `pools`, `upsamples` per layer) — `get_block_ops(name)` maps the
reference block names to the classes.  All point-sized work runs in the library kernels:

    KPConv (gather + influence + K x Cin x Cout contraction)   sgb_kpconv_fwd / _bwd  (tcgen05 contraction)
    ind_max_pool / closest_pool                                sgb_ind_max_pool_* / sgb_closest_pool_*

    unary convolutions [n, Cin] x [Cin, Cout]                  sgb_linear_tf32x3 (tcgen05, TF32 x 3), forward and backward
and batch norm / LeakyReLU are elementwise torch ops on the device.  There is no
CPU path (the operators reject non-CUDA tensors).  Kernel-point dispositions are an explicit input (`config.K_points`,
[K,3] for unit K_radius; the reference regenerates them with an unseeded optimisation): they are scaled by
K_radius = 1.5 * extent exactly as convolution_ops.py:124-131 does.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kpconv_ops as KO


def weight_variable(shape):
    """network_blocks.py:37-41: truncated normal (2 sigma) with stddev sqrt(2 / shape[-1]), rounded to 1e-3."""
    std = (2.0 / shape[-1]) ** 0.5
    w = nn.init.trunc_normal_(torch.empty(*shape), 0.0, std, -2.0 * std, 2.0 * std)
    return nn.Parameter(torch.round(w * 1000.0) / 1000.0)


class BatchNorm(nn.Module):
    """network_blocks.py:147-163 `batch_norm`: tf.layers.batch_normalization(momentum, epsilon=1e-6) or a bias."""

    def __init__(self, dim, use_batch_norm=True, momentum=0.99):
        super().__init__()
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.bn = nn.BatchNorm1d(dim, eps=1e-6, momentum=1.0 - momentum)      # TF momentum = 1 - torch momentum
        else:
            self.offset = nn.Parameter(torch.zeros(dim))

    def forward(self, x, training=True, slope=1.0, residual=None):
        """slope-activation(batch_norm(x) + residual) in one library call (`sgb_bn_act_fwd/_bwd`); slope = 1: no activation.
        Without batch norm (config.use_batch_norm = False) the reference adds a bias: plain torch, as rare as that setting."""
        if not self.use_batch_norm:
            y = x + self.offset
            if residual is not None:
                y = y + residual
            return y if slope == 1.0 else F.leaky_relu(y, slope)
        bn = self.bn
        return KO.bn_act(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, residual=residual, eps=bn.eps, slope=slope,
                         training=training, momentum=bn.momentum)


def leaky_relu(features, alpha=0.2):
    return F.leaky_relu(features, alpha)


def kp_conv(query_points, support_points, neighbors_indices, features, K_values, radius, config):
    """network_blocks.py:84-101 `KPConv`: extent from the layer radius, kernel points scaled to K_radius = 1.5 extent."""
    extent = config.KP_extent * radius / config.density_parameter
    K_points = config.K_points.to(features.device, torch.float32) * (1.5 * extent)
    return KO.KPConv(query_points, support_points, neighbors_indices, features, K_values, fixed=config.fixed_kernel_points,
                     KP_extent=extent, KP_influence=config.KP_influence, aggregation_mode=config.convolution_mode, K_points=K_points)


class UnaryBlock(nn.Module):
    """`unary_block` (176-188): 1x1 convolution + BN + LeakyReLU."""

    def __init__(self, in_dim, fdim, config):
        super().__init__()
        self.w = weight_variable([in_dim, fdim])
        self.bn = BatchNorm(fdim, config.use_batch_norm, config.batch_norm_momentum)

    def forward(self, layer_ind, inputs, features, radius, config, training=True):
        return self.bn(KO.unary_convolution(features, self.w), training, slope=0.2)


class SimpleBlock(nn.Module):
    """`simple_block` (191-213) / `simple_strided_block` (216-238): KPConv + BN + LeakyReLU."""
    strided = False

    def __init__(self, in_dim, fdim, config):
        super().__init__()
        self.w = weight_variable([config.num_kernel_points, in_dim, fdim])
        self.bn = BatchNorm(fdim, config.use_batch_norm, config.batch_norm_momentum)

    def forward(self, layer_ind, inputs, features, radius, config, training=True):
        if self.strided:
            q, s, idx = inputs['points'][layer_ind + 1], inputs['points'][layer_ind], inputs['pools'][layer_ind]
        else:
            q, s, idx = inputs['points'][layer_ind], inputs['points'][layer_ind], inputs['neighbors'][layer_ind]
        return self.bn(kp_conv(q, s, idx, features, self.w, radius, config), training, slope=0.2)


class SimpleStridedBlock(SimpleBlock):
    strided = True


class ResnetbBlock(nn.Module):
    """`resnetb_block` (290-337) / `resnetb_strided_block` (530-581): 1x1 -> KPConv -> 1x1 + shortcut (the strided shortcut is
    a max pooling over the pool neighbourhoods)."""
    strided = False

    def __init__(self, in_dim, fdim, config):
        super().__init__()
        self.conv1_w = weight_variable([in_dim, fdim // 2])
        self.conv1_bn = BatchNorm(fdim // 2, config.use_batch_norm, config.batch_norm_momentum)
        self.conv2_w = weight_variable([config.num_kernel_points, fdim // 2, fdim // 2])
        self.conv2_bn = BatchNorm(fdim // 2, config.use_batch_norm, config.batch_norm_momentum)
        self.conv3_w = weight_variable([fdim // 2, 2 * fdim])
        self.conv3_bn = BatchNorm(2 * fdim, config.use_batch_norm, config.batch_norm_momentum)
        if in_dim != 2 * fdim:
            self.shortcut_w = weight_variable([in_dim, 2 * fdim])
            self.shortcut_bn = BatchNorm(2 * fdim, config.use_batch_norm, config.batch_norm_momentum)
        else:
            self.shortcut_w = None

    def forward(self, layer_ind, inputs, features, radius, config, training=True):
        x = self.conv1_bn(KO.unary_convolution(features, self.conv1_w), training, slope=0.2)
        if self.strided:
            q, s, idx = inputs['points'][layer_ind + 1], inputs['points'][layer_ind], inputs['pools'][layer_ind]
        else:
            q, s, idx = inputs['points'][layer_ind], inputs['points'][layer_ind], inputs['neighbors'][layer_ind]
        x = self.conv2_bn(kp_conv(q, s, idx, x, self.conv2_w, radius, config), training, slope=0.2)
        shortcut = KO.ind_max_pool(features, inputs['pools'][layer_ind]) if self.strided else features
        if self.shortcut_w is not None:
            shortcut = self.shortcut_bn(KO.unary_convolution(shortcut, self.shortcut_w), training)
        # leaky_relu(batch_norm(conv3) + shortcut): the residual join rides in the same kernel as the last normalisation
        return self.conv3_bn(KO.unary_convolution(x, self.conv3_w), training, slope=0.2, residual=shortcut)


class ResnetbStridedBlock(ResnetbBlock):
    strided = True


class ResnetbDeformableBlock(ResnetbBlock):
    """`resnetb_deformable_block` (393-440) / `resnetb_deformable_strided_block` (641-692): the bottleneck block whose middle
    convolution is a deformable KPConv (convolution_ops.py:252-493).  Extra variables, as the reference creates them (zeros):
    `conv2_offset_w` [K, fdim/2, 3K (+K if modulated)] and `conv2_offset_b`.  The offsets of the last forward are kept in
    `self.last_offsets` (with the operands of the layer) for the 'fitting' regulariser, `offsets_loss()`."""

    def __init__(self, in_dim, fdim, config):
        super().__init__(in_dim, fdim, config)
        K = config.num_kernel_points
        odim = (4 if getattr(config, "modulated", False) else 3) * K
        self.conv2_offset_w = nn.Parameter(torch.zeros(K, fdim // 2, odim))
        self.conv2_offset_b = nn.Parameter(torch.zeros(odim))
        self.last_offsets = None

    def forward(self, layer_ind, inputs, features, radius, config, training=True):
        x = self.conv1_bn(KO.unary_convolution(features, self.conv1_w), training, slope=0.2)
        if self.strided:
            q, s, idx = inputs['points'][layer_ind + 1], inputs['points'][layer_ind], inputs['pools'][layer_ind]
        else:
            q, s, idx = inputs['points'][layer_ind], inputs['points'][layer_ind], inputs['neighbors'][layer_ind]
        extent = config.KP_extent * radius / config.density_parameter
        K_points = config.K_points.to(features.device, torch.float32) * (1.5 * extent)
        x, offsets = KO.KPConv_deformable(q, s, idx, x, self.conv2_w, self.conv2_offset_w, self.conv2_offset_b, fixed=config.fixed_kernel_points,
                                          KP_extent=extent, KP_influence=config.KP_influence, aggregation_mode=config.convolution_mode,
                                          modulated=getattr(config, "modulated", False), K_points=K_points)
        self.last_offsets = (q, s, idx, K_points, offsets, extent)
        x = self.conv2_bn(x, training, slope=0.2)
        shortcut = KO.ind_max_pool(features, inputs['pools'][layer_ind]) if self.strided else features
        if self.shortcut_w is not None:
            shortcut = self.shortcut_bn(KO.unary_convolution(shortcut, self.shortcut_w), training)
        # leaky_relu(batch_norm(conv3) + shortcut): the residual join rides in the same kernel as the last normalisation
        return self.conv3_bn(KO.unary_convolution(x, self.conv3_w), training, slope=0.2, residual=shortcut)

    def offsets_loss(self, loss_type="fitting"):
        """KPFCNN_model.py:218-286 for this layer (multiply by config.offsets_decay and add to the training loss)."""
        q, s, idx, K_points, offsets, extent = self.last_offsets
        return KO.deformable_offsets_loss(q, s, idx, K_points, offsets, extent, loss_type)


class ResnetbDeformableStridedBlock(ResnetbDeformableBlock):
    strided = True


class MaxPoolBlock(nn.Module):
    """`max_pool_block` (824-832)."""

    def __init__(self, in_dim=None, fdim=None, config=None):
        super().__init__()

    def forward(self, layer_ind, inputs, features, radius=None, config=None, training=True):
        return KO.ind_max_pool(features, inputs['pools'][layer_ind])


class NearestUpsampleBlock(nn.Module):
    """`nearest_upsample_block` (940-948)."""

    def __init__(self, in_dim=None, fdim=None, config=None):
        super().__init__()

    def forward(self, layer_ind, inputs, features, radius=None, config=None, training=True):
        return KO.closest_pool(features, inputs['upsamples'][layer_ind - 1])


_BLOCKS = {
    'unary': UnaryBlock, 'simple': SimpleBlock, 'simple_strided': SimpleStridedBlock, 'resnetb': ResnetbBlock,
    'resnetb_strided': ResnetbStridedBlock, 'max_pool': MaxPoolBlock, 'max_pool_wide': MaxPoolBlock, 'nearest_upsample': NearestUpsampleBlock,
    'resnetb_deformable': ResnetbDeformableBlock, 'resnetb_deformable_strided': ResnetbDeformableStridedBlock,
}


def get_block_ops(block_name):
    """network_blocks.py:951-1015 for the rigid blocks and the deformable bottleneck blocks of the ScanNet architecture
    (training_Scannet.py:78-98); the inception / `_v2` development variants are not built."""
    if block_name not in _BLOCKS:
        raise ValueError('Unknown block name in the architecture definition : ' + block_name)
    return _BLOCKS[block_name]
