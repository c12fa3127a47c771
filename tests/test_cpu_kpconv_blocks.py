"""Host-side logic of the KPFCNN block modules (seggroup_b200/kpconv_blocks.py): variable shapes and
initialisation of network_blocks.py:37-47, the block-name registry of :951-1015, and the oracle restatement's batch norm.
No kernel is called here (the operators themselves are CUDA only: tests/test_gpu_kpconv.py)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch


def cfg():
    return SimpleNamespace(KP_extent=1.0, density_parameter=5.0, KP_influence="linear", convolution_mode="sum", num_kernel_points=15,
                           use_batch_norm=True, batch_norm_momentum=0.99, fixed_kernel_points="center", K_points=torch.zeros(15, 3))


def test_weight_variable_is_truncated_normal_rounded_to_1e3():
    from seggroup_b200.kpconv_blocks import weight_variable
    torch.manual_seed(0)
    w = weight_variable([15, 64, 128]).detach()
    std = (2.0 / 128) ** 0.5
    assert float(w.abs().max()) <= 2 * std + 5e-4                                     # tf.truncated_normal: 2 sigma
    assert torch.allclose(w * 1000, torch.round(w * 1000), atol=1e-3)                 # tf.round(initial * 1000) / 1000
    assert abs(float(w.std()) - 0.88 * std) < 0.05 * std                              # std of a 2-sigma truncated normal


def test_block_registry_and_variable_shapes():
    from seggroup_b200 import kpconv_blocks as B
    c = cfg()
    with pytest.raises(ValueError, match="Unknown block name in the architecture definition : inception_deformable"):
        B.get_block_ops("inception_deformable")
    b = B.get_block_ops("resnetb_strided")(64, 64, c)
    shapes = {k: tuple(v.shape) for k, v in b.named_parameters()}
    assert shapes["conv1_w"] == (64, 32) and shapes["conv2_w"] == (15, 32, 32) and shapes["conv3_w"] == (32, 128)
    assert shapes["shortcut_w"] == (64, 128)
    assert B.get_block_ops("resnetb")(128, 64, c).shortcut_w is None                  # in_dim == 2 fdim: identity shortcut
    assert B.get_block_ops("simple")(4, 32, c).w.shape == (15, 4, 32)
    d = B.get_block_ops("resnetb_deformable_strided")(64, 64, c)                      # N3: offset convolution variables (zeros, as the reference)
    assert tuple(d.conv2_offset_w.shape) == (15, 32, 45) and tuple(d.conv2_offset_b.shape) == (45,) and float(d.conv2_offset_w.abs().sum()) == 0
    bn = b.conv1_bn.bn
    assert bn.eps == 1e-6 and abs(bn.momentum - 0.01) < 1e-12                         # TF momentum 0.99
    c.use_batch_norm = False
    assert tuple(B.get_block_ops("unary")(8, 16, c).bn.offset.shape) == (16,)


def test_oracle_batch_norm_is_tf_training_mode():
    from oracle import kpconv_oracle as K
    g = torch.Generator().manual_seed(1)
    x = torch.randn(500, 7, generator=g, dtype=torch.float64) * 3 + 1
    gamma, beta = torch.rand(7, generator=g, dtype=torch.float64) + 0.5, torch.randn(7, generator=g, dtype=torch.float64)
    y = K.batch_norm_train(x, gamma, beta)
    ref = torch.nn.functional.batch_norm(x, None, None, gamma, beta, True, 0.01, 1e-6)
    assert torch.allclose(y, ref, atol=1e-12)
