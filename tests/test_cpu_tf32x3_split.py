"""The operand convention of the tcgen05 kernels (seggroup_b200/csrc/tc_common.cuh), restated in numpy: x = hi + lo with hi = x with the low 13
mantissa bits cleared and lo = tf32(x - hi); a product is hi*hi + lo*hi + hi*lo.  Checks the accuracy claim the parity tolerances rest on
(~21 mantissa bits per product) on the first EdgeConv layer as ec2_tc1_kernel evaluates it: E = (x_j - x_i, x_i, 1, 0...) times
W1s = (scale1 W1 | folded BatchNorm-1 bias | 0), K = 24 (reference: seggroup/model.py:121-133, conv1 -> BN -> LeakyReLU)."""
import numpy as np


def tf32_hi(x):
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    hi = tf32_hi(x)
    lo = tf32_hi(np.asarray(x, np.float32) - hi)
    return hi, lo


def tf32x3(a, b):
    """a [M,K] · b [N,K]^T with the three-term split, products exact (fp64), accumulation in fp32 like TMEM."""
    ah, al = split(a)
    bh, bl = split(b)
    acc = np.zeros((a.shape[0], b.shape[0]), np.float32)
    for k in range(a.shape[1]):                       # K order of the MMA instructions is irrelevant at this tolerance
        for x, y in ((ah, bh), (al, bh), (ah, bl)):
            acc = (acc.astype(np.float64) + np.outer(x[:, k].astype(np.float64), y[:, k].astype(np.float64))).astype(np.float32)
    return acc


def test_split_is_exact_up_to_22_bits():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-8, 8, 100000))).astype(np.float32)
    hi, lo = split(x)
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)
    rel = np.abs((hi.astype(np.float64) + lo) - x) / np.abs(x)
    assert rel.max() < 2.0 ** -21


def test_first_layer_as_one_gemm_with_a_bias_column():
    rng = np.random.default_rng(1)
    n, cin, cout = 4000, 18, 64
    xi = rng.uniform(-1, 1, (n, 9)).astype(np.float32)
    xj = rng.uniform(-1, 1, (n, 9)).astype(np.float32)
    valid = (rng.uniform(size=n) > 0.05).astype(np.float32)          # rows outside a CTA's range are all-zero, bias column included
    W1 = (rng.standard_normal((cout, cin)) * 0.3).astype(np.float32)
    mean, invstd = rng.standard_normal(cout).astype(np.float32), rng.uniform(0.5, 2.0, cout).astype(np.float32)
    gamma, beta = rng.uniform(0.5, 1.5, cout).astype(np.float32), (0.2 * rng.standard_normal(cout)).astype(np.float32)
    scale = gamma * invstd
    E = np.zeros((n, 24), np.float32)
    E[:, :9], E[:, 9:18], E[:, 18] = xj - xi, xi, 1.0
    E *= valid[:, None]
    W1s = np.zeros((cout, 24), np.float32)
    W1s[:, :18] = W1 * scale[:, None]
    W1s[:, 18] = beta - scale * mean
    got = tf32x3(E, W1s)
    e64 = np.concatenate([xj - xi, xi], 1).astype(np.float64)
    ref = ((e64 @ W1.astype(np.float64).T - mean) * invstd * gamma + beta) * valid[:, None]      # BN(conv1(e)), 0 for an invalid row
    scale_ref = np.abs(e64) @ np.abs(W1s[:, :18].astype(np.float64)).T + np.abs(W1s[:, 18])
    assert np.max(np.abs(got - ref) / scale_ref) < 4e-6            # fp32-level: the folded weights are themselves rounded to fp32
    assert np.all(got[valid == 0] == 0.0)                            # -> h = lrelu(0) = 0: an out-of-range edge contributes nothing
    plain = (tf32_hi(E).astype(np.float64) @ tf32_hi(W1s).astype(np.float64).T)
    assert np.max(np.abs(plain - ref) / scale_ref) > 1e-4           # a single TF32 pass is two orders of magnitude worse
