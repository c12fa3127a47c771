"""TEST INFRASTRUCTURE — golden vectors for the KPConv operator set minted by EXECUTING the unmodified reference files
`kpconv/kernels/convolution_ops.py` and `kpconv/models/network_blocks.py` on the torch-backed TensorFlow stand-in of
oracle/tf_shim.py (TensorFlow itself is not installable in the build container).  The reference code decides every gather,
formula, transpose and matmul; the stand-in only supplies the tf.* primitives, evaluated in float64.

    python -m oracle.make_golden_kpconv            (build container only: needs /root/reference)

writes tests/golden/kpconv_ref_ops.npz     KPConv_ops (6 influence / aggregation modes), KPConv_deform_ops (+ modulations),
                                           KPConv_deformable (offset convolution + deformed convolution), ind_max_pool, closest_pool
       tests/golden/kpconv_ref_blocks.npz  network_blocks.py blocks in training mode: unary, simple, simple_strided, resnetb,
                                           resnetb_strided, max_pool, nearest_upsample, resnetb_deformable, resnetb_deformable_strided
each with the inputs, the variables, the outputs and the gradients (input features + every trainable variable) for a
fixed upstream gradient.  `cases(...)` is shared with tests/test_kpconv_reference_pin.py, which replays the same cases live
when the reference tree is present.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import kpconv_oracle as K  # noqa: E402
from oracle import tf_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
OPS_MODES = [("linear", "sum"), ("linear", "closest"), ("constant", "sum"), ("constant", "closest"), ("gaussian", "sum"), ("gaussian", "closest")]
BLOCKS = ["unary", "simple", "simple_strided", "resnetb", "resnetb_strided", "max_pool", "nearest_upsample",
          "resnetb_deformable", "resnetb_deformable_strided"]


def geometry(seed=3, n=420):
    """A small two-level geometry: points on three faces of a box, grid-subsampled twice, radius neighbours from the oracle
    (itself pinned by the compiled reference cores).  All float32 / int32 numpy."""
    rng = np.random.default_rng(seed)
    face = rng.integers(0, 3, n)
    uv = rng.random((n, 2)).astype(np.float32) * 0.6
    pts = np.zeros((n, 3), np.float32)
    for f in range(3):
        m = face == f
        pts[m] = np.insert(uv[m], f, 0.0, axis=1)
    pts += rng.normal(0, 0.002, pts.shape).astype(np.float32)
    lens = np.array([n], np.int32)
    p0, l0 = K.batch_grid_subsampling(pts, lens, 0.04)
    p1, l1 = K.batch_grid_subsampling(p0, l0, 0.08)
    r0, r1 = 0.10, 0.20
    g = dict(p0=p0, p1=p1, r0=np.float32(r0), r1=np.float32(r1),
             nb0=K.batch_neighbors(p0, p0, l0, l0, r0), nb1=K.batch_neighbors(p1, p1, l1, l1, r1),
             pool0=K.batch_neighbors(p1, p0, l1, l0, r0), up0=K.batch_neighbors(p0, p1, l0, l1, r1)[:, :1].copy())
    kp = rng.random((15, 3)) * 2 - 1
    kp = kp / np.linalg.norm(kp, axis=1, keepdims=True) * rng.random((15, 1)) ** (1 / 3)
    kp[0] = 0
    g["kp_unit"] = kp.astype(np.float32)                       # unit-ball disposition; a layer uses kp_unit * 1.5 * extent
    return g


def config(K_points):
    """The attributes network_blocks.py reads (training_Scannet.py:78-110 values)."""
    return SimpleNamespace(KP_extent=1.0, density_parameter=5.0, KP_influence="linear", convolution_mode="sum", num_kernel_points=15,
                           use_batch_norm=True, batch_norm_momentum=0.99, fixed_kernel_points="center", modulated=False,
                           K_points=torch.as_tensor(K_points))


def ops_inputs(g, seed=5, cin=16, cout=24):
    rng = np.random.default_rng(seed)
    n0 = len(g["p0"])
    ext = 0.05
    return dict(q=g["p1"], s=g["p0"], idx=g["pool0"], feats=rng.standard_normal((n0, cin)).astype(np.float32),
                kp=(g["kp_unit"] * 1.5 * ext).astype(np.float32), kv=(rng.standard_normal((15, cin, cout)) / np.sqrt(15 * cin)).astype(np.float32),
                extent=np.float32(ext), go=rng.standard_normal((len(g["p1"]), cout)).astype(np.float32),
                offsets=(rng.standard_normal((len(g["p1"]), 15, 3)) * 0.3 * ext).astype(np.float32),
                modulations=(2 / (1 + np.exp(-rng.standard_normal((len(g["p1"]), 15))))).astype(np.float32),
                kv0=(rng.standard_normal((15, cin, 45)) * 0.2 / np.sqrt(15 * cin)).astype(np.float32), b0=(rng.standard_normal(45) * 0.05).astype(np.float32),
                kv0m=(rng.standard_normal((15, cin, 60)) * 0.2 / np.sqrt(15 * cin)).astype(np.float32), b0m=(rng.standard_normal(60) * 0.05).astype(np.float32))


def run_reference_ops(g, x):
    """-> dict name -> float64 numpy (outputs and gradients) from the unmodified reference under the stand-in."""
    out = {}
    w = tf_shim.wrap
    with tf_shim.installed(dtype=torch.float64):
        co = tf_shim.import_reference("kernels.convolution_ops")
        nbk = tf_shim.import_reference("models.network_blocks")
        go = w(x["go"])
        for infl, mode in OPS_MODES:
            f = w(x["feats"]).requires_grad_(True); kv = w(x["kv"]).requires_grad_(True)
            y = co.KPConv_ops(w(x["q"]), w(x["s"]), w(x["idx"]), f, w(x["kp"]), kv, float(x["extent"]), infl, mode)
            (y * go).sum().backward()
            tag = "ops/%s_%s/" % (infl, mode)
            out[tag + "out"], out[tag + "dfeats"], out[tag + "dkv"] = y.detach().numpy(), f.grad.numpy(), kv.grad.numpy()
        # deformable operator with explicit offsets (and modulations), convolution_ops.py:371-493
        for infl, mode, modulated in [("linear", "sum", False), ("linear", "sum", True), ("gaussian", "sum", False), ("constant", "sum", False),
                                      ("linear", "closest", False)]:
            f = w(x["feats"]).requires_grad_(True); kv = w(x["kv"]).requires_grad_(True)
            off = w(x["offsets"]).requires_grad_(True)
            mod = w(x["modulations"]).requires_grad_(True) if modulated else None
            y = co.KPConv_deform_ops(w(x["q"]), w(x["s"]), w(x["idx"]), f, w(x["kp"]), off, mod, kv, float(x["extent"]), infl, mode)
            (y * go).sum().backward()
            tag = "deform_ops/%s_%s_%d/" % (infl, mode, int(modulated))
            out[tag + "out"], out[tag + "dfeats"], out[tag + "dkv"] = y.detach().numpy(), f.grad.numpy(), kv.grad.numpy()
            # 'constant' influence is a step function of the distance: no gradient reaches the offsets
            out[tag + "doffsets"] = off.grad.numpy() if off.grad is not None else np.zeros(tuple(off.shape))
            if modulated:
                out[tag + "dmod"] = mod.grad.numpy()
        # KPConv_deformable end to end (offset convolution K_values0 / b0 -> offsets * extent -> deformed convolution), :252-368.
        # The kernel disposition generator (random, unseeded: kernel_points.py:182-278) is replaced by the given K_points.
        for modulated in (False, True):
            co.create_kernel_points = lambda radius, num_kpoints, num_kernels, dimension, fixed: (x["kp"].astype(np.float64) / (1.5 * float(x["extent"])) * radius)[None]
            f = w(x["feats"]).requires_grad_(True); kv = w(x["kv"]).requires_grad_(True)
            tf_shim.reset_variables()
            with tf_shim.presets({"offset_conv_weights": x["kv0m" if modulated else "kv0"], "offset_conv_bias": x["b0m" if modulated else "b0"]}):
                y = co.KPConv_deformable(w(x["q"]), w(x["s"]), w(x["idx"]), f, kv, fixed="center", KP_extent=float(x["extent"]), KP_influence="linear",
                                         aggregation_mode="sum", modulated=modulated)
            (y * go).sum().backward()
            v = dict(tf_shim.variables())
            tag = "deformable/%d/" % int(modulated)
            out[tag + "out"], out[tag + "dfeats"], out[tag + "dkv"] = y.detach().numpy(), f.grad.numpy(), kv.grad.numpy()
            out[tag + "dkv0"], out[tag + "db0"] = v["offset_conv_weights"].grad.numpy(), v["offset_conv_bias"].grad.numpy()
        # index pooling, network_blocks.py:49-81
        for name, fn, idx in (("ind_max_pool", nbk.ind_max_pool, x["idx"]), ("closest_pool", nbk.closest_pool, g["up0"])):
            src = x["feats"] if name == "ind_max_pool" else x["go"][:, :16]
            f = w(src).requires_grad_(True)
            y = fn(f, w(idx))
            gg = w(np.random.default_rng(9).standard_normal(tuple(y.shape)))
            (y * gg).sum().backward()
            out[name + "/out"], out[name + "/dx"], out[name + "/go"] = y.detach().numpy(), f.grad.numpy(), gg.numpy()
    return out


# ---- blocks ------------------------------------------------------------------------------------------------------------
def block_case(name, g, seed=11):
    """-> (layer_ind, fdim, radius, features, variables {reference scoped name: value}) for one block."""
    rng = np.random.default_rng(seed + BLOCKS.index(name))
    n0, n1 = len(g["p0"]), len(g["p1"])
    cin, fdim = 16, 16
    r0 = float(g["r0"])

    def wv(*shape):                                            # weight_variable values (any values do: they are presets)
        return (np.round(rng.standard_normal(shape) * np.sqrt(2 / shape[-1]) * 1000) / 1000).astype(np.float32)

    def bn(scope, c):
        pre = scope + "/" if scope else ""
        return {pre + "batch_normalization/gamma": (0.5 + rng.random(c)).astype(np.float32), pre + "batch_normalization/beta": (0.2 * rng.standard_normal(c)).astype(np.float32)}
    feats = rng.standard_normal((n0, cin)).astype(np.float32)
    V = {}
    if name == "unary":
        V.update({"weights": wv(cin, fdim)}); V.update(bn("", fdim))
    elif name in ("simple", "simple_strided"):
        V.update({"weights": wv(15, cin, fdim)}); V.update(bn("", fdim))
    elif name.startswith("resnetb"):
        V.update({"conv1/weights": wv(cin, fdim // 2), "conv2/weights": wv(15, fdim // 2, fdim // 2), "conv3/weights": wv(fdim // 2, 2 * fdim),
                  "shortcut/weights": wv(cin, 2 * fdim)})
        for sc, c in (("conv1", fdim // 2), ("conv2", fdim // 2), ("conv3", 2 * fdim), ("shortcut", 2 * fdim)):
            V.update(bn(sc, c))
    if "deformable" in name:                                   # offset convolution of KPConv_deformable (zeros in the reference: no deformation)
        pre, c = "conv2/", fdim // 2
        V[pre + "offset_conv_weights"] = (rng.standard_normal((15, c, 45)) * 0.3 / np.sqrt(15 * c)).astype(np.float32)
        V[pre + "offset_conv_bias"] = (rng.standard_normal(45) * 0.05).astype(np.float32)
    if name == "nearest_upsample":
        feats = rng.standard_normal((n1, cin)).astype(np.float32)
        return 1, fdim, float(g["r1"]), feats, V
    return 0, fdim, r0, feats, V


def reference_inputs(g):
    w = tf_shim.wrap
    return {"points": [w(g["p0"]), w(g["p1"])], "neighbors": [w(g["nb0"]), w(g["nb1"])], "pools": [w(g["pool0"])], "upsamples": [w(g["up0"])]}


def run_reference_blocks(g):
    out = {}
    cfg = config(g["kp_unit"])
    with tf_shim.installed(dtype=torch.float64):
        co = tf_shim.import_reference("kernels.convolution_ops")
        nbk = tf_shim.import_reference("models.network_blocks")
        # K_points are an explicit input: the reference's generator is an unseeded random optimisation (kernel_points.py:182-278)
        co.create_kernel_points = lambda radius, num_kpoints, num_kernels, dimension, fixed: (g["kp_unit"].astype(np.float64) * radius)[None]
        for name in BLOCKS:
            li, fdim, radius, feats, V = block_case(name, g)
            tf_shim.reset_variables()
            f = tf_shim.wrap(feats).requires_grad_(True)
            with tf_shim.presets(V):
                y = nbk.get_block_ops(name)(li, reference_inputs(g), f, radius, fdim, cfg, True)
            go = tf_shim.wrap(np.random.default_rng(21).standard_normal(tuple(y.shape)))
            (y * go).sum().backward()
            tag = "block/%s/" % name
            out[tag + "out"], out[tag + "dfeats"], out[tag + "go"] = y.detach().numpy(), f.grad.numpy(), go.numpy()
            created = dict(tf_shim.variables())
            missing = set(V) - set(created)
            assert not missing, (name, missing, list(created))
            for k in V:
                out[tag + "d/" + k] = created[k].grad.numpy()
    return out


def main():
    g = geometry()
    x = ops_inputs(g)
    ops = run_reference_ops(g, x)
    # stored as float32 (the values are float64 evaluations; 6e-8 relative storage rounding against comparison bars of 1e-6 .. 1e-4)
    f32 = lambda d: {k: v.astype(np.float32) for k, v in d.items()}
    np.savez_compressed(os.path.join(GOLDEN, "kpconv_ref_ops.npz"), **f32(ops))
    blocks = run_reference_blocks(g)
    np.savez_compressed(os.path.join(GOLDEN, "kpconv_ref_blocks.npz"), **f32(blocks))
    for fn in ("kpconv_ref_ops.npz", "kpconv_ref_blocks.npz"):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)), "bytes")
    print("geometry: n0 %d, n1 %d, W nb0 %d, pool0 %d" % (len(g["p0"]), len(g["p1"]), g["nb0"].shape[1], g["pool0"].shape[1]))


if __name__ == "__main__":
    main()
