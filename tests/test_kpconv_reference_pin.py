"""Pins the KPConv restatement (oracle/kpconv_oracle.py: `kpconv_ops`, `kpconv_deform_ops`, `kpconv_deformable`,
`ind_max_pool`, `closest_pool`, `block_forward`) against golden vectors minted by EXECUTING the unmodified reference
(`kpconv/kernels/convolution_ops.py`, `kpconv/models/network_blocks.py`) on the torch-backed TensorFlow stand-in
(oracle/tf_shim.py, oracle/make_golden_kpconv.py).  Runs on CPU.  In the build container (reference tree present) the
reference is also re-executed live and compared with the committed fixtures."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
T64 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64)


def close(a, b, tol=2e-6):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-12)


@pytest.fixture(scope="module")
def setup():
    from oracle import make_golden_kpconv as M
    g = M.geometry()
    return M, g, M.ops_inputs(g), np.load(os.path.join(GOLDEN, "kpconv_ref_ops.npz")), np.load(os.path.join(GOLDEN, "kpconv_ref_blocks.npz"))


def test_kpconv_ops_restatement_matches_reference_vectors(setup):
    from oracle import kpconv_oracle as K
    M, g, x, gold, _ = setup
    for infl, mode in M.OPS_MODES:
        f = T64(x["feats"]).requires_grad_(True); kv = T64(x["kv"]).requires_grad_(True)
        y = K.kpconv_ops(T64(x["q"]), T64(x["s"]), torch.as_tensor(x["idx"]), f, T64(x["kp"]), kv, float(x["extent"]), infl, mode, dtype=torch.float64)
        (y * T64(x["go"])).sum().backward()
        tag = "ops/%s_%s/" % (infl, mode)
        assert close(y.detach(), gold[tag + "out"]), tag
        assert close(f.grad, gold[tag + "dfeats"]) and close(kv.grad, gold[tag + "dkv"]), tag


def test_deformable_restatement_matches_reference_vectors(setup):
    from oracle import kpconv_oracle as K
    M, g, x, gold, _ = setup
    for infl, mode, modulated in [("linear", "sum", False), ("linear", "sum", True), ("gaussian", "sum", False), ("constant", "sum", False),
                                  ("linear", "closest", False)]:
        f = T64(x["feats"]).requires_grad_(True); kv = T64(x["kv"]).requires_grad_(True)
        off = T64(x["offsets"]).requires_grad_(True)
        mod = T64(x["modulations"]).requires_grad_(True) if modulated else None
        y = K.kpconv_deform_ops(T64(x["q"]), T64(x["s"]), torch.as_tensor(x["idx"]), f, T64(x["kp"]), off, mod, kv, float(x["extent"]), infl, mode)
        (y * T64(x["go"])).sum().backward()
        tag = "deform_ops/%s_%s_%d/" % (infl, mode, int(modulated))
        assert close(y.detach(), gold[tag + "out"]), tag
        assert close(f.grad, gold[tag + "dfeats"]) and close(kv.grad, gold[tag + "dkv"]), tag
        assert close(off.grad if off.grad is not None else torch.zeros_like(off), gold[tag + "doffsets"]), tag
        if modulated:
            assert close(mod.grad, gold[tag + "dmod"]), tag
    for modulated in (False, True):
        f = T64(x["feats"]).requires_grad_(True); kv = T64(x["kv"]).requires_grad_(True)
        kv0 = T64(x["kv0m" if modulated else "kv0"]).requires_grad_(True); b0 = T64(x["b0m" if modulated else "b0"]).requires_grad_(True)
        y = K.kpconv_deformable(T64(x["q"]), T64(x["s"]), torch.as_tensor(x["idx"]), f, T64(x["kp"]), kv, kv0, b0, float(x["extent"]), "linear", "sum", modulated)
        (y * T64(x["go"])).sum().backward()
        tag = "deformable/%d/" % int(modulated)
        assert close(y.detach(), gold[tag + "out"]), tag
        for a, k in ((f.grad, "dfeats"), (kv.grad, "dkv"), (kv0.grad, "dkv0"), (b0.grad, "db0")):
            assert close(a, gold[tag + k]), (tag, k)


def test_index_pooling_restatement_matches_reference_vectors(setup):
    from oracle import kpconv_oracle as K
    M, g, x, gold, _ = setup
    for name, fn, idx, src in (("ind_max_pool", K.ind_max_pool, x["idx"], x["feats"]), ("closest_pool", K.closest_pool, g["up0"], x["go"][:, :16])):
        f = T64(src).requires_grad_(True)
        y = fn(f, torch.as_tensor(idx))
        (y * T64(gold[name + "/go"])).sum().backward()
        assert close(y.detach(), gold[name + "/out"]) and close(f.grad, gold[name + "/dx"]), name


def _restatement_params(name, V):
    """reference scoped variable names -> the names oracle.kpconv_oracle.block_forward / seggroup_b200.kpconv_blocks use"""
    P = {}
    for k, v in V.items():
        sc, _, leaf = k.rpartition("/")
        sc = sc.replace("/batch_normalization", "").replace("batch_normalization", "")
        if leaf == "weights":
            P[(sc + "_w") if sc else "w"] = v
        elif leaf in ("gamma", "beta"):
            P[((sc + "_bn") if sc else "bn") + ".bn." + ("weight" if leaf == "gamma" else "bias")] = v
        elif leaf == "offset_conv_weights":
            P[(sc + "_" if sc else "") + "offset_w"] = v
        elif leaf == "offset_conv_bias":
            P[(sc + "_" if sc else "") + "offset_b"] = v
    return P


def test_blocks_restatement_matches_reference_vectors(setup):
    from oracle import kpconv_oracle as K
    M, g, x, _, gold = setup
    cfg = M.config(g["kp_unit"])
    inputs = {"points": [T64(g["p0"]), T64(g["p1"])], "neighbors": [torch.as_tensor(g["nb0"]), torch.as_tensor(g["nb1"])],
              "pools": [torch.as_tensor(g["pool0"])], "upsamples": [torch.as_tensor(g["up0"])]}
    for name in M.BLOCKS:
        li, fdim, radius, feats, V = M.block_case(name, g)
        Vt = {k: T64(v).requires_grad_(True) for k, v in V.items()}
        f = T64(feats).requires_grad_(True)
        y = K.block_forward(name, _restatement_params(name, Vt), li, inputs, f, radius, cfg)
        tag = "block/%s/" % name
        (y * T64(gold[tag + "go"])).sum().backward()
        assert close(y.detach(), gold[tag + "out"]), name
        assert close(f.grad, gold[tag + "dfeats"], 5e-6), name
        for k, v in Vt.items():
            assert v.grad is not None, (name, k)
            assert close(v.grad, gold[tag + "d/" + k], 5e-6), (name, k)


def test_reference_under_stand_in_reproduces_fixtures_when_present(setup):
    """Build container only: re-execute the unmodified reference and compare with the committed fixtures."""
    from oracle import tf_shim
    if not os.path.isdir(tf_shim.REFERENCE_KPCONV):
        pytest.skip("reference tree not present on this machine")
    M, g, x, gold_ops, gold_blocks = setup
    ops = M.run_reference_ops(g, x)
    for k in gold_ops.files:
        assert close(ops[k], gold_ops[k], 1e-6), k
    blocks = M.run_reference_blocks(g)
    for k in gold_blocks.files:
        assert close(blocks[k], gold_blocks[k], 1e-6), k
