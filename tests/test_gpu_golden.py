"""CUDA path against the golden vectors minted from the unmodified reference (tests/golden/): pseudo-label ids
bit-exact, loss / metrics / gradients within 1e-4 (gradients 1e-3) relative."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from test_oracle_golden import CASES, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,g", CASES)
def test_cuda_matches_reference_golden(scene8k, mode, g):
    from seggroup_b200 import pipeline
    from seggroup_b200.params import TRAINABLE, init_params
    gold = load_golden(mode, g)
    p = {k: v.cuda() for k, v in init_params(1, g).items()}
    mask = None
    if mode == "train":
        for k in TRAINABLE:
            p[k].requires_grad_(True)
        n_inst = int(gold["out/0"][0, 1])
        torch.manual_seed(1001)
        mask = (F.dropout(torch.ones(n_inst, 128), 0.5, True) != 0).cuda()
    sc = pipeline.SceneDevice.from_host(scene8k)
    with torch.set_grad_enabled(mode == "train"):
        res = pipeline.forward_scene(sc, p, mode=mode, dropout_mask=mask)
    assert res.status == 0
    for k in gold.files:
        if k.startswith("label/"):
            got = res.labels[k[6:]].cpu().numpy()
            assert np.array_equal(got, gold[k]), "%s: %d vertices differ" % (k, (got != gold[k]).sum())
    metrics = [gold["out/%d" % i] for i in range(4 if mode == "train" else 3)]
    if mode == "train":
        assert np.allclose(res.loss_raw.detach().cpu().numpy(), metrics[0], rtol=1e-4)
        metrics = metrics[1:]
        (res.loss_raw[:, 0].sum() / res.loss_raw[:, 1].sum()).backward()
        for k in TRAINABLE:
            if "grad/" + k in gold.files:
                gr = gold["grad/" + k]
                err = np.abs(p[k].grad.cpu().numpy() - gr).max() / (np.abs(gr).max() + 1e-30)
                assert err < 1e-3, (k, err)
    for a, b in zip(res.metrics, metrics):
        assert np.allclose(a.cpu().numpy(), b, atol=1e-6)


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


from test_oracle_golden import BIG  # noqa: E402


@pytest.mark.parametrize("stem", list(BIG))
@pytest.mark.parametrize("mode", ["ins_infer", "train"])
def test_cuda_matches_reference_golden_50k(stem, mode):
    """BASELINE.json sizes — one synthetic ScanNet-shaped scene of 50,000 points / ~350 segments (configs[0]; two scenes) and one
    of 150,000 points / ~1,050 segments (the reference's real scene size): the unmodified reference on CPU (fixture) against
    the CUDA path.  All 14 label vectors bit-exact, cluster counts per level, metrics, loss 1e-4.

    Gradients (training mode) are pinned in two steps, following the measurement in profiles/r02c_diag_*.txt:
    (1) against the oracle evaluated with the kernels' own kNN tie rule (score desc, member position asc): every index list of
        the forward is BIT-IDENTICAL (FPS picks, kNN lists of MLP1/2/3, cluster maps), and all 19 parameter gradients agree
        within 1e-4 relative L2 (measured: <= 1e-5 at 50k, <= 6e-5 at 150k) for a common ReLU active set in the two GCN layers.
        The derivative of ReLU jumps at 0: the scene of seed 7 has ONE pre-activation of gcn_2 at |Z| = 1.0e-6 (max |Z| = 12.7)
        whose sign differs between two fp32 evaluations, and that entry alone moves every gradient upstream of it by 1-3 %.
        The test asserts that the two active sets differ only at entries with |Z| < 1e-5 max|Z|.
    (2) against the torch-CPU fixture (reference tie rule = whatever torch.topk returns among exactly tied candidates, model.py:35):
        classifier head 2e-3 relative L2; below the head the documented tie-rule / kink distance (5e-2)."""
    from test_oracle_golden import big_scene, load_golden_50k
    from seggroup_b200 import pipeline
    from seggroup_b200.params import TRAINABLE, init_params
    gold = load_golden_50k(mode, stem)
    scene = big_scene(stem)
    p = {k: v.cuda() for k, v in init_params(1, 4.0).items()}
    mask = None
    if mode == "train":
        for k in TRAINABLE:
            p[k].requires_grad_(True)
        n_inst = int(gold["out/0"][0, 1])
        torch.manual_seed(1001)
        mask = (F.dropout(torch.ones(n_inst, 128), 0.5, True) != 0).cuda()
    sc = pipeline.SceneDevice.from_host(scene)
    with torch.set_grad_enabled(mode == "train"):
        res = pipeline.forward_scene(sc, p, mode=mode, dropout_mask=mask, keep_aux=(mode == "train"))
    assert res.status == 0
    assert [L.S for L in res.levels][1:4] == list(gold["n_clusters"]), ([L.S for L in res.levels], gold["n_clusters"])
    for k in gold.files:
        if k.startswith("label/"):
            got = res.labels[k[6:]].cpu().numpy()
            assert np.array_equal(got, gold[k]), "%s: %d vertices differ" % (k, (got != gold[k]).sum())
    metrics = [gold["out/%d" % i] for i in range(4 if mode == "train" else 3)]
    if mode == "train":
        assert np.allclose(res.loss_raw.detach().cpu().numpy(), metrics[0], rtol=1e-4)
        metrics = metrics[1:]
        (res.loss_raw[:, 0].sum() / res.loss_raw[:, 1].sum()).backward()
        from oracle import seggroup_oracle as O
        live = res.aux["_live"]
        relu_sets = {t: (live["gcn_" + t].detach() > 0).cpu() for t in ("2", "3")}      # relu(Z) > 0 <=> Z > 0 (the ReLU is fused into the GEMM epilogue)
        params_cpu = O.init_params(1, 4.0)
        ref = O.forward(scene, params_cpu, mode="train", tie="canonical", dropout_mask=mask.cpu(), want_grads=True, relu_masks=relu_sets)
        # (1a) index lists: bit-identical
        assert np.array_equal(res.aux["cloud_idx_1"].cpu().numpy(), ref["cloud_idx_1"]), "FPS picks"
        assert np.array_equal(res.aux["knn_1"].cpu().numpy(), ref["knn_1"].numpy()), "kNN(10) of the 64-point clouds"
        for t in ("2", "3"):
            assert np.array_equal(res.aux["knn_" + t].cpu().numpy(), ref["knn_" + t].numpy()), "kNN(20) of layer " + t
        for Lc, Lr in zip(res.levels, ref["levels"]):
            assert Lc.S == Lr.S and np.array_equal(Lc.seg2cl.cpu().numpy()[:len(Lr.seg2cluster)], Lr.seg2cluster)
        # (1b) the ReLU active sets may differ only within rounding of zero
        for t in ("2", "3"):
            zr = ref["_live"]["Z_" + t].detach()
            flipped = (zr > 0) != relu_sets[t]
            assert int(flipped.sum()) <= 4
            if flipped.any():
                assert float(zr[flipped].abs().max()) < 1e-5 * float(zr.abs().max()), t
        # (1c) all 19 parameter gradients within 1e-4 relative L2 of the canonical-rule oracle
        for k in TRAINABLE:
            gc = ref["grads"][k]
            if gc is not None:
                assert _rel(p[k].grad.cpu().numpy(), gc.numpy()) < 1e-4, ("vs the canonical kNN tie rule", k)
        # (2) the torch-CPU fixture
        for k in TRAINABLE:
            if "grad/" + k in gold.files:
                gr = gold["grad/" + k]
                gg = p[k].grad.cpu().numpy()
                if k.startswith("classifier."):
                    assert _rel(gg, gr) < 2e-3, k        # measured 1e-4 .. 7e-4 (profiles/r02c_diag_*.txt)
                else:
                    assert _rel(gg, gr) < 5e-2, ("vs the torch-CPU fixture", k)
    for a, b in zip(res.metrics, metrics):
        assert np.allclose(a.cpu().numpy(), b, atol=1e-6)
