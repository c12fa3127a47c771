"""Pins the oracle (oracle/seggroup_oracle.py) against golden vectors minted from the unmodified reference
(tests/golden/, see oracle/make_golden.py).  Runs on CPU."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("train", None), ("train", 2.0), ("train", 4.0), ("sem_infer", 4.0), ("ins_infer", 4.0)]


def load_golden(mode, g):
    return np.load(os.path.join(GOLDEN, "seggroup_%s_g%s.npz" % (mode, "none" if g is None else ("%g" % g))))


@pytest.fixture(scope="module")
def scene():
    from seggroup_b200 import synth
    return synth.make_scene(5, 8000, n_small_segs=2)


@pytest.mark.parametrize("mode,g", CASES)
@pytest.mark.parametrize("tie", ["torch", "canonical"])
def test_oracle_matches_reference_golden(scene, mode, g, tie):
    from oracle import seggroup_oracle as O
    gold = load_golden(mode, g)
    params = O.init_params(1, g)
    torch.manual_seed(1001)                     # dropout RNG position of the reference run (ref_harness.run_reference)
    out = O.forward(scene, params, mode=mode, tie=tie, want_grads=(mode == "train" and g == 4.0))
    for k in gold.files:
        if k.startswith("label/"):
            assert np.array_equal(out["labels"][k[6:]], gold[k]), k
    metrics = [gold["out/%d" % i] for i in range(4 if mode == "train" else 3)]
    if mode == "train":
        assert np.allclose(out["loss_raw"], metrics[0], rtol=1e-5)
        metrics = metrics[1:]
    for a, b in zip(out["metrics"], metrics):
        assert np.allclose(a, b, atol=1e-6)
    if "grads" in out:
        for k in O.TRAINABLE:
            if "grad/" + k in gold.files:
                gr = gold["grad/" + k]
                assert np.abs(out["grads"][k].numpy() - gr).max() <= 1e-5 * np.abs(gr).max() + 1e-9, k


def test_reference_still_agrees_when_present(scene, tmp_path):
    """In the build container (where /root/reference exists) re-run the reference itself against one golden file."""
    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("reference tree not present on this machine")
    from seggroup_b200 import synth
    synth.write_scene_tree(str(tmp_path), [scene])
    res = ref_harness.run_reference(str(tmp_path), 0, mode="sem_infer", seed=1, bn_gamma_scale=4.0, exp_name="chk")
    gold = load_golden("sem_infer", 4.0)
    for k, v in res["labels"].items():
        assert np.array_equal(v, gold["label/" + k]), k


# ---- BASELINE.json sizes: one scene of 50,000 points / ~350 segments (configs[0]) and of 150,000 points / ~1,100 segments
# (the reference's real scene size, prepare_data.py:29), minted by oracle/make_golden_50k.py / oracle/make_golden_scene.py
# fixture stem -> (scene seed, points)
BIG = {"seggroup50k_%s_g4": (7, 50000), "seggroup50k_s9_%s_g4": (9, 50000), "seggroup150k_s12_%s_g4": (12, 150000)}


def load_golden_50k(mode, stem="seggroup50k_%s_g4"):
    return np.load(os.path.join(GOLDEN, (stem % mode) + ".npz"))


_scenes = {}


def big_scene(stem):
    from seggroup_b200 import synth
    if stem not in _scenes:
        _scenes[stem] = synth.make_scene(*BIG[stem])
    return _scenes[stem]


@pytest.mark.parametrize("stem", list(BIG))
@pytest.mark.parametrize("mode", ["ins_infer", "train"])
def test_oracle_matches_reference_golden_50k(stem, mode):
    from oracle import seggroup_oracle as O
    if stem.startswith("seggroup150k") and mode == "ins_infer":
        pytest.skip("150k: the training fixture carries the same 14 label vectors")
    gold = load_golden_50k(mode, stem)
    params = O.init_params(1, 4.0)
    torch.manual_seed(1001)
    out = O.forward(big_scene(stem), params, mode=mode, tie="torch", want_grads=(mode == "train"))
    for k in gold.files:
        if k.startswith("label/"):
            assert np.array_equal(out["labels"][k[6:]], gold[k]), k
    metrics = [gold["out/%d" % i] for i in range(4 if mode == "train" else 3)]
    if mode == "train":
        assert np.allclose(out["loss_raw"], metrics[0], rtol=1e-5)
        metrics = metrics[1:]
        for k in O.TRAINABLE:
            if "grad/" + k in gold.files:
                gr = gold["grad/" + k]
                assert np.abs(out["grads"][k].numpy() - gr).max() <= 1e-5 * np.abs(gr).max() + 1e-9, k
    for a, b in zip(out["metrics"], metrics):
        assert np.allclose(a, b, atol=1e-6)
