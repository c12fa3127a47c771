"""TEST INFRASTRUCTURE — mints the golden fixtures under tests/golden/ by executing the UNMODIFIED reference
(`/root/reference/seggroup/model.py`, via oracle/ref_harness.py) on seeded synthetic scenes.

    python -m oracle.make_golden          (build container only: needs /root/reference)

The reference ships no golden vectors for this path, so these files are what pins the oracle
(tests/test_oracle_golden.py) and, through it and directly, the CUDA path (tests/test_gpu_golden.py).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from seggroup_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENE_ARGS = dict(seed=5, n_points=8000, n_small_segs=2)
CASES = [("train", None), ("train", 2.0), ("train", 4.0), ("sem_infer", 4.0), ("ins_infer", 4.0)]


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    scene = synth.make_scene(SCENE_ARGS["seed"], SCENE_ARGS["n_points"], n_small_segs=SCENE_ARGS["n_small_segs"])
    tree = tempfile.mkdtemp(prefix="sgb_golden_")
    synth.write_scene_tree(tree, [scene])
    for mode, g in CASES:
        # classifier dropout p=0.5 draws from torch's CPU RNG: seed fixed inside run_reference (seed+1000)
        res = ref_harness.run_reference(tree, 0, mode=mode, seed=1, bn_gamma_scale=g, exp_name="golden_%s_%s" % (mode, g))
        out = {"label/" + k: v.astype(np.int32) for k, v in res["labels"].items()}
        for i, o in enumerate(res["out"]):
            out["out/%d" % i] = o.numpy()
        if mode == "train":
            out["loss"] = np.float64(res["loss"])
            if g == 4.0:
                for k, v in res["grads"].items():
                    if v is not None:
                        out["grad/" + k] = v.numpy()
        cap = res["capture"]
        out["n_clusters"] = np.array([len(np.unique(c)) for c in cap["group_nearby_clusters"]], np.int64)
        name = "seggroup_%s_g%s.npz" % (mode, "none" if g is None else ("%g" % g))
        np.savez_compressed(os.path.join(GOLDEN, name), **out)
        print(name, {k: v.shape for k, v in out.items() if not k.startswith("label/") and not k.startswith("grad/")}, out["n_clusters"])
    with open(os.path.join(GOLDEN, "README.md"), "w") as f:
        f.write("Golden vectors minted by `python -m oracle.make_golden` from the unmodified reference\n"
                "(`/root/reference/seggroup/model.py`, torch %s CPU) on `synth.make_scene(%r)`;\n"
                "weights = `torch.manual_seed(1)` default init with `mlp_1.bn1.weight *= g`; dropout RNG seed 1001.\n"
                % (torch.__version__, SCENE_ARGS))


if __name__ == "__main__":
    main()
