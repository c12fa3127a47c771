"""TEST INFRASTRUCTURE — golden fixture at the size of BASELINE.json configs[0] (one scene, 50,000 points, ~300 segments):
the UNMODIFIED reference (`/root/reference/seggroup/model.py` via oracle/ref_harness.py) in `ins_infer` mode and one
training pass (loss, metrics, 19 parameter gradients).

    python -m oracle.make_golden_50k          (build container only: needs /root/reference; ~2 min of CPU)
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from seggroup_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENE_ARGS = dict(seed=7, n_points=50000)
G = 4.0


def main():
    scene = synth.make_scene(SCENE_ARGS["seed"], SCENE_ARGS["n_points"])
    tree = tempfile.mkdtemp(prefix="sgb_golden50k_")
    synth.write_scene_tree(tree, [scene])
    for mode in ("ins_infer", "train"):
        t0 = time.time()
        res = ref_harness.run_reference(tree, 0, mode=mode, seed=1, bn_gamma_scale=G, exp_name="golden50k_%s" % mode)
        out = {"label/" + k: v.astype(np.int32) for k, v in res["labels"].items()}
        for i, o in enumerate(res["out"]):
            out["out/%d" % i] = o.numpy()
        if mode == "train":
            out["loss"] = np.float64(res["loss"])
            for k, v in res["grads"].items():
                if v is not None:
                    out["grad/" + k] = v.numpy()
        out["n_clusters"] = np.array([len(np.unique(c)) for c in res["capture"]["group_nearby_clusters"]], np.int64)
        out["n_segments"] = np.int64(scene.n_segments if hasattr(scene, "n_segments") else len(scene.seg_offsets) - 1)
        name = "seggroup50k_%s_g4.npz" % mode
        np.savez_compressed(os.path.join(GOLDEN, name), **out)
        print(name, "reference CPU time %.1f s" % (time.time() - t0), out["n_clusters"], os.path.getsize(os.path.join(GOLDEN, name)))


if __name__ == "__main__":
    main()
