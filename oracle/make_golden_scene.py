"""TEST INFRASTRUCTURE — golden fixtures at BASELINE.json sizes: the UNMODIFIED reference (`/root/reference/seggroup/model.py`
via oracle/ref_harness.py) on one synthetic scene, `ins_infer` mode and one training pass (14 label vectors, cluster counts
per level, metrics, loss, 19 parameter gradients, the ReLU margins of the GCN layers, the reference's CPU time).

    python -m oracle.make_golden_scene --points 150000 --seed 11        (build container only: needs /root/reference)

writes tests/golden/seggroup{points/1000}k_s{seed}_{mode}_g4.npz.  150,000 points is the reference's real scene size
(prepare_data.py:29; its export buffers are 150,000 entries, model.py:525).
"""
from __future__ import annotations

import argparse
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
G = 4.0


def mint(seed, n_points, modes=("ins_infer", "train"), write=True):
    scene = synth.make_scene(seed, n_points)
    tree = tempfile.mkdtemp(prefix="sgb_golden_")
    synth.write_scene_tree(tree, [scene])
    info = {}
    for mode in modes:
        t0 = time.time()
        res = ref_harness.run_reference(tree, 0, mode=mode, seed=1, bn_gamma_scale=G, exp_name="golden_%s" % mode)
        dt = time.time() - t0
        out = {"label/" + k: v.astype(np.int32) for k, v in res["labels"].items()}
        for i, o in enumerate(res["out"]):
            out["out/%d" % i] = o.numpy()
        if mode == "train":
            out["loss"] = np.float64(res["loss"])
            for k, v in res["grads"].items():
                if v is not None:
                    out["grad/" + k] = v.numpy()
        out["n_clusters"] = np.array([len(np.unique(c)) for c in res["capture"]["group_nearby_clusters"]], np.int64)
        out["n_segments"] = np.int64(len(scene.seg_offsets) - 1)
        out["relu_margin"] = np.array([res["relu_margin"].get(g, (np.nan, np.nan)) for g in ("gcn_2", "gcn_3")], np.float64)   # (min|Z|, max|Z|)
        out["reference_cpu_seconds"] = np.float64(dt)
        out["reference_cpu_threads"] = np.int64(os.cpu_count() or 1)
        out["scene"] = np.array([seed, n_points], np.int64)
        name = "seggroup%dk_s%d_%s_g4.npz" % (n_points // 1000, seed, mode)
        if write:
            np.savez_compressed(os.path.join(GOLDEN, name), **out)
        info[mode] = dict(seconds=dt, n_clusters=out["n_clusters"].tolist(), relu_margin=out["relu_margin"].tolist())
        print(name, "reference CPU time %.1f s (%.0f points/s)" % (dt, n_points / dt), out["n_clusters"], "relu margin (min|Z|, max|Z|)",
              out["relu_margin"].tolist(), os.path.getsize(os.path.join(GOLDEN, name)) if write else "", flush=True)
    return info


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=150000)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--modes", default="ins_infer,train")
    ap.add_argument("--dry", action="store_true", help="report margins / counts only")
    a = ap.parse_args()
    mint(a.seed, a.points, tuple(a.modes.split(",")), write=not a.dry)
