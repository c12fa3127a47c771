"""Host I/O around the path: the binary CSR cache of the side files SegModel.forward reads must return
exactly what the JSON / .pth parse returns, be rebuilt when a source file changes, and survive a corrupt cache file."""
import os
import time

import numpy as np


def test_scene_cache_round_trip(tmp_path):
    from seggroup_b200 import synth
    from seggroup_b200.model import load_scene_files
    scene = synth.make_scene(3, 6000)
    tree = str(tmp_path / "tree")
    synth.write_scene_tree(tree, [scene])
    root = os.path.join(tree, "dataset", "scannet")
    ref = load_scene_files(scene.name, root)
    assert np.array_equal(ref[2], scene.seg_offsets) and np.array_equal(ref[3], scene.seg_members)
    cache = str(tmp_path / "cache")
    first = load_scene_files(scene.name, root, cache)                    # builds the cache
    cpath = os.path.join(cache, scene.name + ".sgbcache.npz")
    assert os.path.isfile(cpath)
    t_built = os.path.getmtime(cpath)
    second = load_scene_files(scene.name, root, cache)                   # served from the cache
    assert os.path.getmtime(cpath) == t_built
    for a, b, c in zip(ref, first, second):
        assert a.dtype == b.dtype == c.dtype and np.array_equal(a, b) and np.array_equal(a, c)
    # a newer source file invalidates the cache
    seg_json = os.path.join(root, "label", "real", "resampled", scene.name, scene.name + ".seg.json")
    os.utime(seg_json, (time.time() + 5, time.time() + 5))
    load_scene_files(scene.name, root, cache)
    assert os.path.getmtime(cpath) > t_built or os.path.getmtime(cpath) >= os.path.getmtime(seg_json) - 5
    # a corrupt cache file is rebuilt, not trusted
    with open(cpath, "wb") as f:
        f.write(b"not an npz")
    os.utime(cpath, (time.time() + 60, time.time() + 60))
    again = load_scene_files(scene.name, root, cache)
    for a, b in zip(ref, again):
        assert np.array_equal(a, b)
