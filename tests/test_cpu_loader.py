"""the shard loader (seggroup_b200/loader.py) reads the reference's on-disk formats (data.py:28-38, model.py:696-724,
610-614) through the binary CSR cache and collates block-diagonal batches.  Host logic only (no kernel call)."""
import os

import numpy as np
import torch


def test_shard_loader_collates_block_diagonal_batches(tmp_path):
    from seggroup_b200 import synth
    from seggroup_b200.loader import SceneShardLoader
    scenes = [synth.make_scene(51 + i, 6000 + 500 * i, name="ld_%d" % i) for i in range(5)]
    synth.write_scene_tree(str(tmp_path), scenes)
    root = os.path.join(str(tmp_path), "dataset", "scannet")
    names = open(os.path.join(root, "scannetv2_train.txt")).readlines()
    cache = os.path.join(str(tmp_path), "cache")
    seen = []
    for rank in range(2):
        ld = SceneShardLoader(names, data_root=root, rank=rank, world=2, batch_size=2, cache_dir=cache, prefetch=1)
        assert len(ld) == (2 if rank == 0 else 1)
        for hb in ld:
            seen += hb.index
            off = 0
            for b, i in enumerate(hb.index):
                s = scenes[i]
                p0, p1 = hb.pt_off[b], hb.pt_off[b + 1]
                assert p1 - p0 == s.n_points and p0 == off
                assert np.array_equal(hb.data[p0:p1].numpy(), s.data)
                assert np.array_equal(hb.weak_label[p0:p1].numpy(), s.weak_label.astype(np.int32))
                assert np.array_equal(hb.seg_members[p0:p1].numpy() - p0, s.seg_members)
                g0, g1 = hb.seg_cnt_off[b], hb.seg_cnt_off[b + 1]
                assert np.array_equal(hb.seg_off[g0:g1 + 1].numpy() - p0, s.seg_offsets)
                r0, r1 = hb.raw_off[b], hb.raw_off[b + 1]
                assert np.array_equal(hb.unmap[r0:r1].numpy() - p0, s.unmap)
                assert np.array_equal(hb.real_label[r0:r1].numpy(), s.real_label)
                off = p1
            adj = hb.adj0.numpy()
            assert (np.diff(adj[:, 0]) >= 0).all()                                     # still lexicographic over the batch
            expect = np.concatenate([scenes[i].adj + hb.pt_off[b] for b, i in enumerate(hb.index)])
            assert np.array_equal(adj, expect.astype(np.int32))
    assert sorted(seen) == [0, 1, 2, 3, 4]
    assert len(os.listdir(cache)) == 5                                                 # one binary CSR file per scene
    # second pass is served from the cache (seg.json no longer needed)
    os.remove(os.path.join(root, "label", "real", "resampled", "ld_0", "ld_0.seg.json"))
    open(os.path.join(root, "label", "real", "resampled", "ld_0", "ld_0.seg.json"), "w").write("[]")
    os.utime(os.path.join(root, "label", "real", "resampled", "ld_0", "ld_0.seg.json"), (0, 0))   # older than the cache
    hb = next(iter(SceneShardLoader(names, data_root=root, batch_size=1, cache_dir=cache)))
    assert hb.seg_off.numel() == scenes[0].n_segments + 1


def test_loader_surfaces_missing_files(tmp_path):
    import pytest
    from seggroup_b200.loader import SceneShardLoader
    with pytest.raises(Exception):
        list(SceneShardLoader(["nope\n"], data_root=str(tmp_path), batch_size=1))


def test_scene_pack_round_trip_and_rebuild(tmp_path):
    """seggroup_b200/scene_pack.py: a pack returns exactly what the reference's files parse to (dtypes as uploaded), is used while
    it is newer than its sources, rebuilt when a source changes, and never trusted when it does not parse."""
    import time
    from seggroup_b200 import scene_pack, synth
    sc = synth.make_scene(77, 7000, name="pk_0")
    synth.write_scene_tree(str(tmp_path), [sc])
    root = os.path.join(str(tmp_path), "dataset", "scannet")
    cache = os.path.join(str(tmp_path), "packs")
    ref = scene_pack.load_scene("pk_0", root, "manual", None)                 # parsed from the reference's files
    first = scene_pack.load_scene("pk_0", root, "manual", cache)              # parsed + packed
    p = scene_pack.pack_path(cache, "pk_0", "manual")
    assert os.path.isfile(p)
    t_built = os.path.getmtime(p)
    got = scene_pack.load_scene("pk_0", root, "manual", cache)                # served from the pack
    assert os.path.getmtime(p) == t_built
    want = dict(data=(np.float32, sc.data), weak=(np.int32, sc.weak_label), seg_off=(np.int32, sc.seg_offsets), seg_members=(np.int32, sc.seg_members),
                adj=(np.int32, sc.adj), unmap=(np.int64, sc.unmap), real=(np.int64, sc.real_label))
    for k, (dt, val) in want.items():
        for d in (ref, first, got):
            assert d[k].dtype == dt and np.array_equal(d[k], val), k
    src = scene_pack.source_paths("pk_0", root, "manual")[0]
    os.utime(src, (time.time() + 5, time.time() + 5))                         # newer source -> rebuilt
    scene_pack.load_scene("pk_0", root, "manual", cache)
    assert os.path.getmtime(p) >= os.path.getmtime(src) - 5 and os.path.getmtime(p) != t_built or os.path.getmtime(p) > t_built
    for junk in (b"short", b"SGBPACK1" + b"\\x00" * 100):                      # foreign / truncated file -> rebuilt, not trusted
        with open(p, "wb") as f:
            f.write(junk)
        os.utime(p, (time.time() + 60, time.time() + 60))
        again = scene_pack.load_scene("pk_0", root, "manual", cache)
        assert np.array_equal(again["seg_members"], sc.seg_members) and np.array_equal(again["real"], sc.real_label)


def test_side_pack_matches_the_side_files(tmp_path):
    """scene_pack.load_side: what SegModel reads per scene (adj, unmap, seg CSR, real labels), packed without the point cloud."""
    from seggroup_b200 import scene_pack, synth
    from seggroup_b200.model import load_scene_files
    sc = synth.make_scene(78, 6500, name="sd_0")
    synth.write_scene_tree(str(tmp_path), [sc])
    root = os.path.join(str(tmp_path), "dataset", "scannet")
    cache = os.path.join(str(tmp_path), "packs")
    adj, unmap, so, sm = load_scene_files("sd_0", root)
    for _ in range(2):                                     # built, then served from the pack
        d = scene_pack.load_side("sd_0", root, cache)
        assert np.array_equal(d["adj"], adj) and np.array_equal(d["unmap"], unmap) and np.array_equal(d["seg_off"], so)
        assert np.array_equal(d["seg_members"], sm) and np.array_equal(d["real"], sc.real_label) and d["data"].shape == (0, 6)
    assert os.listdir(cache) == ["sd_0.side.sgbpack"]
