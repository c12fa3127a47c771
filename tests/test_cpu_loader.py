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
