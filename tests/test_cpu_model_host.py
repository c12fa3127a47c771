"""Host-side logic of seggroup_b200.model.SegModel that needs no GPU."""
import types

import torch


def test_running_statistics_closed_form_matches_sequential_updates(tmp_path, monkeypatch):
    """SegModel._update_bn applies the n per-scene BatchNorm running-statistics updates of a scene batch in closed form; it must equal
    the reference's one-update-per-forward sequence (torch BatchNorm: momentum 0.1, unbiased variance; seggroup/model.py never calls .eval())."""
    from seggroup_b200.model import SegModel
    torch.manual_seed(0)
    (tmp_path / "dataset" / "scannet").mkdir(parents=True)
    (tmp_path / "dataset" / "scannet" / "scannetv2_train.txt").write_text("scene0000_00\n")     # the constructor reads the scene list (model.py:668)
    monkeypatch.chdir(tmp_path)
    model = SegModel(exp_name="t")
    model.train()
    B = 5
    counts = [3000000, 2999980, 1200, 40, 7]
    mean, var = torch.randn(B, 64), torch.rand(B, 64) + 0.1
    icount = torch.tensor([12.0, 3.0, 1.0, 40.0, 2.0])
    hm, hv = torch.randn(B, 128), torch.rand(B, 128)
    res = types.SimpleNamespace(bn_stats_scenes={"mlp_2.bn1": (mean, var, counts), "classifier.bn1": (hm, hv, icount), "unknown": (mean, var, counts)})
    want = {}
    for name, bn, (mu, v, c) in (("mlp_2", model.mlp_2.bn1, (mean, var, counts)), ("head", model.classifier.bn1, (hm, hv, icount))):
        rm, rv = bn.running_mean.clone(), bn.running_var.clone()
        for b in range(B):
            cb = float(c[b])
            rm = 0.9 * rm + 0.1 * mu[b]
            rv = 0.9 * rv + 0.1 * v[b] * (cb / max(cb - 1, 1))
        want[name] = (rm, rv)
    for _ in range(2):                                   # second call: cached weights
        for bn in (model.mlp_2.bn1, model.classifier.bn1):
            bn.reset_running_stats()
        model._update_bn(res)
        for name, bn in (("mlp_2", model.mlp_2.bn1), ("head", model.classifier.bn1)):
            assert torch.allclose(bn.running_mean, want[name][0], rtol=1e-5, atol=1e-6)
            assert torch.allclose(bn.running_var, want[name][1], rtol=1e-5, atol=1e-6)
            assert int(bn.num_batches_tracked) == B
    assert int(model.mlp_3.bn1.num_batches_tracked) == 0


def test_scene_batch_concat_offsets():
    """SceneDevice.concat: ids of scene b are shifted by the first point of scene b (one expanded offset vector per array) — equal to
    the per-scene formula, dtypes kept, inputs untouched."""
    from seggroup_b200 import pipeline, synth
    scs = [synth.make_scene(40 + i, 8000 + 500 * i) for i in range(3)]
    sd = [pipeline.SceneDevice.from_host(s, device="cpu") for s in scs]
    keep = [s.seg_members.clone() for s in sd]
    b = pipeline.SceneDevice.concat(sd)
    pt = [0]
    for s in sd:
        pt.append(pt[-1] + s.n_points)
    assert torch.equal(b.seg_members, torch.cat([s.seg_members + pt[i] for i, s in enumerate(sd)])) and b.seg_members.dtype == torch.int32
    assert torch.equal(b.adj0, torch.cat([s.adj0 + pt[i] for i, s in enumerate(sd)]))
    assert torch.equal(b.unmap, torch.cat([s.unmap + pt[i] for i, s in enumerate(sd)])) and b.unmap.dtype == torch.int64
    assert torch.equal(b.seg_off, torch.cat([sd[0].seg_off[:1]] + [s.seg_off[1:] + pt[i] for i, s in enumerate(sd)]))
    assert b.split.pt_off == pt and b.n_scenes == 3
    assert all(torch.equal(a, s.seg_members) for a, s in zip(keep, sd))


def test_batch_index_vectors_without_repeat_interleave():
    """kpconv_inputs.get_batch_inds (common.py:386-430) and the offset vectors of SceneDevice.concat come from a parallel binary search
    over the batch boundaries; equal to torch.repeat_interleave, empty batch entries included."""
    from seggroup_b200.kpconv_inputs import get_batch_inds
    for lens in ([3, 2, 5], [0, 4, 0, 0, 7, 1], [5], [0], [1, 1, 1, 1]):
        t = torch.tensor(lens, dtype=torch.int32)
        ref = torch.repeat_interleave(torch.arange(len(lens), dtype=torch.int32), t.long())
        got = get_batch_inds(t)
        assert torch.equal(ref, got) and got.dtype == torch.int32, lens
