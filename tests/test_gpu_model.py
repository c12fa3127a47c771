"""Boundary B1: the nn.Module drop-in (`seggroup_b200.model.SegModel`) driven the way seggroup/train.py and
infer.py drive the reference: tensors shaped [1,N,6] / [1,N,2] / [1,1], side files read by relative path, the
14 label files written, loss tuple returned, gradients in .grad after backward, BN buffers updated."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tree(tmp_path, scene8k):
    from seggroup_b200 import synth
    synth.write_scene_tree(str(tmp_path), [scene8k])
    old = os.getcwd()
    os.chdir(tmp_path)
    yield tmp_path
    os.chdir(old)


def _inputs(scene):
    data = torch.from_numpy(scene.data.copy()).unsqueeze(0).cuda()
    weak = torch.from_numpy(scene.weak_label.copy()).unsqueeze(0).cuda()
    info = torch.tensor([[0]]).cuda()       # DDP moves `info` to the GPU as well
    return data, weak, info


def test_train_step_matches_oracle(tree, scene8k):
    from oracle import seggroup_oracle as O
    from seggroup_b200.model import SegModel
    torch.manual_seed(1)
    model = SegModel(exp_name="t").to("cuda")
    with torch.no_grad():
        model.mlp_1.bn1.weight.mul_(4.0)
    model.classifier.dp1.p = 0.0            # identity dropout so the oracle can follow (mask 0.5 * 2 = 1)
    model.train()
    model.epoch = "1"
    params = O.from_reference_state({k: v.detach().cpu() for k, v in model.state_dict().items()})
    out = model(*_inputs(scene8k))
    assert len(out) == 4 and out[0].shape == (1, 2) and out[1].shape == (1, 2, 40) and out[3].shape == (4,)
    loss = torch.sum(out[0][:, 0]) / torch.sum(out[0][:, 1])       # train.py:165-167
    loss.backward()
    model.flush_exports()
    probe = O.forward(scene8k, params, mode="ins_infer", tie="canonical")
    n_inst = len(np.unique(probe["levels"][-1].ins))
    ref = O.forward(scene8k, params, mode="train", tie="canonical", dropout_mask=torch.full((n_inst, 128), 0.5), want_grads=True)
    assert abs(float(out[0][0, 0]) - ref["loss_raw"][0, 0]) < 1e-4 * abs(ref["loss_raw"][0, 0])
    root = os.path.join("results", "t", scene8k.name, "epoch_1")
    files = sorted(os.listdir(root))
    assert len(files) == 14
    for k, v in ref["labels"].items():
        got = np.loadtxt(os.path.join(root, k + ".txt"), dtype=np.int64)
        assert np.array_equal(got, v), k
    named = dict(model.named_parameters())
    for k in O.TRAINABLE:
        gr = ref["grads"][k]
        if gr is None:
            continue
        g = named[k].grad.cpu()
        assert float((g - gr).abs().max()) <= 1e-3 * float(gr.abs().max()) + 1e-7, k
    for a, b in zip(out[1:], ref["metrics"]):
        assert np.allclose(a.cpu().numpy(), b, atol=1e-6)
    assert int(model.mlp_2.bn1.num_batches_tracked) == 1 and float(model.mlp_2.bn1.running_mean.abs().sum()) > 0


def test_infer_modes_and_optimizer(tree, scene8k):
    from seggroup_b200.model import SegModel
    torch.manual_seed(1)
    for flags, tag, n_files in ((dict(sem_infer=True), "sem_infer", 6), (dict(ins_infer=True), "ins_infer", 14)):
        model = SegModel(exp_name="i", **flags).to("cuda")
        model.epoch = tag
        with torch.no_grad():
            out = model(*_inputs(scene8k))
        model.flush_exports()
        assert len(out) == 3
        assert len(os.listdir(os.path.join("results", "i", scene8k.name, tag))) == n_files
    # two SGD steps as train.py runs them (lr 0.1, momentum 0.9, wd 1e-4): the loss must stay finite and parameters move
    model = SegModel(exp_name="o").to("cuda")
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    model.epoch = "1"
    w0 = model.mlp_3.conv2[0].weight.detach().clone()
    for _ in range(2):
        loss_raw = model(*_inputs(scene8k))[0]
        loss = torch.sum(loss_raw[:, 0]) / torch.sum(loss_raw[:, 1])
        opt.zero_grad(); loss.backward(); opt.step()
        assert torch.isfinite(loss)
    assert float((model.mlp_3.conv2[0].weight - w0).abs().max()) > 0
    sd = model.state_dict()
    model2 = SegModel(exp_name="o2")
    model2.load_state_dict(sd)


def test_cpu_input_is_refused(tree, scene8k):
    from seggroup_b200._lib import SgbError
    from seggroup_b200.model import SegModel
    model = SegModel(exp_name="c")
    d, w, i = _inputs(scene8k)
    with pytest.raises(SgbError):
        model(d.cpu(), w.cpu(), i.cpu())


def test_batched_forward_equals_per_scene_calls(tmp_path):
    """B > 1 through the nn.Module (a DataLoader with batch_size = B): loss / metrics per scene and the 14 files of every scene
    equal what B separate calls produce."""
    from seggroup_b200 import synth
    from seggroup_b200.model import SegModel
    scenes = [synth.make_scene(31, 6000, name="scene_a"), synth.make_scene(32, 6000, name="scene_b"), synth.make_scene(33, 6000, name="scene_c")]
    synth.write_scene_tree(str(tmp_path), scenes)
    old = os.getcwd()
    os.chdir(tmp_path)
    try:
        torch.manual_seed(1)
        model = SegModel(exp_name="b").to("cuda")
        with torch.no_grad():
            model.mlp_1.bn1.weight.mul_(4.0)
        model.classifier.dp1.p = 0.0
        model.epoch = "1"
        data = torch.stack([torch.from_numpy(s.data.copy()) for s in scenes]).cuda()
        weak = torch.stack([torch.from_numpy(s.weak_label.copy()) for s in scenes]).cuda()
        info = torch.arange(3).view(3, 1).cuda()
        single = [model(data[b:b + 1], weak[b:b + 1], info[b:b + 1]) for b in range(3)]
        model.flush_exports()
        files = {}
        for s in scenes:
            root = os.path.join("results", "b", s.name, "epoch_1")
            files[s.name] = {f: open(os.path.join(root, f), "rb").read() for f in sorted(os.listdir(root))}
            assert len(files[s.name]) == 14
        model.exp_name = "b2"
        out = model(data, weak, info)
        model.flush_exports()
        assert out[0].shape == (3, 2) and out[1].shape == (3, 2, 40) and out[2].shape == (3, 2, 40) and out[3].shape == (3, 4)
        for b, s in enumerate(scenes):
            assert torch.allclose(out[0][b], single[b][0][0], rtol=1e-6)
            assert torch.equal(out[1][b], single[b][1][0]) and torch.equal(out[2][b], single[b][2][0]) and torch.equal(out[3][b], single[b][3])
            root = os.path.join("results", "b2", s.name, "epoch_1")
            for f, blob in files[s.name].items():
                assert open(os.path.join(root, f), "rb").read() == blob, (s.name, f)
        loss = (out[0][:, 0] / out[0][:, 1]).mean()
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
        assert int(model.mlp_2.bn1.num_batches_tracked) == 6               # 3 single calls + one batch of 3
    finally:
        os.chdir(old)


def test_export_failure_is_raised(tree, scene8k):
    """A label file that cannot be written must surface (the reference writes synchronously and raises)."""
    from seggroup_b200._lib import SgbError
    from seggroup_b200.model import SegModel
    model = SegModel(exp_name="x", ins_infer=True).to("cuda")
    model.epoch = "ins_infer"
    root = os.path.join("results", "x", scene8k.name, "ins_infer")
    os.makedirs(root)
    os.makedirs(os.path.join(root, "final.sem.txt"))                       # a directory where a file has to go
    with torch.no_grad():
        model(*_inputs(scene8k))
    with pytest.raises(SgbError):
        model.flush_exports()


def test_side_pack_cold_path_equals_the_file_parse(tmp_path):
    """SegModel with a cache directory and the HBM scene cache off (first epoch / inference: every scene is read at its forward): the
    side files come from the scene packs, staged in pinned memory and uploaded with one copy per scene — outputs and the 14 label
    files equal the plain parse of the reference's files, on the first call (pack built) and on the second (pack mapped)."""
    from seggroup_b200 import synth
    from seggroup_b200.model import SegModel
    scenes = [synth.make_scene(61 + i, 6000 + 300 * i, name="pk_%d" % i) for i in range(3)]
    synth.write_scene_tree(str(tmp_path), scenes)
    old = os.getcwd()
    os.chdir(tmp_path)
    try:
        outs, files = [], []
        for exp, cache in (("plain", None), ("packed", str(tmp_path / "packs")), ("packed2", str(tmp_path / "packs"))):
            torch.manual_seed(1)
            model = SegModel(exp_name=exp, ins_infer=True).to("cuda")
            with torch.no_grad():
                model.mlp_1.bn1.weight.mul_(4.0)
            model.epoch = "ins_infer"
            model.cache_scenes = False
            model.scene_cache_dir = cache
            res = []
            for b, s in enumerate(scenes):                  # one scene per call (sizes differ) ...
                d = torch.from_numpy(s.data.copy()).unsqueeze(0).cuda()
                w = torch.from_numpy(s.weak_label.copy()).unsqueeze(0).cuda()
                res.append(model(d, w, torch.tensor([[b]]).cuda()))
            model.flush_exports()
            outs.append(res)
            got = {}
            for s in scenes:
                root = os.path.join("results", exp, s.name, "ins_infer")
                got[s.name] = {f: open(os.path.join(root, f), "rb").read() for f in sorted(os.listdir(root))}
                assert len(got[s.name]) == 14
            files.append(got)
        assert sorted(os.listdir(tmp_path / "packs")) == ["pk_%d.side.sgbpack" % i for i in range(3)]
        for other in (1, 2):
            assert files[other] == files[0]
            for a, b in zip(outs[0], outs[other]):
                for x, y in zip(a, b):
                    assert torch.equal(x, y)
    finally:
        os.chdir(old)
