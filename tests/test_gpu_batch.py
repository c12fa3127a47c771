"""Scene batches (pipeline.SceneDevice.concat): several scenes concatenated into one block-diagonal graph must give, for every
scene, exactly what the scene gives alone — the reference runs one scene per forward (train.py:92, model.py:684-693), so every
per-scene rule has to survive the batching: BatchNorm statistics, the order-dependent grouping replays (one CTA per scene),
arg-min columns of the per-scene distance matrix, kNN lists (ids relative to the scene, unfilled columns -> the scene's point 0),
segment labels (root point id inside the scene), metrics, the classifier head and its loss."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scenes():
    from seggroup_b200 import synth
    a = synth.make_scene(5, 8000, n_small_segs=2)
    b = synth.make_scene(21, 9001, n_small_segs=1)                       # odd point count: slices of the batch arrays are not 16-byte multiples
    c = synth.make_scene(22, 6000)
    # scene d: its first cluster is isolated and unlabeled, so it survives phase A (row arg-min = itself = the FIRST cluster of
    # its own scene) and phase B has to assign it (tests/test_gpu_pipeline.py::test_final_clustering_phase_b) — inside a batch
    d = copy.deepcopy(synth.make_scene(23, 7000))
    seg_pts = d.seg_members[d.seg_offsets[0]:d.seg_offsets[1]]
    d.weak_label = d.weak_label.copy(); d.weak_label[seg_pts] = -1
    in_seg = np.zeros(d.n_points, bool); in_seg[seg_pts] = True
    d.adj = d.adj[~(in_seg[d.adj[:, 0]] | in_seg[d.adj[:, 1]])]
    return [a, b, d, c]


def _params(train):
    from seggroup_b200.params import TRAINABLE, init_params
    p = {k: v.cuda() for k, v in init_params(1, 4.0).items()}
    if train:
        for k in TRAINABLE:
            p[k].requires_grad_(True)
    return p


@pytest.mark.parametrize("mode", ["ins_infer", "sem_infer"])
def test_batch_labels_equal_single_scene_labels(mode):
    from seggroup_b200 import pipeline
    scenes = _scenes()
    p = _params(False)
    dev = [pipeline.SceneDevice.from_host(s) for s in scenes]
    with torch.no_grad():
        single = [pipeline.forward_scene(d, p, mode=mode, keep_aux=True) for d in dev]
        batch = pipeline.forward_scene(pipeline.SceneDevice.concat(dev), p, mode=mode, keep_aux=True)
    assert batch.status == 0
    if mode == "ins_infer":
        assert single[2].levels[-1].S == single[2].aux["phaseA_clusters"] - 1, "phase B did not fire in the third scene"
    pt = batch.levels[0].scene_cl_off
    assert len(pt) == len(scenes) + 1
    for b, r in enumerate(single):
        lab = batch.scene_labels(b)
        assert set(lab) == set(r.labels)
        for k, v in r.labels.items():
            assert torch.equal(lab[k], v), (b, k, int((lab[k] != v).sum()))
        for x, y in zip(batch.metrics_scenes[b], r.metrics):
            assert torch.equal(x, y), b
        for Lb, Ls in zip(batch.levels, r.levels):                        # cluster counts of the scene at every level
            assert Lb.scene_cl_off[b + 1] - Lb.scene_cl_off[b] == Ls.S
    if mode == "ins_infer":                                                # kNN lists: ids relative to the scene
        lo = 0
        for r, d in zip(single, dev):
            for t in ("2", "3"):
                assert torch.equal(batch.aux["knn_" + t][lo:lo + d.n_points], r.aux["knn_" + t]), t
            lo += d.n_points


def test_batch_training_step_equals_mean_of_single_scene_steps():
    """loss_raw per scene and d(mean_b loss_b)/d(params) of the batch against the scenes run one by one (same dropout masks)."""
    from seggroup_b200 import pipeline
    from seggroup_b200.params import TRAINABLE
    scenes = _scenes()
    dev = [pipeline.SceneDevice.from_host(s) for s in scenes]
    g = torch.Generator().manual_seed(7)
    p = _params(True)
    with torch.no_grad():
        n_inst = [int(torch.unique(pipeline.forward_scene(d, p, mode="ins_infer").levels[-1].cl_ins).numel()) for d in dev]
    masks = [(torch.rand(n, 128, generator=g) > 0.5).cuda() for n in n_inst]
    grads, losses = [], []
    for d, m in zip(dev, masks):
        r = pipeline.forward_scene(d, p, mode="train", dropout_mask=m)
        losses.append(r.loss_raw[0].detach())
        gs = torch.autograd.grad(r.loss_raw[0, 0] / r.loss_raw[0, 1], [p[k] for k in TRAINABLE], allow_unused=True)
        grads.append(gs)
    rb = pipeline.forward_scene(pipeline.SceneDevice.concat(dev), p, mode="train", dropout_mask=masks)
    assert rb.loss_raw.shape == (len(scenes), 2)
    for b, l in enumerate(losses):
        assert torch.allclose(rb.loss_raw[b].detach(), l, rtol=1e-6, atol=0), (b, rb.loss_raw[b], l)
    gb = torch.autograd.grad((rb.loss_raw[:, 0] / rb.loss_raw[:, 1]).mean(), [p[k] for k in TRAINABLE], allow_unused=True)
    for i, k in enumerate(TRAINABLE):
        if gb[i] is None:
            assert all(gs[i] is None for gs in grads), k
            continue
        ref = sum(gs[i] for gs in grads) / len(scenes)
        err = float((gb[i] - ref).abs().max() / (ref.abs().max() + 1e-30))
        assert err < 2e-5, (k, err)


def test_executor_fused_batch_matches_stream_executor():
    from seggroup_b200 import engine, pipeline
    from seggroup_b200.params import TRAINABLE
    scenes = _scenes()[:3]
    dev = [pipeline.SceneDevice.from_host(s) for s in scenes]
    p = _params(True)
    torch.manual_seed(3)
    ex = engine.SceneExecutor(n_streams=2, reserve_bytes_per_stream=0, fused=False)
    la = ex.train_batch(dev, p, list(TRAINABLE))
    ga = {k: p[k].grad.clone() for k in TRAINABLE if p[k].grad is not None}
    ex.close()
    for k in TRAINABLE:
        p[k].grad = None
    exf = engine.SceneExecutor(reserve_bytes_per_stream=0, fused=True)
    lb = exf.train_batch(dev, p, list(TRAINABLE))
    exf.close()
    # the classifier's dropout masks differ between the two runs (torch RNG): compare everything the mask does not touch
    assert exf.last_result.loss_raw.shape == (3, 2)
    assert torch.isfinite(la) and torch.isfinite(lb)
    for k in ga:
        assert p[k].grad is not None and p[k].grad.shape == ga[k].shape


def test_shard_loader_batch_runs_like_resident_batch(tmp_path):
    """N2: a batch collated by the shard loader from the on-disk tree (pinned host arrays, ids already offset) gives the labels of
    the same scenes uploaded one by one and concatenated on the device."""
    import os
    from seggroup_b200 import pipeline, synth
    from seggroup_b200.loader import SceneShardLoader
    scenes = [synth.make_scene(61 + i, 6000 + 250 * i, name="lg_%d" % i) for i in range(3)]
    synth.write_scene_tree(str(tmp_path), scenes)
    root = os.path.join(str(tmp_path), "dataset", "scannet")
    names = open(os.path.join(root, "scannetv2_train.txt")).readlines()
    hb = next(iter(SceneShardLoader(names, data_root=root, batch_size=3, cache_dir=os.path.join(str(tmp_path), "c"))))
    assert hb.data.is_pinned()
    p = _params(False)
    with torch.no_grad():
        a = pipeline.forward_scene(hb.to_device("cuda"), p, mode="ins_infer")
        b = pipeline.forward_scene(pipeline.SceneDevice.concat([pipeline.SceneDevice.from_host(s) for s in scenes]), p, mode="ins_infer")
    for k, v in b.labels.items():
        assert torch.equal(a.labels[k], v), k
    for x, y in zip(a.metrics_scenes, b.metrics_scenes):
        assert all(torch.equal(u, w) for u, w in zip(x, y))
