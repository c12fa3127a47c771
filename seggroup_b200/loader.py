"""Shard loader for the on-disk formats either side of the hot path.

The reference feeds `SegModel.forward` from `seggroup/data.py:18-41` (`ScanNet(Dataset)`: torch.load of `<scene>.pcl.pth`,
`.label.pth`, `.info.pth` per item, DataLoader with batch size 1) and the model then re-reads and re-parses the scene's side files
on EVERY forward (`adj.pth`, `unmap.pth`, and `seg.json` — a JSON list of 150,000 Python lists, model.py:696-724) plus the real
labels for `evaluate` (model.py:610-614).  On a B200 that host work is an order of magnitude slower than the step itself.

`SceneShardLoader` does all of it ahead of the step, for a whole shard of scenes:
  * the scene list is partitioned `rank::world` (engine.shard_scenes; padded when the consumer issues a collective per step);
  * a background thread loads the next batches: point cloud, weak labels, side files through the binary CSR cache
    (`model.load_scene_files(..., cache_dir)`: seg.json is parsed once per scene, ever), real labels;
  * every batch is collated into ONE set of pinned host arrays in the block-diagonal layout `pipeline.SceneDevice.concat` produces
    (point / segment ids offset by the scene's base), so the step starts with seven large H2D copies and no device-side fix-up.

    for hb in SceneShardLoader(names, rank=r, world=w, batch_size=8, cache_dir="csr_cache"):
        batch = hb.to_device("cuda")                    # pipeline.SceneDevice (a scene batch)
        res = pipeline.forward_scene(batch, params, mode="ins_infer")
"""
from __future__ import annotations

import os
import queue
import threading
from dataclasses import dataclass

import numpy as np
import torch

from . import engine
from .model import load_scene_files


@dataclass
class HostBatch:
    """One collated batch in (pinned) host memory; ids already offset by the scene's base."""
    names: list
    index: list                   # positions of the scenes in the scene list (the `info` tensor of seggroup/data.py:37)
    data: torch.Tensor            # [sum N, 6] f32
    weak_label: torch.Tensor      # [sum N, 2] i32
    seg_off: torch.Tensor         # [sum S + 1] i32
    seg_members: torch.Tensor     # [sum N] i32
    adj0: torch.Tensor            # [sum E0, 2] i32
    unmap: torch.Tensor           # [sum N_raw] i64
    real_label: torch.Tensor      # [sum N_raw, 2] i64
    pt_off: list
    seg_cnt_off: list
    raw_off: list

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.data, self.weak_label, self.seg_off, self.seg_members, self.adj0, self.unmap, self.real_label))

    def to_device(self, device="cuda"):
        from . import ops, pipeline
        t = lambda a: a.to(device, non_blocking=True)
        split = ops.SceneSplit(self.pt_off, self.seg_cnt_off, self.raw_off, device) if len(self.names) > 1 else None
        return pipeline.SceneDevice(data=t(self.data), weak_label=t(self.weak_label), seg_off=t(self.seg_off), seg_members=t(self.seg_members),
                                    adj0=t(self.adj0), unmap=t(self.unmap), real_label=t(self.real_label), name="+".join(self.names), split=split)


def load_scene(name, data_root=os.path.join("dataset", "scannet"), label_style="manual", cache_dir=None):
    """Everything one forward of one scene needs, as numpy arrays (file layout of SURVEY.md 9.1)."""
    d = os.path.join(data_root, "data", "resampled", name)
    data = torch.load(os.path.join(d, name + ".pcl.pth")).numpy().astype(np.float32, copy=False)
    weak = torch.load(os.path.join(data_root, "label", "seg", label_style, "resampled", name, name + ".label.pth")).numpy().astype(np.int32)
    adj, unmap, seg_off, seg_members = load_scene_files(name, data_root, cache_dir)
    real = torch.load(os.path.join(data_root, "label", "real", "raw", name, name + ".label.pth")).numpy().astype(np.int64, copy=False)
    return dict(data=data, weak=weak, adj=adj, unmap=unmap, seg_off=seg_off, seg_members=seg_members, real=real)


def collate(scenes, names, index, pin=True):
    """Block-diagonal concatenation of per-scene arrays into one set of (pinned) host tensors."""
    pt, sg, rw, ed = [0], [0], [0], [0]
    for s in scenes:
        pt.append(pt[-1] + s["data"].shape[0]); sg.append(sg[-1] + len(s["seg_off"]) - 1)
        rw.append(rw[-1] + s["unmap"].shape[0]); ed.append(ed[-1] + s["adj"].shape[0])
    pin = pin and torch.cuda.is_available()
    new = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pin)
    hb = HostBatch(names=list(names), index=list(index), data=new((pt[-1], 6), torch.float32), weak_label=new((pt[-1], 2), torch.int32),
                   seg_off=new((sg[-1] + 1,), torch.int32), seg_members=new((pt[-1],), torch.int32), adj0=new((ed[-1], 2), torch.int32),
                   unmap=new((rw[-1],), torch.int64), real_label=new((rw[-1], 2), torch.int64), pt_off=pt, seg_cnt_off=sg, raw_off=rw)
    hb.seg_off[0] = 0
    for i, s in enumerate(scenes):
        p0, p1 = pt[i], pt[i + 1]
        hb.data[p0:p1] = torch.from_numpy(s["data"])
        hb.weak_label[p0:p1] = torch.from_numpy(s["weak"])
        hb.seg_members[p0:p1] = torch.from_numpy(s["seg_members"] + p0)
        hb.seg_off[sg[i] + 1:sg[i + 1] + 1] = torch.from_numpy(s["seg_off"][1:] + p0)
        hb.adj0[ed[i]:ed[i + 1]] = torch.from_numpy(s["adj"] + p0)
        hb.unmap[rw[i]:rw[i + 1]] = torch.from_numpy(s["unmap"] + p0)
        hb.real_label[rw[i]:rw[i + 1]] = torch.from_numpy(s["real"])
    return hb


class SceneShardLoader:
    """Iterates over the batches of this rank's shard; a daemon thread keeps `prefetch` collated batches ready."""

    def __init__(self, scene_names, data_root=os.path.join("dataset", "scannet"), label_style="manual", rank=0, world=1, batch_size=8,
                 cache_dir=None, pad=False, pin=True, prefetch=2, epochs=1):
        self.names = [n.strip() for n in scene_names]
        self.data_root, self.label_style, self.cache_dir = data_root, label_style, cache_dir
        self.batch_size, self.pin, self.prefetch, self.epochs = int(batch_size), pin, int(prefetch), int(epochs)
        self.shard = engine.shard_scenes(len(self.names), rank, world, pad=pad)

    def __len__(self):
        return self.epochs * ((len(self.shard) + self.batch_size - 1) // self.batch_size)

    def _batches(self):
        for _ in range(self.epochs):
            for b0 in range(0, len(self.shard), self.batch_size):
                yield self.shard[b0:b0 + self.batch_size]

    def _load(self, idx):
        names = [self.names[i] for i in idx]
        scenes = [load_scene(n, self.data_root, self.label_style, self.cache_dir) for n in names]
        return collate(scenes, names, idx, self.pin)

    def __iter__(self):
        q = queue.Queue(maxsize=max(1, self.prefetch))
        stop = threading.Event()

        def work():
            try:
                for idx in self._batches():
                    if stop.is_set():
                        return
                    q.put(self._load(idx))
                q.put(None)
            except BaseException as e:               # surface loader failures in the consumer
                q.put(e)

        th = threading.Thread(target=work, daemon=True, name="sgb-shard-loader")
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while not q.empty():
                try:
                    q.get_nowait()
                except queue.Empty:
                    break
