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
from . import scene_pack


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

    def recycle(self):
        """Give the pinned buffers back to the loader's pool (call once the H2D copies of this batch have completed)."""
        owned, self._owned = getattr(self, "_owned", []), []
        with _pinned_lock:
            for key, base in owned:
                _pinned_pool.setdefault(key, []).append(base)

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
    """Everything one forward of one scene needs, as numpy arrays (scene_pack.load_scene: one raw binary read per scene when
    `cache_dir` is given, the reference's pickles + JSON otherwise)."""
    return scene_pack.load_scene(name, data_root, label_style, cache_dir)


_pinned_pool = {}
_pinned_lock = threading.Lock()


def _pinned(tag, n_elems, dtype, cols, pin):
    """[n, cols] host tensor; pinned buffers are drawn from a pool of capacity-rounded allocations (cudaHostAlloc of a few tens of
    MB costs milliseconds, a batch needs seven of them): a HostBatch gives its buffers back with `recycle()`."""
    if not pin:
        return torch.empty((n_elems, cols) if cols > 1 else (n_elems,), dtype=dtype), None
    cap = max(1, -(-n_elems // 65536)) * 65536
    key = (tag, cap, cols, dtype)
    with _pinned_lock:
        pool = _pinned_pool.setdefault(key, [])
        base = pool.pop() if pool else None
    if base is None:
        base = torch.empty((cap, cols) if cols > 1 else (cap,), dtype=dtype, pin_memory=True)
    return base[:n_elems], (key, base)


def collate(scenes, names, index, pin=True, pool=None):
    """Block-diagonal concatenation of per-scene arrays into one set of (pinned) host tensors; ids are offset by the first point
    of their scene while they are copied (no temporaries).  pool: optional executor — the per-scene copies (numpy, GIL released)
    run on its threads."""
    pt, sg, rw, ed = [0], [0], [0], [0]
    for s in scenes:
        pt.append(pt[-1] + s["data"].shape[0]); sg.append(sg[-1] + len(s["seg_off"]) - 1)
        rw.append(rw[-1] + s["unmap"].shape[0]); ed.append(ed[-1] + s["adj"].shape[0])
    pin = pin and torch.cuda.is_available()
    owned = []

    def new(tag, n, dt, cols=1):
        t, own = _pinned(tag, n, dt, cols, pin)
        if own is not None:
            owned.append(own)
        return t
    hb = HostBatch(names=list(names), index=list(index), data=new("data", pt[-1], torch.float32, 6), weak_label=new("weak", pt[-1], torch.int32, 2),
                   seg_off=new("seg_off", sg[-1] + 1, torch.int32), seg_members=new("seg_members", pt[-1], torch.int32),
                   adj0=new("adj", ed[-1], torch.int32, 2), unmap=new("unmap", rw[-1], torch.int64), real_label=new("real", rw[-1], torch.int64, 2),
                   pt_off=pt, seg_cnt_off=sg, raw_off=rw)
    hb._owned = owned
    v = {k: getattr(hb, k).numpy() for k in ("data", "weak_label", "seg_off", "seg_members", "adj0", "unmap", "real_label")}
    v["seg_off"][0] = 0

    def copy_scene(i):
        s = scenes[i]
        p0, p1 = pt[i], pt[i + 1]
        v["data"][p0:p1] = s["data"]
        v["weak_label"][p0:p1] = s["weak"]
        np.add(s["seg_members"], p0, out=v["seg_members"][p0:p1], casting="unsafe")
        np.add(s["seg_off"][1:], p0, out=v["seg_off"][sg[i] + 1:sg[i + 1] + 1], casting="unsafe")
        np.add(s["adj"], p0, out=v["adj0"][ed[i]:ed[i + 1]], casting="unsafe")
        np.add(s["unmap"], p0, out=v["unmap"][rw[i]:rw[i + 1]], casting="unsafe")
        v["real_label"][rw[i]:rw[i + 1]] = s["real"]

    if pool is not None and len(scenes) > 1:
        list(pool.map(copy_scene, range(len(scenes))))
    else:
        for i in range(len(scenes)):
            copy_scene(i)
    return hb


class SceneShardLoader:
    """Iterates over the batches of this rank's shard; a daemon thread keeps `prefetch` collated batches ready."""

    def __init__(self, scene_names, data_root=os.path.join("dataset", "scannet"), label_style="manual", rank=0, world=1, batch_size=8,
                 cache_dir=None, pad=False, pin=True, prefetch=2, epochs=1):
        self.names = [n.strip() for n in scene_names]
        self.data_root, self.label_style, self.cache_dir = data_root, label_style, cache_dir
        self.batch_size, self.pin, self.prefetch, self.epochs = int(batch_size), pin, int(prefetch), int(epochs)
        self.shard = engine.shard_scenes(len(self.names), rank, world, pad=pad)
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=max(1, min(4, self.batch_size)), thread_name_prefix="sgb-scene-read")

    def __len__(self):
        return self.epochs * ((len(self.shard) + self.batch_size - 1) // self.batch_size)

    def _batches(self):
        for _ in range(self.epochs):
            for b0 in range(0, len(self.shard), self.batch_size):
                yield self.shard[b0:b0 + self.batch_size]

    def _load(self, idx):
        names = [self.names[i] for i in idx]
        one = lambda n: load_scene(n, self.data_root, self.label_style, self.cache_dir)
        scenes = list(self._pool.map(one, names)) if len(names) > 1 else [one(names[0])]       # file reads release the GIL
        return collate(scenes, names, idx, self.pin, pool=self._pool)

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
