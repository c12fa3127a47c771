"""Scene packs: one raw binary file per scene with every array a forward needs, in the dtypes that are uploaded.

The reference stores a scene as five `torch.save` pickles plus a JSON list of 150,000 Python lists (seggroup/data.py:28-38,
model.py:696-724, 610-614); reading them costs 11 ms per 150k-point scene even when the JSON has been parsed before (three zip
archives with CRC checks, a `.npz`), which a B200 step does not hide.  A pack is written once next to the binary CSR cache and
mapped back with ONE `np.memmap` (views into the page cache, no parsing, no copy): 64-byte header, then the arrays 64-byte aligned.

    header: b"SGBPACK1" + int64[7] = rows of (data, weak, seg_off, seg_members, adj, unmap, real)
    data f32 [N,6] | weak i32 [N,2] | seg_off i32 [S1+1] | seg_members i32 [N] | adj i32 [E0,2] | unmap i64 [N_raw] | real i64 [N_raw,2]

A pack is rebuilt when any of its source files is newer, or when it does not parse (truncated, foreign, other version).
"""
from __future__ import annotations

import os

import numpy as np

MAGIC = b"SGBPACK1"
FIELDS = (("data", np.float32, 6), ("weak", np.int32, 2), ("seg_off", np.int32, 1), ("seg_members", np.int32, 1),
          ("adj", np.int32, 2), ("unmap", np.int64, 1), ("real", np.int64, 2))
_ALIGN = 64


def source_paths(name, data_root, label_style):
    j = os.path.join
    return (j(data_root, "data", "resampled", name, name + ".pcl.pth"),
            j(data_root, "label", "seg", label_style, "resampled", name, name + ".label.pth"),
            j(data_root, "adj", "mesh", "resampled", name, name + ".adj.pth"),
            j(data_root, "data", "resampled", name, name + ".unmap.pth"),
            j(data_root, "label", "real", "resampled", name, name + ".seg.json"),
            j(data_root, "label", "real", "raw", name, name + ".label.pth"))


def pack_path(cache_dir, name, label_style):
    return os.path.join(cache_dir, "%s.%s.sgbpack" % (name, label_style))


def _layout(rows):
    off, out = _ALIGN, []
    for (key, dt, cols), r in zip(FIELDS, rows):
        nbytes = int(r) * cols * np.dtype(dt).itemsize
        out.append((key, dt, cols, int(r), off, nbytes))
        off = (off + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    return out, off


def write_pack(path, arrays):
    """arrays: dict with the keys of FIELDS.  Atomic (temporary file + rename): concurrent ranks may build the same scene."""
    rows = [int(np.asarray(arrays[k]).shape[0]) for k, _, _ in FIELDS]
    lay, total = _layout(rows)
    buf = np.zeros(total, np.uint8)
    buf[:8] = np.frombuffer(MAGIC, np.uint8)
    buf[8:8 + 56].view(np.int64)[:] = rows
    for key, dt, cols, r, off, nbytes in lay:
        a = np.ascontiguousarray(np.asarray(arrays[key]).astype(dt, copy=False)).reshape(r, cols) if cols > 1 else \
            np.ascontiguousarray(np.asarray(arrays[key]).astype(dt, copy=False)).reshape(r)
        buf[off:off + nbytes] = a.view(np.uint8).reshape(-1)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    tmp = "%s.tmp%d" % (path, os.getpid())
    buf.tofile(tmp)
    os.replace(tmp, path)


def read_pack(path):
    """-> dict of numpy views into one buffer; raises ValueError on anything that is not a complete pack."""
    buf = np.memmap(path, dtype=np.uint8, mode="r")          # views into the page cache: the collate copy is the only copy
    if buf.size < _ALIGN or bytes(buf[:8]) != MAGIC:
        raise ValueError("not a scene pack: " + path)
    rows = buf[8:8 + 56].view(np.int64)
    if (rows < 0).any():
        raise ValueError("corrupt scene pack: " + path)
    lay, total = _layout(rows)
    if buf.size != total:
        raise ValueError("truncated scene pack: " + path)
    out = {}
    for key, dt, cols, r, off, nbytes in lay:
        a = buf[off:off + nbytes].view(dt)
        out[key] = a.reshape(r, cols) if cols > 1 else a
    return out


def load_scene(name, data_root=os.path.join("dataset", "scannet"), label_style="manual", cache_dir=None):
    """Everything one forward of one scene needs, as numpy arrays (file layout of SURVEY.md 9.1): from the scene pack when
    `cache_dir` holds a current one, else parsed from the reference's files (and packed for the next time)."""
    import torch
    from .model import load_scene_files
    if cache_dir is not None:
        p = pack_path(cache_dir, name, label_style)
        try:
            newest = max(os.path.getmtime(s) for s in source_paths(name, data_root, label_style))
            if os.path.getmtime(p) >= newest:
                return read_pack(p)
        except (OSError, ValueError):                     # missing, stale, truncated or foreign: rebuild
            pass
    d = os.path.join(data_root, "data", "resampled", name)
    data = torch.load(os.path.join(d, name + ".pcl.pth")).numpy().astype(np.float32, copy=False)
    weak = torch.load(os.path.join(data_root, "label", "seg", label_style, "resampled", name, name + ".label.pth")).numpy().astype(np.int32)
    adj, unmap, seg_off, seg_members = load_scene_files(name, data_root, None)
    real = torch.load(os.path.join(data_root, "label", "real", "raw", name, name + ".label.pth")).numpy().astype(np.int64, copy=False)
    out = dict(data=data, weak=weak, adj=adj, unmap=unmap, seg_off=seg_off, seg_members=seg_members, real=real)
    if cache_dir is not None:
        write_pack(pack_path(cache_dir, name, label_style), out)
    return out


def side_sources(name, data_root):
    j = os.path.join
    return (j(data_root, "adj", "mesh", "resampled", name, name + ".adj.pth"), j(data_root, "data", "resampled", name, name + ".unmap.pth"),
            j(data_root, "label", "real", "resampled", name, name + ".seg.json"), j(data_root, "label", "real", "raw", name, name + ".label.pth"))


def load_side(name, data_root=os.path.join("dataset", "scannet"), cache_dir=None):
    """The files SegModel.forward itself reads for a scene (model.py:696-699, 713-724, 610-614: adj, unmap, seg.json, real labels)
    as a pack without the point cloud / weak labels (those arrive through the loader): dict with adj, unmap, seg_off,
    seg_members, real."""
    import torch
    from .model import load_scene_files
    if cache_dir is not None:
        p = os.path.join(cache_dir, name + ".side.sgbpack")
        try:
            if os.path.getmtime(p) >= max(os.path.getmtime(s) for s in side_sources(name, data_root)):
                return read_pack(p)
        except (OSError, ValueError):
            pass
    adj, unmap, seg_off, seg_members = load_scene_files(name, data_root, None)
    real = torch.load(os.path.join(data_root, "label", "real", "raw", name, name + ".label.pth")).numpy().astype(np.int64, copy=False)
    out = dict(data=np.zeros((0, 6), np.float32), weak=np.zeros((0, 2), np.int32), adj=adj, unmap=unmap, seg_off=seg_off,
               seg_members=seg_members, real=real)
    if cache_dir is not None:
        write_pack(os.path.join(cache_dir, name + ".side.sgbpack"), out)
    return out


def pack_layout(buf):
    """{key: (byte offset, rows, cols, numpy dtype)} of a pack held in a uint8 buffer (see read_pack for the checks)."""
    rows = np.asarray(buf[8:8 + 56]).view(np.int64)
    lay, _ = _layout(rows)
    return {key: (off, r, cols, dt) for key, dt, cols, r, off, nbytes in lay}
