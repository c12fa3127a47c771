"""Synthetic ScanNet-shaped scenes.

The reference ships no data; every test, the oracle and bench.py draw their inputs from this
seeded generator.  A scene is a box room (floor, 4 walls, axis-aligned furniture boxes) sampled
at N surface points, over-segmented into S ~ N/165 segments, with a symmetric 3-NN point
adjacency, seg-level weak labels (largest segment of each instance labelled) and the identity
resampling maps.  `write_scene_tree` lays the files out exactly where the reference reads them
(seggroup/data.py:28-38, seggroup/model.py:696-699, 610-611).

Numpy/scipy only: runs on the CPU box and on the GPU box alike.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np

SEM_VALID_CLASS_IDS = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39])


@dataclass
class Scene:
    name: str
    data: np.ndarray          # [N,6] f32  xyz (m) + rgb in [-1,1]
    weak_label: np.ndarray    # [N,2] i64  (sem 0..39, ins 0..I-1), -1 = unlabeled
    seg_members: np.ndarray   # [N]   i32  point ids, segments concatenated in ascending-root order
    seg_offsets: np.ndarray   # [S+1] i32  CSR offsets into seg_members
    adj: np.ndarray           # [E0,2] i64 row-sorted, unique-lexicographic point adjacency
    real_label: np.ndarray    # [N,2] i64  (sem 1..40, ins 1..I)
    unmap: np.ndarray         # [N_raw] i64 raw vertex -> resampled point
    map: np.ndarray           # [N] i64 resampled point -> raw vertex
    meta: dict = field(default_factory=dict)

    @property
    def n_points(self):
        return self.data.shape[0]

    @property
    def n_segments(self):
        return self.seg_offsets.shape[0] - 1

    def seg_json(self):
        """List of length N; entry i = member list of the segment rooted at i, else [] (dataset/scannet/util.py:205-220)."""
        out = [[] for _ in range(self.n_points)]
        for s in range(self.n_segments):
            m = self.seg_members[self.seg_offsets[s]:self.seg_offsets[s + 1]]
            out[int(m[0])] = [int(v) for v in m]
        return out


def _sample_rect(rng, n, origin, u, v):
    a = rng.random((n, 1))
    b = rng.random((n, 1))
    return origin[None, :] + a * u[None, :] + b * v[None, :], np.concatenate([a * np.linalg.norm(u), b * np.linalg.norm(v)], 1)


def make_scene(seed: int, n_points: int = 50000, *, dup_frac: float = 0.03, pts_per_seg: int = 165,
               n_small_segs: int = 0, name: str | None = None, noise: float = 0.004) -> Scene:
    """Generate one scene.  Deterministic in (seed, n_points, options)."""
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    side = 6.0 * np.sqrt(n_points / 50000.0)
    H = 3.0
    n_box = int(np.clip(12 * n_points / 50000.0, 4, 40))

    # ---- surfaces: (instance id, origin, u, v) rectangles --------------------------------------
    rects = []
    rects.append((0, np.array([0, 0, 0.0]), np.array([side, 0, 0.0]), np.array([0, side, 0.0])))  # floor
    wall_defs = [
        (np.array([0, 0, 0.0]), np.array([side, 0, 0.0])),
        (np.array([0, side, 0.0]), np.array([side, 0, 0.0])),
        (np.array([0, 0, 0.0]), np.array([0, side, 0.0])),
        (np.array([side, 0, 0.0]), np.array([0, side, 0.0])),
    ]
    for w, (o, u) in enumerate(wall_defs):
        rects.append((1 + w, o, u, np.array([0, 0, H])))
    for b in range(n_box):
        sx, sy, sz = rng.uniform(0.4, 1.4), rng.uniform(0.4, 1.4), rng.uniform(0.3, 1.2)
        ox, oy = rng.uniform(0.1, side - sx - 0.1), rng.uniform(0.1, side - sy - 0.1)
        ins = 5 + b
        o = np.array([ox, oy, 0.0])
        rects.append((ins, o + np.array([0, 0, sz]), np.array([sx, 0, 0.0]), np.array([0, sy, 0.0])))  # top
        rects.append((ins, o, np.array([sx, 0, 0.0]), np.array([0, 0, sz])))
        rects.append((ins, o + np.array([0, sy, 0.0]), np.array([sx, 0, 0.0]), np.array([0, 0, sz])))
        rects.append((ins, o, np.array([0, sy, 0.0]), np.array([0, 0, sz])))
        rects.append((ins, o + np.array([sx, 0, 0.0]), np.array([0, sy, 0.0]), np.array([0, 0, sz])))

    areas = np.array([np.linalg.norm(np.cross(u, v)) for _, _, u, v in rects])
    n_unique = n_points - int(round(dup_frac * n_points))
    counts = np.floor(areas / areas.sum() * n_unique).astype(np.int64)
    counts[0] += n_unique - counts.sum()

    xyz_l, ins_l, seg_key_l = [], [], []
    # segment cell size: ~pts_per_seg points per cell at this scene's surface density (0.35-0.5 m at the
    # reference density)
    g0 = float(np.sqrt(pts_per_seg / (n_unique / areas.sum())))
    for r, (ins, o, u, v) in enumerate(rects):
        p, uv = _sample_rect(rng, int(counts[r]), o, u, v)
        g = g0 * rng.uniform(0.9, 1.15)
        cell = np.floor(uv / g).astype(np.int64)
        xyz_l.append(p)
        ins_l.append(np.full(p.shape[0], ins, np.int64))
        seg_key_l.append(r * 1_000_000 + cell[:, 0] * 1000 + cell[:, 1])
    xyz = np.concatenate(xyz_l, 0)
    xyz += rng.normal(0.0, noise, xyz.shape)
    ins = np.concatenate(ins_l)
    seg_key = np.concatenate(seg_key_l)

    # exact duplicate points (the reference resamples every scene to a fixed size by tiling vertices,
    # dataset/scannet/util.py:669-681, so duplicates are normal and exercise every tie rule)
    n_dup = n_points - n_unique
    if n_dup > 0:
        src = rng.integers(0, n_unique, n_dup)
        xyz = np.concatenate([xyz, xyz[src]], 0)
        ins = np.concatenate([ins, ins[src]])
        seg_key = np.concatenate([seg_key, seg_key[src]])
    rgb_src = rng.uniform(-1, 1, (n_unique, 3))
    rgb = np.concatenate([rgb_src, rgb_src[src]], 0) if n_dup > 0 else rgb_src

    perm = rng.permutation(n_points)
    xyz, ins, seg_key, rgb = xyz[perm], ins[perm], seg_key[perm], rgb[perm]
    xyz = xyz.astype(np.float32)
    rgb = rgb.astype(np.float32)

    # ---- over-segmentation: merge undersized segments into the nearest big one of the same instance
    _, seg = np.unique(seg_key, return_inverse=True)
    min_pts = max(20, pts_per_seg // 8)
    for _ in range(8):
        cnt = np.bincount(seg)
        small = np.nonzero((cnt < min_pts) & (cnt > 0))[0]
        if small.size == 0:
            break
        S = cnt.shape[0]
        cen = np.stack([np.bincount(seg, xyz[:, d].astype(np.float64), S) for d in range(3)], 1) / np.maximum(cnt, 1)[:, None]
        seg_ins = np.zeros(S, np.int64)
        seg_ins[seg] = ins
        big = np.nonzero(cnt >= min_pts)[0]
        remap = np.arange(S)
        for s in small:
            cand = big[seg_ins[big] == seg_ins[s]]
            if cand.size == 0:
                cand = big
            remap[s] = cand[np.argmin(((cen[cand] - cen[s]) ** 2).sum(1))]
        seg = remap[seg]
        _, seg = np.unique(seg, return_inverse=True)

    # optional tiny (<5 point) segments to exercise the small-cluster sweep (model.py:228-239): carve
    # 3 nearby points out of an interior of a big segment; they stay unlabeled so a merge is always
    # admissible (no label veto) and the reference loop terminates.
    if n_small_segs > 0:
        tree0 = cKDTree(xyz.astype(np.float64))
        nxt = seg.max() + 1
        used = np.zeros(n_points, bool)
        centers = rng.choice(n_points, n_small_segs * 4, replace=False)
        made = 0
        for c in centers:
            _, nb = tree0.query(xyz[c].astype(np.float64), k=3)
            nb = np.unique(nb)
            if nb.size < 3 or used[nb].any() or np.unique(seg[nb]).size != 1:
                continue
            used[nb] = True
            seg[nb] = nxt
            nxt += 1
            made += 1
            if made == n_small_segs:
                break
        _, seg = np.unique(seg, return_inverse=True)

    # segments enumerated by ascending smallest member (= root point id), members ascending
    order = np.argsort(seg, kind="stable")
    seg_sorted = seg[order]
    starts = np.concatenate([[0], np.nonzero(np.diff(seg_sorted))[0] + 1, [n_points]])
    roots = order[starts[:-1]]
    seg_rank = np.argsort(np.argsort(roots))
    new_id = seg_rank[seg]                     # segment index in ascending-root order
    order = np.argsort(new_id, kind="stable")
    cnt = np.bincount(new_id)
    seg_offsets = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    seg_members = order.astype(np.int32)
    S = cnt.shape[0]

    # ---- point adjacency: symmetric 3-NN, rows sorted, unique-lexicographic (util.py:771-811)
    tree = cKDTree(xyz.astype(np.float64))
    _, nb = tree.query(xyz.astype(np.float64), k=4)
    src_i = np.repeat(np.arange(n_points), 4)
    dst_i = nb.reshape(-1)
    keep = src_i != dst_i
    e = np.stack([np.minimum(src_i[keep], dst_i[keep]), np.maximum(src_i[keep], dst_i[keep])], 1)
    adj = np.unique(e, axis=0).astype(np.int64)

    # ---- labels
    n_ins = int(ins.max()) + 1
    sem_of_ins = SEM_VALID_CLASS_IDS[np.arange(n_ins) % 20]     # 1..40
    real_label = np.stack([sem_of_ins[ins], ins + 1], 1).astype(np.int64)
    weak = np.full((n_points, 2), -1, np.int64)
    # an instance may span segments that the merge step re-homed; label the largest segment that
    # is majority-owned by the instance (style "maxseg", util.py:332-335)
    seg_ins_major = np.zeros(S, np.int64)
    for s in range(S):
        m = seg_members[seg_offsets[s]:seg_offsets[s + 1]]
        seg_ins_major[s] = np.bincount(ins[m]).argmax()
    for i in range(n_ins):
        cand = np.nonzero((seg_ins_major == i) & (cnt >= 5))[0]
        if cand.size == 0:
            continue
        s = cand[np.argmax(cnt[cand])]
        m = seg_members[seg_offsets[s]:seg_offsets[s + 1]]
        weak[m, 0] = sem_of_ins[i] - 1
        weak[m, 1] = i

    data = np.concatenate([xyz, rgb], 1).astype(np.float32)
    ident = np.arange(n_points, dtype=np.int64)
    return Scene(name=name or f"scene{seed:04d}_00", data=data, weak_label=weak, seg_members=seg_members,
                 seg_offsets=seg_offsets, adj=adj, real_label=real_label, unmap=ident.copy(), map=ident.copy(),
                 meta=dict(seed=seed, n_points=n_points, n_segments=int(S), n_instances=n_ins, side=float(side)))


def write_scene_tree(root: str, scenes: list[Scene], label_style: str = "manual") -> None:
    """Write the on-disk tree SegModel.forward and ScanNet(Dataset) read."""
    import torch

    base = os.path.join(root, "dataset", "scannet")
    os.makedirs(base, exist_ok=True)
    with open(os.path.join(base, "scannetv2_train.txt"), "w") as f:
        for sc in scenes:
            f.write(sc.name + "\n")
    for idx, sc in enumerate(scenes):
        n = sc.name

        def d(*parts):
            p = os.path.join(base, *parts, n)
            os.makedirs(p, exist_ok=True)
            return p

        p = d("data", "resampled")
        torch.save(torch.from_numpy(sc.data.copy()), os.path.join(p, n + ".pcl.pth"))
        torch.save(torch.tensor([idx], dtype=torch.long), os.path.join(p, n + ".info.pth"))
        torch.save(torch.from_numpy(sc.map.copy()), os.path.join(p, n + ".map.pth"))
        torch.save(torch.from_numpy(sc.unmap.copy()), os.path.join(p, n + ".unmap.pth"))
        torch.save(torch.from_numpy(sc.adj.copy()), os.path.join(d("adj", "mesh", "resampled"), n + ".adj.pth"))
        with open(os.path.join(d("label", "real", "resampled"), n + ".seg.json"), "w") as f:
            json.dump(sc.seg_json(), f)
        torch.save(torch.from_numpy(sc.weak_label.copy()), os.path.join(d("label", "seg", label_style, "resampled"), n + ".label.pth"))
        torch.save(torch.from_numpy(sc.real_label.copy()), os.path.join(d("label", "real", "raw"), n + ".label.pth"))


def make_cloud(seed: int, n_points: int, *, batches: int = 1):
    """Noisy-sheet cloud for the KPConv operator set (grid subsample / radius neighbours)."""
    out, lens = [], []
    for b in range(batches):
        sc = make_scene(seed * 131 + b, n_points, dup_frac=0.0)
        out.append(sc.data[:, :3])
        lens.append(n_points)
    return np.concatenate(out, 0).astype(np.float32), np.array(lens, np.int32)
