"""Functional SegGroup forward (seggroup/model.py:684-932) on the CUDA kernels.

`forward_scene(scene, params, mode)` runs one scene — or a scene batch (`SceneDevice.concat`: several scenes as ONE block-diagonal
scene, every launch serving all of them, everything the reference defines per scene kept per scene) — through graph init ->
structural grouping layer -> two semantic grouping layers -> final clustering -> classifier, returning the loss tensors (train),
the per-layer pseudo labels and the metrics.  Differentiable ops are `torch.autograd.Function`s whose forward AND backward are
kernels of libseggroup_b200.so (EdgeConv, pooling, edge distances, GCN aggregation, the GCN `fc` GEMMs on tcgen05, the classifier
head + label-smoothed cross entropy); torch itself is used for allocation and elementwise glue on [S, C] tensors.

Host synchronisation: two 4-byte read-backs per clustering level (cluster count and edge count fix the shapes of everything
downstream) — the reference does the whole grouping on the host.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

I32 = torch.int32


@dataclass
class SceneDevice:
    """One scene resident in HBM (what `seggroup/data.py` + the side files of model.py:696-699 provide)."""
    data: torch.Tensor          # [N,6] f32
    weak_label: torch.Tensor    # [N,2] i32 (sem, ins)
    seg_off: torch.Tensor       # [S1+1] i32
    seg_members: torch.Tensor   # [N] i32
    adj0: torch.Tensor          # [E0,2] i32
    unmap: torch.Tensor | None  # [N_raw] i64 or None (identity)
    real_label: torch.Tensor | None = None   # [N_raw,2] i64 (sem, ins) for evaluate()
    name: str = ""
    split: ops.SceneSplit | None = None      # scene batch: several scenes concatenated block-diagonally (None = one scene)

    @property
    def n_points(self):
        return self.data.shape[0]

    @property
    def n_scenes(self):
        return 1 if self.split is None else self.split.n

    def ranges(self):
        """(point, level-1 segment, raw vertex) offsets of the scenes: three host lists of n_scenes + 1 entries."""
        if self.split is None:
            n_raw = self.unmap.numel() if self.unmap is not None else self.n_points
            return [0, self.n_points], [0, self.seg_off.numel() - 1], [0, n_raw]
        return self.split.pt_off, self.split.seg_off, self.split.raw_off

    @staticmethod
    def concat(scenes, data=None, weak_label=None):
        """Several resident scenes -> ONE block-diagonal scene batch (device-side concatenation, point / segment ids offset by
        the scene's base).  Scenes are independent in the reference (one scene per rank per step, train.py:92; BatchNorm
        statistics, graphs and labels are per scene): a batch shares launches, not state.
        data / weak_label: the already concatenated [sum N, 6] f32 / [sum N, 2] i32 rows of the scenes (a loader's batch tensor
        viewed flat) — then the per-scene copies are not concatenated again."""
        if len(scenes) == 1:
            return scenes[0]
        dev = scenes[0].data.device
        pt, sg, rw = [0], [0], [0]
        for sc in scenes:
            pt.append(pt[-1] + sc.n_points)
            sg.append(sg[-1] + sc.seg_off.numel() - 1)
            rw.append(rw[-1] + (sc.unmap.numel() if sc.unmap is not None else sc.n_points))
        cat = torch.cat
        # id arrays: one concatenation + ONE offset addition each (the offset of element e = first point of its scene, expanded from
        # a [B] vector) instead of an addition per scene and array: ~20 launches instead of ~45 for 8 scenes
        counts = [[sc.seg_off.numel() - 1 for sc in scenes], [sc.seg_members.numel() for sc in scenes], [sc.adj0.shape[0] for sc in scenes],
                  [sc.unmap.numel() if sc.unmap is not None else sc.n_points for sc in scenes]]
        ends = [np.cumsum(c).tolist() for c in counts]
        meta = torch.tensor(ends + [pt[:-1]], dtype=torch.int64).to(dev, non_blocking=True)         # one small upload
        base = meta[4]

        def shifted(parts, row, width=1):
            # scene of element e = number of scene ends <= e (a binary search per element over B boundaries: fully parallel).
            # torch.repeat_interleave(base, counts) builds the same vector with ONE thread block: 150-300 us per 1.2 M-element id array
            # (`compute_cuda_kernel<long>`, 1.8 % of the device time in profiles/r03z_launches_bench.md).
            v = cat(parts)
            n = v.shape[0]
            off = base[torch.bucketize(_arange(n, dev), meta[row], right=True)]
            return v.add_(off.to(v.dtype) if width == 1 else off.to(v.dtype).unsqueeze(1))

        seg_off = cat([scenes[0].seg_off[:1], shifted([sc.seg_off[1:] for sc in scenes], 0)])
        seg_members = shifted([sc.seg_members for sc in scenes], 1)
        adj0 = shifted([sc.adj0 for sc in scenes], 2, width=2)
        unmap = shifted([(sc.unmap if sc.unmap is not None else torch.arange(sc.n_points, device=dev)) for sc in scenes], 3)
        real = cat([sc.real_label for sc in scenes]) if all(sc.real_label is not None for sc in scenes) else None
        return SceneDevice(data=cat([sc.data for sc in scenes]) if data is None else data,
                           weak_label=cat([sc.weak_label for sc in scenes]) if weak_label is None else weak_label, seg_off=seg_off,
                           seg_members=seg_members, adj0=adj0, unmap=unmap, real_label=real,
                           name="+".join(sc.name for sc in scenes), split=ops.SceneSplit(pt, sg, rw, dev))

    @staticmethod
    def from_host(scene, device="cuda", non_blocking=False):
        """scene: seggroup_b200.synth.Scene (numpy) -> device tensors (int64 inputs narrowed to int32)."""
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(device, non_blocking=non_blocking)
        return SceneDevice(data=t(scene.data, torch.float32), weak_label=t(scene.weak_label, I32),
                           seg_off=t(scene.seg_offsets, I32), seg_members=t(scene.seg_members, I32),
                           adj0=t(scene.adj, I32), unmap=t(scene.unmap, torch.int64),
                           real_label=t(scene.real_label, torch.int64), name=scene.name)


_arange_cache = {}


def _arange(n, dev):
    """[0, n) int64 on `dev`, cached by the largest size asked for (batch assembly asks for the same few sizes every step)."""
    t = _arange_cache.get(str(dev))
    if t is None or t.numel() < n:
        t = _arange_cache[str(dev)] = torch.arange(n, dtype=torch.int64, device=dev)
    return t[:n]


# ------------------------------------------------------------------------------------------------
# autograd wrappers (forward and backward are both C-ABI kernels)
# ------------------------------------------------------------------------------------------------
class SegmentMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, offsets, members):
        out, arg = ops.segment_pool_max(feat.contiguous(), offsets, members)
        ctx.save_for_backward(arg)
        ctx.n_rows = feat.shape[0]
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g, _):
        (arg,) = ctx.saved_tensors
        return ops.segment_pool_max_bwd(g.contiguous(), arg, ctx.n_rows), None, None


class SegmentMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, offsets, members):
        feat = feat.contiguous()
        ctx.save_for_backward(offsets, members if members is not None else offsets.new_empty(0))
        ctx.n_rows, ctx.has_members = feat.shape[0], members is not None
        return ops.segment_pool_mean(feat, offsets, members)

    @staticmethod
    def backward(ctx, g):
        offsets, members = ctx.saved_tensors
        return ops.segment_pool_mean_bwd(g.contiguous(), offsets, members if ctx.has_members else None, ctx.n_rows), None, None


def aggregate_cluster_feature(Feat_old, offsets, members=None, use_avg=False):
    """seggroup/model.py:278-288 on a CSR of the new clusters (offsets [S+1], members = old-cluster rows in member-list order):
    [S,C] max features, or [S,2C] = (max, mean) with use_avg=True (a flag no caller of the reference sets)."""
    f1, _ = SegmentMaxFn.apply(Feat_old, offsets, members)
    if not use_avg:
        return f1
    return torch.cat([f1, SegmentMeanFn.apply(Feat_old, offsets, members)], dim=-1)


class EdgeDistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, adj, csr_off, csr_nbr, csr_eid):
        feat = feat.contiguous()
        d = ops.edge_dist(feat, adj)
        ctx.save_for_backward(feat, adj, d, csr_off, csr_nbr, csr_eid)
        return d

    @staticmethod
    def backward(ctx, g):
        feat, adj, d, off, nbr, eid = ctx.saved_tensors
        gf = torch.zeros_like(feat)
        ops.edge_dist_bwd(feat, adj, d, g.contiguous(), (off, nbr, eid), gf)
        return gf, None, None, None, None


class GcnAggFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, sims, adj, csr_off, csr_nbr, csr_eid):
        X = X.contiguous(); sims = sims.contiguous()
        AX, rs = ops.gcn_agg(X, sims, (csr_off, csr_nbr, csr_eid))
        ctx.save_for_backward(X, sims, AX, rs, adj, csr_off, csr_nbr, csr_eid)
        return AX

    @staticmethod
    def backward(ctx, g):
        X, sims, AX, rs, adj, off, nbr, eid = ctx.saved_tensors
        dX, dsims = ops.gcn_agg_bwd(g.contiguous(), X, AX, sims, rs, adj, (off, nbr, eid))
        return dX, dsims, None, None, None, None


class LinearReluFn(torch.autograd.Function):
    """relu(fc(x)) of the GCN layer (model.py:146-151, fc = nn.Linear(bias=False)) on the tcgen05 tensor cores (TF32 x 3, ReLU
    fused into the TMEM epilogue); backward = two more GEMMs of the same kernel:
        dX = dZ W          (A = dZ [S,Co],   B = W^T  [Ci,Co])
        dW = dZ^T X        (A = dZ^T [Co,S], B = X^T  [Ci,S], S zero-padded to a multiple of 4)."""
    @staticmethod
    def forward(ctx, x, W):
        x = x.contiguous()
        y = ops.gemm_tf32x3(x, W.contiguous(), relu=True)
        ctx.save_for_backward(x, W, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, y = ctx.saved_tensors
        dz = gy * (y > 0).to(gy.dtype)
        dx = ops.gemm_tf32x3(dz, W.t().contiguous())
        S = x.shape[0]
        pad = (-S) % 4
        dzt, xt = dz.t(), x.t()
        if pad:
            dzt, xt = F.pad(dzt, (0, pad)), F.pad(xt, (0, pad))
        dW = ops.gemm_tf32x3(dzt.contiguous(), xt.contiguous())
        return dx, dW


class ClassifierHeadFn(torch.autograd.Function):
    """Classifier head + label-smoothed cross entropy (model.py:154-166, 902-932, util.py:12-29) for all scenes of a batch in one
    launch (one CTA per scene: BatchNorm1d statistics are per scene), forward and backward.  Returns loss_raw [B,2], logits,
    and the per-scene batch statistics (for the running-stat update of the module's BatchNorm1d)."""
    @staticmethod
    def forward(ctx, feat6, W1, gamma, beta, W2, b2, g_off, gold, mask, drop_scale):
        feat6 = feat6.contiguous()
        o = ops.classifier_head_fwd(feat6, g_off, gold, W1.contiguous(), gamma.contiguous(), beta.contiguous(), W2.contiguous(), b2.contiguous(),
                                    mask, drop_scale)
        ctx.save_for_backward(feat6, W1, gamma, beta, W2, g_off, gold, o["hpre"], o["stats"], o["logits"], mask if mask is not None else feat6.new_empty(0))
        ctx.drop_scale, ctx.has_mask = drop_scale, mask is not None
        ctx.mark_non_differentiable(o["logits"], o["stats"])
        return o["loss_raw"], o["logits"], o["stats"]

    @staticmethod
    def backward(ctx, g_loss, *_):
        feat6, W1, gamma, beta, W2, g_off, gold, hpre, stats, logits, mask = ctx.saved_tensors
        gl = g_loss[:, 0].contiguous()
        dfeat, dW1, dg, db, dW2, db2 = ops.classifier_head_bwd(feat6, g_off, gold, W1.contiguous(), gamma.contiguous(), beta.contiguous(), W2.contiguous(),
                                                               mask if ctx.has_mask else None, ctx.drop_scale, hpre, stats, logits, gl)
        return dfeat, dW1, dg, db, dW2, db2, None, None, None, None


def _sum_scenes(parts):
    """Sum of the per-scene parameter gradients of a batch in scene order: one stack + one reduction per parameter instead of a
    chain of B - 1 additions (the step is host-bound between the clustering levels: 120 tiny launches per step otherwise)."""
    return parts[0] if len(parts) == 1 else torch.stack(parts).sum(0)


class Mlp1Fn(torch.autograd.Function):
    """MLP1 (model.py:65-80) on the segment clouds; BN statistics are returned for the running-stat update.
    seg_ranges: level-1 segment range of every scene of the batch — BatchNorm statistics are per scene (one scene per forward
    in the reference), so the kernels run once per range on row slices of the batch arrays."""
    @staticmethod
    def forward(ctx, clouds, W, gamma, beta, seg_ranges):
        S = clouds.shape[0]
        dev = clouds.device
        Wc, gc, bc = W.contiguous(), gamma.contiguous(), beta.contiguous()
        feat = torch.empty(S, 128, dtype=torch.float32, device=dev)
        knn = torch.empty(S, 64, 10, dtype=I32, device=dev)
        arg_pt = torch.empty(S, 64, dtype=I32, device=dev)
        stats, var, mom = [], [], []
        for lo, hi in seg_ranges:
            o = ops.mlp1_fwd(clouds[lo:hi], Wc, gc, bc, feat=feat[lo:hi], knn=knn[lo:hi], arg_pt=arg_pt[lo:hi])
            stats.append(o["stats"]); var.append(o["var"]); mom.append(o["mom"])
        stats, var, mom = torch.stack(stats), torch.stack(var), torch.stack(mom)
        ctx.save_for_backward(clouds, W, knn, arg_pt, stats, mom)
        ctx.seg_ranges = seg_ranges
        ctx.mark_non_differentiable(knn, stats, var)
        return feat, knn, stats, var                      # stats [B,4,64], var [B,64]

    @staticmethod
    def backward(ctx, g, *_):
        clouds, W, knn, arg_pt, stats, mom = ctx.saved_tensors
        g = g.contiguous()
        gW, gg, gb = [], [], []
        for b, (lo, hi) in enumerate(ctx.seg_ranges):       # fixed scene order -> deterministic sums
            w_, g_, b_ = ops.mlp1_bwd(g[lo:hi], clouds[lo:hi], knn[lo:hi], arg_pt[lo:hi], W, stats[b], mom[b])
            gW.append(w_); gg.append(g_); gb.append(b_)
        return None, _sum_scenes(gW).view_as(W), _sum_scenes(gg), _sum_scenes(gb), None


class EdgeConvPoolFn(torch.autograd.Function):
    """MLP2 / MLP3 (model.py:106-138) fused with the point -> segment max pooling that always follows
    (model.py:793, 834): returns the pooled [S,64] features.  pt_ranges / cl_ranges: point and cluster range of every scene of
    the batch; the EdgeConv kernels (BatchNorm over one scene's N*20 edges) run once per scene on row slices, `knn` holds
    scene-relative ids, the pooling runs once over the whole batch."""
    @staticmethod
    def forward(ctx, x9, knn, cl_pt_off, order, W1, g1, b1, W2, g2, b2, pt_ranges, cl_ranges):
        two = W2 is not None
        N = x9.shape[0]
        dev = x9.device
        out = torch.empty(N, 64, dtype=torch.float32, device=dev)
        argk = torch.empty(N, 64, dtype=torch.uint8, device=dev)
        cW1, cg1, cb1 = W1.contiguous(), g1.contiguous(), b1.contiguous()
        cW2, cg2, cb2 = (W2.contiguous(), g2.contiguous(), b2.contiguous()) if two else (None, None, None)
        per = []
        for lo, hi in pt_ranges:
            per.append(ops.edgeconv_fwd(x9[lo:hi], knn[lo:hi], cW1, cg1, cb1, cW2, cg2, cb2, out=out[lo:hi], argk=argk[lo:hi]))
        pooled, arg = ops.segment_pool_max(out, cl_pt_off, order)
        ctx.two, ctx.pt_ranges, ctx.cl_ranges = two, pt_ranges, cl_ranges
        st = lambda k: torch.stack([o[k] for o in per])
        saved = [x9, knn, arg, argk, W1, st("stats1"), st("mom1"), st("ctr")]
        if two:
            saved += [W2, st("stats2"), st("mom2")]
        ctx.save_for_backward(*saved)
        stats = torch.stack([st("stats1"), st("stats2")], 1) if two else st("stats1").unsqueeze(1)      # [B,L,4,64]
        var = torch.stack([st("var1"), st("var2")], 1) if two else st("var1").unsqueeze(1)              # [B,L,64]
        ctx.mark_non_differentiable(stats, var, out)
        return pooled, out, stats, var

    @staticmethod
    def backward(ctx, g, *_):
        sv = ctx.saved_tensors
        x9, knn, arg, argk, W1, stats1, mom1, ctr = sv[:8]
        two = ctx.two
        if two:
            W2, stats2, mom2 = sv[8:]
        g = g.contiguous()
        acc = {}
        for b, ((lo, hi), (c0, c1)) in enumerate(zip(ctx.pt_ranges, ctx.cl_ranges)):
            arg_b = arg[c0:c1] if lo == 0 else arg[c0:c1] - lo          # arg-max rows relative to the scene's first point
            if two:
                r = ops.edgeconv_bwd(g[c0:c1], arg_b, argk[lo:hi], x9[lo:hi], knn[lo:hi], W1, stats1[b], mom1[b], ctr[b], W2, stats2[b], mom2[b])
            else:
                r = ops.edgeconv_bwd(g[c0:c1], arg_b, argk[lo:hi], x9[lo:hi], knn[lo:hi], W1, stats1[b], mom1[b], ctr[b])
            for k, v in r.items():
                acc.setdefault(k, []).append(v)
        acc = {k: _sum_scenes(v) for k, v in acc.items()}
        if two:
            return (None, None, None, None, acc["gW1"].view_as(W1), acc["gg1"], acc["gb1"], acc["gW2"].view_as(W2), acc["gg2"], acc["gb2"], None, None)
        return (None, None, None, None, acc["gW1"].view_as(W1), acc["gg1"], acc["gb1"], None, None, None, None, None)


# ------------------------------------------------------------------------------------------------
# metrics (model.py:608-655) — sgb_evaluate (evaluate.cu): shared-memory histograms over classes / instance ids
# ------------------------------------------------------------------------------------------------
SEM_VALID = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39]
INS_VALID = [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39]


def evaluate(real_label, sem_pred, ins_pred, status=None):
    """-> (IoU_sem [1,2,40], IoU_ins [1,2,40], acc [4]) float32 on the device of the inputs (one fused library pass)."""
    o = ops.evaluate(real_label.contiguous(), sem_pred.contiguous(), ins_pred.contiguous(), SEM_VALID, INS_VALID, status)
    return o[:80].view(1, 2, 40), o[80:160].view(1, 2, 40), o[160:164]


# ------------------------------------------------------------------------------------------------
@dataclass
class ForwardResult:
    labels: dict = field(default_factory=dict)      # 'layer_1.seg' ... 'final.sem' -> int32 [N_raw] (device; a batch: all scenes concatenated)
    metrics: tuple | None = None                    # (IoU_sem, IoU_ins, acc) of the scene (a batch: of scene 0; all in metrics_scenes)
    loss_raw: torch.Tensor | None = None            # [B,2] (sum, count) per scene
    levels: list = field(default_factory=list)
    bn_stats: dict = field(default_factory=dict)    # prefix -> (batch mean [64], biased var [64], count) (a batch: scene 0)
    aux: dict = field(default_factory=dict)         # intermediates for the parity tests
    status: int = 0
    metrics_scenes: list = field(default_factory=list)   # per scene (IoU_sem, IoU_ins, acc)
    raw_off: list = field(default_factory=list)     # raw-vertex offsets of the scenes inside the label vectors
    bn_stats_scenes: dict = field(default_factory=dict)  # prefix -> (mean [B,64], var [B,64], counts list)

    def scene_labels(self, b):
        """label vectors of scene b of a batch"""
        lo, hi = self.raw_off[b], self.raw_off[b + 1]
        return {k: v[lo:hi] for k, v in self.labels.items()}


def _gcn(p, key, Fc, adj, csr, keep=None, tag=""):
    keep = keep or (lambda name, t: t)
    d = keep("d_" + tag, EdgeDistFn.apply(Fc, adj, *csr))
    sims = keep("sims_" + tag, torch.exp(-d * (1 / 8)))
    AX = keep("AX_" + tag, GcnAggFn.apply(Fc, sims, adj, *csr))
    return LinearReluFn.apply(AX, p[key])


def _pairs(off):
    return list(zip(off[:-1], off[1:]))


def forward_scene(sc: SceneDevice, p: dict, mode: str = "train", keep_aux: bool = False, dropout_mask=None,
                  export: bool = True, classifier=None, sweep_cap: int = 64) -> ForwardResult:
    """p: dict of parameter tensors keyed like the reference state_dict (mlp_1.conv1.0.weight, ...).
    classifier: optional callable Feat_6 -> logits (the nn.Module head of SegModel, which then owns its
    BatchNorm1d buffers and dropout RNG); without it the head is evaluated functionally from `p`.
    sc may be a scene batch (SceneDevice.concat): every launch of the graph / kNN / pooling / export kernels then serves all
    scenes, while everything the reference defines per scene stays per scene (BatchNorm statistics, the order-dependent
    grouping replays, arg-min columns, kNN lists, segment labels, metrics, the classifier head and its loss).
    dropout_mask: [I,128] bool (one scene) or a list of them (batch)."""
    res = ForwardResult()
    sem_infer = mode == "sem_infer"
    dev = sc.data.device
    status = torch.zeros(1, dtype=I32, device=dev)
    split = sc.split
    B = sc.n_scenes
    pt_off, seg_off_s, raw_off = sc.ranges()
    pt_ranges, seg_ranges = _pairs(pt_off), _pairs(seg_off_s)
    res.raw_off = list(raw_off)
    live = res.aux.setdefault("_live", {}) if keep_aux else None

    def keep(name, t):
        """parity tests: keep the (non-detached) stage tensor and its gradient"""
        if live is not None:
            if t.requires_grad:
                t.retain_grad()
            live[name] = t
        return t

    def put_labels(tag, L, seg=True):
        if not export:
            return
        s, i, m = ops.export_labels(sc.unmap, sop, L, want_seg=seg, split=split)
        if seg:
            res.labels[tag + ".seg"] = s
        res.labels[tag + ".ins"], res.labels[tag + ".sem"] = i, m

    def put_metrics(sem_key, ins_key):
        if sc.real_label is None or not export:
            return
        o = ops.evaluate_scenes(sc.real_label, res.labels[sem_key], res.labels[ins_key], raw_off, SEM_VALID, INS_VALID, status)
        res.metrics_scenes += [(r[:80].view(1, 2, 40), r[80:160].view(1, 2, 40), r[160:164]) for r in o]
        res.metrics = res.metrics_scenes[0]

    def put_bn(prefix, mean, var, counts):
        """mean / var [B,64]"""
        res.bn_stats[prefix] = (mean[0], var[0], counts[0])
        res.bn_stats_scenes[prefix] = (mean, var, counts)

    # ---- graph initialisation (model.py:712-738)
    sop, sos, uf = ops.scene_init(sc.seg_off, sc.seg_members, sc.weak_label)
    step = lambda mode, **kw: ops.level_step(mode, uf, sc.seg_off, sc.seg_members, sos, status, sweep_cap=sweep_cap, split=split, **kw)
    L1 = step(2, edges=sc.adj0, mapping=sop)
    adj_1 = L1.adj
    put_labels("layer_1", L1)
    res.levels.append(L1)

    # ---- structural grouping layer (model.py:747-783)
    cloud_idx, _ = ops.cluster_cloud_indices(sc.data, L1.order, L1.cl_pt_off, 64, status=status)
    clouds = ops.cluster_cloud_transform(sc.data, cloud_idx)
    Feat_1, knn_1, stats, var = Mlp1Fn.apply(clouds, p["mlp_1.conv1.0.weight"], p["mlp_1.bn1.weight"], p["mlp_1.bn1.bias"], seg_ranges)
    put_bn("mlp_1.bn1", stats[:, 0], var, [(hi - lo) * 640 for lo, hi in seg_ranges])
    d1 = ops.edge_dist(Feat_1.detach(), adj_1)
    L2 = step(0, old=L1, dist=d1, th=3.0 if sem_infer else 6.0)
    adj_2 = L2.adj
    keep("Feat_1", Feat_1)
    Feat_2, arg_2 = SegmentMaxFn.apply(Feat_1, L2.ch_off, L2.ch_list)
    keep("Feat_2", Feat_2)
    put_labels("layer_2", L2)
    res.levels.append(L2)
    if keep_aux:
        res.aux.update(adj_1=adj_1, cloud_idx_1=cloud_idx, data_1=clouds, knn_1=knn_1, Feat_1=Feat_1.detach(), dists_1=d1, adj_2=adj_2)
    if sem_infer:
        put_metrics("layer_2.sem", "layer_2.ins")
        res.status = L2.status
        return res

    # ---- semantic grouping layers (model.py:788-865)
    def semantic_layer(Lc, Feat_c, pre, gcn_key, tag, two):
        adj_c = Lc.adj
        knn = ops.cluster_knn(sc.data, Lc.order, Lc.cl_pt_off, 20, scene_pt_off=split.d_pt_off if B > 1 else None)
        x9 = ops.centralize(sc.data, Lc.order, Lc.cl_pt_off)
        W2 = p[pre + ".conv2.0.weight"] if two else None
        g2 = p[pre + ".bn2.weight"] if two else None
        b2 = p[pre + ".bn2.bias"] if two else None
        W1, g1, b1 = p[pre + ".conv1.0.weight"], p[pre + ".bn1.weight"], p[pre + ".bn1.bias"]
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (W1, g1, b1, W2, g2, b2)):
            fm, feat_pts, stats, var = EdgeConvPoolFn.apply(x9, knn, Lc.cl_pt_off, Lc.order, W1, g1, b1, W2, g2, b2,
                                                            pt_ranges, _pairs(Lc.scene_cl_off))
        else:
            # inference: nothing is kept for a backward pass
            feat_pts = torch.empty(x9.shape[0], 64, dtype=torch.float32, device=dev)
            per = [ops.edgeconv_fwd(x9[lo:hi], knn[lo:hi], W1.contiguous(), g1.contiguous(), b1.contiguous(),
                                    W2.contiguous() if two else None, g2.contiguous() if two else None, b2.contiguous() if two else None,
                                    want_argk=False, want_backward=False, out=feat_pts[lo:hi]) for lo, hi in pt_ranges]
            fm, _ = ops.segment_pool_max(feat_pts, Lc.cl_pt_off, Lc.order, want_argmax=False)
            st = lambda k: torch.stack([o[k] for o in per])
            stats = torch.stack([st("stats1"), st("stats2")], 1) if two else st("stats1").unsqueeze(1)
            var = torch.stack([st("var1"), st("var2")], 1) if two else st("var1").unsqueeze(1)
        counts = [(hi - lo) * 20 for lo, hi in pt_ranges]
        put_bn(pre + ".bn1", stats[:, 0, 0], var[:, 0], counts)
        if two:
            put_bn(pre + ".bn2", stats[:, 1, 0], var[:, 1], counts)
        keep("pool_" + tag, fm)
        Fc = keep("cat_" + tag, torch.cat([Feat_c, fm], dim=-1))
        Fg = keep("gcn_" + tag, _gcn(p, gcn_key, Fc, adj_c, Lc.csr, keep, tag))
        dd = ops.edge_dist(Fg.detach().contiguous(), adj_c)
        Ln = step(0, old=Lc, dist=dd, th=2.0)
        Fn, _ = SegmentMaxFn.apply(Fg, Ln.ch_off, Ln.ch_list)
        keep("Feat_" + str(int(tag) + 1), Fn)
        if keep_aux:
            res.aux.update({"knn_" + tag: knn, "Feat_mlp_" + tag: feat_pts, "Feat_gcn_" + tag: Fg.detach(), "dists_" + tag: dd,
                            "x9_" + tag: x9})
        return Ln, Fn

    L3, Feat_3 = semantic_layer(L2, Feat_2, "mlp_2", "gcn_2.fc.weight", "2", False)
    put_labels("layer_3", L3)
    L4, Feat_4 = semantic_layer(L3, Feat_3, "mlp_3", "gcn_3.fc.weight", "3", True)
    put_labels("layer_4", L4)
    res.levels += [L3, L4]
    if keep_aux:
        res.aux.update(adj_3=L3.adj, adj_4=L4.adj)

    # ---- final clustering, phase A (model.py:439-470).  In a batch the loop runs until NO scene changes: an iteration on a
    # ---- scene that has converged unions nothing (its unlabeled clusters have no neighbour left or none exist).
    Lo, Feat = L4, Feat_4
    count_old = Lo.S
    while True:
        dd = ops.edge_dist(Feat.detach().contiguous(), Lo.adj)
        Ln = step(1, old=Lo, dist=dd)
        Feat, _ = SegmentMaxFn.apply(Feat, Ln.ch_off, Ln.ch_list)
        Lo = Ln
        if Lo.S == count_old:
            break
        count_old = Lo.S
    res.aux["phaseA_clusters"] = Lo.S
    # phase B (model.py:472-509) only runs when an unlabeled cluster survives phase A
    if Lo.n_unlabeled > 0:
        Lo, Feat = _phase_b(sc, uf, sos, status, Lo, Feat, split, sweep_cap)
    L5, Feat_5 = Lo, Feat
    res.levels.append(L5)
    put_labels("final", L5, seg=False)
    if keep_aux:
        res.aux["Feat_5"] = Feat_5.detach()
        keep("Feat_5", Feat_5)
    put_metrics("final.sem", "final.ins")
    res.status = L5.status
    if mode == "ins_infer":
        return res

    # ---- classifier (model.py:902-932): per-instance max, MLP head, label-smoothed CE (sum) — per scene
    order, off, sem_gt, g_off_d, n_groups, g_min = ops.classifier_groups(L5.cl_ins, L5.cl_sem, L5.d_scene_cl_off, B)
    Feat_6, _ = SegmentMaxFn.apply(Feat_5, off, order)
    if g_min < 2:
        raise ValueError("Expected more than 1 value per channel when training (a scene with a single instance group: "
                         "BatchNorm1d of the classifier, model.py:157)")
    # dropout (model.py:159): Bernoulli keep mask drawn with torch's generator, one launch for the whole batch
    if classifier is not None:
        cp = {"classifier." + k: v for k, v in classifier.named_parameters()}
        drop_p = float(classifier.dp1.p) if classifier.training else 0.0
    else:
        cp, drop_p = p, 0.5
    if isinstance(dropout_mask, (list, tuple)):
        mask = torch.cat([m.to(torch.float32) for m in dropout_mask]).contiguous()
        drop_scale = 2.0
    elif dropout_mask is not None:
        mask, drop_scale = dropout_mask.to(torch.float32).contiguous(), 2.0
    elif drop_p > 0.0:
        mask = (torch.rand(n_groups, 128, device=dev) >= drop_p).to(torch.float32)
        drop_scale = 1.0 / (1.0 - drop_p)
    else:
        mask, drop_scale = None, 1.0
    loss_raw, logits, cstats = ClassifierHeadFn.apply(Feat_6, cp["classifier.linear1.weight"], cp["classifier.bn1.weight"], cp["classifier.bn1.bias"],
                                                      cp["classifier.linear2.weight"], cp["classifier.linear2.bias"], g_off_d, sem_gt, mask, drop_scale)
    res.loss_raw = loss_raw                                            # [B,2] = (sum, count) per scene (model.py:932 returns [1,2])
    res.bn_stats_scenes["classifier.bn1"] = (cstats[:, :128], cstats[:, 128:], loss_raw[:, 1].detach())
    if keep_aux:
        res.aux["logits"] = logits.detach()
    return res


def batch_loss(loss_raw, n_scenes=None):
    """mean over the scenes of loss_sum / loss_num (train.py:165-170 on one scene per rank + DDP's gradient average) from the
    [B,2] tensor the classifier head returns.  The instance count is a constant of the step: dividing by the detached column keeps
    the backward graph to one multiplication (the plain `(l[:,0] / l[:,1]).mean()` differentiates the count column as well: seven
    more launches on a host-bound stretch of the step)."""
    n = loss_raw.shape[0] if n_scenes is None else n_scenes
    w = torch.reciprocal(loss_raw[:, 1].detach() * float(n))
    return (loss_raw[:, 0] * w).sum()


def _phase_b(sc, uf, sos, status, Lo, Feat, split, sweep_cap=64):
    """model.py:472-509 — nearest labelled cluster by sampled-cloud distance for clusters phase A left unlabeled.
    Clouds: FPS kernel; candidate ranking: sgb_phase_b_rank (one CTA per unlabeled cluster, its own scene's clusters only);
    the order-dependent unions: one sequential kernel."""
    P = 1024
    cloud_idx, _ = ops.cluster_cloud_indices(sc.data, Lo.order, Lo.cl_pt_off, P, status=status)
    unl = torch.nonzero(Lo.cl_ins == -1).view(-1).to(I32)                  # ascending dense ids
    width = max(b - a for a, b in _pairs(Lo.scene_cl_off))
    if width > 4096:
        raise _lib_error("phase B with more than 4096 clusters in a scene is not supported")
    cand = ops.phase_b_rank(sc.data, cloud_idx, unl, Lo, width, split)
    ops.group_unlabeled_phase_b(unl, cand, Lo.roots, uf)
    Ln = ops.level_step(2, uf, sc.seg_off, sc.seg_members, sos, status, old=Lo, sweep_cap=sweep_cap, split=split)
    Feat, _ = SegmentMaxFn.apply(Feat, Ln.ch_off, Ln.ch_list)
    return Ln, Feat


def _lib_error(msg):
    from ._lib import SgbError
    return SgbError(msg)
