"""Functional SegGroup forward (seggroup/model.py:684-932) on the CUDA kernels.

`forward_scene(scene, params, mode)` runs one scene: graph init -> structural grouping layer ->
two semantic grouping layers -> final clustering -> classifier, returning the loss tensors (train),
the per-layer pseudo labels and the metrics.  Differentiable ops are `torch.autograd.Function`s whose
forward AND backward are kernels of libseggroup_b200.so; torch itself is used for allocation, the tiny
dense GEMMs of the GCN / classifier heads (plain library GEMMs) and elementwise glue on [S, C] tensors.

Host synchronisation: one 4-byte read-back per clustering level (the cluster count fixes the shapes of
everything downstream) and one per adjacency update — the reference does the whole grouping on the host.
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

    @property
    def n_points(self):
        return self.data.shape[0]

    @staticmethod
    def from_host(scene, device="cuda", non_blocking=False):
        """scene: seggroup_b200.synth.Scene (numpy) -> device tensors (int64 inputs narrowed to int32)."""
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(device, non_blocking=non_blocking)
        return SceneDevice(data=t(scene.data, torch.float32), weak_label=t(scene.weak_label, I32),
                           seg_off=t(scene.seg_offsets, I32), seg_members=t(scene.seg_members, I32),
                           adj0=t(scene.adj, I32), unmap=t(scene.unmap, torch.int64),
                           real_label=t(scene.real_label, torch.int64), name=scene.name)


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


class Mlp1Fn(torch.autograd.Function):
    """MLP1 (model.py:65-80) on the segment clouds; BN statistics are returned for the running-stat update."""
    @staticmethod
    def forward(ctx, clouds, W, gamma, beta):
        o = ops.mlp1_fwd(clouds, W.contiguous(), gamma.contiguous(), beta.contiguous())
        ctx.save_for_backward(clouds, W, o["knn"], o["arg_pt"], o["stats"], o["mom"])
        ctx.mark_non_differentiable(o["knn"], o["stats"], o["var"])
        return o["feat"], o["knn"], o["stats"], o["var"]

    @staticmethod
    def backward(ctx, g, *_):
        clouds, W, knn, arg_pt, stats, mom = ctx.saved_tensors
        gW, gg, gb = ops.mlp1_bwd(g.contiguous(), clouds, knn, arg_pt, W, stats, mom)
        return None, gW.view_as(W), gg, gb


class EdgeConvPoolFn(torch.autograd.Function):
    """MLP2 / MLP3 (model.py:106-138) fused with the point -> segment max pooling that always follows
    (model.py:793, 834): returns the pooled [S,64] features."""
    @staticmethod
    def forward(ctx, x9, knn, cl_pt_off, order, W1, g1, b1, W2, g2, b2):
        two = W2 is not None
        o = ops.edgeconv_fwd(x9, knn, W1.contiguous(), g1.contiguous(), b1.contiguous(),
                             W2.contiguous() if two else None, g2.contiguous() if two else None, b2.contiguous() if two else None)
        pooled, arg = ops.segment_pool_max(o["out"], cl_pt_off, order)
        ctx.two = two
        saved = [x9, knn, arg, o["argk"], W1, o["stats1"], o["mom1"], o["ctr"]]
        if two:
            saved += [W2, o["stats2"], o["mom2"]]
        ctx.save_for_backward(*saved)
        stats = torch.stack([o["stats1"], o["stats2"]]) if two else o["stats1"].unsqueeze(0)
        var = torch.stack([o["var1"], o["var2"]]) if two else o["var1"].unsqueeze(0)
        ctx.mark_non_differentiable(stats, var, o["out"])
        return pooled, o["out"], stats, var

    @staticmethod
    def backward(ctx, g, *_):
        sv = ctx.saved_tensors
        x9, knn, arg, argk, W1, stats1, mom1, ctr = sv[:8]
        if ctx.two:
            W2, stats2, mom2 = sv[8:]
            r = ops.edgeconv_bwd(g.contiguous(), arg, argk, x9, knn, W1, stats1, mom1, ctr, W2, stats2, mom2)
            return (None, None, None, None, r["gW1"].view_as(W1), r["gg1"], r["gb1"], r["gW2"].view_as(W2), r["gg2"], r["gb2"])
        r = ops.edgeconv_bwd(g.contiguous(), arg, argk, x9, knn, W1, stats1, mom1, ctr)
        return (None, None, None, None, r["gW1"].view_as(W1), r["gg1"], r["gb1"], None, None, None)


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
    labels: dict = field(default_factory=dict)      # 'layer_1.seg' ... 'final.sem' -> int32 [N_raw] (device)
    metrics: tuple | None = None                    # (IoU_sem, IoU_ins, acc)
    loss_raw: torch.Tensor | None = None            # [1,2] (sum, count)
    levels: list = field(default_factory=list)
    bn_stats: dict = field(default_factory=dict)    # prefix -> (batch mean [64], biased var [64], count)
    aux: dict = field(default_factory=dict)         # intermediates for the parity tests
    status: int = 0


def _gcn(p, key, Fc, adj, csr, keep=None, tag=""):
    keep = keep or (lambda name, t: t)
    d = keep("d_" + tag, EdgeDistFn.apply(Fc, adj, *csr))
    sims = keep("sims_" + tag, torch.exp(-d * (1 / 8)))
    AX = keep("AX_" + tag, GcnAggFn.apply(Fc, sims, adj, *csr))
    return F.relu(keep("Z_" + tag, F.linear(AX, p[key])))


def forward_scene(sc: SceneDevice, p: dict, mode: str = "train", keep_aux: bool = False, dropout_mask=None,
                  export: bool = True, classifier=None) -> ForwardResult:
    """p: dict of parameter tensors keyed like the reference state_dict (mlp_1.conv1.0.weight, ...).
    classifier: optional callable Feat_6 -> logits (the nn.Module head of SegModel, which then owns its
    BatchNorm1d buffers and dropout RNG); without it the head is evaluated functionally from `p`."""
    res = ForwardResult()
    sem_infer = mode == "sem_infer"
    dev = sc.data.device
    status = torch.zeros(1, dtype=I32, device=dev)
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
        s, i, m = ops.export_labels(sc.unmap, sop, L, want_seg=seg)
        if seg:
            res.labels[tag + ".seg"] = s
        res.labels[tag + ".ins"], res.labels[tag + ".sem"] = i, m

    # ---- graph initialisation (model.py:712-738)
    sop, sos, uf = ops.scene_init(sc.seg_off, sc.seg_members, sc.weak_label)
    step = lambda mode, **kw: ops.level_step(mode, uf, sc.seg_off, sc.seg_members, sos, status, **kw)
    L1 = step(2, edges=sc.adj0, mapping=sop)
    adj_1 = L1.adj
    put_labels("layer_1", L1)
    res.levels.append(L1)

    # ---- structural grouping layer (model.py:747-783)
    cloud_idx, _ = ops.cluster_cloud_indices(sc.data, L1.order, L1.cl_pt_off, 64, status=status)
    clouds = ops.cluster_cloud_transform(sc.data, cloud_idx)
    Feat_1, knn_1, stats, var = Mlp1Fn.apply(clouds, p["mlp_1.conv1.0.weight"], p["mlp_1.bn1.weight"], p["mlp_1.bn1.bias"])
    res.bn_stats["mlp_1.bn1"] = (stats[0], var, L1.S * 640)
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
        if sc.real_label is not None and export:
            res.metrics = evaluate(sc.real_label, res.labels["layer_2.sem"], res.labels["layer_2.ins"], status)
        res.status = L2.status
        return res

    # ---- semantic grouping layers (model.py:788-865)
    def semantic_layer(Lc, Feat_c, pre, gcn_key, tag, two):
        adj_c = Lc.adj
        knn = ops.cluster_knn(sc.data, Lc.order, Lc.cl_pt_off, 20)
        x9 = ops.centralize(sc.data, Lc.order, Lc.cl_pt_off)
        W2 = p[pre + ".conv2.0.weight"] if two else None
        g2 = p[pre + ".bn2.weight"] if two else None
        b2 = p[pre + ".bn2.bias"] if two else None
        W1, g1, b1 = p[pre + ".conv1.0.weight"], p[pre + ".bn1.weight"], p[pre + ".bn1.bias"]
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (W1, g1, b1, W2, g2, b2)):
            fm, feat_pts, stats, var = EdgeConvPoolFn.apply(x9, knn, Lc.cl_pt_off, Lc.order, W1, g1, b1, W2, g2, b2)
        else:
            # inference: nothing is kept for a backward pass
            o = ops.edgeconv_fwd(x9, knn, W1.contiguous(), g1.contiguous(), b1.contiguous(),
                                 W2.contiguous() if two else None, g2.contiguous() if two else None, b2.contiguous() if two else None,
                                 want_argk=False, want_backward=False)
            feat_pts = o["out"]
            fm, _ = ops.segment_pool_max(feat_pts, Lc.cl_pt_off, Lc.order, want_argmax=False)
            stats = torch.stack([o["stats1"], o["stats2"]]) if two else o["stats1"].unsqueeze(0)
            var = torch.stack([o["var1"], o["var2"]]) if two else o["var1"].unsqueeze(0)
        res.bn_stats[pre + ".bn1"] = (stats[0, 0], var[0], sc.n_points * 20)
        if two:
            res.bn_stats[pre + ".bn2"] = (stats[1, 0], var[1], sc.n_points * 20)
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

    # ---- final clustering, phase A (model.py:439-470)
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
        Lo, Feat = _phase_b(sc, uf, sos, status, Lo, Feat)
    L5, Feat_5 = Lo, Feat
    res.levels.append(L5)
    put_labels("final", L5, seg=False)
    if keep_aux:
        res.aux["Feat_5"] = Feat_5.detach()
        keep("Feat_5", Feat_5)
    if sc.real_label is not None and export:
        res.metrics = evaluate(sc.real_label, res.labels["final.sem"], res.labels["final.ins"], status)
    res.status = L5.status
    if mode == "ins_infer":
        return res

    # ---- classifier (model.py:902-932): per-instance max, MLP head, label-smoothed CE (sum)
    ins = L5.cl_ins.cpu().numpy()
    sem = L5.cl_sem.cpu().numpy()
    uniq = np.unique(ins)
    order = np.argsort(ins, kind="stable").astype(np.int32)
    off = np.concatenate([[0], np.cumsum([(ins == u).sum() for u in uniq])]).astype(np.int32)
    sem_gt = torch.as_tensor(np.array([sem[order[off[i]]] for i in range(len(uniq))], np.int64), device=dev)
    Feat_6, _ = SegmentMaxFn.apply(Feat_5, torch.as_tensor(off, device=dev), torch.as_tensor(order, device=dev))
    if classifier is not None:
        logits = classifier(Feat_6)
    else:
        h = F.linear(Feat_6, p["classifier.linear1.weight"])
        h = F.batch_norm(h, None, None, p["classifier.bn1.weight"], p["classifier.bn1.bias"], True, 0.1, 1e-5)
        h = F.leaky_relu(h, 0.2)
        if dropout_mask is None:
            h = F.dropout(h, 0.5, True)
        else:
            h = h * dropout_mask.to(h.dtype) * 2.0
        logits = F.linear(h, p["classifier.linear2.weight"], p["classifier.linear2.bias"])
    eps, n_class = 0.2, logits.size(1)
    one_hot = torch.zeros_like(logits).scatter(1, sem_gt.view(-1, 1), 1)
    one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)
    loss_sum = -(one_hot * F.log_softmax(logits, dim=1)).sum()
    res.loss_raw = torch.cat([loss_sum.view(1), torch.tensor([float(len(uniq))], device=dev)]).unsqueeze(0)
    if keep_aux:
        res.aux["logits"] = logits.detach()
    return res


def _phase_b(sc, uf, sos, status, Lo, Feat):
    """model.py:472-509 — nearest labelled cluster by sampled-cloud distance for clusters phase A left unlabeled.
    Clouds: FPS kernel; distances / sort: torch device ops on [n_unl, S, 1024]; the order-dependent unions: one kernel."""
    P = 1024
    cloud_idx, _ = ops.cluster_cloud_indices(sc.data, Lo.order, Lo.cl_pt_off, P, status=status)
    pts = sc.data[:, :3][cloud_idx.long().view(-1)].view(Lo.S, P, 3)
    unl = torch.nonzero(Lo.cl_ins == -1).view(-1)                           # ascending dense ids
    mean = pts[unl].mean(1).view(-1, 1, 1, 3)                               # [n_unl,1,1,3]
    dmin = ((mean - pts.unsqueeze(0)) ** 2).sum(3).min(-1)[0]               # [n_unl,S]
    cand = torch.sort(dmin, dim=1)[1].to(I32).contiguous()
    ops.group_unlabeled_phase_b(unl.to(I32).contiguous(), cand, Lo.roots, uf)
    Ln = ops.level_step(2, uf, sc.seg_off, sc.seg_members, sos, status, old=Lo)
    Feat, _ = SegmentMaxFn.apply(Feat, Ln.ch_off, Ln.ch_list)
    return Ln, Feat
