"""TEST INFRASTRUCTURE — CPU restatement (oracle) of the SegGroup pseudo-label hot path.

This file is the parity checker for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product (`seggroup_b200/`)
never does.  It restates `/root/reference/seggroup/model.py::SegModel.forward` (file:line cited per
function) on torch-CPU / numpy, written around a SEGMENT-level union-find instead of the reference's
point-level `DisjointSet` (a cluster's point list is always a concatenation of whole level-1 segment
lists, model.py:191, so the two are equivalent; `tests/test_oracle_vs_reference.py` pins that).

Pinning: the reference ships no golden vectors for this path -> the oracle is pinned
against outputs of the reference itself, executed in the build container by `oracle/ref_harness.py`;
the resulting fixtures are committed under `tests/golden/` (`oracle/make_golden.py` is the script).

Numerics follow the reference's CPU arithmetic: kNN scores use the exact fp32 form torch/MKL
produce for K=3 (fused multiply-add chain, probed bit-exact), FPS uses non-fused fp32 squared
distances, BatchNorm always uses batch statistics (the reference never calls .eval()).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

SEM_VALID_CLASS_IDS = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39])
INS_VALID_CLASS_IDS = np.array([3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 28, 33, 34, 36, 39])

SMALL_SWEEP_CAP = 64     # the reference loop (model.py:228-239) can spin forever; the oracle raises instead


# ------------------------------------------------------------------------------------------------
# parameters
# ------------------------------------------------------------------------------------------------
def init_params(seed: int = 1, bn_gamma_scale: float | None = None) -> dict:
    """Default-initialised parameters with the reference's state_dict keys (model.py:65-166, 676-681).

    Built from the same torch.nn layers in the same construction order as SegModel.__init__, so
    `torch.manual_seed(seed)` reproduces the reference's initial weights bit for bit."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd = {}

    def bn(prefix, c):
        sd[prefix + ".weight"] = torch.ones(c)
        sd[prefix + ".bias"] = torch.zeros(c)
        sd[prefix + ".running_mean"] = torch.zeros(c)
        sd[prefix + ".running_var"] = torch.ones(c)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    # mlp_1
    c = nn.Conv2d(6, 64, 1, bias=False)
    bn("mlp_1.bn1", 64); sd["mlp_1.conv1.0.weight"] = c.weight.detach().clone()
    c = nn.Conv2d(18, 64, 1, bias=False)
    bn("mlp_2.bn1", 64); sd["mlp_2.conv1.0.weight"] = c.weight.detach().clone()
    l = nn.Linear(192, 192, bias=False); sd["gcn_2.fc.weight"] = l.weight.detach().clone()
    c = nn.Conv2d(18, 64, 1, bias=False)
    bn("mlp_3.bn1", 64); sd["mlp_3.conv1.0.weight"] = c.weight.detach().clone()
    c = nn.Conv2d(64, 64, 1, bias=False)
    bn("mlp_3.bn2", 64); sd["mlp_3.conv2.0.weight"] = c.weight.detach().clone()
    l = nn.Linear(256, 256, bias=False); sd["gcn_3.fc.weight"] = l.weight.detach().clone()
    l = nn.Linear(256, 128, bias=False); sd["classifier.linear1.weight"] = l.weight.detach().clone()
    bn("classifier.bn1", 128)
    l = nn.Linear(128, 40); sd["classifier.linear2.weight"] = l.weight.detach().clone(); sd["classifier.linear2.bias"] = l.bias.detach().clone()
    if bn_gamma_scale is not None:
        sd["mlp_1.bn1.weight"] = sd["mlp_1.bn1.weight"] * bn_gamma_scale
    return sd


TRAINABLE = ["mlp_1.bn1.weight", "mlp_1.bn1.bias", "mlp_1.conv1.0.weight",
             "mlp_2.bn1.weight", "mlp_2.bn1.bias", "mlp_2.conv1.0.weight",
             "gcn_2.fc.weight",
             "mlp_3.bn1.weight", "mlp_3.bn1.bias", "mlp_3.conv1.0.weight",
             "mlp_3.bn2.weight", "mlp_3.bn2.bias", "mlp_3.conv2.0.weight",
             "gcn_3.fc.weight",
             "classifier.linear1.weight", "classifier.bn1.weight", "classifier.bn1.bias",
             "classifier.linear2.weight", "classifier.linear2.bias"]


def from_reference_state(sd: dict) -> dict:
    """Reference state_dicts hold every BN twice (`mlp_k.bn1.*` and `mlp_k.conv1.1.*`)."""
    return {k: v.detach().clone() for k, v in sd.items() if ".conv1.1." not in k and ".conv2.1." not in k}


# ------------------------------------------------------------------------------------------------
# segment-level union-find  (model.py:169-214)
# ------------------------------------------------------------------------------------------------
class SegUnionFind:
    """State per level-1 segment s (root point id = first member).  `cid[s]` = root segment of the
    cluster that contains s; `members[r]` = level-1 segments of root r in the reference's list order."""

    def __init__(self, seg_offsets, seg_members, weak_label):
        self.S = len(seg_offsets) - 1
        self.seg_offsets = np.asarray(seg_offsets, np.int64)
        self.seg_members = np.asarray(seg_members, np.int64)
        self.root_point = self.seg_members[self.seg_offsets[:-1]]
        self.cid = np.arange(self.S)
        self.members = [[s] for s in range(self.S)]
        self.pnum = np.diff(self.seg_offsets).astype(np.float64)          # model.py:176,721
        self.ins = np.array(weak_label[self.root_point, 1], np.int64)     # model.py:174,712
        self.sem = np.array(weak_label[self.root_point, 0], np.int64)

    def union(self, a, b):                                                # model.py:181-192
        if a == b:
            return
        if self.ins[a] != -1 and self.ins[b] != -1 and self.ins[a] != self.ins[b]:
            return
        for s in self.members[a]:
            self.cid[s] = b
        self.pnum[b] += self.pnum[a]
        if self.ins[a] != self.ins[b]:
            self.ins[b] = -self.ins[a] * self.ins[b]
            self.sem[b] = -self.sem[a] * self.sem[b]
        self.members[b].extend(self.members[a])
        self.members[a] = []

    def roots(self):
        """Ascending root segment == ascending root point id (model.py:209-214)."""
        return np.array([r for r in range(self.S) if self.members[r]], np.int64)

    def level(self):
        """(roots, dense index of every level-1 segment, per-cluster segment lists)."""
        roots = self.roots()
        dense = np.full(self.S, -1, np.int64)
        dense[roots] = np.arange(len(roots))
        return roots, dense[self.cid], [list(self.members[r]) for r in roots]

    def cluster_points(self, seg_list):
        return np.concatenate([self.seg_members[self.seg_offsets[s]:self.seg_offsets[s + 1]] for s in seg_list])


def update_adj(adj_old, new_of_old):
    """model.py:291-302 — map, drop self edges, sort each pair, unique rows (lexicographic)."""
    if len(adj_old) == 0:
        return np.zeros((0, 2), np.int64)
    e = new_of_old[np.asarray(adj_old, np.int64)]
    e = e[e[:, 0] != e[:, 1]]
    if len(e) == 0:
        return np.zeros((0, 2), np.int64)
    e = np.sort(e, axis=1)
    return np.unique(e, axis=0)


def group_nearby(uf: SegUnionFind, dist, adj, roots_old, th):
    """model.py:218-258.  dist: float32 array [A]; adj: [A,2] dense indices of the current level."""
    dist = np.asarray(dist, np.float32)
    thf = np.float32(th)
    ru = roots_old[adj[:, 0]] if len(adj) else np.zeros(0, np.int64)
    rv = roots_old[adj[:, 1]] if len(adj) else np.zeros(0, np.int64)
    for i in range(len(adj)):
        if dist[i] > thf:
            continue
        uf.union(uf.cid[ru[i]], uf.cid[rv[i]])
    sweeps = 0
    while True:
        attempted = False
        for i in range(len(adj)):
            c1, c2 = uf.cid[ru[i]], uf.cid[rv[i]]
            if uf.pnum[c1] < 5 or uf.pnum[c2] < 5:
                uf.union(c1, c2)
                attempted = True
        if not attempted:
            break
        sweeps += 1
        if sweeps > SMALL_SWEEP_CAP:
            raise RuntimeError("small-cluster sweep does not terminate (reference would hang, model.py:228-239)")
    if len(adj) == 0:
        return adj, adj
    conn = uf.cid[ru] == uf.cid[rv]
    return adj[conn], adj[~conn]


# ------------------------------------------------------------------------------------------------
# geometry helpers
# ------------------------------------------------------------------------------------------------
def fps_indices(pts, k):
    """model.py:329-395 with initial_idx=0, skip_initial=True.  pts [n,3] float32 -> k local indices."""
    pts = np.asarray(pts, np.float32)
    sel = np.zeros(k, np.int32)

    def sq(i):
        return ((pts[i][None, :] - pts) ** 2).sum(axis=1)

    mind = sq(0)
    sel[0] = np.argmax(mind)
    mind = sq(sel[0])
    for i in range(1, k):
        sel[i] = np.argmax(mind)
        mind = np.minimum(mind, sq(sel[i]))
    return sel


def cluster_cloud_indices(n, pts, P):
    """model.py:398-420 — local member indices making up the P-row cloud of one cluster."""
    rep, rem = P // n, P % n
    parts = [np.tile(np.arange(n, dtype=np.int64), rep)] if rep else []
    if rem > 0:
        choice = fps_indices(pts, rem)
        if choice[-1] == 0:                                   # trailing-zero fix, model.py:407-412
            j = 1
            while j <= rem and choice[-j] == 0:
                j += 1
            invalid = j - 1 if j <= rem else rem - 1
            # reference: loop runs j=1..rem and breaks at the first non-zero from the end; if none breaks j=rem
            if invalid == 0:
                raise ValueError("degenerate cluster (all picks 0) — the reference raises here")
            choice[-invalid:] = choice[:invalid]
        parts.append(choice.astype(np.int64))
    return np.concatenate(parts)


def knn_scores(x):
    """model.py:30-33 — x [B,3,n] -> [B,n,n] scores, larger = closer."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


def topk_idx(scores, k, tie="torch"):
    """tie='torch': torch.topk as the reference calls it (model.py:35; tie order implementation-defined);
    tie='canonical': (score desc, index asc) — the rule the CUDA path implements."""
    if tie == "torch":
        return scores.topk(k=k, dim=-1)[1]
    order = torch.sort(scores, dim=-1, descending=True, stable=True)[1]
    return order[..., :k]


def cluster_knn(xyz, clusters, k=20, tie="torch", row_chunk=4096):
    """model.py:512-522 — xyz [N,3] float32 tensor, clusters: list of point-id arrays -> [N,k] int64."""
    N = xyz.shape[0]
    out = torch.zeros(N, k, dtype=torch.long)
    for m in clusters:
        m_t = torch.as_tensor(m, dtype=torch.long)
        n = len(m)
        if k >= n:
            out[m_t, :n] = m_t.unsqueeze(0).repeat(n, 1)
        else:
            x = xyz[m_t].unsqueeze(0).transpose(2, 1)           # [1,3,n], same strides as the reference
            if n <= row_chunk:
                idx = topk_idx(knn_scores(x), k, tie).squeeze(0)
            else:
                # rows are independent and each score is a fixed fp32 expression of (i,j) -> chunking rows is exact
                xx = torch.sum(x ** 2, dim=1, keepdim=True)
                parts = []
                for s in range(0, n, row_chunk):
                    xr = x[:, :, s:s + row_chunk]
                    inner = -2 * torch.matmul(xr.transpose(2, 1), x)
                    sc = -xx - inner - xx[:, :, s:s + row_chunk].transpose(2, 1)
                    parts.append(topk_idx(sc, k, tie).squeeze(0))
                idx = torch.cat(parts, 0)
            out[m_t, :k] = m_t[idx.reshape(-1)].view(-1, k)
    return out


# ------------------------------------------------------------------------------------------------
# network pieces (functional; params = dict with the reference state_dict keys)
# ------------------------------------------------------------------------------------------------
def _bn2d(x, p, prefix, stats=None):
    out = F.batch_norm(x, None, None, p[prefix + ".weight"], p[prefix + ".bias"], True, 0.1, 1e-5)
    return out


def mlp1_forward(p, clouds, tie="torch"):
    """model.py:39-80.  clouds [S,64,6] -> [S,128]; also returns the kNN indices [S,64,10]."""
    x = clouds.transpose(2, 1)                                   # [S,6,64]
    B, C, n = x.shape
    idx = topk_idx(knn_scores(x[:, :3]), 10, tie)                 # [S,64,10]
    flat = (idx + torch.arange(B).view(-1, 1, 1) * n).view(-1)
    xt = x.transpose(2, 1).contiguous()
    feat = xt.view(B * n, -1)[flat, :].view(B, n, 10, C).permute(0, 3, 1, 2)
    feat[:, :3] = feat[:, :3] - torch.mean(feat[:, :3], dim=-1, keepdim=True).repeat(1, 1, 1, 10)
    feat[:, :3] *= 10
    y = F.leaky_relu(_bn2d(F.conv2d(feat, p["mlp_1.conv1.0.weight"]), p, "mlp_1.bn1"), 0.2)
    y = y.max(dim=-1)[0]
    return torch.cat([y.max(dim=-1)[0], y.mean(dim=-1)], dim=-1), idx


def edge_features(x9, idx):
    """model.py:83-103.  x9 [N,9], idx [N,k] -> [1,18,N,k]."""
    N, k = idx.shape
    nb = x9[idx.reshape(-1)].view(N, k, 9)
    ctr = x9.view(N, 1, 9).repeat(1, k, 1)
    return torch.cat((nb - ctr, ctr), dim=2).unsqueeze(0).permute(0, 3, 1, 2)


def mlp2_forward(p, x9, idx):
    """model.py:106-118 -> [N,64]"""
    y = F.leaky_relu(_bn2d(F.conv2d(edge_features(x9, idx), p["mlp_2.conv1.0.weight"]), p, "mlp_2.bn1"), 0.2)
    return y.max(dim=-1)[0].squeeze(0).transpose(1, 0).contiguous()


def mlp3_forward(p, x9, idx):
    """model.py:121-138 -> [N,64]"""
    y = F.leaky_relu(_bn2d(F.conv2d(edge_features(x9, idx), p["mlp_3.conv1.0.weight"]), p, "mlp_3.bn1"), 0.2)
    y = F.leaky_relu(_bn2d(F.conv2d(y, p["mlp_3.conv2.0.weight"]), p, "mlp_3.bn2"), 0.2)
    return y.max(dim=-1)[0].squeeze(0).transpose(1, 0).contiguous()


def segment_max(feat, groups):
    """model.py:278-288 (use_avg is False at every call site).  groups: list of index arrays."""
    return torch.cat([torch.max(feat[torch.as_tensor(g, dtype=torch.long)], dim=0, keepdim=True)[0] for g in groups], dim=0)


def edge_distance(feat, adj):
    """model.py:269-274 — F.pairwise_distance adds eps=1e-6 to the difference."""
    adj = torch.as_tensor(adj, dtype=torch.long)
    if adj.numel() == 0:
        return feat.new_zeros(0)
    return F.pairwise_distance(feat[adj[:, 0], :], feat[adj[:, 1], :])


class _ReluWithActiveSet(torch.autograd.Function):
    """relu(z) whose BACKWARD uses a given active set instead of z > 0.  The derivative of ReLU jumps at 0, so two
    implementations whose pre-activations differ in the last bits disagree on the gradient of an entry that is zero to
    rounding; gradient parity is therefore defined for a common active set (the tests check that the two sets differ only where
    |z| is within rounding of zero)."""
    @staticmethod
    def forward(ctx, z, mask):
        ctx.save_for_backward(mask)
        return F.relu(z)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask.to(g.dtype), None


def gcn_forward(w, feat, adj, sims, keep=None, tag="", relu_mask=None):
    """model.py:305-309 + 141-151 — dense A = I + sym(sims), row-normalise, relu(fc(A X))."""
    keep = keep or (lambda name, t: t)
    S = feat.shape[0]
    A = torch.eye(S)
    adj = torch.as_tensor(adj, dtype=torch.long)
    if adj.numel():
        A[adj[:, 0], adj[:, 1]] = sims
        A[adj[:, 1], adj[:, 0]] = sims
    A = A / A.sum(1, keepdim=True).repeat(1, S)
    AX = keep("AX_" + tag, A.mm(feat))
    Z = keep("Z_" + tag, F.linear(AX, w))
    return F.relu(Z) if relu_mask is None else _ReluWithActiveSet.apply(Z, relu_mask)


def centralized(data, clusters):
    """model.py:429-436 — append xyz minus the cluster mean."""
    ctr = data[:, :3].clone()
    for m in clusters:
        m_t = torch.as_tensor(m, dtype=torch.long)
        ctr[m_t] -= ctr[m_t].mean(0)
    return torch.cat([data, ctr], dim=1)


def smoothed_ce_sum(pred, gold):
    """seggroup/util.py:12-29 (eps 0.2, summed)."""
    eps, n_class = 0.2, pred.size(1)
    one_hot = torch.zeros_like(pred).scatter(1, gold.view(-1, 1), 1)
    one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)
    return -(one_hot * F.log_softmax(pred, dim=1)).sum()


def evaluate(real_label, sem_pred, ins_pred):
    """model.py:608-655 -> (IoU_sem [1,2,40], IoU_ins [1,2,40], acc [4]) as float32 numpy."""
    sem_true, ins_true = real_label[:, 0], real_label[:, 1]
    v = sem_true != 0
    sem_true, ins_true, sem_pred, ins_pred = sem_true[v], ins_true[v], sem_pred[v], ins_pred[v]
    iou_sem = np.zeros((1, 2, 40), np.float32)
    for c in range(40):
        iou_sem[0, 0, c] = np.sum((sem_pred == c + 1) & (sem_true == c + 1))
        iou_sem[0, 1, c] = np.sum((sem_pred == c + 1) | (sem_true == c + 1))
    iou_ins = np.zeros((1, 2, 40), np.float32)
    for ins in np.unique(ins_pred):
        if ins == -1:
            continue
        sem = sem_pred[np.where(ins_pred == ins)[0][0]]
        iou_ins[0, 0, sem - 1] += np.sum((ins_pred == ins) & (ins_true == ins))
        iou_ins[0, 1, sem - 1] += np.sum((ins_pred == ins) | (ins_true == ins))

    def acc(t, p):
        return float(np.mean(t == p)) if len(t) else float("nan")

    sv = np.isin(sem_true, SEM_VALID_CLASS_IDS)
    # the reference concatenates per-class index lists; the mean over that multiset equals the masked mean
    iv = np.isin(ins_true, INS_VALID_CLASS_IDS)
    a = np.array([acc(sem_true, sem_pred), acc(ins_true, ins_pred), acc(sem_true[sv], sem_pred[sv]), acc(ins_true[iv], ins_pred[iv])], np.float32)
    return iou_sem, iou_ins, a


# ------------------------------------------------------------------------------------------------
# the forward pass
# ------------------------------------------------------------------------------------------------
class Level:
    def __init__(self, uf: SegUnionFind):
        self.roots, self.seg2cluster, self.seg_lists = uf.level()
        self.S = len(self.roots)
        self.points = [uf.cluster_points(sl) for sl in self.seg_lists]
        self.ins = uf.ins[self.roots].copy()
        self.sem = uf.sem[self.roots].copy()
        self.root_point = uf.root_point[self.roots]

    def point_labels(self, N, unmap):
        """The three label vectors the reference writes per layer (model.py:525-605), raw-vertex order."""
        seg = np.full(N, -1, np.int64); ins = np.full(N, -1, np.int64); sem = np.full(N, -1, np.int64)
        for c in range(self.S):
            seg[self.points[c]] = self.root_point[c]
            if self.ins[c] != -1:
                ins[self.points[c]] = self.ins[c] + 1
            if self.sem[c] != -1:
                sem[self.points[c]] = self.sem[c] + 1
        return seg[unmap], ins[unmap], sem[unmap]


def children_groups(new: Level, old: Level):
    """cluster_new_to_old (model.py:760-768): old dense indices per new cluster, ascending."""
    g = [[] for _ in range(new.S)]
    for j in range(old.S):
        g[new.seg2cluster[old.roots[j]]].append(j)
    return g


def forward(scene, params, mode="train", tie="torch", dropout_mask=None, want_grads=False, relu_masks=None):
    """Restatement of SegModel.forward (model.py:684-932) for one scene.

    scene: seggroup_b200.synth.Scene (or anything with the same fields).  params: state_dict-keyed dict.
    relu_masks: optional {"2": bool [S2,192], "3": bool [S3,256]} active sets for the BACKWARD of the two GCN ReLUs.
    Returns a dict of outputs and intermediates; with want_grads the trainable params get .grad."""
    p = {k: v.clone() for k, v in params.items()}
    if want_grads:
        for k in TRAINABLE:
            p[k].requires_grad_(True)
    out = {}
    live = out["_live"] = {}        # non-detached stage tensors (retain_grad) for stage-by-stage gradient comparisons

    def keep(name, t):
        if want_grads and t.requires_grad:
            t.retain_grad()
        live[name] = t
        return t
    data = torch.from_numpy(np.ascontiguousarray(scene.data))
    N = data.shape[0]
    unmap = np.asarray(scene.unmap, np.int64)
    labels = {}
    sem_infer = mode == "sem_infer"

    # graph initialisation, model.py:712-733
    uf = SegUnionFind(scene.seg_offsets, scene.seg_members, np.asarray(scene.weak_label))
    L1 = Level(uf)
    seg_of_point = np.empty(N, np.int64)
    for s in range(uf.S):
        seg_of_point[uf.seg_members[uf.seg_offsets[s]:uf.seg_offsets[s + 1]]] = s
    adj_1 = update_adj(scene.adj, seg_of_point)
    labels["layer_1.seg"], labels["layer_1.ins"], labels["layer_1.sem"] = L1.point_labels(N, unmap)
    out["adj_1"] = adj_1

    # structural grouping layer, model.py:747-770
    cloud_idx = []
    clouds = []
    for m in L1.points:
        li = cluster_cloud_indices(len(m), scene.data[m, :3], 64)
        cloud_idx.append(m[li])
        c = data[torch.as_tensor(m[li])].clone()
        c[:, :3] -= c[:, :3].mean(0)
        c[:, :3] /= torch.abs(c[:, :3]).max()
        clouds.append(c.unsqueeze(0))
    clouds = torch.cat(clouds, 0)
    out["cloud_idx_1"] = np.stack(cloud_idx)
    out["data_1"] = clouds.clone()
    Feat_1, knn_1 = mlp1_forward(p, clouds, tie)
    out["knn_1"] = knn_1
    out["Feat_1"] = Feat_1.detach().clone()
    keep("Feat_1", Feat_1)
    d1 = edge_distance(Feat_1, adj_1)
    out["dists_1"] = d1.detach().clone()
    _, adj_unc = group_nearby(uf, d1.detach().numpy(), adj_1, L1.roots, 3 if sem_infer else 6)
    L2 = Level(uf)
    adj_2 = update_adj(adj_unc, L2.seg2cluster[L1.roots])
    Feat_2 = keep("Feat_2", segment_max(Feat_1, children_groups(L2, L1)))
    labels["layer_2.seg"], labels["layer_2.ins"], labels["layer_2.sem"] = L2.point_labels(N, unmap)
    out["adj_2"] = adj_2
    out["levels"] = [L1, L2]
    if sem_infer:
        out["labels"] = labels
        out["metrics"] = evaluate(np.asarray(scene.real_label), labels["layer_2.sem"], labels["layer_2.ins"])
        return out

    def semantic_layer(Lc, Feat_c, adj_c, mlp, gcn_key, tag):
        """model.py:788-815 / 829-856"""
        knn = cluster_knn(data[:, :3], Lc.points, 20, tie)
        x9 = centralized(data, Lc.points)
        fm = mlp(p, x9, knn)
        out["knn_" + tag] = knn
        out["Feat_mlp_" + tag] = fm.detach().clone()
        fm = keep("pool_" + tag, segment_max(fm, Lc.points))
        Fc = keep("cat_" + tag, torch.cat([Feat_c, fm], dim=-1))
        sims = keep("sims_" + tag, torch.exp(-keep("d_" + tag, edge_distance(Fc, adj_c)) * (1 / 8)))
        Fc = keep("gcn_" + tag, gcn_forward(p[gcn_key], Fc, adj_c, sims, keep, tag, (relu_masks or {}).get(tag)))
        out["Feat_gcn_" + tag] = Fc.detach().clone()
        dd = edge_distance(Fc, adj_c)
        out["dists_" + tag] = dd.detach().clone()
        _, unc = group_nearby(uf, dd.detach().numpy(), adj_c, Lc.roots, 2)
        Ln = Level(uf)
        adj_n = update_adj(unc, Ln.seg2cluster[Lc.roots])
        Fn = keep("Feat_" + str(int(tag) + 1), segment_max(Fc, children_groups(Ln, Lc)))
        return Ln, Fn, adj_n

    L3, Feat_3, adj_3 = semantic_layer(L2, Feat_2, adj_2, mlp2_forward, "gcn_2.fc.weight", "2")
    labels["layer_3.seg"], labels["layer_3.ins"], labels["layer_3.sem"] = L3.point_labels(N, unmap)
    L4, Feat_4, adj_4 = semantic_layer(L3, Feat_3, adj_3, mlp3_forward, "gcn_3.fc.weight", "3")
    labels["layer_4.seg"], labels["layer_4.ins"], labels["layer_4.sem"] = L4.point_labels(N, unmap)
    out["adj_3"], out["adj_4"] = adj_3, adj_4
    out["levels"] += [L3, L4]

    # final clustering, model.py:439-509
    Lo, Feat, adj = L4, Feat_4, adj_4
    count_old = Feat.shape[0]
    while True:
        S = Feat.shape[0]
        dm = torch.ones(S, S) * 1000
        if len(adj):
            dd = edge_distance(Feat, adj)
            a = torch.as_tensor(adj, dtype=torch.long)
            dm[a[:, 0], a[:, 1]] = dd
            dm[a[:, 1], a[:, 0]] = dd
        amin = torch.min(dm, dim=-1)[1].numpy()
        for i in range(S):
            c1 = uf.cid[Lo.roots[i]]
            if uf.ins[c1] != -1:
                continue
            uf.union(c1, uf.cid[Lo.roots[amin[i]]])
        Ln = Level(uf)
        adj = update_adj(adj, Ln.seg2cluster[Lo.roots])
        Feat = segment_max(Feat, children_groups(Ln, Lo))
        Lo = Ln
        if Feat.shape[0] == count_old:
            break
        count_old = Feat.shape[0]
    out["phaseA_clusters"] = Lo.S
    unlabeled = [i for i in range(Lo.S) if uf.ins[uf.cid[Lo.roots[i]]] == -1]
    if unlabeled:
        pts1024 = []
        for m in Lo.points:
            li = cluster_cloud_indices(len(m), scene.data[m, :3], 1024)
            pts1024.append(data[torch.as_tensor(m[li]), :3].unsqueeze(0))
        pts1024 = torch.cat(pts1024, 0)                                      # [S,1024,3]
        for i in range(Lo.S):
            c1 = uf.cid[Lo.roots[i]]
            if uf.ins[uf.cid[c1]] != -1:
                continue
            mean = torch.mean(pts1024[i], dim=0).unsqueeze(0)
            dmin = torch.min(((mean - pts1024) ** 2).sum(dim=2), dim=-1)[0]
            for j in torch.sort(dmin)[1].tolist():
                if j == i:
                    continue
                c2 = uf.cid[Lo.roots[j]]
                if uf.ins[uf.cid[c2]] == -1:
                    continue
                uf.union(c1, c2)
        Ln = Level(uf)
        adj = update_adj(adj, Ln.seg2cluster[Lo.roots])
        Feat = segment_max(Feat, children_groups(Ln, Lo))
        Lo = Ln
    L5, Feat_5 = Lo, Feat
    out["levels"].append(L5)
    out["Feat_5"] = Feat_5.detach().clone()
    keep("Feat_5", Feat_5)
    _, labels["final.ins"], labels["final.sem"] = L5.point_labels(N, unmap)
    out["labels"] = labels
    out["metrics"] = evaluate(np.asarray(scene.real_label), labels["final.sem"], labels["final.ins"])
    if mode == "ins_infer":
        return out

    # classifier, model.py:902-932
    ins_list = L5.ins
    groups, sem_gt = [], []
    for ins in np.unique(ins_list):
        idx = np.where(ins_list == ins)[0]
        groups.append(idx)
        sem_gt.append(int(L5.sem[idx[0]]))
    Feat_6 = torch.cat([torch.max(Feat_5[torch.as_tensor(g)], dim=0, keepdim=True)[0] if len(g) > 1 else Feat_5[torch.as_tensor(g)] for g in groups], 0)
    h = F.linear(Feat_6, p["classifier.linear1.weight"])
    h = F.batch_norm(h, None, None, p["classifier.bn1.weight"], p["classifier.bn1.bias"], True, 0.1, 1e-5)
    h = F.leaky_relu(h, 0.2)
    if dropout_mask is None:
        h = F.dropout(h, 0.5, True)
    else:
        h = h * dropout_mask.to(h.dtype) * 2.0
    logits = F.linear(h, p["classifier.linear2.weight"], p["classifier.linear2.bias"])
    loss_sum = smoothed_ce_sum(logits, torch.as_tensor(sem_gt, dtype=torch.long))
    out["logits"] = logits.detach().clone()
    out["loss_raw"] = np.array([[float(loss_sum.detach()), float(len(groups))]], np.float32)
    if want_grads:
        (loss_sum / len(groups)).backward()
        out["grads"] = {k: (p[k].grad.clone() if p[k].grad is not None else None) for k in TRAINABLE}
    return out
