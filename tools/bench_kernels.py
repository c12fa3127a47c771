"""Per-kernel roofline micro-benchmark of the hot-path kernels (SURVEY.md 8d algorithmic bytes / flops).

    python tools/bench_kernels.py [--points 150000,500000] [--reps 20] [--only pool,kpconv] [--out gpurun_out/kernels.json]

Every timed launch is preceded by a 256 MiB write (L2 flush, outside the CUDA-event bracket), timed with CUDA events on
the launching stream; the median over `--reps` launches is reported next to the algorithmic bytes (HBM-bound kernels,
peak = MEASURED_PEAKS.json hbm_gbs) or flops (the tcgen05 contractions, peak = bf16_tflops / 2 for kind::tf32, and the
TF32 x 3 split spends 3 MMAs per useful product).  One JSON object per kernel, one per line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seggroup_b200 import _lib, ops, synth  # noqa: E402
from seggroup_b200 import kpconv_ops as KO  # noqa: E402


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


class Timer:
    def __init__(self, reps, warm=2):
        self.reps, self.warm = reps, warm
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def __call__(self, fn):
        for _ in range(self.warm):
            fn()
        ts = []
        for _ in range(self.reps):
            self.flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), float(np.min(ts))


def emit(rows, name, N, ms, ms_min, nbytes=None, flops=None, note="", **kw):
    P = peaks()
    r = {"kernel": name, "points": N, "ms_median": round(ms, 5), "ms_min": round(ms_min, 5)}
    if nbytes is not None:
        r.update(bound="hbm", algorithmic_bytes=int(nbytes), achieved_gbs=round(nbytes / ms / 1e6, 1), peak_gbs=P["hbm_gbs"],
                 frac=round(nbytes / ms / 1e6 / P["hbm_gbs"], 4))
    if flops is not None:
        r.update(algorithmic_flops=int(flops), achieved_tflops=round(flops / ms / 1e9, 2))
    if note:
        r["note"] = note
    r.update(kw)
    rows.append(r)
    print(json.dumps(r), flush=True)


def scene_arrays(N, seed=11):
    sc = synth.make_scene(seed, N)
    dev = "cuda"
    t = lambda a, dt=None: (torch.as_tensor(np.ascontiguousarray(a)) if dt is None else torch.as_tensor(np.ascontiguousarray(a)).to(dt)).to(dev)
    return sc, dict(data=t(sc.data), seg_off=t(sc.seg_offsets, torch.int32), seg_members=t(sc.seg_members, torch.int32),
                    adj=t(sc.adj, torch.int32), weak=t(sc.weak_label, torch.int32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="150000,500000")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    want = lambda k: not only or k in only
    _lib.load()
    T = Timer(args.reps, args.warm)
    rows = []
    g = torch.Generator().manual_seed(0)
    for N in [int(x) for x in args.points.split(",")]:
        sc, d = scene_arrays(N)
        S = d["seg_off"].numel() - 1
        C = 64
        if want("pool"):
            feat = torch.randn(N, C, generator=g).cuda()
            ms, mn = T(lambda: ops.segment_pool_max(feat, d["seg_off"], d["seg_members"]))
            emit(rows, "segment_pool_max_fwd [N,64] (a10)", N, ms, mn, nbytes=4 * N * C + 4 * N + 16 * S * C, segments=S)
            out, arg = ops.segment_pool_max(feat, d["seg_off"], d["seg_members"])
            go = torch.randn(S, C, generator=g).cuda()
            ms, mn = T(lambda: ops.segment_pool_max_bwd(go, arg, N))
            emit(rows, "segment_pool_max_bwd [N,64] (zero-fill + scatter)", N, ms, mn, nbytes=8 * S * C + 4 * N * C, segments=S)
        if want("pool") and N <= 200000:
            # the same kernel over a batch of 8 scenes' member lists concatenated (BASELINE configs[1] batch): 8 x more rows per launch
            B = 8
            featB = torch.randn(B * N, C, generator=g).cuda()
            offB = torch.cat([d["seg_off"][:-1] + i * N for i in range(B)] + [torch.tensor([B * N], dtype=torch.int32, device="cuda")]).to(torch.int32)
            memB = torch.cat([d["seg_members"] + i * N for i in range(B)]).to(torch.int32)
            ms, mn = T(lambda: ops.segment_pool_max(featB, offB, memB))
            emit(rows, "segment_pool_max_fwd [8N,64] (a10, 8 scenes in one launch)", B * N, ms, mn, nbytes=B * (4 * N * C + 4 * N + 16 * S * C), segments=B * S)
            outB, argB = ops.segment_pool_max(featB, offB, memB)
            goB = torch.randn(B * S, C, generator=g).cuda()
            ms, mn = T(lambda: ops.segment_pool_max_bwd(goB, argB, B * N))
            emit(rows, "segment_pool_max_bwd [8N,64] (8 scenes in one launch)", B * N, ms, mn, nbytes=B * (8 * S * C + 4 * N * C), segments=B * S)
            del featB, outB, argB, goB
        if want("ceiling"):
            # what the memory system gives the same access pattern without any arithmetic: a library row gather of the
            # SAME 256-byte rows in the SAME (random) member order, and a plain streaming copy (read + write bytes counted)
            featC = torch.randn(N, C, generator=g).cuda()
            idx = d["seg_members"].long()
            dst = torch.empty_like(featC)
            ms, mn = T(lambda: torch.index_select(featC, 0, idx, out=dst))
            emit(rows, "reference: torch.index_select of the [N,64] rows in member order (random 256-B gather + streaming write)", N, ms, mn,
                 nbytes=2 * 4 * N * C + 8 * N)
            ms, mn = T(lambda: dst.copy_(featC))
            emit(rows, "reference: streaming copy of [N,64] (read + write)", N, ms, mn, nbytes=2 * 4 * N * C)
            if N <= 200000:
                featB = torch.randn(8 * N, C, generator=g).cuda()
                idxB = torch.cat([d["seg_members"] + i * N for i in range(8)]).long()
                dstB = torch.empty_like(featB)
                ms, mn = T(lambda: torch.index_select(featB, 0, idxB, out=dstB))
                emit(rows, "reference: torch.index_select of [8N,64] rows in member order", 8 * N, ms, mn, nbytes=8 * (2 * 4 * N * C + 8 * N))
                ms, mn = T(lambda: dstB.copy_(featB))
                emit(rows, "reference: streaming copy of [8N,64] (read + write)", 8 * N, ms, mn, nbytes=8 * 2 * 4 * N * C)
                del featB, idxB, dstB
            del featC, dst
        if want("centralize") or want("knn") or want("edgeconv"):
            order = d["seg_members"]
            x9 = ops.centralize(d["data"], order, d["seg_off"])
        if want("centralize"):
            ms, mn = T(lambda: ops.centralize(d["data"], order, d["seg_off"]))
            emit(rows, "centralize (a8)", N, ms, mn, nbytes=24 * N + 4 * N + 36 * N)
        if want("knn") or want("edgeconv"):
            knn = ops.cluster_knn(x9, order, d["seg_off"], 20)
        if want("knn"):
            ms, mn = T(lambda: ops.cluster_knn(x9, order, d["seg_off"], 20))
            emit(rows, "cluster_knn k=20 (a7)", N, ms, mn, nbytes=16 * N + 80 * N, note="compute bound: exact per-cluster top-k")
        if want("edgeconv"):
            from seggroup_b200.params import init_params
            p = {k: v.cuda() for k, v in init_params(1, 4.0).items()}
            a2 = (x9, knn, p["mlp_2.conv1.0.weight"], p["mlp_2.bn1.weight"], p["mlp_2.bn1.bias"])
            ms, mn = T(lambda: ops.edgeconv_fwd(*a2, want_backward=False))
            emit(rows, "edgeconv_fwd MLP2 infer (a9)", N, ms, mn, nbytes=36 * N + 80 * N + 256 * N + 64 * N, flops=2.0 * N * 20 * 18 * 64)
            a3 = a2[:2] + (p["mlp_3.conv1.0.weight"], p["mlp_3.bn1.weight"], p["mlp_3.bn1.bias"], p["mlp_3.conv2.0.weight"], p["mlp_3.bn2.weight"],
                           p["mlp_3.bn2.bias"])
            for wb in (False, True):
                ms, mn = T(lambda: ops.edgeconv_fwd(*a3, want_backward=wb))
                emit(rows, "edgeconv_fwd MLP3 %s (a9, tcgen05 64x64 layer)" % ("train" if wb else "infer"), N, ms, mn,
                     nbytes=36 * N + 80 * N + 256 * N + 64 * N, flops=2.0 * N * 20 * (18 * 64 + 64 * 64))
            # backward of MLP3 fused with the point -> segment pooling (sparse arg-max edges + dense tcgen05 pass)
            o3 = ops.edgeconv_fwd(*a3, want_backward=True)
            pooled, arg = ops.segment_pool_max(o3["out"], d["seg_off"], d["seg_members"])
            gp = torch.randn(S, 64, generator=g).cuda()
            ms, mn = T(lambda: ops.edgeconv_bwd(gp, arg, o3["argk"], x9, knn, a3[2], o3["stats1"], o3["mom1"], o3["ctr"], a3[5], o3["stats2"], o3["mom2"]))
            emit(rows, "edgeconv_bwd MLP3 (a9, dense pass on tcgen05)", N, ms, mn, nbytes=36 * N + 80 * N, flops=2.0 * N * 20 * (18 * 64 + 64 * 64 + 64 * 18))
        if want("export"):
            from seggroup_b200 import pipeline
            # label export needs a level; run the model-free part of the pipeline: scene_init + level_build
            seg_of_point, sos, uf = ops.scene_init(d["seg_off"], d["seg_members"], d["weak"])
            L = ops.level_build(uf, d["seg_off"], d["seg_members"], sos)
            unmap = torch.arange(N, dtype=torch.int64, device="cuda")
            ms, mn = T(lambda: ops.export_labels(unmap, seg_of_point, L))
            emit(rows, "export_labels (a16)", N, ms, mn, nbytes=8 * N + 4 * N + 12 * N)
        # ---- KPConv operator set on a 4 cm subsample of a noisy-sheet cloud (SURVEY.md 8d config 5 geometry)
        if want("grid") or want("neighbors") or want("kpconv") or want("kppool"):
            pts, lens = synth.make_cloud(5, N)
            P = torch.as_tensor(pts).cuda()
            Lb = torch.as_tensor(lens).to(torch.int32).cuda()
            sub, sb = KO.batch_grid_subsampling(P, Lb, 0.04)
            M = sub.shape[0]
        if want("grid"):
            ms, mn = T(lambda: KO.batch_grid_subsampling(P, Lb, 0.04))
            emit(rows, "grid_subsample dl=0.04 (a18, incl. the size read-back)", N, ms, mn, nbytes=12 * N + 12 * M, voxels=M)
        if want("neighbors") or want("kpconv") or want("kppool"):
            nb = KO.batch_ordered_neighbors(sub, sub, sb, sb, 0.10)
            W = nb.shape[1]
        if want("neighbors"):
            ms, mn = T(lambda: KO.batch_ordered_neighbors(sub, sub, sb, sb, 0.10))
            emit(rows, "radius_neighbors r=0.10 (a19, count + read-back + fill)", M, ms, mn, nbytes=24 * M + 4 * M * W, width=W)
            for rr in (0.05, 0.20):                      # radius sweep on the dl = 0.04 subsample
                Wr = KO.batch_ordered_neighbors(sub, sub, sb, sb, rr).shape[1]
                ms, mn = T(lambda: KO.batch_ordered_neighbors(sub, sub, sb, sb, rr))
                emit(rows, "radius_neighbors r=%.2f (a19, count + read-back + fill)" % rr, M, ms, mn, nbytes=24 * M + 4 * M * Wr, width=Wr)
        if want("kpconv"):
            K = 15
            extent = 0.04
            kp = torch.randn(K, 3, generator=g)
            kp = kp / kp.norm(dim=1, keepdim=True) * 0.06 * torch.rand(K, 1, generator=g) ** (1 / 3)
            kp[0] = 0
            kp = kp.cuda()
            nbc = nb[:, :min(W, 64)].contiguous()
            Wc = nbc.shape[1]
            if N >= 400000 and want("kpconv"):           # config 5: neighbour-count cap sweep at 64x64 (W cap 16 / 32 / 64 / 128)
                nb20 = KO.batch_ordered_neighbors(sub, sub, sb, sb, 0.20)
                feats = torch.randn(M, 64, generator=g).cuda()
                kv = (torch.randn(K, 64, 64, generator=g) / np.sqrt(K * 64)).cuda()
                for cap in (16, 32, 64, 128):
                    nbw = nb20[:, :min(nb20.shape[1], cap)].contiguous()
                    ms, mn = T(lambda: KO.KPConv_ops(sub, sub, nbw, feats, kp, kv, 0.08, "linear", "sum", tensor_cores=True))
                    emit(rows, "kpconv_fwd 64x64 %s, r=0.20 neighbours capped at W=%d (a20)" % ("tcgen05" if nbw.shape[1] <= 64 else "fp32 SIMT (W > 64)", nbw.shape[1]), M, ms, mn,
                         nbytes=4 * M * nbw.shape[1] + 24 * M + 4 * M * 64 * 2 + 4 * K * 64 * 64, flops=2.0 * M * K * 64 * 64, width=nbw.shape[1])
                del nb20
            for cin, cout in [(64, 64), (32, 32), (128, 128)]:
                feats = torch.randn(M, cin, generator=g).cuda()
                kv = (torch.randn(K, cin, cout, generator=g) / np.sqrt(K * cin)).cuda()
                by = 4 * M * Wc + 24 * M + 4 * M * cin + 4 * M * cout + 4 * K * cin * cout
                fl = 2.0 * M * K * cin * cout
                for tc in (False, True):
                    ms, mn = T(lambda: KO.KPConv_ops(sub, sub, nbc, feats, kp, kv, extent, "linear", "sum", tensor_cores=tc))
                    emit(rows, "kpconv_fwd %dx%d %s (a20)" % (cin, cout, "tcgen05 contraction" if tc else "fp32 SIMT"), M, ms, mn, nbytes=by, flops=fl,
                         width=Wc, note="flops = the K*Cin*Cout contraction only")
                fd = feats.clone().requires_grad_(True)
                kd = kv.clone().requires_grad_(True)
                out = KO.KPConv_ops(sub, sub, nbc, fd, kp, kd, extent, "linear", "sum")
                go = torch.randn_like(out)

                def bwd():
                    fd.grad = None; kd.grad = None
                    out.backward(go, retain_graph=True)
                ms, mn = T(bwd)
                emit(rows, "kpconv_bwd %dx%d tcgen05 (a20: dK = WF^T g, GW = g K^T, scatter)" % (cin, cout), M, ms, mn, flops=2 * fl, width=Wc)
                fs = feats.clone().requires_grad_(True)
                ks = kv.clone().requires_grad_(True)
                outs = KO.KPConv_ops(sub, sub, nbc, fs, kp, ks, extent, "linear", "sum", tensor_cores=False)

                def bwd_simt():
                    fs.grad = None; ks.grad = None
                    outs.backward(go, retain_graph=True)
                ms, mn = T(bwd_simt)
                emit(rows, "kpconv_bwd %dx%d fp32 SIMT (a20)" % (cin, cout), M, ms, mn, flops=2 * fl, width=Wc)
        if want("kppool"):
            # a21: strided-block shortcut pooling on the same geometry: pool the dl=0.04 features onto a dl=0.08 subsample
            sub2, sb2 = KO.batch_grid_subsampling(sub, sb, 0.08)
            pool_inds = KO.batch_ordered_neighbors(sub2, sub, sb2, sb, 0.08)
            M2, Wp = pool_inds.shape
            for dch in (64, 128):
                xf = torch.randn(M, dch, generator=g).cuda()
                ms, mn = T(lambda: KO.ind_max_pool(xf, pool_inds))
                emit(rows, "ind_max_pool d=%d (a21)" % dch, M, ms, mn, nbytes=4 * M * dch + 4 * M2 * Wp + 4 * M2 * dch, width=Wp, pooled=M2,
                     note="gathered rows 4*M2*W*d = %d B are L2 traffic" % (4 * M2 * Wp * dch))
                up_inds = KO.batch_ordered_neighbors(sub, sub2, sb, sb2, 0.08)[:, :1].contiguous()
                xc = torch.randn(M2, dch, generator=g).cuda()
                ms, mn = T(lambda: KO.closest_pool(xc, up_inds))
                emit(rows, "closest_pool d=%d (a21, upsampling)" % dch, M, ms, mn, nbytes=4 * M2 * dch + 4 * M + 4 * M * dch, pooled=M2)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
