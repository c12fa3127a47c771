#!/usr/bin/env python
"""Stage-by-stage parity diagnostic: CUDA path vs the oracle with the kernels' own (canonical) kNN tie rule.

    python tools/diag_parity.py [--points 50000] [--seed 7] [--out gpurun_out/diag_parity.txt]

Prints (i) bit-identity of every index list the forward produces (FPS picks, kNN lists of MLP1/2/3, cluster maps, labels),
(ii) relative error of the stage features, (iii) relative L2 error of dLoss/d(stage tensor) walking back from the loss, and of
the parameter gradients — i.e. WHERE along the backward pass the two implementations separate.  Test infrastructure: it
imports oracle/ and is not part of the product path.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=50000)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--g", type=float, default=4.0)
    ap.add_argument("--fixture", default="", help="tests/golden stem with %s for the mode, e.g. seggroup50k_s9_%s_g4")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "diag_parity.txt"))
    a = ap.parse_args()
    from oracle import seggroup_oracle as O
    from seggroup_b200 import pipeline, synth
    from seggroup_b200.params import TRAINABLE, init_params

    lines = []

    def say(*x):
        s = " ".join(str(v) for v in x)
        print(s, flush=True)
        lines.append(s)

    scene = synth.make_scene(a.seed, a.points)
    say("scene seed %d: N=%d S1=%d" % (a.seed, a.points, len(scene.seg_offsets) - 1))
    # ---- oracle, canonical tie rule
    params_cpu = O.init_params(1, a.g)
    t0 = time.time()
    ref = O.forward(scene, params_cpu, mode="train", tie="canonical", want_grads=False)   # first pass: number of instances
    n_inst = int(ref["loss_raw"][0, 1])
    torch.manual_seed(1001)
    mask = F.dropout(torch.ones(n_inst, 128), 0.5, True) != 0
    ref = O.forward(scene, params_cpu, mode="train", tie="canonical", dropout_mask=mask, want_grads=True)
    say("oracle (canonical ties) 2 passes: %.1f s; loss_raw %s; levels %s" % (time.time() - t0, ref["loss_raw"].tolist(), [L.S for L in ref["levels"]]))
    # ---- CUDA
    p = {k: v.cuda() for k, v in init_params(1, a.g).items()}
    for k in TRAINABLE:
        p[k].requires_grad_(True)
    sc = pipeline.SceneDevice.from_host(scene)
    res = pipeline.forward_scene(sc, p, mode="train", dropout_mask=mask.cuda(), keep_aux=True)
    loss = res.loss_raw[0, 0] / res.loss_raw[0, 1]
    loss.backward()
    say("cuda loss_raw %s; levels %s" % (res.loss_raw.detach().cpu().numpy().tolist(), [L.S for L in res.levels]))

    # ---- (i) index lists
    say("\n== index lists (entries that differ / total)")
    ci = res.aux["cloud_idx_1"].cpu().numpy(); ro = ref["cloud_idx_1"]
    say("cloud_idx_1 (FPS picks)      %d / %d" % ((ci != ro).sum(), ci.size))
    k1 = res.aux["knn_1"].cpu().numpy(); r1 = ref["knn_1"].numpy()
    say("knn_1 [S,64,10] ordered      %d / %d ; as sets per row: %d rows differ" % (
        (k1 != r1).sum(), k1.size, (np.sort(k1, -1) != np.sort(r1, -1)).any(-1).sum()))
    for tag in ("2", "3"):
        kc = res.aux["knn_" + tag].cpu().numpy(); kr = ref["knn_" + tag].numpy()
        say("knn_%s [N,20] ordered         %d / %d ; as sets per row: %d rows differ" % (
            tag, (kc != kr).sum(), kc.size, (np.sort(kc, -1) != np.sort(kr, -1)).any(-1).sum()))
    for i, (Lc, Lr) in enumerate(zip(res.levels, ref["levels"])):
        same = Lc.S == Lr.S and np.array_equal(Lc.seg2cl.cpu().numpy()[:len(Lr.seg2cluster)], Lr.seg2cluster)
        say("level %d: S %d vs %d, seg->cluster map identical: %s" % (i + 1, Lc.S, Lr.S, same))
    for k in ("adj_1", "adj_2", "adj_3", "adj_4"):
        say("%s identical: %s" % (k, np.array_equal(res.aux[k].cpu().numpy(), np.asarray(ref[k]).reshape(-1, 2))))
    nl = sum(int((res.labels[k].cpu().numpy() != v).sum()) for k, v in ref["labels"].items())
    say("label vectors (14): %d vertices differ" % nl)

    # ---- (ii) features
    say("\n== stage features, relative L2 error (max abs error)")
    for k in ("data_1", "Feat_1", "Feat_mlp_2", "Feat_gcn_2", "Feat_mlp_3", "Feat_gcn_3", "Feat_5", "logits", "dists_1", "dists_2", "dists_3"):
        x = res.aux[k].detach().cpu().numpy(); y = ref[k].numpy()
        say("%-12s %.3e  (%.3e)" % (k, rel(x, y), np.abs(x - y).max()))
    # arg-max rows of the point -> segment pooling, each side from its own features (first maximal row in member order)
    for tag, Lr in (("2", ref["levels"][1]), ("3", ref["levels"][2])):
        fc = res.aux["Feat_mlp_" + tag].cpu().numpy(); fr = ref["Feat_mlp_" + tag].numpy()
        diff = tot = 0
        for m in Lr.points:
            diff += int((np.argmax(fc[m], 0) != np.argmax(fr[m], 0)).sum()); tot += 64
        say("pool arg-max rows of layer %s: %d / %d differ" % (tag, diff, tot))

    # ---- (iii) gradients
    say("\n== dLoss/d(stage), relative L2 error, walking back from the loss")
    lc, lr = res.aux["_live"], ref["_live"]
    for k in ("Feat_5", "Feat_4", "gcn_3", "Z_3", "AX_3", "sims_3", "d_3", "cat_3", "pool_3", "Feat_3", "gcn_2", "Z_2", "AX_2", "sims_2", "d_2",
              "cat_2", "pool_2", "Feat_2", "Feat_1"):
        if k in lc and k in lr and lc[k].grad is not None and lr[k].grad is not None:
            gc, gr = lc[k].grad.cpu().numpy(), lr[k].grad.numpy()
            say("%-8s %.3e   |g| %.3e" % (k, rel(gc, gr), np.linalg.norm(gr)))
        else:
            say("%-8s (no grad kept: %s %s)" % (k, k in lc, k in lr))
    # forward values of the GCN internals, and the channel-by-channel picture of the first stage that separates
    say("\n== GCN internals, forward relative L2 error")
    for k in ("d_2", "sims_2", "AX_2", "Z_2", "d_3", "sims_3", "AX_3", "Z_3"):
        if k in lc and k in lr:
            say("%-8s %.3e   min|Z| %.3e" % (k, rel(lc[k].detach().cpu().numpy(), lr[k].detach().numpy()), float(lr[k].detach().abs().min())))
    for k in ("Z_2", "AX_2", "cat_2"):
        if k in lc and lc[k].grad is not None:
            gc, gr = lc[k].grad.cpu().numpy(), lr[k].grad.numpy()
            err = np.abs(gc - gr)
            i, j = np.unravel_index(np.argmax(err), err.shape)
            say("%s grad: worst entry (%d,%d) cuda %.6e oracle %.6e; rows with error > 1e-5*max: %d of %d; per-column-block error %s" % (
                k, i, j, gc[i, j], gr[i, j], int((err.max(1) > 1e-5 * np.abs(gr).max()).sum()), err.shape[0],
                [float("%.2e" % rel(gc[:, a:a + 64], gr[:, a:a + 64])) for a in range(0, err.shape[1], 64)]))
    say("\n== parameter gradients, relative L2 error (and max-abs error / max-abs)")
    for k in TRAINABLE:
        gr = ref["grads"][k]
        if gr is None or p[k].grad is None:
            continue
        gr = gr.numpy(); gc = p[k].grad.cpu().numpy()
        say("%-28s %.3e  (%.3e)" % (k, rel(gc, gr), np.abs(gc - gr).max() / (np.abs(gr).max() + 1e-30)))
    # the same comparison with the CUDA path's ReLU active set imposed on the oracle's backward (one entry of Z within rounding
    # of zero flips the derivative 0 <-> 1 and that alone moves every gradient upstream of it)
    masks = {t: (lc["gcn_" + t].detach() > 0).cpu() for t in ("2", "3")}
    for t in ("2", "3"):
        zr = lr["Z_" + t].detach()
        flipped = (zr > 0) != masks[t]
        say("ReLU of gcn_%s: %d of %d activations differ; largest |Z| among them %.3e (max |Z| %.3e, min |Z| %.3e)" % (
            t, int(flipped.sum()), flipped.numel(), float(zr[flipped].abs().max()) if flipped.any() else 0.0, float(zr.abs().max()), float(zr.abs().min())))
    ref2 = O.forward(scene, params_cpu, mode="train", tie="canonical", dropout_mask=mask, want_grads=True, relu_masks=masks)
    gold = None
    if a.fixture:
        gold = np.load(os.path.join(ROOT, "tests", "golden", (a.fixture % "train") + ".npz"))
        say("fixture %s: relu margin %s, reference CPU seconds %.1f" % (a.fixture, gold["relu_margin"].tolist() if "relu_margin" in gold.files else None,
                                                                         float(gold["reference_cpu_seconds"]) if "reference_cpu_seconds" in gold.files else -1))
    say("\n== parameter gradients, relative L2 error: vs oracle(canonical) | vs oracle(canonical, CUDA's ReLU active set) | vs the reference fixture")
    for k in TRAINABLE:
        if ref["grads"][k] is None or p[k].grad is None:
            continue
        gc = p[k].grad.cpu().numpy()
        say("%-28s %.3e | %.3e | %s" % (k, rel(gc, ref["grads"][k].numpy()), rel(gc, ref2["grads"][k].numpy()),
                                       ("%.3e" % rel(gc, gold["grad/" + k])) if gold is not None and ("grad/" + k) in gold.files else "-"))
    # smallest edge distances feeding d(dist)/d(feat) = diff / d  (ill-conditioned when d ~ eps * sqrt(C))
    say("\n== smallest edge distances per level (pairwise_distance eps = 1e-6)")
    for k in ("dists_1", "dists_2", "dists_3"):
        d = np.sort(ref[k].numpy())
        say("%s: n=%d min %.3e, 5 smallest %s" % (k, d.size, d[0] if d.size else float("nan"), d[:5].tolist()))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
