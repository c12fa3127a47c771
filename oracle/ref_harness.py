"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference `seggroup/model.py` on CPU.

Only usable where `/root/reference` exists (the build container, not the GPU box).  It is used to
(1) pin `oracle/seggroup_oracle.py` against the real reference and (2) mint the golden fixtures
under `tests/golden/` (see `oracle/make_golden.py`).  Nothing in the product imports this.

How the reference is made runnable without touching it:
  * `chainer.cuda.get_array_module` (model.py:19,362) and `plyfile.PlyData` (model.py:20) are absent
    from this image -> tiny `sys.modules` stubs;
  * `dataset.scannet.util.visualize_labels` (model.py:24) is only called when `visualize=True` -> stub;
  * model.py:50,90 hard-code `torch.device('cuda')` for an index offset -> the module global `torch`
    is replaced by a proxy whose `device('cuda')` answers `cpu`;
  * all paths in model.py are relative -> cwd must be a scratch directory holding the synthetic
    `dataset/scannet/**` tree (`seggroup_b200.synth.write_scene_tree`).
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_SEGGROUP = os.environ.get("SEGGROUP_REFERENCE", "/root/reference/seggroup")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_SEGGROUP, "model.py"))


class _TorchCpuProxy:
    """Stands in for the `torch` module inside the reference: device('cuda') -> cpu."""

    def __init__(self):
        self._t = torch

    def device(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            return torch.device("cpu")
        return torch.device(*a, **k)

    def __getattr__(self, name):
        return getattr(self._t, name)


_ref_mod = None


def load_reference():
    """Import the reference model module (once) and return it."""
    global _ref_mod
    if _ref_mod is not None:
        return _ref_mod
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REF_SEGGROUP)
    chainer = types.ModuleType("chainer")
    chainer_cuda = types.ModuleType("chainer.cuda")
    chainer_cuda.get_array_module = lambda *a, **k: np
    chainer.cuda = chainer_cuda
    plyfile = types.ModuleType("plyfile")
    plyfile.PlyData = object
    ds = types.ModuleType("dataset")
    ds_sc = types.ModuleType("dataset.scannet")
    ds_util = types.ModuleType("dataset.scannet.util")
    ds_util.visualize_labels = lambda *a, **k: None
    for name, mod in [("chainer", chainer), ("chainer.cuda", chainer_cuda), ("plyfile", plyfile),
                      ("dataset", ds), ("dataset.scannet", ds_sc), ("dataset.scannet.util", ds_util)]:
        sys.modules.setdefault(name, mod)
    # the reference does `from data import ScanNet`, `from util import ...`: load them by path under
    # those exact names so nothing else on sys.path can shadow them
    for name in ("data", "util"):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_SEGGROUP, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
    spec = importlib.util.spec_from_file_location("seggroup_reference_model", os.path.join(REF_SEGGROUP, "model.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["seggroup_reference_model"] = mod
    spec.loader.exec_module(mod)
    if not torch.cuda.is_available():
        mod.torch = _TorchCpuProxy()
    _ref_mod = mod
    return mod


@contextlib.contextmanager
def _chdir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


class Capture:
    """Records intermediates of one reference forward by wrapping its module-level functions."""

    NAMES = ["get_knn", "get_cluster_pointcloud", "calculate_distance", "aggregate_cluster_feature",
             "update_adj", "group_nearby_clusters", "group_unlabeled_clusters"]

    def __init__(self, mod):
        self.mod = mod
        self.log = {n: [] for n in self.NAMES}
        self._orig = {}

    def __enter__(self):
        for n in self.NAMES:
            self._orig[n] = getattr(self.mod, n)
            setattr(self.mod, n, self._wrap(n, self._orig[n]))
        return self

    def __exit__(self, *exc):
        for n, f in self._orig.items():
            setattr(self.mod, n, f)

    def _wrap(self, name, fn):
        def inner(*a, **k):
            out = fn(*a, **k)
            if name == "group_nearby_clusters":
                self.log[name].append(np.array(out[0].cluster_id).copy())
            elif name == "group_unlabeled_clusters":
                self.log[name].append((np.array(out[0].cluster_id).copy(), out[1].detach().clone(), out[2].clone()))
            elif isinstance(out, torch.Tensor):
                self.log[name].append(out.detach().clone())
            else:
                self.log[name].append(out)
            return out
        return inner


def run_reference(tree_root: str, scene_index: int, *, mode: str = "train", seed: int = 1, state_dict=None,
                  bn_gamma_scale: float | None = None, exp_name: str = "ref", backward: bool = True,
                  capture: bool = True, n_threads: int | None = None):
    """Run the reference SegModel on scene `scene_index` of the tree at `tree_root`.

    mode: 'train' | 'ins_infer' | 'sem_infer'.  Returns a dict with outputs, label files (as int arrays),
    captured intermediates and (train mode) parameter gradients.
    """
    mod = load_reference()
    if n_threads:
        torch.set_num_threads(n_threads)
    with _chdir(tree_root):
        torch.manual_seed(seed)
        model = mod.SegModel(exp_name=exp_name, cuda=False, sem_infer=(mode == "sem_infer"), ins_infer=(mode == "ins_infer"))
        if state_dict is not None:
            model.load_state_dict(state_dict)
        if bn_gamma_scale is not None:
            with torch.no_grad():
                model.mlp_1.bn1.weight.mul_(bn_gamma_scale)
        model.train()
        model.epoch = mode if mode != "train" else "1"
        init_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
        dataset = mod.ScanNet(label_style="manual")
        data, weak_label, info = dataset[scene_index]
        # the reference's export buffers default to 150000 entries (model.py:525,552,580)
        if data.shape[0] > 150000:
            for fn in ("export_segment_label", "export_instance_label", "export_semantic_label"):
                f = getattr(mod, fn)
                f.__defaults__ = (data.shape[0],)
        cap = Capture(mod) if capture else contextlib.nullcontext()
        # ReLU margin of the two GCN layers (model.py:151): a pre-activation within rounding of zero makes the gradient
        # implementation-defined (the derivative jumps there), so fixtures record how close the reference got to the kink.
        # Forward hooks observe; they do not change what the reference computes.
        margins = {}
        hooks = [getattr(model, g).fc.register_forward_hook(
            lambda m, i, o, g=g: margins.__setitem__(g, (float(o.detach().abs().min()), float(o.detach().abs().max()))))
            for g in ("gcn_2", "gcn_3")]
        torch.manual_seed(seed + 1000)   # dropout mask seed (classifier, model.py:159)
        with cap:
            if mode == "train":
                out = model(data.unsqueeze(0), weak_label.unsqueeze(0), info.unsqueeze(0))
            else:
                with torch.no_grad():
                    out = model(data.unsqueeze(0), weak_label.unsqueeze(0), info.unsqueeze(0))
        for h in hooks:
            h.remove()
        res = {"init_state": init_state, "out": [o.detach().clone() for o in out], "relu_margin": margins}
        if mode == "train" and backward:
            loss_raw = out[0]
            loss = torch.sum(loss_raw[:, 0]) / torch.sum(loss_raw[:, 1])
            loss.backward()
            res["grads"] = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
            res["loss"] = float(loss)
        scene_name = model.scene_list[int(info)][:-1]
        stage = mode if mode != "train" else "epoch_1"
        out_root = os.path.join("results", exp_name, scene_name, stage)
        labels = {}
        for fn in sorted(os.listdir(out_root)):
            if fn.endswith(".txt"):
                labels[fn[:-4]] = np.loadtxt(os.path.join(out_root, fn), dtype=np.int64)
        res["labels"] = labels
        res["final_state"] = {k: v.detach().clone() for k, v in model.state_dict().items()}
        if capture:
            res["capture"] = cap.log
    return res
