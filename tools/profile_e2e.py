"""Where does the end-to-end step (SegModel.forward from pinned host tensors, bench.py's `e2e` leg) spend its time beyond the
resident step?  Development aid: per-phase wall time with a device synchronize after every phase, then the device idle gaps
of one un-synchronised step (torch profiler).  usage: python tools/profile_e2e.py [scenes] [points]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from seggroup_b200 import _lib, engine, pipeline, synth
from seggroup_b200.model import SegModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 150000
dev = torch.device("cuda", 0)
_lib.load()
scenes_host = [synth.make_scene(i, N, name="scene%04d_%02d" % (0, i)) for i in range(B)]
tree = tempfile.mkdtemp(prefix="sgb_e2e_")
synth.write_scene_tree(tree, scenes_host)
os.chdir(tree)
torch.manual_seed(1)
model = SegModel(exp_name="bench").to(dev)
with torch.no_grad():
    model.mlp_1.bn1.weight.mul_(4.0)
model.epoch = "1"
model.scene_cache_dir = os.path.join(tree, "csr_cache")
opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
data_h = torch.stack([torch.from_numpy(s.data) for s in scenes_host]).pin_memory()
weak_h = torch.stack([torch.from_numpy(s.weak_label.astype(np.int64)) for s in scenes_host]).pin_memory()
info_h = torch.arange(B).view(B, 1)
engine.reserve_current_stream(6 << 30, device=dev)
sync = torch.cuda.synchronize


def step(phases=None):
    def mark(name, t0):
        if phases is not None:
            sync()
            phases[name] = phases.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    t = time.perf_counter()
    d, w = data_h.to(dev, non_blocking=True), weak_h.to(dev, non_blocking=True)
    t = mark("h2d", t)
    out = model(d, w, info_h)
    t = mark("model.forward (scene objects, concat, forward_scene, BN buffers, export enqueue)", t)
    loss = (out[0][:, 0] / out[0][:, 1]).mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    t = mark("backward", t)
    opt.step()
    t = mark("sgd", t)
    v = float(loss.item())
    mark("loss.item", t)
    return v


for _ in range(3):
    step()
model.flush_exports()
sync()
# the pieces of model.forward, synchronised
import seggroup_b200.model as M
orig = {k: getattr(SegModel, k) for k in ("_scene", "_update_bn", "_export")}
acc = {}


def wrap(name):
    f = orig[name]

    def g(self, *a, **kw):
        sync(); t0 = time.perf_counter()
        r = f(self, *a, **kw)
        sync(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    return g


for k in orig:
    setattr(SegModel, k, wrap(k))
fs = pipeline.forward_scene
cc = pipeline.SceneDevice.concat


def fs_t(*a, **kw):
    sync(); t0 = time.perf_counter(); r = fs(*a, **kw); sync(); acc["forward_scene"] = acc.get("forward_scene", 0.0) + (time.perf_counter() - t0) * 1e3
    return r


pipeline.forward_scene = fs_t
R = 3
ph = {}
for _ in range(R):
    step(ph)
model.flush_exports()
print("synchronised phases, ms per step (%d scenes x %d points):" % (B, N))
for k, v in ph.items():
    print("  %-90s %8.3f" % (k, v / R))
for k, v in acc.items():
    print("    inside forward: %-72s %8.3f" % (k, v / R))
for k in orig:
    setattr(SegModel, k, orig[k])
pipeline.forward_scene = fs
sync()
t0 = time.perf_counter()
for _ in range(5):
    step()
model.flush_exports()
sync()
print("un-synchronised: %.3f ms per step" % ((time.perf_counter() - t0) * 1e3 / 5))
from torch.profiler import profile as _tp, ProfilerActivity as _PA
with _tp(activities=[_PA.CUDA]) as prof:
    step(); sync()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs = sorted(ev, key=lambda e: e.time_range.start)
busy = sum(e.device_time for e in evs)
gaps = {}
t_end = evs[0].time_range.end
total_gap = 0.0
for e in evs[1:]:
    g = e.time_range.start - t_end
    if g > 2.0:
        a = gaps.setdefault(e.name[:60], [0, 0.0]); a[0] += 1; a[1] += g
        total_gap += g
    t_end = max(t_end, e.time_range.end)
print("one step: %d device activities, %.3f ms summed, idle gaps %.3f ms over a span of %.3f ms" % (len(evs), busy / 1e3, total_gap / 1e3, (t_end - evs[0].time_range.start) / 1e3))
for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:20]:
    print("  gap before %-60s n %4d  %8.3f ms" % (k, v[0], v[1] / 1e3))
agg = {}
for e in evs:
    if "Memcpy" in e.name or "Memset" in e.name or "at::native" in e.name:
        a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += e.device_time
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print("  %-70s n %4d  %9.3f ms" % (k, v[0], v[1] / 1e3))
