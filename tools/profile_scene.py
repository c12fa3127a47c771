"""Per-entry-point device time of one scene forward(+backward) — development aid, not the bench.
usage: python tools/profile_scene.py [N] [mode] [gscale] [scenes]     (scenes > 1: one block-diagonal batch)"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seggroup_b200 import synth, pipeline, _lib
from seggroup_b200.params import init_params, TRAINABLE

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
g = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
t = time.time(); scenes = [synth.make_scene(11 + i, N) for i in range(B)]; scene = scenes[0]; print("gen %.1fs" % (time.time() - t), scene.meta)
p = {k: v.cuda() for k, v in init_params(1, g).items()}
if mode == "train":
    for k in TRAINABLE: p[k].requires_grad_(True)
sc = pipeline.SceneDevice.concat([pipeline.SceneDevice.from_host(s) for s in scenes])
def step():
    with torch.set_grad_enabled(mode == "train"):
        r = pipeline.forward_scene(sc, p, mode=mode)
        if mode == "train":
            pipeline.batch_loss(r.loss_raw).backward()
    return r
if os.environ.get("SGB_PROFILE_FAST"):       # under `ncu --set full` (tools/gpu_final.sh): one warm-up step, one step to capture, nothing else
    r = step(); torch.cuda.synchronize()
    r = step(); torch.cuda.synchronize()
    print("levels", [L.S for L in r.levels])
    sys.exit(0)
for _ in range(2): r = step()
torch.cuda.synchronize()
t = time.time()
for _ in range(3): r = step()
torch.cuda.synchronize()
wall = (time.time() - t) / 3
print("levels", [L.S for L in r.levels], "wall ms/batch %.2f  -> %.0f points/s" % (wall * 1e3, B * N / wall))
l0 = _lib.launch_count(); step(); torch.cuda.synchronize(); print("library launches per step", _lib.launch_count() - l0)
from torch.profiler import profile as _tp, ProfilerActivity as _PA
with _tp(activities=[_PA.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot_k = sum(e.device_time for e in ev) if ev and hasattr(ev[0], "device_time") else sum(e.cuda_time for e in ev)
agg = {}
for e in ev:
    a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += (e.device_time if hasattr(e, "device_time") else e.cuda_time)
print("torch profiler: %d device activities, %.3f ms summed" % (len(ev), tot_k / 1e3))
# idle gaps of the device between consecutive activities, attributed to the activity that FOLLOWS the gap
evs = sorted(ev, key=lambda e: e.time_range.start)
if not evs:                                   # another profiler owns the device (ncu): no activity records
    sys.exit(0)
gaps = {}
t_end = evs[0].time_range.end
total_gap = 0.0
for e in evs[1:]:
    g = e.time_range.start - t_end
    if g > 2.0:
        a = gaps.setdefault(e.name[:60], [0, 0.0]); a[0] += 1; a[1] += g
        total_gap += g
    t_end = max(t_end, e.time_range.end)
print("device idle gaps > 2 us: %.3f ms in total over a span of %.3f ms" % (total_gap / 1e3, (t_end - evs[0].time_range.start) / 1e3))
for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print("  gap before %-60s n %4d  %8.3f ms" % (k, v[0], v[1] / 1e3))
if os.environ.get("SGB_TIMELINE"):          # ordered list of the idle gaps: what ran before, what ran after
    t_end, prev = evs[0].time_range.end, evs[0].name
    t0 = evs[0].time_range.start
    for e in evs[1:]:
        g = e.time_range.start - t_end
        if g > 4.0:
            print("  t=%8.3f ms gap %7.1f us  after %-42s before %s" % ((e.time_range.start - t0) / 1e3, g, prev[:42], e.name[:60]))
        if e.time_range.end >= t_end:
            t_end, prev = e.time_range.end, e.name
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("  %-70s n %4d  %9.3f ms" % (k, v[0], v[1] / 1e3))
_lib.enable_profile()
step(); torch.cuda.synchronize()
tot = sum(v[1] for v in _lib.profile.values())
for k, v in sorted(_lib.profile.items(), key=lambda kv: -kv[1][1]):
    print("%-32s calls %3d  %8.3f ms  %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("sum of kernel ms %.3f" % tot)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"N": N, "mode": mode, "scenes": B, "wall_ms": wall * 1e3, "kernels": _lib.profile}, open("gpurun_out/profile_scene_%dx%d_%s.json" % (B, N, mode), "w"), indent=1)
