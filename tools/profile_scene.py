"""Per-entry-point device time of one scene forward(+backward) — development aid, not the bench.
usage: python tools/profile_scene.py [N] [mode] [gscale]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seggroup_b200 import synth, pipeline, _lib
from seggroup_b200.params import init_params, TRAINABLE

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
g = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
t = time.time(); scene = synth.make_scene(11, N); print("gen %.1fs" % (time.time() - t), scene.meta)
p = {k: v.cuda() for k, v in init_params(1, g).items()}
if mode == "train":
    for k in TRAINABLE: p[k].requires_grad_(True)
sc = pipeline.SceneDevice.from_host(scene)
def step():
    with torch.set_grad_enabled(mode == "train"):
        r = pipeline.forward_scene(sc, p, mode=mode)
        if mode == "train":
            (r.loss_raw[:, 0].sum() / r.loss_raw[:, 1].sum()).backward()
    return r
for _ in range(2): r = step()
torch.cuda.synchronize()
t = time.time()
for _ in range(3): r = step()
torch.cuda.synchronize()
wall = (time.time() - t) / 3
print("levels", [L.S for L in r.levels], "wall ms/scene %.2f  -> %.0f points/s" % (wall * 1e3, N / wall))
_lib.enable_profile()
step(); torch.cuda.synchronize()
tot = sum(v[1] for v in _lib.profile.values())
for k, v in sorted(_lib.profile.items(), key=lambda kv: -kv[1][1]):
    print("%-32s calls %3d  %8.3f ms  %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("sum of kernel ms %.3f" % tot)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"N": N, "mode": mode, "wall_ms": wall * 1e3, "kernels": _lib.profile}, open("gpurun_out/profile_scene_%d_%s.json" % (N, mode), "w"), indent=1)
