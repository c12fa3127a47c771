"""Host-side (Python) cost of one training scene pass: cProfile over a few iterations."""
import sys, os, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seggroup_b200 import synth, pipeline, _lib
from seggroup_b200.params import init_params, TRAINABLE
N = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
scene = synth.make_scene(11, N)
p = {k: v.cuda() for k, v in init_params(1, 4.0).items()}
for k in TRAINABLE: p[k].requires_grad_(True)
sc = pipeline.SceneDevice.from_host(scene)
def step():
    r = pipeline.forward_scene(sc, p, mode="train")
    (r.loss_raw[:, 0].sum() / r.loss_raw[:, 1].sum()).backward()
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t = time.time()
for _ in range(5): step()
torch.cuda.synchronize()
wall = (time.time() - t) / 5
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
print("wall ms/scene", wall * 1e3)
