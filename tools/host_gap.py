"""Where does the host spend its time in one forward of a scene batch? (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seggroup_b200 import synth, pipeline, ops, _lib
from seggroup_b200.params import init_params
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scenes = [synth.make_scene(11 + i, 150000) for i in range(B)]
p = {k: v.cuda() for k, v in init_params(1, 4.0).items()}
sc = pipeline.SceneDevice.concat([pipeline.SceneDevice.from_host(s) for s in scenes])
marks = []
orig = _lib._call
def traced(name, *a):
    t0 = time.perf_counter(); r = orig(name, *a); t1 = time.perf_counter()
    marks.append((name, t0, t1)); return r
with torch.no_grad():
    for _ in range(3): pipeline.forward_scene(sc, p, mode="ins_infer")
    torch.cuda.synchronize()
    _lib._call = traced
    for it in range(2):
        marks.clear()
        torch.cuda.synchronize(); T0 = time.perf_counter()
        r = pipeline.forward_scene(sc, p, mode="ins_infer")
        T1 = time.perf_counter(); torch.cuda.synchronize(); T2 = time.perf_counter()
        in_lib = sum(b - a for _, a, b in marks)
        print("iter %d: host %.2f ms (of which inside library calls %.2f ms over %d calls), device done at %.2f ms" % (it, (T1 - T0) * 1e3, in_lib * 1e3, len(marks), (T2 - T0) * 1e3))
        prev = T0
        rows = []
        for name, a, b in marks:
            rows.append((name, (a - prev) * 1e3, (b - a) * 1e3)); prev = b
        rows.append(("<return>", (T1 - prev) * 1e3, 0.0))
        print("  largest python gaps BEFORE a library call (ms) and the call's own duration:")
        for name, gap, dur in sorted(rows, key=lambda r: -r[1])[:12]:
            print("    %-34s gap %.3f  call %.3f" % (name, gap, dur))
        print("  longest library calls:")
        for name, gap, dur in sorted(rows, key=lambda r: -r[2])[:8]:
            print("    %-34s call %.3f" % (name, dur))
