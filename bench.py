#!/usr/bin/env python
"""bench.py — points/sec of the SegGroup hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scenes 8] [--points 150000]

A "step" is one training step of BASELINE.json configs[1]: forward + backward + SGD update over a batch of
`--scenes` synthetic ScanNet-v2-shaped scenes of `--points` points each (per-scene BatchNorm statistics,
gradient = mean over scenes of loss_sum/loss_num, reference semantics of train.py:165-170 at 8 ranks).
With N > 1 (torchrun, one rank per GPU) every rank runs its own batch (weak scaling) and the flat fp32
gradient buffer is all-reduced once per step over NCCL.

One JSON line is printed by rank 0:
  value        whole-job points/sec of the step, the scene batch already resident in HBM, CUDA-event timed, max over ranks, nothing
               instrumented inside the timed region
  e2e          the same step through the reference-facing plugin call `SegModel.forward(data, weak_label, info)` on a scene tree on
               disk: pinned host tensors copied H2D, side files parsed by the model (kept in HBM after their first use), 14 label
               files per scene copied D2H and written, loss read back; the region ends when the files are on disk
  inference    pseudo-label generation (ins_infer) over the same batch: resident value and the end-to-end plugin value
  roofline     dominant entry point (EdgeConv forward of MLP3, tcgen05): useful flops / CUDA-event time (measured in a separate
               pass) vs the measured tensor peak; `traffic` / `ncu` come from the committed ncu summary of this round (profiles/)
  roofline_more the HBM-bound gather / scatter entries vs the measured copy bandwidth, the tcgen05 backward
  cpu_baseline the oracle port of the reference CPU path on ONE scene of the batch's size (150,000 points) on this box's host
               cores, with the unmodified reference's recorded time beside it
  e2e_first_epoch / config4_one_scene_per_gpu / config3_sharded_inference: BASELINE.json configs 3-4 and the cold-cache step
`--impl reference` times the CPU port alone, as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "points/sec SegGroup fwd+bwd and inference at 1/2/4/8 B200; % HBM roofline"
UNIT = "points/s"
GSCALE = 4.0          # calibrated weight set of SURVEY.md 8d: every grouping stage does non-trivial work


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML in-process (a thread that wakes
    every 50 ms; an external `nvidia-smi -lms` poller takes driver locks that the four launching threads also need and
    was measured to slow the step it is supposed to observe), nvidia-smi only if NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, interval=0.01):
        self.index, self.interval, self.rows, self.proc, self.nv, self.stop_flag = index, interval, [], None, None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((sm, mx, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            self.stop_flag.wait(self.interval)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)
            sm = [r[0] for r in self.rows]
            mx = [r[1] for r in self.rows if r[1] is not None]
            reasons = sorted({k for r in self.rows for k in r[2]})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


def cpu_port_points_per_sec(n_points, steps=1, warmup=0):
    """The oracle (CPU restatement of the reference's torch/numpy path) fwd+bwd on one scene: points/sec."""
    from oracle import seggroup_oracle as O
    from seggroup_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    scene = synth.make_scene(101, n_points)
    params = O.init_params(1, GSCALE)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.forward(scene, params, mode="train", tie="torch", want_grads=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return n_points / float(np.mean(times)), float(np.mean(times)), torch.get_num_threads()


def unmodified_reference_record():
    """Points/s of the UNMODIFIED reference (seggroup/model.py on torch-CPU) as measured in the build container when the
    150k-point golden fixture was minted (oracle/make_golden_scene.py stores the time next to the vectors): the Python
    reference cannot travel to the GPU box, so this is a recorded number, not one measured in this run."""
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "seggroup150k_s12_train_g4.npz"))
        sec, thr, n = float(z["reference_cpu_seconds"]), int(z["reference_cpu_threads"]), int(z["scene"][1])
        return {"value": n / sec, "unit": UNIT, "seconds_per_scene": sec, "cores": thr, "where": "build container (8 vCPU), fwd+bwd of one 150,000-point scene",
                "source": "tests/golden/seggroup150k_s12_train_g4.npz"}
    except Exception:
        return None


def run_reference_arm(args):
    """The reference's CPU implementation of the path on this box's host cores: the oracle port (the unmodified Python reference
    lives in /root/reference, which does not exist here).  One step = forward + backward of ONE scene of the bench workload's
    size (150,000 points), i.e. one eighth of the 8-scene batch; at most --ref-max-steps steps are timed so that the arm ends
    within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.points
    steps = max(1, min(args.steps, args.ref_max_steps))
    pps, sec, cores = cpu_port_points_per_sec(n, steps=steps, warmup=min(args.warmup, 1))
    sample = "1 scene x %d points fwd+bwd per step (one of the %d scenes of the batch), %d steps timed" % (n, args.scenes, steps)
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SegGroup training step fwd+bwd+SGD, %d scenes x %d points per GPU (BASELINE configs[1])" % (args.scenes, args.points),
                       "reference_arm": "oracle port of seggroup/model.py (torch-CPU + numpy) on host cores; the Python reference does not travel to the GPU box"},
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "unmodified_reference": unmodified_reference_record()},
            "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def ncu_record(kernel_substr):
    """Per-launch metrics of a kernel from the committed `ncu --set full` summary of THIS round (tools/ncu_summary.py --json),
    or None: nothing is typed in by hand."""
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True):
        if name[:3] in ("r02", "r03") and name.endswith("_ncu.json"):
            try:
                rows = json.load(open(os.path.join(ROOT, "profiles", name)))
            except Exception:
                continue
            hits = [r for r in rows if kernel_substr in r.get("kernel", "")]
            if hits:
                r = max(hits, key=lambda r: r.get("time_us", 0.0))
                return dict(r, source="profiles/" + name)
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--points", type=int, default=150000)
    ap.add_argument("--ref-max-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 3 / config 4 / first-epoch legs")
    ap.add_argument("--shard-scenes", type=int, default=1201, help="config 3: scenes of the whole job (len(scannetv2_train.txt))")
    ap.add_argument("--lanes", type=int, default=1, help="scene batches in flight per GPU: the batch is split into this many block-diagonal "
                    "sub-batches, each driven by its own host thread / CUDA stream (1 = one batch on the caller's stream)")
    ap.add_argument("--sampler", default="nvml", choices=["nvml", "smi", "off"], help="clock sampler during the timed region")
    ap.add_argument("--gc", default="frozen", choices=["frozen", "default"],
                    help="frozen: collect + gc.freeze() after setup and no cyclic collections inside the timed regions")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import gc
    import shutil
    import tempfile
    from seggroup_b200 import _lib, engine, pipeline, synth
    from seggroup_b200.model import SegModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the seggroup_b200 product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL's own banner / debug lines (NCCL_DEBUG=VERSION|INFO in the environment) go to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/seggroup_b200_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    B, N = args.scenes, args.points

    # ---- synthetic batch (seeded per rank) written as the on-disk tree the reference's loader / model read
    scenes_host = [synth.make_scene(1000 * rank + i, N, name="scene%04d_%02d" % (rank, i)) for i in range(B)]
    # the scene tree and the label files live on tmpfs when the box has one with room: the container's root filesystem serialises
    # file creation + 600 KB writes (measured in the build container: 112 label files take 73 ms on /tmp whatever the number of
    # writer threads, 18 ms on /dev/shm), which would time the overlay filesystem rather than the plugin
    tree_root = None
    for cand in (os.environ.get("SGB_BENCH_TREE"), "/dev/shm"):
        if cand and os.path.isdir(cand) and os.access(cand, os.W_OK):
            st_ = os.statvfs(cand)
            if st_.f_bavail * st_.f_frsize > (4 << 30):
                tree_root = cand
                break
    tree = tempfile.mkdtemp(prefix="sgb_bench_r%d_" % rank, dir=tree_root)
    synth.write_scene_tree(tree, scenes_host)
    os.chdir(tree)
    torch.manual_seed(1)
    model = SegModel(exp_name="bench").to(dev)
    with torch.no_grad():
        model.mlp_1.bn1.weight.mul_(GSCALE)
    model.epoch = "1"
    model.scene_cache_dir = os.path.join(tree, "csr_cache")                                   # binary CSR cache of the side files (N2)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)       # train.py:96-97
    params = model._params()
    grads_of = [p for p in model.parameters()]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                         # > 126 MB L2
    # what seggroup/data.py hands the step: pinned [B,N,6] f32, [B,N,2] i64, [B,1] i64
    data_h = torch.stack([torch.from_numpy(s.data) for s in scenes_host]).pin_memory()
    weak_h = torch.stack([torch.from_numpy(s.weak_label.astype(np.int64)) for s in scenes_host]).pin_memory()
    info_h = torch.arange(B).view(B, 1)
    h2d_bytes = data_h.numel() * 4 + weak_h.numel() * 8 + info_h.numel() * 8
    resident = pipeline.SceneDevice.concat([pipeline.SceneDevice.from_host(s) for s in scenes_host])
    engine.reserve_current_stream(6 << 30, device=dev)
    metrics_buf = torch.zeros(165, device=dev)

    def finish_step(loss_raw, metrics):
        """loss = mean over scenes of loss_sum / loss_num (train.py:165-170 at B ranks), backward, ONE all-reduce of the flat
        gradient buffer with the 165 logging floats of train.py:172-175 appended, SGD step."""
        loss = pipeline.batch_loss(loss_raw)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if dist is not None:
            extra = torch.cat([loss.detach().view(1), torch.stack([m[0] for m in metrics]).sum(0).view(-1), torch.stack([m[1] for m in metrics]).sum(0).view(-1),
                               torch.stack([m[2] for m in metrics]).sum(0).view(-1)])
            engine.allreduce_flat([p.grad for p in grads_of if p.grad is not None], dist, average=True, extra=extra)
        opt.step()
        return loss.detach()

    lanes = max(1, min(args.lanes, B))
    ex = None
    if lanes > 1:
        ex = engine.SceneExecutor(dev, n_streams=lanes, fused=False, reserve_bytes_per_stream=4 << 30)
        per = (B + lanes - 1) // lanes
        lane_batches = [pipeline.SceneDevice.concat([pipeline.SceneDevice.from_host(s) for s in scenes_host[i:i + per]]) for i in range(0, B, per)]
        train_keys = list(params.keys()) + ["classifier.linear1.weight", "classifier.bn1.weight", "classifier.bn1.bias", "classifier.linear2.weight", "classifier.linear2.bias"]
        all_params = dict(params, **{"classifier." + k: v for k, v in model.classifier.named_parameters()})

    def step_resident(batch=resident, single_lane=False):
        flush_buf.fill_(0)                                  # evict L2 between steps (inside the timed region, ~40 us)
        if ex is not None and batch is resident and not single_lane:
            opt.zero_grad(set_to_none=True)
            loss = ex.train_batch(lane_batches, all_params, train_keys, classifier=model.classifier)
            if dist is not None:
                engine.allreduce_flat([p.grad for p in grads_of if p.grad is not None], dist, average=True)
            opt.step()
            return loss
        r = pipeline.forward_scene(batch, params, mode="train", classifier=model.classifier)
        return finish_step(r.loss_raw, r.metrics_scenes)

    # What a DataLoader(pin_memory=True) + prefetching iterator gives train.py:161-163: the H2D copies of the NEXT step's batch are
    # issued on a copy stream while the current step computes; every step still copies its own inputs from pinned host memory
    # inside the timed region (h2d_bytes_per_step), the compute stream waits for them with an event.
    copy_stream = torch.cuda.Stream(device=dev)
    inflight = []

    def h2d_enqueue():
        with torch.cuda.stream(copy_stream):
            d, w = data_h.to(dev, non_blocking=True), weak_h.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        inflight.append((d, w, ev))

    def h2d_next():
        if not inflight:
            h2d_enqueue()
        d, w, ev = inflight.pop(0)
        torch.cuda.current_stream(dev).wait_event(ev)
        d.record_stream(torch.cuda.current_stream(dev)); w.record_stream(torch.cuda.current_stream(dev))
        h2d_enqueue()                                       # the following step's batch
        return d, w

    def step_e2e():
        flush_buf.fill_(0)
        d, w = h2d_next()
        out = model(d, w, info_h)                           # the plugin call of train.py:163 (B scenes per call)
        loss = finish_step(out[0], model.last_result.metrics_scenes)
        return float(loss.item())                           # device -> host read of the step's loss

    def timed(fn, n_steps, after=None):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(n_steps):
            last = fn()
        if after is not None:
            after()                                         # e.g. wait for the label files of the region
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    def freeze():
        if args.gc == "frozen":
            # The step is driven from Python; a generation-2 cyclic collection over the interpreter's whole heap (torch, numpy, the
            # scene objects: measured 65-400 ms) inside a timed step would be charged to that step.  Everything that exists after
            # set-up and warm-up is moved to the permanent generation and the collector is paused (reference counting still frees
            # the per-step tensors; a collection runs between the regions).
            gc.collect()
            gc.freeze()
            gc.disable()

    pts_per_step = B * N * world
    # ---- timed region 1 (headline `value`): inputs resident in HBM, no instrumentation inside
    for _ in range(args.warmup):
        step_resident()
    freeze()
    sampler = ClockSampler(local)
    if args.sampler == "smi":
        sampler.nv = None
    if rank == 0 and args.sampler != "off":
        sampler.start()
    launches0 = _lib.launch_count()
    ms, _ = timed(step_resident, args.steps)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if (rank == 0 and args.sampler != "off") else None
    value = pts_per_step * args.steps / (ms * 1e-3)

    # ---- kernel timing pass (separate from the headline region, directly after it: the later legs leave loader / writer threads and a
    # second model behind, and with them the entry times came out 1.7x longer than in this position): CUDA events around the entry points named below
    _lib.time_entry = {"sgb_edgeconv_fwd", "sgb_edgeconv_bwd", "sgb_segment_pool_max_fwd", "sgb_segment_pool_max_bwd", "sgb_cluster_knn_scenes",
                       "sgb_gcn_agg_fwd", "sgb_centralize", "sgb_export_labels_scenes"}
    _lib.timed_events = []
    for _ in range(2):
        step_resident(single_lane=True)                     # one batch on one stream: the events bracket one entry at a time
    torch.cuda.synchronize()
    kernel_events = _lib.timed_events
    _lib.time_entry = None


    # ---- timed region 2 (`e2e`): the reference-facing plugin call with HOST buffers — SegModel.forward on the scene tree (side
    # files parsed by the model and kept in HBM after their first use, 14 label files per scene written by the writer threads,
    # the region ends when they are on disk), H2D of the loader's tensors and D2H of labels + loss inside the region
    gc.collect()
    for _ in range(max(2, min(args.warmup, 3))):
        step_e2e()
    model.flush_exports()
    d2h0 = model.d2h_bytes
    ms_e2e, _ = timed(step_e2e, args.steps, after=model.flush_exports)
    d2h_bytes = (model.d2h_bytes - d2h0) // args.steps + 4
    e2e = pts_per_step * args.steps / (ms_e2e * 1e-3)

    # ---- pseudo-label inference (ins_infer): resident, and end to end through the plugin (labels copied to the host + written)
    gc.collect()
    model_inf = SegModel(exp_name="bench_inf", ins_infer=True).to(dev)
    model_inf.load_state_dict(model.state_dict())
    model_inf.epoch = "ins_infer"
    model_inf.scene_cache_dir = model.scene_cache_dir

    def infer_resident():
        flush_buf.fill_(0)
        with torch.no_grad():
            return pipeline.forward_scene(resident, params, mode="ins_infer")

    def infer_e2e():
        flush_buf.fill_(0)
        with torch.no_grad():
            d, w = h2d_next()
            return model_inf(d, w, info_h)

    for _ in range(2):
        infer_resident(); infer_e2e()
    model_inf.flush_exports()
    ms_inf, _ = timed(infer_resident, args.steps)
    d2h0 = model_inf.d2h_bytes
    ms_inf_e2e, _ = timed(infer_e2e, args.steps, after=model_inf.flush_exports)
    inf_d2h = (model_inf.d2h_bytes - d2h0) // args.steps

    extra = {}
    if not args.no_extra:
        # ---- first epoch: nothing cached in HBM, side files read from the binary CSR cache every step (N2)
        model.cache_scenes = False
        step_e2e()
        ms_cold, _ = timed(step_e2e, 2, after=model.flush_exports)
        model.cache_scenes = True
        extra["e2e_first_epoch"] = {"value": pts_per_step * 2 / (ms_cold * 1e-3), "unit": UNIT, "ms_per_step": ms_cold / 2,
                                    "what": "same plugin call with the HBM scene cache off: adj / unmap / seg CSR read from the on-disk binary cache and uploaded every step"}
        # ---- loader-fed step (N2): a background thread reads the tree (binary CSR cache), collates the 8 scenes into pinned
        # block-diagonal host arrays and the step uploads them — nothing cached in HBM, every input crosses PCIe every step
        from seggroup_b200.loader import SceneShardLoader
        n_ld = max(3, args.steps)
        ld = iter(SceneShardLoader(model.scene_list, data_root=model.data_root, batch_size=B, cache_dir=model.scene_cache_dir,
                                   prefetch=2, epochs=n_ld + 2))
        ld_bytes = []

        def step_loader():
            hb = next(ld)
            ld_bytes.append(hb.nbytes)
            loss = float(step_resident(hb.to_device(dev)).item())
            hb.recycle()                                    # the step has completed: the pinned buffers go back to the loader's pool
            return loss
        step_loader(); step_loader()
        ms_ld, _ = timed(step_loader, n_ld)
        extra["e2e_shard_loader"] = {"value": pts_per_step * n_ld / (ms_ld * 1e-3), "unit": UNIT, "ms_per_step": ms_ld / n_ld, "h2d_bytes_per_step": int(ld_bytes[-1]),
                                     "what": "training step fed by seggroup_b200.loader.SceneShardLoader (prefetch thread, binary CSR cache, pinned collated "
                                             "batch, 7 H2D copies per step, loss read back); no label files"}
        # ---- BASELINE configs[3] unit: 1 scene per GPU per step (the reference's own batch size, train.py:92)
        one = pipeline.SceneDevice.from_host(scenes_host[0])
        for _ in range(2):
            step_resident(one)
        ms1, _ = timed(lambda: step_resident(one), max(3, args.steps))
        extra["config4_one_scene_per_gpu"] = {"value": N * world * max(3, args.steps) / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1 / max(3, args.steps),
                                              "workload": "data-parallel training, 1 scene x %d points per GPU per step, one flat all-reduce" % N}
        # ---- BASELINE configs[2]: pseudo-label generation over 1,201 scenes sharded rank::world, no data-path collective, one
        # final metric reduce.  The 8 resident scenes stand in for the rank's shard (scene i of the shard = resident scene i % 8).
        mine = engine.shard_scenes(args.shard_scenes, rank, world)
        n_batches = (len(mine) + B - 1) // B
        acc = torch.zeros(164, device=dev)

        def shard_pass():
            acc.zero_()
            with torch.no_grad():
                for _ in range(n_batches):
                    r = pipeline.forward_scene(resident, params, mode="ins_infer")
                    for m in r.metrics_scenes:
                        acc[:80] += m[0].view(-1); acc[80:160] += m[1].view(-1); acc[160:] += m[2]
            if dist is not None:
                dist.all_reduce(acc)                        # the one collective of the job
            return acc
        ms_sh, _ = timed(shard_pass, 1)
        done = n_batches * B                                # scenes processed by this rank (last batch padded to B)
        extra["config3_sharded_inference"] = {"value": done * world * N / (ms_sh * 1e-3), "unit": UNIT, "seconds": ms_sh * 1e-3,
                                              "scenes_total": args.shard_scenes, "scenes_per_rank": len(mine), "batches_per_rank": n_batches,
                                              "workload": "ins_infer over %d scenes x %d points sharded rank::world in batches of %d, labels exported to HBM, "
                                                          "no data-path collective, one final all-reduce of the metrics" % (args.shard_scenes, N, B)}

    if rank == 0 and os.environ.get("SGB_BENCH_DEBUG"):
        byk = {}
        for (n_, t_, a, c) in kernel_events:
            byk.setdefault((n_, t_), []).append(a.elapsed_time(c))
        for k, v in sorted(byk.items(), key=lambda kv: -sum(kv[1])):
            sys.stderr.write("[entry] %-28s tag %-9s n %3d  mean %.4f  min %.4f  max %.4f ms\n" % (k[0], k[1], len(v), np.mean(v), min(v), max(v)))
    if rank == 0:
        peaks, peak_kind = measured_peaks()

        def ev_ms(name, tag=None):
            if tag == "max":                                    # the widest call of that entry point (the point-sized pooling of a batch)
                tags = [t_ for (n_, t_, a, c) in kernel_events if n_ == name]
                tag = max(tags) if tags else None
            v = [a.elapsed_time(c) for (n_, t_, a, c) in kernel_events if n_ == name and (tag is None or t_ == tag)]
            return (float(np.mean(v)), len(v)) if v else (None, 0)

        # Dominant entry point by device time (profiles/): EdgeConv forward of MLP3 = first-layer moments (SIMT) + the fused
        # tcgen05 kernel + BN2/LeakyReLU apply, one call per scene (BatchNorm statistics are per scene).  GEMM-shaped work ->
        # tensor roofline: useful flops = both 1x1 convolutions over the N*20 edges; the 64x64 one runs as TF32 x 3 (fp32
        # parity), i.e. 6x the cost of the same contraction in bf16.
        k_ms, k_n = ev_ms("sgb_edgeconv_fwd", 1)
        flops = 2.0 * N * 20 * (18 * 64 + 64 * 64)
        achieved = flops / (k_ms * 1e-3) / 1e12 if k_ms else None
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        ncu = ncu_record("ec2_tc1_kernel<1, 1>") or ncu_record("ec2_tc1_kernel<true, true>")
        traffic = None
        if ncu and ncu.get("dram_bytes") is not None:
            traffic = ncu["dram_bytes"] * (N / float(ncu.get("points", 150000)))
        roofline = {"bound": "tensor", "kernel": "sgb_edgeconv_fwd two_layer (gram1_pt_kernel + ec2_tc1_kernel<ARG,GRAM> [both layers on tcgen05, kind::tf32 x3] + ec2_apply_kernel), one launch per scene",
                    "achieved": achieved, "peak": peak_tf, "peak_kind": peak_kind + " bf16 dense, sustained (tf32 runs at half of it, the x3 split costs 3 MMAs)",
                    "unit": "TFLOP/s", "frac": achieved / peak_tf if achieved else None, "traffic": traffic, "ms_per_launch": k_ms,
                    "algorithmic_flops_per_launch": flops, "launches_timed": k_n,
                    "frac_of_tf32x3_ceiling": achieved / (peak_tf / 6.0) if achieved else None, "ncu": ncu}
        roofline_more = []

        def hbm_row(entry, label, nbytes, tag=None, ncu_key=None):
            t_ms, n_ = ev_ms(entry, tag)
            rec = ncu_record(ncu_key) if ncu_key else None
            roofline_more.append({"bound": "hbm", "kernel": label, "achieved": nbytes / (t_ms * 1e-3) / 1e9 if t_ms else None, "peak": peaks["hbm_gbs"],
                                  "unit": "GB/s", "frac": nbytes / (t_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if t_ms else None, "ms_per_launch": t_ms,
                                  "algorithmic_bytes_per_launch": int(nbytes), "launches_timed": n_,
                                  "traffic": rec.get("dram_bytes") if rec else None, "ncu": rec})
        S_tot = sum(s.n_segments for s in scenes_host)
        NT = B * N
        hbm_row("sgb_segment_pool_max_fwd", "sgb_segment_pool_max_fwd [B*N,64] point -> segment max pooling with arg-max, one launch per batch (gather)",
                NT * (4 * 64 + 4) + 16 * S_tot * 64, tag="max", ncu_key="segment_pool_staged_kernel")
        hbm_row("sgb_cluster_knn_scenes", "sgb_cluster_knn_scenes [B*N] exact per-cluster kNN(20) (FP32-issue bound; bytes = 16 N + 4 N k)", NT * 96, ncu_key="knn_sweep_kernel")
        hbm_row("sgb_centralize", "sgb_centralize [B*N] cluster means + per-point subtraction (scatter + stream)", NT * 64, ncu_key="centralize_kernel")
        hbm_row("sgb_export_labels_scenes", "sgb_export_labels_scenes [B*N_raw] label gather through unmap", NT * 24, ncu_key="export_labels_kernel")
        # backward of MLP3: sparse arg-max edges + the dense pass over all N*20 edges on tcgen05 (Bm h per edge, TF32 x 3)
        b_ms, b_n = ev_ms("sgb_edgeconv_bwd", 1)
        bflops = 2.0 * N * 20 * (18 * 64 + 64 * 64 + 64 * 18)
        roofline_more.append({"bound": "tensor", "kernel": "sgb_edgeconv_bwd two_layer (bwd_sparse + ec2_bwd_tc_kernel [tcgen05 kind::tf32 x3] + finalize), one launch per scene",
                              "achieved": bflops / (b_ms * 1e-3) / 1e12 if b_ms else None, "peak": peak_tf, "unit": "TFLOP/s",
                              "frac": bflops / (b_ms * 1e-3) / 1e12 / peak_tf if b_ms else None, "ms_per_launch": b_ms,
                              "algorithmic_flops_per_launch": bflops, "launches_timed": b_n, "ncu": ncu_record("ec2_bwd_tc_kernel")})
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            pps, sec, cores = cpu_port_points_per_sec(N)
            cpu = {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "1 scene x %d points fwd+bwd (%.1f s): one of the %d scenes of the batch" % (N, sec, B),
                   "unmodified_reference": unmodified_reference_record()}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "SegGroup training step fwd+bwd+SGD, %d scenes x %d points per GPU (BASELINE configs[1])" % (B, N),
                           "weights": "torch.manual_seed(1) default init, mlp_1.bn1.weight x %g" % GSCALE, "l2": "256 MiB flush buffer written every step",
                           "parallelism": "dp%d" % world if world > 1 else "single",
                           "tree": "scene tree + label files under %s" % (tree_root or tempfile.gettempdir()),
                           "batching": "the %d scenes run as %d block-diagonal scene batch(es) (per-scene BatchNorm / grouping / labels)" % (B, lanes)},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e / args.steps,
                        "api": "seggroup_b200.model.SegModel.forward(data [B,N,6], weak_label [B,N,2], info [B,1]) on a scene tree on disk: pinned host "
                               "tensors -> H2D (every step, on a copy stream one step ahead, as a prefetching pinned DataLoader does), forward, backward, "
                               "all-reduce, SGD, loss.item(); 14 label files per scene copied D2H and written; "
                               "the region ends when the files are on disk; parsed side files stay in HBM after their first use"},
                "inference": {"value": pts_per_step * args.steps / (ms_inf * 1e-3), "unit": UNIT, "ms_per_step": ms_inf / args.steps,
                              "workload": "pseudo-label generation (ins_infer, 14 label vectors per scene exported to HBM) over the same batch",
                              "e2e": {"value": pts_per_step * args.steps / (ms_inf_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_inf_e2e / args.steps,
                                      "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(inf_d2h),
                                      "api": "SegModel(ins_infer=True).forward from pinned host tensors; label vectors copied to the host and the 14 files per scene written"}},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_more": roofline_more, "cpu_baseline": cpu}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    os.chdir(ROOT)
    shutil.rmtree(tree, ignore_errors=True)


if __name__ == "__main__":
    main()
