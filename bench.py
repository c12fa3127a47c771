#!/usr/bin/env python
"""bench.py — points/sec of the SegGroup hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scenes 8] [--points 150000]

A "step" is one training step of BASELINE.json configs[1]: forward + backward + SGD update over a batch of
`--scenes` synthetic ScanNet-v2-shaped scenes of `--points` points each (per-scene BatchNorm statistics,
gradient = mean over scenes of loss_sum/loss_num, reference semantics of train.py:165-170 at 8 ranks).
With N > 1 (torchrun, one rank per GPU) every rank runs its own batch (weak scaling) and the flat fp32
gradient buffer is all-reduced once per step over NCCL.

One JSON line is printed by rank 0:
  value        whole-job points/sec, scene inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the host-facing call: every step copies the scenes' inputs from pinned host
               memory and reads the loss back
  roofline     dominant entry point (EdgeConv forward of MLP3, tcgen05): useful flops / CUDA-event time vs the measured tensor peak;
               roofline_more: the HBM-bound gather kernel (segment pooling) vs the measured copy bandwidth, the tcgen05 backward
  inference    pseudo-label generation (ins_infer) points/sec over the same batch
  cpu_baseline the oracle port of the reference CPU path, timed on a bounded sample on this box's host cores
`--impl reference` times that CPU port alone, as the reference arm.
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

    def __init__(self, index=0, interval=0.05):
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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_points
    pps, sec, cores = cpu_port_points_per_sec(n, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    sample = "1 scene x %d points fwd+bwd per step (bounded sample of the %d x %d batch)" % (n, args.scenes, args.points)
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SegGroup training step fwd+bwd, %d scenes x %d points per GPU (configs[1])" % (args.scenes, args.points),
                       "reference_arm": "oracle port of seggroup/model.py on host cores (the Python reference does not travel to the GPU box)"},
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--points", type=int, default=150000)
    ap.add_argument("--ref-points", type=int, default=30000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=4, help="scenes in flight per GPU (CUDA streams / host threads)")
    ap.add_argument("--sampler", default="nvml", choices=["nvml", "smi", "off"], help="clock sampler during the timed region")
    ap.add_argument("--switch-interval", type=float, default=0.0, help="sys.setswitchinterval for the scene threads (0 = leave)")
    ap.add_argument("--gc", default="frozen", choices=["frozen", "default"],
                    help="frozen: collect + gc.freeze() after setup and no cyclic collections inside the timed regions")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.switch_interval > 0:
        sys.setswitchinterval(args.switch_interval)

    from seggroup_b200 import _lib, engine, pipeline, synth
    from seggroup_b200.params import TRAINABLE, init_params

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

    # ---- synthetic batch (seeded per rank), pinned on the host and resident on the device
    scenes_host = [synth.make_scene(1000 * rank + i, args.points) for i in range(args.scenes)]
    pinned = []
    for s in scenes_host:
        pinned.append({k: torch.as_tensor(np.ascontiguousarray(v)).pin_memory() for k, v in
                       dict(data=s.data, weak=s.weak_label.astype(np.int32), seg_off=s.seg_offsets.astype(np.int32),
                            seg_members=s.seg_members.astype(np.int32), adj=s.adj.astype(np.int32), unmap=s.unmap, real=s.real_label).items()})
    h2d_bytes = sum(sum(t.numel() * t.element_size() for t in d.values()) for d in pinned)

    def upload(d):
        t = {k: v.to(dev, non_blocking=True) for k, v in d.items()}
        return pipeline.SceneDevice(data=t["data"], weak_label=t["weak"], seg_off=t["seg_off"], seg_members=t["seg_members"], adj0=t["adj"],
                                    unmap=t["unmap"], real_label=t["real"])

    resident = [upload(d) for d in pinned]
    torch.manual_seed(1)
    p = {k: v.to(dev) for k, v in init_params(1, GSCALE).items()}
    train_keys = list(TRAINABLE)
    for k in train_keys:
        p[k].requires_grad_(True)
    opt = torch.optim.SGD([p[k] for k in train_keys], lr=0.1, momentum=0.9, weight_decay=1e-4)      # train.py:96-97
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)                                 # > 126 MB L2

    def allreduce_grads():
        if dist is not None:
            engine.allreduce_flat([p[k].grad for k in train_keys], dist, average=True)

    ex = engine.SceneExecutor(dev, n_streams=args.streams)

    def step(scenes, from_host):
        flush_buf.fill_(0)                                  # evict L2 between steps (inside the timed region, ~40 us)
        opt.zero_grad(set_to_none=False)
        total = ex.train_batch(scenes, p, train_keys, upload=upload if from_host else None)
        allreduce_grads()
        opt.step()
        return total

    def timed(scenes, from_host, n_steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        marks = [time.perf_counter()]
        mallocs = []
        for _ in range(n_steps):
            last = step(scenes, from_host)
            if from_host:
                last = float(last.item())                   # device -> host read of the step's loss
            marks.append(time.perf_counter())
            if os.environ.get("SGB_BENCH_DEBUG"):
                ms_ = torch.cuda.memory_stats(dev)
                mallocs.append((ms_.get("num_device_alloc", 0), ms_.get("num_device_free", 0), ms_.get("num_alloc_retries", 0)))
        if os.environ.get("SGB_BENCH_DEBUG") and rank == 0:  # host-side progress per step (not part of the measurement)
            sys.stderr.write("host ms per step (%s): %s\n" % ("e2e" if from_host else "resident",
                                                               " ".join("%.1f" % ((b - a) * 1e3) for a, b in zip(marks[:-1], marks[1:]))))
            sys.stderr.write("cudaMalloc/cudaFree/retries after each step: %s\n" % " ".join("%d/%d/%d" % m for m in mallocs))
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), last

    import gc
    if os.environ.get("SGB_BENCH_DEBUG"):
        t_gc = {}

        def _gc_cb(phase, info):
            if phase == "start":
                t_gc["t"] = time.perf_counter()
            else:
                dt = (time.perf_counter() - t_gc.get("t", time.perf_counter())) * 1e3
                if dt > 5.0:
                    sys.stderr.write("gc gen%d took %.1f ms (collected %d)\n" % (info["generation"], dt, info["collected"]))
        gc.callbacks.append(_gc_cb)
    for _ in range(args.warmup):
        step(resident, False)
    if args.gc == "frozen":
        # The step is driven by Python threads; a generation-2 cyclic collection over the interpreter's whole heap (torch,
        # numpy, the scene objects: measured 100-400 ms) inside a timed step would be charged to that step.  Everything that
        # exists after set-up and warm-up is moved to the permanent generation, and the collector is paused during the timed
        # regions (reference counting still frees the per-step tensors; a collection runs between the regions).
        gc.collect()
        gc.freeze()
        gc.disable()
    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local)
    if args.sampler == "smi":
        sampler.nv = None
    if rank == 0 and args.sampler != "off":
        sampler.start()
    launches0 = _lib.launch_count()
    _lib.time_entry = {"sgb_edgeconv_fwd", "sgb_edgeconv_bwd", "sgb_segment_pool_max_fwd"}
    _lib.timed_events = []
    ms, _ = timed(resident, False, args.steps)
    launches = _lib.launch_count() - launches0
    kernel_events = _lib.timed_events
    _lib.time_entry = None
    clocks = sampler.stop() if (rank == 0 and args.sampler != "off") else None
    pts_per_step = args.scenes * args.points * world
    value = pts_per_step * args.steps / (ms * 1e-3)
    # ---- timed region 2: end to end from pinned host buffers
    if args.gc == "frozen":
        gc.collect()
    step(pinned, True)
    ms_e2e, _ = timed(pinned, True, args.steps)
    e2e = pts_per_step * args.steps / (ms_e2e * 1e-3)

    # ---- pseudo-label inference (ins_infer) over the same resident batch: reported beside the training number
    if args.gc == "frozen":
        gc.collect()
    with torch.no_grad():
        ex.infer_batch(resident, p)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            flush_buf.fill_(0)
            ex.infer_batch(resident, p)
        e1.record()
        torch.cuda.synchronize()
    ms_inf = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist is not None:
        dist.all_reduce(ms_inf, op=dist.ReduceOp.MAX)
    ms_inf = float(ms_inf.item())

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        N = args.points
        # Dominant entry point by device time (profiles/): EdgeConv forward of MLP3 = first-layer moments (SIMT) + the fused
        # tcgen05 kernel + BN2/LeakyReLU apply.  GEMM-shaped work -> tensor roofline: useful flops = both 1x1 convolutions over
        # the N*20 edges; the 64x64 one runs as TF32 x 3 (fp32 parity), i.e. 6x the cost of the same contraction in bf16.
        ev = [a.elapsed_time(c) for (name, tag, a, c) in kernel_events if name == "sgb_edgeconv_fwd" and tag == 1]
        k_ms = float(np.mean(ev)) if ev else None
        flops = 2.0 * N * 20 * (18 * 64 + 64 * 64)
        achieved = flops / (k_ms * 1e-3) / 1e12 if k_ms else None
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        # traffic / pipe utilisation of the dominant kernel of this entry (ec2_tc_kernel<1,1>) from the ncu --set full capture of
        # the same workload (profiles/r01i_ncu_full_train_150k.txt: dram__bytes_read.sum + dram__bytes_write.sum per launch at
        # N = 150,000; scaled linearly in N for other sizes)
        traffic = (19.34e6 + 91.79e6) * (N / 150000.0)
        roofline = {"bound": "tensor", "kernel": "sgb_edgeconv_fwd two_layer (gram1_pt_kernel + ec2_tc_kernel<ARG,GRAM> [tcgen05 kind::tf32 x3] + ec2_apply_kernel)",
                    "achieved": achieved, "peak": peak_tf, "peak_kind": peak_kind + " bf16 dense (tf32 runs at half of it, the x3 split costs 3 MMAs)",
                    "unit": "TFLOP/s", "frac": achieved / peak_tf if achieved else None, "traffic": traffic, "ms_per_launch": k_ms,
                    "algorithmic_flops_per_launch": flops, "launches_timed": len(ev),
                    "frac_of_tf32x3_ceiling": achieved / (peak_tf / 6.0) if achieved else None,
                    "ncu": {"source": "profiles/r01i_ncu_full_train_150k.txt", "kernel": "ec2_tc_kernel<1,1>", "us_alone": 855.1,
                            "tensor_pipe_active_pct": 22.6, "l1tex_throughput_pct": 94.1, "limiter": "shared-memory bandwidth (operand tiles + transposed Gram copy)"},
                    "note": "timed with CUDA events on the launching stream while %d scenes are in flight (alone: 1.00 ms, "
                            "profiles/r01i_kernels_150k_500k.json)" % args.streams}
        # the HBM-bound gather/scatter kernel of the path: point -> segment max pooling (sgb_segment_pool_max_fwd on [N,64])
        evp = [a.elapsed_time(c) for (name, tag, a, c) in kernel_events if name == "sgb_segment_pool_max_fwd" and tag == N]
        p_ms = float(np.mean(evp)) if evp else None
        pool_bytes = N * (4 * 64 + 4) + 16 * 1100 * 64
        roofline_more = [{"bound": "hbm", "kernel": "sgb_segment_pool_max_fwd [N,64] (memset + segment_pool_staged_kernel + decode)",
                          "traffic": 39.38e6 * (N / 150000.0), "ncu_kernel_us_alone": 16.4,
                          "achieved": pool_bytes / (p_ms * 1e-3) / 1e9 if p_ms else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": pool_bytes / (p_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if p_ms else None, "ms_per_launch": p_ms,
                          "algorithmic_bytes_per_launch": pool_bytes, "launches_timed": len(evp)}]
        # backward of MLP3: sparse arg-max edges + the dense pass over all N*20 edges on tcgen05 (Bm h per edge, TF32 x 3)
        evb = [a.elapsed_time(c) for (name, tag, a, c) in kernel_events if name == "sgb_edgeconv_bwd" and tag == 1]
        b_ms = float(np.mean(evb)) if evb else None
        bflops = 2.0 * N * 20 * (18 * 64 + 64 * 64 + 64 * 18)
        roofline_more.append({"bound": "tensor", "kernel": "sgb_edgeconv_bwd two_layer (bwd_sparse + ec2_bwd_tc_kernel [tcgen05 kind::tf32 x3] + finalize)",
                              "achieved": bflops / (b_ms * 1e-3) / 1e12 if b_ms else None, "peak": peak_tf, "unit": "TFLOP/s",
                              "frac": bflops / (b_ms * 1e-3) / 1e12 / peak_tf if b_ms else None, "ms_per_launch": b_ms,
                              "algorithmic_flops_per_launch": bflops, "launches_timed": len(evb)})
        cpu = None
        if not args.no_cpu_baseline:
            pps, sec, cores = cpu_port_points_per_sec(args.ref_points)
            cpu = {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "1 scene x %d points fwd+bwd (%.1f s) of the %d x %d batch" % (args.ref_points, sec, args.scenes, args.points)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "SegGroup training step fwd+bwd+SGD, %d scenes x %d points per GPU (BASELINE configs[1])" % (args.scenes, args.points),
                           "weights": "torch.manual_seed(1) default init, mlp_1.bn1.weight x %g" % GSCALE, "l2": "256 MiB flush buffer written every step",
                           "parallelism": "dp%d" % world if world > 1 else "single", "scenes_in_flight": args.streams},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                "inference": {"value": pts_per_step * args.steps / (ms_inf * 1e-3), "unit": UNIT, "ms_per_step": ms_inf / args.steps,
                              "workload": "pseudo-label generation (ins_infer incl. label export to HBM) over the same batch"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_more": roofline_more, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
