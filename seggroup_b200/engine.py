"""Scene executor for one GPU + the multi-GPU plumbing.

A SegGroup scene is a chain of ~600 kernels of which many are latency bound (single-CTA sequential union replays, CSR builds over
a few thousand rows) and a handful are wide (EdgeConv, kNN).  One scene therefore cannot fill 148 SMs; the unit of parallelism
on a B200 is the scene (scenes are independent, BatchNorm statistics are per scene).

`SceneExecutor` (default `fused=True`): the scenes of a step are concatenated into ONE block-diagonal scene batch
(`pipeline.SceneDevice.concat`) and driven by one forward / backward on the caller's stream — every graph / kNN / pooling / export
launch serves all scenes, the order-dependent replays run one CTA per scene (round 2: 56 -> 25 ms per 8 x 150k step).
`fused=False` is the round-1 executor: `n_streams` host threads, each bound to its own CUDA stream, every thread driving whole
scenes (or sub-batches) so that the narrow kernels of one overlap the wide kernels of the others.

Gradient semantics of a training batch = the reference at `len(scenes)` ranks (train.py:165-170 + DDP average): mean over scenes
of loss_sum / loss_num, per-scene contributions summed in scene order (independent of thread scheduling).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import torch

from . import pipeline


_reserved = set()      # (device index, raw stream) pairs whose allocator pool has been pre-sized


def reserve_current_stream(nbytes: int = 1 << 30, device=None):
    """Pre-size the caching allocator's pool of the CURRENT stream once (see SceneExecutor.reserve for why)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    if key in _reserved or nbytes <= 0:
        return
    _reserved.add(key)
    big = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    small = [torch.empty(1 << 20, dtype=torch.uint8, device=dev) for _ in range(64)]
    tiny = [torch.empty(256 << 10, dtype=torch.uint8, device=dev) for _ in range(128)]
    del big, small, tiny


class SceneExecutor:
    def __init__(self, device=None, n_streams: int = 4, reserve_bytes_per_stream: int = 3 << 30, fused: bool = True):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.fused = bool(fused)
        self.last_result = None
        if self.fused:
            n_streams = 0                      # fused batches run on the caller's stream: no lanes, no host threads
        self.n_streams = max(0 if self.fused else 1, int(n_streams))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_streams)]
        self._pool = ThreadPoolExecutor(max_workers=max(1, self.n_streams), thread_name_prefix="sgb-scene")
        self.reserve(reserve_bytes_per_stream)

    def reserve(self, nbytes: int):
        """Pre-size the caching allocator's per-stream pools.  A scene needs ~130 MB of transient buffers at 150k points
        (~0.5 GB at 500k) out of 180 GB of HBM, but the allocator only learns that one cudaMalloc at a time, and a
        cudaMalloc costs 20-130 ms on this platform (measured: every step that triggered one took 1.5-3x as long).  One
        large block per lane stream (split on demand) and a stock of small-pool segments are allocated once and handed
        back to the allocator's cache, so steady-state steps never reach the driver."""
        if nbytes <= 0:
            return
        main = torch.cuda.current_stream(self.device)
        for st in self.streams + [main]:
            with torch.cuda.stream(st):
                big = torch.empty(nbytes if (st is not main or self.fused) else nbytes // 4, dtype=torch.uint8, device=self.device)
                small = [torch.empty(1 << 20, dtype=torch.uint8, device=self.device) for _ in range(64)]      # small-pool segments (2 MB each hold two)
                tiny = [torch.empty(256 << 10, dtype=torch.uint8, device=self.device) for _ in range(128)]
                del big, small, tiny
        torch.cuda.synchronize(self.device)

    def close(self):
        self._pool.shutdown(wait=True)

    # ------------------------------------------------------------------------------------------
    def _run(self, fn, items):
        """fn(item, index) is executed for every item on one of the executor's streams; results in item order.
        Item i always runs on stream i % n_streams (lane), in index order within the lane: the caching allocator keeps
        one pool per stream, so a fixed item -> stream map lets every step after the first reuse the previous step's
        blocks instead of calling cudaMalloc (a device-wide synchronisation) whenever a stream meets a new shape."""
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        n_lanes = min(self.n_streams, len(items))

        def lane(sid):
            torch.cuda.set_device(self.device)
            st = self.streams[sid]
            st.wait_event(ready)                           # parameters / inputs produced on the caller's stream
            outs = []
            with torch.cuda.stream(st):
                for i in range(sid, len(items), n_lanes):
                    out = fn(items[i], i)
                    done = torch.cuda.Event()
                    done.record(st)
                    outs.append((i, out, done))
            return outs

        futs = [self._pool.submit(lane, sid) for sid in range(n_lanes)]
        res = [None] * len(items)
        for f in futs:
            for i, out, done in f.result():
                res[i] = (out, done)
        outs = []
        for out, done in res:
            main.wait_event(done)                          # device-side join, no host synchronisation
            outs.append(out)
        return outs

    # ------------------------------------------------------------------------------------------
    def infer_batch(self, scenes, params, mode="ins_infer", upload=None, fused=None):
        """-> list of ForwardResult (fused: ONE ForwardResult for the whole batch, see train_batch).  `upload(scene)` (optional)
        turns a host-side scene into a SceneDevice on the worker's stream (end-to-end path: the H2D copies of one scene overlap
        the kernels of the others)."""
        if self.fused if fused is None else fused:
            batch = scenes if isinstance(scenes, pipeline.SceneDevice) else pipeline.SceneDevice.concat(
                [upload(s) if upload is not None else s for s in scenes])
            with torch.no_grad():
                return pipeline.forward_scene(batch, params, mode=mode)

        def fn(sc, i):
            if upload is not None:
                sc = upload(sc)
            with torch.no_grad():
                return pipeline.forward_scene(sc, params, mode=mode)
        return self._run(fn, scenes)

    def train_batch(self, scenes, params, train_keys, upload=None, fused=None, classifier=None):
        """Forward + backward of every scene; accumulates d(mean_i loss_i)/d(param) into `param.grad` (fixed scene
        order) and returns the mean loss (0-dim device tensor).  params: dict name -> tensor; train_keys: names
        of the leaves that require grad.
        fused (default): the scenes are concatenated into ONE block-diagonal batch (pipeline.SceneDevice.concat; `scenes` may
        already be such a batch) and driven by one forward / backward on the caller's stream — every graph / kNN / pooling /
        export launch serves all scenes, the order-dependent replays run one CTA per scene, BatchNorm statistics stay per
        scene.  fused=False: one scene per CUDA stream / host thread (the round-1 executor)."""
        leaves = [params[k] for k in train_keys]
        if self.fused if fused is None else fused:
            batch = scenes if isinstance(scenes, pipeline.SceneDevice) else pipeline.SceneDevice.concat(
                [upload(s) if upload is not None else s for s in scenes])
            r = pipeline.forward_scene(batch, params, mode="train", classifier=classifier)
            loss = pipeline.batch_loss(r.loss_raw)                 # mean over scenes of loss_sum / loss_num (train.py:165-170 + DDP)
            grads = torch.autograd.grad(loss, leaves, allow_unused=True)
            for pp, g in zip(leaves, grads):
                if g is None:
                    continue
                if pp.grad is None:
                    pp.grad = g
                else:
                    pp.grad.add_(g)
            self.last_result = r
            return loss.detach()
        # items may themselves be scene batches (lanes of several scenes each): the mean is over ALL scenes
        n = sum(getattr(sc, "n_scenes", 1) for sc in scenes)
        main = torch.cuda.current_stream(self.device)

        def fn(sc, i):
            if upload is not None:
                sc = upload(sc)
            r = pipeline.forward_scene(sc, params, mode="train", classifier=classifier)
            loss = pipeline.batch_loss(r.loss_raw, n)
            grads = torch.autograd.grad(loss, leaves, allow_unused=True)
            for g in grads:
                if g is not None:
                    g.record_stream(main)
            loss = loss.detach()
            loss.record_stream(main)
            return loss, grads

        outs = self._run(fn, scenes)
        total = torch.zeros((), device=self.device)
        for loss, grads in outs:
            total += loss
            for p, g in zip(leaves, grads):
                if g is None:
                    continue
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.add_(g)
        return total


# ------------------------------------------------------------------------------------------------
# multi-GPU plumbing: scenes are the shard unit, one process per GPU
# ------------------------------------------------------------------------------------------------
def shard_scenes(n_scenes: int, rank: int, world: int, pad: bool = False):
    """Scene indices of `rank`: round-robin `rank::world`, by default WITHOUT the padding duplicates of the reference's
    DistributedSampler (train.py:102 pads 1,201 scenes to 1,208 at 8 ranks, so 7 scenes are counted twice).
    HAZARD: without padding the ranks get unequal counts (1,201 scenes on 8 ranks: 151 on rank 0, 150 elsewhere).  That is
    what inference wants (no collective until the final reduce), but a loop that issues a collective PER STEP (training with
    allreduce_flat) would leave the rank with the extra scene blocked in NCCL forever: use pad=True there — it repeats scenes
    from the head of the list exactly like DistributedSampler, so every rank takes ceil(n / world) steps."""
    if not pad:
        return list(range(rank, n_scenes, world))
    per = -(-n_scenes // world)
    idx = list(range(n_scenes)) + list(range(per * world - n_scenes))
    return idx[rank:per * world:world]


def allreduce_flat(tensors, dist_module=None, average=True, extra=None):
    """ONE all-reduce for a list of tensors (the 147,880 gradients of SegModel = 0.59 MB: latency bound, so bucketing or
    overlap would only add launches).  `extra`: optional 1-D tensor (e.g. the 165 logging floats of train.py:172-175)
    appended to the same buffer and returned summed.  In place; works for NCCL (CUDA) and gloo (CPU) process groups."""
    import torch.distributed as dist
    d = dist_module or dist
    if not d.is_available() or not d.is_initialized() or d.get_world_size() == 1:
        return extra.reshape(-1).to(tensors[0].dtype).clone() if extra is not None else None      # same dtype / fresh tensor as the multi-rank path
    flats = [t.reshape(-1) for t in tensors] + ([extra.reshape(-1).to(tensors[0].dtype)] if extra is not None else [])
    flat = torch.cat(flats)
    d.all_reduce(flat)
    n_extra = extra.numel() if extra is not None else 0
    body = flat[:flat.numel() - n_extra]
    if average:
        body /= d.get_world_size()
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(body[o:o + n].view_as(t))
        o += n
    return flat[flat.numel() - n_extra:].clone() if extra is not None else None
