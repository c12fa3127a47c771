"""Data-parallel drop-in proof (reference: seggroup/train.py:94 wraps the model in DistributedDataParallel(find_unused_parameters=True),
one scene per rank per step, :165-175): two ranks over NCCL, each with its own scene, must end the backward pass with the gradients
of the single-GPU two-scene batch (mean over scenes of loss_sum / loss_num).  Needs two GPUs (gpurun --gpus 2); skipped otherwise."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tree, ret):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    os.chdir(tree)
    from seggroup_b200 import engine, synth
    from seggroup_b200.model import SegModel
    scenes = [synth.make_scene(41 + i, 6000, name="ddp_%d" % i) for i in range(world)]
    torch.manual_seed(1)
    model = SegModel(exp_name="ddp%d" % rank).to(rank)
    with torch.no_grad():
        model.mlp_1.bn1.weight.mul_(4.0)
    model.classifier.dp1.p = 0.0
    model.epoch = "1"
    ddp = DistributedDataParallel(model, device_ids=[rank], find_unused_parameters=True)        # train.py:94
    s = scenes[rank]
    data = torch.from_numpy(s.data.copy()).unsqueeze(0)
    weak = torch.from_numpy(s.weak_label.copy()).unsqueeze(0)
    info = torch.tensor([[rank]])
    out = ddp(data.cuda(rank), weak.cuda(rank), info)                                           # train.py:161-163
    loss = torch.sum(out[0][:, 0]) / torch.sum(out[0][:, 1])                                    # train.py:165-167
    loss.backward()                                                                             # DDP averages the gradients
    # the 165 logging floats of train.py:172-175 ride in one all-reduce (engine.allreduce_flat with `extra`)
    logs = torch.cat([loss.detach().view(1), out[1].reshape(-1), out[2].reshape(-1), out[3].reshape(-1)])
    summed = engine.allreduce_flat([torch.zeros(1, device=rank)], dist, average=False, extra=logs)
    model.flush_exports()
    ok = True
    if rank == 0:
        # single-GPU reference: the same two scenes as ONE batch through the same module class
        torch.manual_seed(1)
        ref = SegModel(exp_name="ddp_ref").to(0)
        with torch.no_grad():
            ref.mlp_1.bn1.weight.mul_(4.0)
        ref.classifier.dp1.p = 0.0
        ref.epoch = "1"
        d2 = torch.stack([torch.from_numpy(x.data.copy()) for x in scenes]).cuda(0)
        w2 = torch.stack([torch.from_numpy(x.weak_label.copy()) for x in scenes]).cuda(0)
        o2 = ref(d2, w2, torch.arange(world).view(-1, 1))
        (o2[0][:, 0] / o2[0][:, 1]).mean().backward()
        ref.flush_exports()
        worst = 0.0
        for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            if q.grad is None:
                ok = ok and (p.grad is None or float(p.grad.abs().max()) == 0.0)
                continue
            err = float((p.grad - q.grad).abs().max() / (q.grad.abs().max() + 1e-30))
            worst = max(worst, err)
            ok = ok and err < 2e-5
        ok = ok and abs(float(summed[0]) - float((o2[0][:, 0] / o2[0][:, 1]).sum())) < 1e-4 * abs(float(summed[0]))
        ok = ok and torch.allclose(summed[1:81], o2[1].sum(0).reshape(-1)) and summed.numel() == 165
        ret["worst"] = worst
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_ddp_two_ranks_equal_single_gpu_two_scene_batch(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from seggroup_b200 import synth
    scenes = [synth.make_scene(41 + i, 6000, name="ddp_%d" % i) for i in range(2)]
    synth.write_scene_tree(str(tmp_path), scenes)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), ret), nprocs=2, join=True)
    assert ret[0] and ret[1], dict(ret)
