#!/usr/bin/env python
"""Data-parallel training steps under torchrun (one process per GPU, NCCL): every rank runs the patched reference's training step
on ITS OWN sample, the gradients are averaged with ONE flat all-reduce (parallel.all_reduce_gradients), every rank applies the
same Adam step.  Checks that the replicas stay bit-identical and that the averaged gradient equals the mean of the per-rank
gradients; prints the step time (max over ranks, CUDA events) and the time of the collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_ddp_step.py
"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cds_mvsnet_b200 as C  # noqa: E402
from cds_mvsnet_b200 import losses, parallel, synthetic  # noqa: E402
from oracle import ref_live  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
cfg = dict(W=640, H=512, N=3, ndepths=(48, 32, 8), ratios=(4.0, 1.5, 0.75), B=1, Dtot=192, interval=2.65)
z = np.load(os.path.join(ROOT, "tests", "golden", "weights_both_dtu_blended.npz"))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
s = synthetic.make_sample(cfg, "plane", seed=100 + rank)          # a different work item per rank
imgs, dv = s.imgs.to(dev), s.depth_values.to(dev)
proj = {k: v.to(dev) for k, v in s.proj_matrices.items()}
gt = s.gt_depth.to(dev)
gts = {"stage1": gt[:, ::4, ::4].contiguous(), "stage2": gt[:, ::2, ::2].contiguous(), "stage3": gt, "stage4": gt}
masks = {k: torch.ones_like(v) for k, v in gts.items()}
rmodel, rmodule, _, _ = ref_live.load()
C.patch(rmodel, rmodule, level="leaf")
model = ref_live.build_model(sd, cfg["ndepths"], cfg["ratios"], device=dev, rmodel=rmodel).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
interval = torch.tensor([cfg["interval"]], device=dev)
times, coll, hist = [], [], []
for it in range(4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    e0.record()
    opt.zero_grad(set_to_none=True)
    out = model(imgs, proj, dv, gt_depths=gts, temperature=0.01)
    total, _ = losses.final_loss(out, gts, masks, dlossw=[0.5, 1.0, 2.0], depth_interval=interval)
    total.backward()
    if it == 0 and world > 1:   # the averaged gradient is the mean of the ranks' gradients
        probe = model.cost_regularization[0].conv0.conv.weight.grad.detach().clone()
        bucket = [torch.empty_like(probe) for _ in range(world)]
        dist.all_gather(bucket, probe)
        want = torch.stack(bucket).mean(0)
    e1.record()
    n = parallel.all_reduce_gradients(model.parameters())
    e2.record()
    opt.step()
    e3.record()
    torch.cuda.synchronize()
    if it == 0 and world > 1:
        got = model.cost_regularization[0].conv0.conv.weight.grad
        assert (got - want).abs().max() <= 1e-6 * want.abs().max() + 1e-12, "averaged gradient != mean of the ranks' gradients"
    times.append(parallel.max_over_ranks(e0.elapsed_time(e3), dev))
    coll.append(parallel.max_over_ranks(e1.elapsed_time(e2), dev))
    hist.append(float(total.detach()))
# replicas stay identical: a checksum of all parameters agrees across the ranks
chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
if world > 1:
    all_chk = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(all_chk, chk)
    assert all(torch.equal(c, all_chk[0]) for c in all_chk), "replicas diverged"
if rank == 0:
    print(json.dumps({"n_gpus": world, "config": "640x512 N=3 D=48/32/8, one sample per rank, patch(level=leaf) + final_loss + flat gradient all-reduce + Adam",
                      "ms_per_step": times[1:], "ms_all_reduce": coll[1:], "gradient_elements": n, "loss_rank0": hist,
                      "samples_per_s": world * 1e3 / (sum(times[1:]) / len(times[1:]))}))
if world > 1:
    dist.destroy_process_group()
