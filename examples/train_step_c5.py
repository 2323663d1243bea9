#!/usr/bin/env python
"""BASELINE.json configs[4] in context: a single-view U-Net training step with the fused MixedLoss, the
global batch sharded over the GPUs of one box, CNN gradients all-reduced by DistributedDataParallel (NCCL).

    python examples/train_step_c5.py [--batch-per-gpu 32] [--size 256] [--steps 10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        examples/train_step_c5.py --batch-per-gpu 32

The network is a stand-in of the reference generator's shape and size (examples/unet_standin.py, 79.99 M parameters;
the reference's ``models.py`` is out of scope for this repository and is used unchanged by its own training script;
see INTEGRATION.md).  ``bench.py`` runs the same step as its ``train_step_c5`` leg with the GLOBAL batch fixed at 256.  What this script shows is the hot path in its real position:
``tanh(generator(x))`` goes straight into ``MixedLoss.forward_encoded`` (decode + map-L1 + rendering loss + their
gradient in one kernel), each rank samples the scenes of its own batch slice (``NativeSceneSampler`` keyed by the
global sample index), and the only traffic over NVLink is DDP's gradient all-reduce.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import svbrdf_estimation_b200 as S                                 # noqa: E402
from svbrdf_estimation_b200 import sharding                        # noqa: E402
from svbrdf_estimation_b200.environment import NativeSceneSampler  # noqa: E402
from unet_standin import UNetStandIn                               # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch-per-gpu", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(313)
    model = UNetStandIn().to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True) if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=1e-5)                                    # main.py:74
    B = args.batch_per_gpu
    lo, hi = sharding.shard_range(B * world, rank, world)            # this rank's slice of the global batch
    loss_fn = S.MixedLoss(S.LocalRenderer(), l1_weight=0.1,
                          scene_sampler=NativeSceneSampler(seed=313, first_batch_element=lo))
    g = torch.Generator("cpu").manual_seed(1000 + rank)
    images = torch.rand(B, 3, args.size, args.size, generator=g).to(dev)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    target = bench.synthetic_maps(B, args.size, 2000 + rank).to(dev)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    loss_ms, step_ms, losses = [], [], []
    for it in range(args.warmup + args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev[0].record()
        opt.zero_grad(set_to_none=True)
        enc = net(images)
        ev[1].record()
        loss = loss_fn.forward_encoded(enc, target)
        ev[2].record()
        loss.backward()
        opt.step()
        ev[3].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            step_ms.append((time.perf_counter() - t0) * 1e3)
            loss_ms.append(ev[1].elapsed_time(ev[2]))
            losses.append(float(sharding.global_mean_loss(loss.detach(), B)))     # NCCL all-reduce of the scalar, for logging
    t = torch.tensor([sum(step_ms) / len(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"example": "configs[4] U-Net step with fused MixedLoss", "n_gpus": world, "batch_per_gpu": B,
                          "global_batch": B * world, "size": args.size, "unet_stand_in_params_M": round(n_params / 1e6, 1),
                          "step_ms": float(t.item()), "fused_loss_fwd_bwd_ms": sum(loss_ms) / len(loss_ms),
                          "samples_per_s": B * world / (float(t.item()) * 1e-3), "loss_first": losses[0], "loss_last": losses[-1]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
