#!/usr/bin/env python
"""BASELINE.json configs[4] in context: a single-view U-Net training step with the fused MixedLoss, the
global batch sharded over the GPUs of one box, CNN gradients all-reduced by DistributedDataParallel (NCCL).

    python examples/train_step_c5.py [--batch-per-gpu 32] [--size 256] [--steps 10]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        examples/train_step_c5.py --batch-per-gpu 32

The network is a stand-in: a pix2pix-style 8-down / 8-up U-Net with instance norm that emits the reference's
9-channel encoding (the reference's ``models.py`` is out of scope for this repository and is used unchanged by
its own training script).  What this script shows is the hot path in its real position:
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
import svbrdf_estimation_b200 as S                                 # noqa: E402
from svbrdf_estimation_b200.environment import NativeSceneSampler  # noqa: E402


class Down(nn.Module):
    def __init__(self, cin, cout, norm=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 4, 2, 1)
        self.norm = nn.InstanceNorm2d(cout, affine=True) if norm else nn.Identity()

    def forward(self, x):
        return self.norm(self.conv(nn.functional.leaky_relu(x, 0.2)))


class Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.ConvTranspose2d(cin, cout, 4, 2, 1)
        self.norm = nn.InstanceNorm2d(cout, affine=True)

    def forward(self, x):
        return self.norm(self.conv(nn.functional.relu(x)))


class UNetStandIn(nn.Module):
    """3 -> 9 channels, 8 stride-2 encoders and 8 decoders with skip connections (for 256x256 inputs)."""

    def __init__(self, ngf=64, depth=8):
        super().__init__()
        ch = [min(ngf * 2 ** i, ngf * 8) for i in range(depth)]
        self.first = nn.Conv2d(3, ch[0], 4, 2, 1)
        self.downs = nn.ModuleList([Down(ch[i], ch[i + 1], norm=(i + 1 < depth - 1)) for i in range(depth - 1)])
        ups = []
        for i in range(depth - 1, 0, -1):
            ups.append(Up(ch[i] * (1 if i == depth - 1 else 2), ch[i - 1]))
        self.ups = nn.ModuleList(ups)
        self.last = nn.ConvTranspose2d(ch[0] * 2, 9, 4, 2, 1)

    def forward(self, x):
        feats = [self.first(x)]
        for d in self.downs:
            feats.append(d(feats[-1]))
        y = feats[-1]
        for i, u in enumerate(self.ups):
            y = u(y)
            y = torch.cat((y, feats[-2 - i]), dim=1)
        return torch.tanh(self.last(nn.functional.relu(y)))       # the 9-channel encoding in [-1,1] (models.py:336-338)


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
    model = UNetStandIn(depth=8 if args.size >= 256 else max(3, args.size.bit_length() - 1)).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=1e-5)                                    # main.py:74
    B = args.batch_per_gpu
    loss_fn = S.MixedLoss(S.LocalRenderer(), l1_weight=0.1,
                          scene_sampler=NativeSceneSampler(seed=313, first_batch_element=rank * B))
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
            losses.append(float(loss.detach()))
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
