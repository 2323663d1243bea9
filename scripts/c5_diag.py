#!/usr/bin/env python
"""Where does the configs[4] step spend its time at per-GPU batch 32 vs 256?  (1 GPU, no DDP): step time, summed kernel
time (torch profiler), top kernels; variants: channels_last, bucket sizes are N>1 only."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import bench                                             # noqa: E402
import svbrdf_estimation_b200 as S                       # noqa: E402
from svbrdf_estimation_b200.environment import NativeSceneSampler  # noqa: E402
from unet_standin import UNetStandIn                     # noqa: E402


def run(B, channels_last=False, graph=False, prof=False):
    dev = torch.device("cuda", 0)
    torch.manual_seed(313)
    torch.backends.cudnn.benchmark = True
    model = UNetStandIn().to(dev)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), lr=1e-5, fused=True)
    net = model
    loss_fn = S.MixedLoss(S.LocalRenderer(), scene_sampler=NativeSceneSampler(313))
    images = torch.rand(B, 3, 256, 256, device=dev)
    if channels_last:
        images = images.contiguous(memory_format=torch.channels_last)
    target = bench.synthetic_maps(min(B, 32), 256, 5000).repeat((B + 31) // 32, 1, 1, 1)[:B].to(dev).contiguous()

    if graph:
        # forward and backward of the network as two CUDA graphs (the loss stays outside: fresh scenes per step)
        net = torch.cuda.make_graphed_callables(model, (images,), num_warmup_iters=3)

    def step():
        opt.zero_grad(set_to_none=not graph)
        enc = net(images)
        loss = loss_fn.forward_encoded(enc.contiguous(), target)
        loss.backward()
        opt.step()
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 8
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    out = "B=%d channels_last=%s graph=%s: %.2f ms/step (%.3f ms/sample)" % (B, channels_last, graph, ms, ms / B)
    if prof:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
        ev = p.key_averages()
        tot = sum(e.device_time_total for e in ev) / 2 / 1e3
        out += "; summed device kernel time %.2f ms/step" % tot
        top = sorted(ev, key=lambda e: -e.device_time_total)[:12]
        out += "\n" + "\n".join("    %-90s %8.2f ms  x%d" % (e.key[:90], e.device_time_total / 2 / 1e3, e.count // 2) for e in top)
    print(out, flush=True)
    del model, opt
    torch.cuda.empty_cache()


if __name__ == "__main__":
    if "--graph" in sys.argv:
        run(32)
        run(32, graph=True)
        run(64)
        run(64, graph=True)
        run(256)
        run(256, graph=True)
    else:
        run(32, prof=True)
        run(256, prof=True)
        run(32, channels_last=True)
