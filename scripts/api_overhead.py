#!/usr/bin/env python
"""How much does the Python layer add on top of the kernels?  Times (a) the raw C-ABI step, (b)
RenderingLoss(LocalRenderer()) forward+backward with the reference-order scene sampler, (c) the same
with the native sampler, on BASELINE.json configs[1]; wall clock with a device sync per step."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                             # noqa: E402
import svbrdf_estimation_b200 as S                       # noqa: E402
from svbrdf_estimation_b200 import environment as E     # noqa: E402


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    B, size, N = 64, 256, 9
    inp = bench.synthetic_maps(B, size, 1).cuda().requires_grad_(True)
    tgt = bench.synthetic_maps(B, size, 2).cuda()
    rec = E.sample_loss_configs(B)

    def fixed():
        inp.grad = None
        S.rendering_loss_with_records(inp, tgt, rec).backward()

    ref_order = S.RenderingLoss(S.LocalRenderer())
    native = S.RenderingLoss(S.LocalRenderer(), scene_sampler=E.NativeSceneSampler(1))

    def api(mod):
        def run():
            inp.grad = None
            mod(inp, tgt).backward()
        return run
    print("python API, fixed records        : %.3f ms/step" % timeit(fixed))
    print("python API, reference-order scenes: %.3f ms/step" % timeit(api(ref_order)))
    print("python API, native sampler        : %.3f ms/step" % timeit(api(native)))
    print("sampler alone (reference order)   : %.3f ms" % (timeit(lambda: E.sample_loss_configs(B), 20)))


if __name__ == "__main__":
    main()
