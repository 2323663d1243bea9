#!/usr/bin/env python
"""e2e (host entry point) timing only: python scripts/e2e_probe.py [workload]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from svbrdf_estimation_b200 import _cabi, environment as E
w = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, size, N, nr, ns, _ = bench.WORKLOADS[w]
lib = _cabi.lib()
torch.cuda.init()
ctx = ctypes.c_void_p()
_cabi.check(lib.svbrdf_b200_ctx_create(ctypes.byref(ctx), B, N, size, size))
n = B * 12 * size * size
pin = [lib.svbrdf_b200_ctx_pinned(ctx, k) for k in range(3)]
a, b = bench.synthetic_maps(B, size, 1), bench.synthetic_maps(B, size, 2)
ctypes.memmove(pin[0], a.data_ptr(), n * 4); ctypes.memmove(pin[1], b.data_ptr(), n * 4)
rec = E.sample_loss_configs(B, nr, ns)
loss = ctypes.c_float()
for grad in (True, False):
    for _ in range(3):
        _cabi.check(lib.svbrdf_b200_rendering_loss_host(ctx, pin[0], pin[1], B, rec.data_ptr(), N, ctypes.byref(loss), pin[2] if grad else None))
    t0 = time.perf_counter()
    for _ in range(20):
        _cabi.check(lib.svbrdf_b200_rendering_loss_host(ctx, pin[0], pin[1], B, rec.data_ptr(), N, ctypes.byref(loss), pin[2] if grad else None))
    dt = (time.perf_counter() - t0) / 20
    print("%s grad=%s: %.3f ms/step, H2D %.1f GB/s, %.2f G evals/s, loss %.6f" % (w, grad, dt * 1e3, 2 * n * 4 / dt / 1e9, B * size * size * N / dt / 1e9, loss.value))
lib.svbrdf_b200_ctx_destroy(ctx)
