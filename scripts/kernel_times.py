#!/usr/bin/env python
"""Device time of every C-ABI entry point on one workload (CUDA events, inputs resident in HBM).
    python scripts/kernel_times.py [c2|c3|c4|c5]  ->  JSON lines"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                             # noqa: E402
from svbrdf_estimation_b200 import _cabi                 # noqa: E402
from svbrdf_estimation_b200 import environment as E     # noqa: E402


def main():
    w = sys.argv[1] if len(sys.argv) > 1 else "c2"
    B, size, N, nr, ns, _ = bench.WORKLOADS[w]
    if os.environ.get("SVB_VARIANT_LIB"):                 # time a variant build (scripts/variant_bench.py build ...)
        _cabi.LIB_PATH = os.environ["SVB_VARIANT_LIB"]
    lib = _cabi.lib()
    dev = torch.device("cuda", 0)
    inp, tgt = bench.synthetic_maps(B, size, 1).to(dev), bench.synthetic_maps(B, size, 2).to(dev)
    enc = (torch.rand(B, 9, size, size, device=dev) * 2 - 1)
    rec = E.sample_loss_configs(B, nr, ns)
    lin = torch.linspace(-1, 1, size, device=dev)
    nb = lib.svbrdf_b200_workspace_bytes(B, N, size, size)
    ws = torch.empty(nb // 4 + 1, device=dev)
    out = torch.zeros(3, device=dev)
    grad, genc = torch.empty_like(inp), torch.empty_like(enc)
    images = torch.empty(B, N, 3, size, size, device=dev)
    gimg = torch.randn_like(images)
    st = torch.cuda.current_stream().cuda_stream
    P, HW = B * size * size, size * size
    calls = {
        "loss_forward": (lambda: lib.svbrdf_b200_loss_forward(inp.data_ptr(), tgt.data_ptr(), B, size, size, rec.data_ptr(), N, lin.data_ptr(), out.data_ptr(), ws.data_ptr(), nb, st), 96 * P),
        "loss_forward_backward": (lambda: lib.svbrdf_b200_loss_forward_backward(inp.data_ptr(), tgt.data_ptr(), B, size, size, rec.data_ptr(), N, lin.data_ptr(), out.data_ptr(), grad.data_ptr(), ws.data_ptr(), nb, st), 144 * P),
        "loss_forward_backward_accurate": (lambda: lib.svbrdf_b200_loss_forward_backward_accurate(inp.data_ptr(), tgt.data_ptr(), B, size, size, rec.data_ptr(), N, lin.data_ptr(), out.data_ptr(), grad.data_ptr(), ws.data_ptr(), nb, st), 144 * P),
        "mixed_loss_forward_backward": (lambda: lib.svbrdf_b200_mixed_loss_forward_backward(inp.data_ptr(), tgt.data_ptr(), B, size, size, rec.data_ptr(), N, 0.1, lin.data_ptr(), out.data_ptr(), grad.data_ptr(), ws.data_ptr(), nb, st), 144 * P),
        "mixed_loss_encoded_forward_backward": (lambda: lib.svbrdf_b200_mixed_loss_encoded_forward_backward(enc.data_ptr(), tgt.data_ptr(), B, size, size, rec.data_ptr(), N, 0.1, lin.data_ptr(), out.data_ptr(), genc.data_ptr(), ws.data_ptr(), nb, st), 120 * P),
        "render_forward": (lambda: lib.svbrdf_b200_render_forward(inp.data_ptr(), B, size, size, rec.data_ptr(), N, 1, lin.data_ptr(), images.data_ptr(), st), (48 + 12 * N) * P),
        "render_backward": (lambda: lib.svbrdf_b200_render_backward(inp.data_ptr(), B, size, size, rec.data_ptr(), N, 1, lin.data_ptr(), gimg.data_ptr(), grad.data_ptr(), st), (96 + 12 * N) * P),
    }
    for name, (fn, nbytes) in calls.items():
        for _ in range(5):
            assert fn() == 0, lib.svbrdf_b200_last_error()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(json.dumps({"workload": w, "entry": name, "ms": round(ms, 4), "G_evals_s": round(P * N / ms / 1e6, 1),
                          "algorithmic_GB_s": round(nbytes / ms / 1e6, 1), "hbm_frac": round(nbytes / ms / 1e6 / 6551.0, 3)}))


if __name__ == "__main__":
    main()
