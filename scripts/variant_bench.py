#!/usr/bin/env python
"""Development harness: build several variants of the library (different -D flags) HERE, then time
the fused loss fwd+bwd kernel of each on the GPU box.

  build:  python scripts/variant_bench.py build name1="-DSVB_LOSS_MINB=4" name2="..."
  run  :  python scripts/variant_bench.py run [workload]        (on the GPU box; prints a table)
"""
import ctypes
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "variants")


def build(specs):
    from concurrent.futures import ThreadPoolExecutor
    from svbrdf_estimation_b200 import _build
    os.makedirs(VDIR, exist_ok=True)
    for f in glob.glob(os.path.join(VDIR, "*.so")):
        os.remove(f)

    def one(spec):
        name, _, flags = spec.partition("=")
        try:
            _build.compile_library(os.path.join(VDIR, name + ".so"), extra_flags=flags.split())
            return name, "OK"
        except Exception as exc:
            return name, "FAILED %s" % exc
    with ThreadPoolExecutor(max_workers=3) as pool:
        for name, status in pool.map(one, specs):
            print(name, status)


def run(workload="c2", steps=100, entry="svbrdf_b200_loss_forward_backward"):
    import torch
    import bench
    from svbrdf_estimation_b200 import _cabi
    from svbrdf_estimation_b200 import environment as E
    B, size, N, nr, ns, _ = bench.WORKLOADS[workload]
    dev = torch.device("cuda", 0)
    inp = bench.synthetic_maps(B, size, 1001).to(dev)
    tgt = bench.synthetic_maps(B, size, 2001).to(dev)
    sets = [(inp, tgt, torch.empty_like(inp)), (inp.roll(1, 0).contiguous(), tgt.roll(1, 0).contiguous(), torch.empty_like(inp))]
    torch.manual_seed(313)
    rec = E.sample_loss_configs(B, nr, ns)
    lin = torch.linspace(-1, 1, size, device=dev)
    loss = torch.zeros(1, device=dev)
    evals = B * size * size * N
    rows = []
    for path in sorted(glob.glob(os.path.join(VDIR, "*.so"))):
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in _cabi.PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        wsb = lib.svbrdf_b200_workspace_bytes(B, N, size, size)
        ws = torch.empty(wsb // 4 + 1, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def step(i):
            a, b, g = sets[i % 2]
            rc = getattr(lib, entry)(a.data_ptr(), b.data_ptr(), B, size, size, rec.data_ptr(), N,
                                     lin.data_ptr(), loss.data_ptr(), g.data_ptr(), ws.data_ptr(), wsb, st)
            assert rc == 0, lib.svbrdf_b200_last_error()
        for i in range(5):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        rows.append({"variant": os.path.basename(path)[:-3], "workload": workload, "entry": entry.replace("svbrdf_b200_", ""), "ms": ms, "G_evals_s": evals / ms / 1e6,
                     "loss": float(loss.item()), "gsum": float(sets[(steps - 1) % 2][2].double().abs().sum().item())})
        print(json.dumps(rows[-1]))
    return rows


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        for w in (sys.argv[2:] or ["c2"]):
            run(w)
            if w == "c2":
                run(w, entry="svbrdf_b200_loss_forward_backward_accurate")
