#!/usr/bin/env python
"""Experiment: does the loss kernel reading/writing PINNED HOST memory directly (UVA zero-copy over PCIe) beat the
staged host entry (H2D copies -> kernel -> D2H copy, pipelined in slices)?   python scripts/zero_copy_probe.py [workload]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from svbrdf_estimation_b200 import _cabi, environment as E

w = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, size, N, nr, ns, _ = bench.WORKLOADS[w]
lib = _cabi.lib()
dev = torch.device("cuda:0")
a, b = bench.synthetic_maps(B, size, 1), bench.synthetic_maps(B, size, 2)
h_in, h_tg = a.pin_memory(), b.pin_memory()
h_gr = torch.empty_like(a).pin_memory()
d_in, d_tg, d_gr = h_in.to(dev), h_tg.to(dev), torch.empty_like(a, device=dev)
rec = E.sample_loss_configs(B, nr, ns)
lin = torch.linspace(-1, 1, size, device=dev)
loss = torch.zeros(4, device=dev)
ws_bytes = lib.svbrdf_b200_workspace_bytes(B, N, size, size)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream


def run(pi, pt, pg):
    _cabi.check(lib.svbrdf_b200_loss_forward_backward(pi.data_ptr(), pt.data_ptr(), B, size, size, rec.data_ptr(), N,
                                                      lin.data_ptr(), loss.data_ptr(), pg.data_ptr(), ws.data_ptr(), ws_bytes, st))


cases = {"all device (kernel only)": (d_in, d_tg, d_gr), "zero-copy in+target+grad": (h_in, h_tg, h_gr),
         "zero-copy in+target, grad on device": (h_in, h_tg, d_gr), "zero-copy grad only": (d_in, d_tg, h_gr),
         "zero-copy input only": (h_in, d_tg, d_gr)}
evals = B * size * size * N
for name, (pi, pt, pg) in cases.items():
    for _ in range(2):
        run(pi, pt, pg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        run(pi, pt, pg)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print("%-40s %8.3f ms/step  %7.2f G evals/s  loss %.6f" % (name, dt * 1e3, evals / dt / 1e9, float(loss[0])), flush=True)
torch.cuda.synchronize()
ref = d_gr.clone(); run(d_in, d_tg, ref); run(h_in, h_tg, h_gr); torch.cuda.synchronize()
print("zero-copy gradient equals device gradient:", bool(torch.equal(ref.cpu(), h_gr)))
