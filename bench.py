#!/usr/bin/env python
"""Benchmark of the rendering-loss hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic SVBRDF maps: RenderingLoss
forward + backward (loss value and d loss / d input) for ``B`` maps of ``H x W`` under ``N``
light/view configurations per map = ``B*H*W*N`` pixel-light evaluations.

  value     whole-job G pixel-light evals/s with maps resident in HBM (C-ABI device entry point,
            CUDA events on the launch stream, max over ranks); buffer sets are rotated and each
            set (input+target+grad) is several times larger than L2.
  e2e       same metric through the C-ABI host entry point: pinned HOST maps in, loss + gradient
            back on the host, H2D/D2H copies inside the timed region.
  roofline  the fused kernel against the FP32 bound (330 FLOP per evaluation) and,
            as roofline_hbm, against the HBM bound (144 B per pixel).
  cpu_baseline  the oracle port of the reference (eager PyTorch CPU, all host threads) on a bounded
            sample of the same workload, rank 0 / N=1 only.

``--impl reference`` times that CPU port alone (the reference is pure Python and cannot travel to
the GPU box) and prints the same JSON line with ``"impl": "reference"``.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (per-GPU batch, size, N, n_random, n_specular, BASELINE.json config it is)
    "c1": (8, 256, 9, 3, 6, "configs[0]: single-view RenderingLoss, batch 8 of 256x256 SVBRDF maps, 9 light/view configs (the reference's CPU-runnable case)"),
    "c2": (64, 256, 9, 3, 6, "configs[1]: single-view RenderingLoss fwd+bwd, batch 64 of 256x256 maps, 9 configs"),
    "c3": (32, 256, 27, 9, 18, "configs[2]: batch 32 of 256x256, 27 light/view configs"),
    "c4": (16, 1024, 9, 3, 6, "configs[3]: high-res 1024x1024 maps, batch 16, 9 configs"),
    "c5": (32, 256, 9, 3, 6, "configs[4] loss part: global batch 256 over 8 GPUs = 32 per GPU, 256x256, 9 configs"),
}
FLOP_PER_EVAL = 330.0        # SURVEY.md §8d canonical fwd+bwd count (FMA = 2)
# Register-file operand-bandwidth model of the hot loop (scripts/sass_stats.py on the shipped kernel: cycles one
# scheduler needs to fetch the operands of one record iteration of one warp = 64 pixels)
RF_CYCLES_PER_WARP_RECORD = 440.0
BYTES_PER_PIXEL = 144.0      # read input 48 + read target 48 + write grad 48
METRIC = "rendering-loss fwd+bwd pixel-light evals/s"
UNIT = "G evals/s"


def synthetic_maps(batch, size, seed):
    """unit upper-hemisphere normals, diffuse/specular U(0,1), roughness U(0.1,1) x3."""
    import torch
    g = torch.Generator("cpu").manual_seed(seed)
    xy = torch.randn(batch, 2, size, size, generator=g) * 0.3
    n = torch.cat((xy, torch.ones(batch, 1, size, size)), dim=1)
    n = n / n.norm(dim=1, keepdim=True)
    d = torch.rand(batch, 3, size, size, generator=g)
    s = torch.rand(batch, 3, size, size, generator=g)
    r = (torch.rand(batch, 1, size, size, generator=g) * 0.9 + 0.1).repeat(1, 3, 1, 1)
    return torch.cat((n, d, r, s), dim=1).contiguous()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ---------------------------------------------------------------------------------------------
# clocks: NVML polling thread (nvidia_ml_py), nvidia-smi fallback
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.power = index, [], set(), []
        self.max_mhz, self._stop, self._t, self._nv = None, threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _poll(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self._nv is not None:
            self._stop.clear()
            self._t = threading.Thread(target=self._poll, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self, note):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "how": "unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None, "how": note}


def bind_host_thread_near_gpu(index):
    """Pins the calling thread to the CPUs NVML reports as local to GPU `index` (same NUMA node / PCIe root), so that
    the pinned host buffers allocated next are placed in that node's memory.  With 8 ranks on one box this keeps
    every rank's 604 MB/step of PCIe traffic off the inter-socket link.  Returns a short note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        return "thread bound to %d of %d CPUs local to the GPU (%d-%d)" % (len(after), before, after[0], after[-1])
    except Exception as exc:
        return "not bound (%s)" % (repr(exc)[:80],)


# ---------------------------------------------------------------------------------------------
# CPU port of the reference (oracle) - cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------
def cpu_port_step(inp, tgt, cfg):
    from oracle import reference_port as O   # the checker, used here only as the CPU baseline being timed
    return O.rendering_loss_and_grad(inp, tgt, cfg)


def cpu_reference_impl():
    """-> (step function, kind): the unmodified reference (its bytecode under oracle/_ref, built by oracle/build_ref.py in
    the build container) when it is there, else the restatement in oracle/reference_port.py."""
    try:
        from oracle import ref_loader
        if ref_loader.available() and ref_loader.load() is not None:
            return ref_loader.rendering_loss_and_grad, "reference"
    except Exception:
        pass
    return cpu_port_step, "port"


def time_cpu_port(workload, steps, warmup, budget_s, full_batch=False):
    """Times the reference's CPU implementation of the path (all host threads) on the workload - the full batch when
    (steps + warmup) passes fit ``budget_s``, else the largest leading part of the batch that does.
    Returns (G evals/s, seconds per step, sample batch, threads, kind)."""
    import torch
    from svbrdf_estimation_b200 import environment as E
    B, size, N, nr, ns, _ = WORKLOADS[workload]
    step_fn, kind = cpu_reference_impl()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(313)
    probe_b = 1
    inp, tgt = synthetic_maps(probe_b, size, 1001), synthetic_maps(probe_b, size, 2001)
    cfg = E.sample_loss_configs(probe_b, nr, ns)
    step_fn(inp, tgt, cfg)
    t0 = time.perf_counter()
    step_fn(inp, tgt, cfg)
    t1 = time.perf_counter() - t0
    b = B if full_batch else (8 if size <= 256 else 1)
    while b > 1 and (steps + warmup) * t1 * b * 0.8 > budget_s:      # a batch amortises per-call overhead: ~0.8 t1 per element
        b //= 2
    b = max(1, min(b, B))
    inp, tgt = synthetic_maps(b, size, 1001), synthetic_maps(b, size, 2001)
    cfg = E.sample_loss_configs(b, nr, ns)
    for _ in range(warmup):
        step_fn(inp, tgt, cfg)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step_fn(inp, tgt, cfg)
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    evals = b * size * size * N
    return evals / sec / 1e9, sec, b, torch.get_num_threads(), kind


def time_gpu_eager_port(workload, dev, batch=2, steps=3):
    """The same eager port on the GPU (what the reference does when main.py runs with --gpu-id): ~100 small
    kernels per render call, 2*N render calls per sample -> launch-bound.  Bounded sample of `batch` maps."""
    import torch
    from svbrdf_estimation_b200 import environment as E
    B, size, N, nr, ns, _ = WORKLOADS[workload]
    b = min(batch, B)
    inp, tgt = synthetic_maps(b, size, 1001).to(dev), synthetic_maps(b, size, 2001).to(dev)
    torch.manual_seed(313)
    cfg = E.sample_loss_configs(b, nr, ns)
    cpu_port_step(inp, tgt, cfg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port_step(inp, tgt, cfg)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    return b * size * size * N / sec / 1e9, sec, b


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B, size, N, nr, ns, desc = WORKLOADS[args.workload]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    val, sec, b, threads, kind = time_cpu_port(args.workload, steps, warmup, budget_s=240.0, full_batch=True)
    sample = ("the whole batch of the workload per step" if b == B else "first %d of %d maps of the workload per step" % (b, B)) \
        + " (%dx%d, N=%d), fwd+bwd via autograd" % (size, size, N)
    what = ("the UNMODIFIED reference (development/multiImage_pytorch/{losses,renderers,environment,utils}.py compiled to bytecode "
            "under oracle/_ref by oracle/build_ref.py): RenderingLoss(LocalRenderer()) forward + backward, eager PyTorch on the host cores"
            if kind == "reference" else
            "the reference's eager-PyTorch CPU path restated in oracle/reference_port.py (bit-exact to the reference in fp32, "
            "tests/test_oracle_golden.py); oracle/_ref was not found on this machine")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "batch_per_step": b, "batch_per_gpu": B, "size": size, "N": N,
                       "same_config_as_gpu_arm": b == B},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "note": what}
    args.out.emit(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def kernel_ms_for_roofline(ms_per_step, per_step, world):
    """step = fused kernel + 1-CTA finalize (~2 %): the timed-region mean at N=1, the per-step median under torchrun"""
    return ms_per_step if world == 1 else statistics.median(per_step)


def probe_fp32_peaks(lib, torch, stream):
    """Measured register-resident FP32 rates in TFLOP/s (FMA = 2): scalar FFMA, packed FFMA2, FMUL+FADD mix.
    The probe kernels live in a bench-only object (scripts/fp32_probe.cu -> scripts/libfp32_probe.so), not in the
    product library."""
    path = os.path.join(ROOT, "scripts", "libfp32_probe.so")
    if not os.path.exists(path):
        return {}
    probe = ctypes.CDLL(path)
    probe.fp32_probe_launch.restype = ctypes.c_int
    probe.fp32_probe_launch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    blocks, iters = sms * 16, 4000
    sink = torch.zeros(blocks * 256, device="cuda")
    out = {}
    for kind, name, flop_per_op in ((0, "ffma", 2.0), (1, "ffma2", 2.0), (3, "fmul_fadd", 1.0)):
        ops = ctypes.c_int(0)
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = probe.fp32_probe_launch(kind, blocks, iters, sink.data_ptr(), ctypes.byref(ops), stream)
            if rc != 0:
                return out
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            rate = blocks * 256.0 * iters * ops.value * flop_per_op / (ms * 1e-3) / 1e12
            out[name] = max(out.get(name, 0.0), rate)
    return out


def train_step_c5(args, torch, dist, dev, world, rank, local_rank):
    """BASELINE.json configs[4]: the full single-view training step (main.py:104-118) with the fused MixedLoss, GLOBAL
    batch 256 sharded over the ranks (strong scaling: 256 / world per GPU), the CNN gradients all-reduced by
    DistributedDataParallel over NCCL inside the timed region, Adam(lr=1e-5) (main.py:74) included.  The network is
    examples/unet_standin.py: the reference's generator layer for layer in size and shape (79.99 M parameters -> 320 MB of
    fp32 gradients per step).  Returns the sub-record for the JSON line (rank 0) - timings are the max over ranks."""
    import torch.nn as nn
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from unet_standin import UNetStandIn
    import svbrdf_estimation_b200 as S
    from svbrdf_estimation_b200 import sharding
    from svbrdf_estimation_b200.environment import NativeSceneSampler
    GLOBAL_BATCH, size = args.c5_global_batch, 256
    lo, hi = sharding.shard_range(GLOBAL_BATCH, rank, world)
    B = hi - lo
    torch.manual_seed(313)                                        # same initial weights on every rank
    torch.backends.cudnn.benchmark = True
    model = UNetStandIn().to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], bucket_cap_mb=25, gradient_as_bucket_view=True) if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=1e-5, fused=True)
    # every rank samples the scenes of ITS batch elements: keyed by the global sample index, so the scenes of the job
    # do not depend on how the batch is sharded
    loss_fn = S.MixedLoss(S.LocalRenderer(), l1_weight=0.1, scene_sampler=NativeSceneSampler(seed=313, first_batch_element=lo))
    g = torch.Generator("cpu").manual_seed(4000)
    images = torch.rand(GLOBAL_BATCH, 3, 64, 64, generator=g)[lo:hi]
    images = torch.nn.functional.interpolate(images, size=(size, size), mode="nearest").to(dev).contiguous()
    target = synthetic_maps(GLOBAL_BATCH if GLOBAL_BATCH <= 32 else 32, size, 5000)
    target = target.repeat((GLOBAL_BATCH + target.shape[0] - 1) // target.shape[0], 1, 1, 1)[lo:hi].to(dev).contiguous()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(sync=True):
        import contextlib
        ctx = contextlib.nullcontext() if (sync or world == 1) else net.no_sync()
        with ctx:
            opt.zero_grad(set_to_none=True)
            enc = net(images)
            ev[1].record()
            loss = loss_fn.forward_encoded(enc, target)           # decode + map-L1 + rendering loss + their gradient: ONE kernel
            ev[2].record()
            loss.backward()                                       # DDP: bucketed all-reduce overlapped with the U-Net backward
        opt.step()
        return loss

    def timed(n, sync=True):
        barrier()
        ev[0].record()
        loss_ms = 0.0
        for _ in range(n):
            loss = one_step(sync)
        ev[3].record()
        barrier()
        ms = ev[0].elapsed_time(ev[3]) / n
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), loss

    steps, warm = args.c5_steps, 3
    for _ in range(warm):
        one_step()
    step_ms, loss = timed(steps)
    one_step()
    torch.cuda.synchronize()
    loss_kernel_ms = ev[1].elapsed_time(ev[2])                    # fused MixedLoss forward_encoded (fwd + bwd) of the last step
    global_loss = float(sharding.global_mean_loss(loss.detach(), B))
    rec = {"config": "configs[4]: full single-view U-Net training step with the fused rendering loss, global batch %d sharded over %d B200"
                     % (GLOBAL_BATCH, world),
           "scaling": "strong", "global_batch": GLOBAL_BATCH, "batch_per_gpu": B, "n_gpus": world, "size": size, "N": 9,
           "model": "examples/unet_standin.py (shape and size of the reference's SingleViewModel generator)", "params_M": n_params / 1e6,
           "grad_allreduce_MB": n_params * 4 / 1e6, "dtype": "f32 (cuDNN convolutions with TF32 as torch defaults; loss kernels fp32)",
           "steps": steps, "step_ms": step_ms, "samples_per_s": GLOBAL_BATCH / (step_ms * 1e-3),
           "fused_mixed_loss_fwd_bwd_ms": loss_kernel_ms, "loss_share_of_step": loss_kernel_ms / step_ms,
           "loss_G_evals_per_s_per_gpu": B * size * size * 9 / (loss_kernel_ms * 1e-3) / 1e9, "global_loss": global_loss,
           "timed_region": "zero_grad + U-Net forward + MixedLoss.forward_encoded + backward (DDP bucketed all-reduce, 25 MB buckets) + fused Adam; CUDA events, max over ranks"}
    if world > 1:
        nosync_ms, _ = timed(max(3, steps // 2), sync=False)      # same step without the gradient all-reduce
        one_step()                                                # leave the replicas consistent again (this one all-reduces)
        flat = torch.empty(n_params, device=dev)
        for _ in range(2):
            dist.all_reduce(flat)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dist.all_reduce(flat)
        e1.record()
        barrier()
        ar_ms = e0.elapsed_time(e1) / 5
        t = torch.tensor([ar_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ar_ms = float(t.item())
        rec.update({"collective": "NCCL all-reduce (average) of the CNN gradients, %.0f MB, DistributedDataParallel buckets" % (n_params * 4 / 1e6),
                    "allreduce_alone_ms": ar_ms, "allreduce_bus_GBps": 2.0 * (world - 1) / world * n_params * 4 / (ar_ms * 1e-3) / 1e9,
                    "step_ms_without_allreduce": nosync_ms, "allreduce_exposed_ms": max(0.0, step_ms - nosync_ms),
                    "allreduce_overlapped_fraction": max(0.0, min(1.0, 1.0 - (step_ms - nosync_ms) / ar_ms)) if ar_ms > 0 else None})
        del flat
    del net, model, opt, images, target
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist
    from svbrdf_estimation_b200 import _cabi
    from svbrdf_estimation_b200 import environment as E
    from svbrdf_estimation_b200.renderers import coordinate_table

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()
    B, size, N, nr, ns, desc = WORKLOADS[args.workload]
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    HW = size * size
    evals_per_step = B * HW * N

    # ---- synthetic inputs: two rotating buffer sets, each (in+tgt+grad) > L2 ----------------------
    n_sets = 2
    host_in = synthetic_maps(B, size, 1001 + rank)
    host_tg = synthetic_maps(B, size, 2001 + rank)
    sets = []
    for s in range(n_sets):
        sets.append((host_in.to(dev).roll(s, 0).contiguous(), host_tg.to(dev).roll(s, 0).contiguous(),
                     torch.empty(B, 12, size, size, device=dev)))
    torch.manual_seed(313 + rank)
    records = [E.sample_loss_configs(B, nr, ns) for _ in range(4)]   # pre-sampled scenes (sampler timed separately)
    t0 = time.perf_counter()
    E.sample_loss_configs(B, nr, ns)
    sampler_ms = (time.perf_counter() - t0) * 1e3
    lin = coordinate_table(size, dev)
    ws_bytes = lib.svbrdf_b200_workspace_bytes(B, N, size, size)
    ws = torch.empty(ws_bytes // 4 + 1, device=dev)
    loss = torch.zeros(1, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step(i):
        a, b, g = sets[i % n_sets]
        rec = records[i % len(records)]
        _cabi.check(lib.svbrdf_b200_loss_forward_backward(a.data_ptr(), b.data_ptr(), B, size, size, rec.data_ptr(), N,
                                                          lin.data_ptr(), loss.data_ptr(), g.data_ptr(), ws.data_ptr(),
                                                          ws_bytes, stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        step(i)
    ev1.record()
    barrier()
    sampler.stop()
    clock_note = "NVML polled every 4 ms during the timed region"
    total_ms = ev0.elapsed_time(ev1)
    if len(sampler.samples) < 5:
        # timed region too short for the poller: sample during an extra untimed loop of the same step
        sampler.start()
        t_end = time.perf_counter() + 0.4
        i = 0
        while time.perf_counter() < t_end:
            step(i); i += 1
            if i % 64 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        sampler.stop()
        clock_note = "timed region was %.1f ms: NVML polled during an extra untimed 0.4 s loop of the same step" % total_ms
    barrier()
    # per-launch statistics of the step (fused kernel + finalize), outside the timed region: one event pair per step
    n_stat = min(steps, 50)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_stat + 1)]
    ev[0].record()
    for i in range(n_stat):
        step(i)
        ev[i + 1].record()
    torch.cuda.synchronize()
    per_step = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n_stat))
    loss_value = float(loss.item())
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / steps
    value = world * evals_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- informational: the accurate-highlight variant of the same step -----------------
    accurate = None
    if world == 1:
        def step_acc(i):
            a, b, g = sets[i % n_sets]
            rec = records[i % len(records)]
            _cabi.check(lib.svbrdf_b200_loss_forward_backward_accurate(a.data_ptr(), b.data_ptr(), B, size, size, rec.data_ptr(), N,
                                                                       lin.data_ptr(), loss.data_ptr(), g.data_ptr(), ws.data_ptr(),
                                                                       ws_bytes, stream))
        n_acc = max(10, min(steps, 50))
        for i in range(3):
            step_acc(i)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for i in range(n_acc):
            step_acc(i)
        eb.record()
        torch.cuda.synchronize()
        ms_acc = ea.elapsed_time(eb) / n_acc
        accurate = {"entry": "svbrdf_b200_loss_forward_backward_accurate", "ms_per_step": ms_acc,
                    "value": evals_per_step / (ms_acc * 1e-3) / 1e9, "unit": UNIT, "steps": n_acc,
                    "frac_of_fp32_roofline": None,
                    "note": "gradients ~2e-6 of fp64 instead of ~7e-5 (profiles/r1_accuracy_study.txt); not the headline value"}
        step(0)                      # leave the default kernels' loss in `loss` for the parity check below
        torch.cuda.synchronize()

    # ---- e2e: host buffers through the C-ABI host entry point ---------------------------------------
    e2e = None
    if not args.no_e2e:
        all_cpus = os.sched_getaffinity(0)
        numa_note = bind_host_thread_near_gpu(local_rank)
        ctx = ctypes.c_void_p()
        _cabi.check(lib.svbrdf_b200_ctx_create(ctypes.byref(ctx), B, N, size, size))
        pin = [lib.svbrdf_b200_ctx_pinned(ctx, w) for w in range(3)]
        res3 = (ctypes.c_float * 3)()
        e2e_steps = max(3, min(steps, 20))
        to10 = lambda m: torch.cat((m[:, 0:7], m[:, 9:12]), dim=1).contiguous()
        host_enc = (torch.rand(B, 9, size, size, generator=torch.Generator().manual_seed(77 + rank)) * 1.8 - 0.9)

        def host_leg(a_host, la, b_host, lb, l1w, with_grad, n_steps):
            """Times svbrdf_b200_loss_host with the context's own pinned buffers (filled here, outside the timed region)."""
            ctypes.memmove(pin[0], a_host.data_ptr(), a_host.numel() * 4)
            ctypes.memmove(pin[1], b_host.data_ptr(), b_host.numel() * 4)

            def one(i):
                rec = records[i % len(records)]
                _cabi.check(lib.svbrdf_b200_loss_host(ctx, pin[0], la, pin[1], lb, B, rec.data_ptr(), N, l1w, res3,
                                                      pin[2] if with_grad else None))
            for i in range(3):
                one(i)
            barrier()
            t0 = time.perf_counter()
            for i in range(n_steps):
                one(i)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item()) / n_steps
            h2d = (a_host.numel() + b_host.numel()) * 4
            d2h = (a_host.numel() * 4 if with_grad else 0) + 12
            return {"value": world * evals_per_step / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_steps,
                    "pcie_GBps_per_gpu_both_directions": (h2d + d2h) / dt / 1e9}, float(res3[0])

        in10, tg10 = to10(host_in), to10(host_tg)
        e10, loss10 = host_leg(in10, 10, tg10, 10, -1.0, True, e2e_steps)
        e12, loss12 = host_leg(host_in, 12, host_tg, 12, -1.0, True, e2e_steps)
        e9, _ = host_leg(host_enc, 9, tg10, 10, 0.1, True, e2e_steps)
        eng, _ = host_leg(in10, 10, tg10, 10, -1.0, False, e2e_steps)
        # parity of host and device entry points on the same inputs (bitwise: same kernels, same coordinates)
        step(0)
        torch.cuda.synchronize()
        ctypes.memmove(pin[0], host_in.data_ptr(), host_in.numel() * 4)
        ctypes.memmove(pin[1], host_tg.data_ptr(), host_tg.numel() * 4)
        _cabi.check(lib.svbrdf_b200_loss_host(ctx, pin[0], 12, pin[1], 12, B, records[0].data_ptr(), N, -1.0, res3, pin[2]))
        same = abs(res3[0] - float(loss.item())) <= 1e-7 * abs(float(loss.item())) and loss10 == loss12
        lib.svbrdf_b200_ctx_destroy(ctx)

        # the path a caller of the reference interface takes with PAGEABLE CPU tensors: RenderingLoss(LocalRenderer())(x, t)
        # + backward(), scene sampling (reference order) included; tensors staged through pinned double buffers both ways
        pageable = None
        if world == 1:
            import svbrdf_estimation_b200 as S
            mod = S.RenderingLoss(S.LocalRenderer())
            xin = host_in.clone().requires_grad_(True)

            def api_step():
                xin.grad = None
                mod(xin, host_tg).backward()
            for _ in range(2):
                api_step()
            torch.cuda.synchronize()
            n_api = max(3, min(steps, 8))
            t0 = time.perf_counter()
            for _ in range(n_api):
                api_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n_api
            # what the same tensors cost with torch's own blocking pageable copies (no kernel): tensor.cuda() x2, grad.cpu()
            gdev = torch.empty(B, 12, size, size, device=dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                host_in.cuda(); host_tg.cuda(); gdev.cpu()
            torch.cuda.synchronize()
            plain_ms = (time.perf_counter() - t0) / 3 * 1e3
            del gdev
            pageable = {"torch_blocking_pageable_copies_alone_ms": plain_ms,"value": evals_per_step / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": n_api,
                        "call": "RenderingLoss(LocalRenderer())(cpu_input, cpu_target).backward() with pageable [64,12,256,256] CPU tensors "
                                "(renderers.host_to_device -> staging.upload / download: 32 MB pinned double buffers)",
                        "h2d_bytes_per_step": 2 * host_in.numel() * 4, "d2h_bytes_per_step": host_in.numel() * 4 + 4}
        os.sched_setaffinity(0, all_cpus)           # the CPU baseline below uses every host thread again
        e2e = dict(e10)
        e2e.update({"entry": "svbrdf_b200_loss_host, SVBRDF_LAYOUT_MAPS10 input/target/gradient (roughness stored once: the model and the "
                             "dataset replicate it, utils.py:78-80): host maps -> loss + gradient on the host, 16 batch slices pipelined "
                             "over H2D / compute / D2H streams",
                    "host_buffers": "the context's own PINNED buffers (svbrdf_b200_ctx_pinned), filled before the timed region; every step "
                                    "copies them to the device, computes, and copies loss + gradient back inside the timed region",
                    "matches_device_entry": bool(same), "host_placement": numa_note,
                    "maps12_variant": dict(e12, note="same call with the reference's 12-channel tensors (SVBRDF_LAYOUT_MAPS12): 36 instead of 30 planes per step"),
                    "encoded9_variant": dict(e9, note="MixedLoss on the 9-channel network output + 10-channel target (28 planes per step)"),
                    "loss_only_variant": dict(eng, note="forward only (no gradient computed or downloaded): the upload alone, i.e. the PCIe floor of this entry point"),
                    "python_api_pageable_tensors": pageable})

    # ---- configs[4]: the full training step, global batch 256 strong-scaled over the ranks (all ranks take part) ----
    c5 = None
    if not args.no_c5:
        try:
            c5 = train_step_c5(args, torch, dist, dev, world, rank, local_rank)
        except Exception as exc:                  # never lose the headline line over the informational leg
            if world > 1:
                raise
            c5 = {"unavailable": repr(exc)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline ----------------------------------------------------------------------------------------
    peaks, peak_src = measured_peaks()
    probes = probe_fp32_peaks(lib, torch, stream)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    nominal_fp32 = sms * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    kernel_ms = kernel_ms_for_roofline(ms_per_step, per_step, world)
    ach_tflops = FLOP_PER_EVAL * evals_per_step / (kernel_ms * 1e-3) / 1e12
    ach_gbs = BYTES_PER_PIXEL * B * HW / (kernel_ms * 1e-3) / 1e9
    # profiler-derived facts of the shipped kernel, regenerated by scripts/measure_all.sh + scripts/profile_to_json.py:
    # DRAM bytes per launch (ncu) for this workload, executed FLOP per evaluation (SASS), loop / out-of-loop shares (ncu)
    def load_json(name):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)
        except Exception:
            return {}
    kprof, sass = load_json("kernel_profile.json"), load_json("sass_model.json")
    build_id = lib.svbrdf_b200_build_id().decode()
    traffic = (kprof.get(args.workload) or {}).get("traffic")
    executed = None
    if sass.get("executed_flop_per_eval"):
        fe = float(sass["executed_flop_per_eval"])
        share = kprof.get("loop_share_c2") or {}
        executed = {"flop_per_eval_executed": fe, "mufu_per_eval_executed": sass.get("mufu_per_eval"),
                    "tflops_executed": fe * evals_per_step / (kernel_ms_for_roofline(ms_per_step, per_step, world) * 1e-3) / 1e12,
                    "flop_per_eval_algorithmic": FLOP_PER_EVAL, "registers": sass.get("registers"),
                    "record_loop_instructions": sass.get("record_loop_instructions"),
                    "out_of_loop_instruction_share_c2": share.get("out_of_loop_instruction_share"),
                    "out_of_loop_residency_share_c2": share.get("out_of_loop_residency_share"),
                    "sass_model_matches_this_build": sass.get("build_id") == build_id,
                    "ncu_profile_matches_this_build": kprof.get("build_id") == build_id,
                    "how": "profiles/sass_model.json (scripts/sass_stats.py on the shipped cubin), profiles/kernel_profile.json (ncu)"}
    roofline = {"bound": "fp32", "kernel": "loss_kernel<BWD=1,MIXED=0>", "achieved": ach_tflops, "peak": nominal_fp32,
                "unit": "TFLOP/s", "frac": ach_tflops / nominal_fp32, "traffic": traffic,
                "peak_source": "nominal %d SM x 128 lanes x 2 x %.0f MHz (no FP32 figure in MEASURED_PEAKS.json)" % (sms, sm_max_mhz),
                "peak_measured_ffma": probes.get("ffma"), "frac_of_measured_ffma": ach_tflops / probes["ffma"] if probes.get("ffma") else None,
                "flop_per_eval": FLOP_PER_EVAL, "evals_per_launch": evals_per_step, "kernel_ms": kernel_ms,
                "G_evals_per_s_kernel": evals_per_step / (kernel_ms * 1e-3) / 1e9,
                "G_evals_per_s_at_100pct": nominal_fp32 * 1e12 / FLOP_PER_EVAL / 1e9}
    # the bound that actually binds: operand fetch from the two register-file banks of each scheduler
    warp_records = B * HW * N / 64.0
    rf_cycles = float(sass.get("register_file_cycles_model") or RF_CYCLES_PER_WARP_RECORD)
    rf_ms = rf_cycles * warp_records / (sms * 4) / (sm_max_mhz * 1e6) * 1e3
    roofline_rf = {"bound": "register-file operand bandwidth (a MODEL fitted to micro-benchmarks, profiles/r1_microbench.txt - no hardware counter exposes register-bank conflicts on this part; informational)", "cycles_per_warp_record": rf_cycles,
                   "bound_ms": rf_ms, "frac": rf_ms / kernel_ms,
                   "how": "scripts/sass_stats.py cost model (B300_MICROARCH.md 'RF banking': rt = max(pipe, distinct even, distinct odd source registers)) "
                          "x warp-record iterations / (SMs x 4 schedulers x SM clock); prologue/epilogue not counted"}
    if accurate is not None:
        accurate["frac_of_fp32_roofline"] = FLOP_PER_EVAL * evals_per_step / (accurate["ms_per_step"] * 1e-3) / 1e12 / nominal_fp32
    roofline_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": float(peaks["hbm_gbs"]), "unit": "GB/s",
                    "frac": ach_gbs / float(peaks["hbm_gbs"]), "peak_source": peak_src + " MEASURED_PEAKS.json hbm_gbs",
                    "bytes_per_pixel": BYTES_PER_PIXEL, "traffic": traffic}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # in a fresh process (the same code path as --impl reference): inside this one - CUDA context, pinned pools, NCCL and
        # NVML threads alive - the same CPU passes were measured 1.8x slower than on their own
        import subprocess
        cpu = None
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload,
                                  "--steps", "3", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                                 timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
            ref = json.loads([l for l in out.stdout.splitlines() if l.strip().startswith("{")][-1])
            cpu = dict(ref["cpu_baseline"])
            cpu["sample"] += "; 1 warm-up + 3 timed passes of %.2f s in a separate CPU-only process" % (ref["ms_per_step"] / 1e3)
        except Exception as exc:
            cpu = {"unavailable": repr(exc)[:200]}
    eager = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, sec, b = time_gpu_eager_port(args.workload, dev)
            eager = {"value": v, "unit": UNIT, "kind": "port", "device": "same B200, eager PyTorch",
                     "sample": "first %d of %d maps, fwd+bwd via autograd, %.2f s per pass" % (b, B, sec)}
        except Exception as exc:          # informational leg only
            eager = {"unavailable": repr(exc)[:200]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "batch_per_gpu": B, "size": size, "N": N,
                       "parallelism": "batch-sharded x%d, no data-path collective" % world,
                       "l2": "2 rotating buffer sets; each set (input+target+grad) is %.0f MB > 126 MB L2" % (3 * B * 12 * HW * 4 / 1e6),
                       "scene_sampler_ms_per_step_host": sampler_ms},
            "value_per_gpu": value / world, "loss": loss_value,
            "roofline": roofline, "executed_work": executed, "roofline_hbm": roofline_hbm, "roofline_register_file": roofline_rf, "fp32_probes_tflops": probes,
            "accurate_variant": accurate, "train_step_c5": c5,
            "cpu_baseline": cpu, "reference_port_eager_on_gpu": eager, "e2e": e2e, "gpu_launches": 2 * steps,
            "gpu_launches_note": "per step: 1 fused loss fwd+bwd kernel + 1 single-CTA finalize kernel",
            "clocks": sampler.summary(clock_note),
            "step_ms_with_event_per_step": {"min": per_step[0], "median": statistics.median(per_step), "max": per_step[-1], "n": len(per_step)}}
    args.out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


class OnlyJsonOnStdout:
    """Everything any library writes to file descriptor 1 during the run (NCCL prints its version banner there) goes to
    stderr; `emit` writes the one JSON line to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the configs[4] training-step leg")
    ap.add_argument("--c5-global-batch", type=int, default=256)
    ap.add_argument("--c5-steps", type=int, default=6)
    args = ap.parse_args()
    args.out = OnlyJsonOnStdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
