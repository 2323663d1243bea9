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


def time_cpu_port(workload, steps, warmup, budget_s):
    """Times the oracle port on a bounded sample (first ``b`` batch elements) of the workload.
    Returns (G evals/s, seconds per step, sample batch, threads)."""
    import torch
    from svbrdf_estimation_b200 import environment as E
    B, size, N, nr, ns, _ = WORKLOADS[workload]
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(313)
    probe_b = 1
    inp, tgt = synthetic_maps(probe_b, size, 1001), synthetic_maps(probe_b, size, 2001)
    cfg = E.sample_loss_configs(probe_b, nr, ns)
    cpu_port_step(inp, tgt, cfg)
    t0 = time.perf_counter()
    cpu_port_step(inp, tgt, cfg)
    t1 = time.perf_counter() - t0
    b = 8 if size <= 256 else 1
    while b > 1 and (steps + warmup) * t1 * b > budget_s:
        b //= 2
    b = min(b, B)
    inp, tgt = synthetic_maps(b, size, 1001), synthetic_maps(b, size, 2001)
    cfg = E.sample_loss_configs(b, nr, ns)
    for _ in range(warmup):
        cpu_port_step(inp, tgt, cfg)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_port_step(inp, tgt, cfg)
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    evals = b * size * size * N
    return evals / sec / 1e9, sec, b, torch.get_num_threads()


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
    val, sec, b, threads = time_cpu_port(args.workload, steps, warmup, budget_s=150.0)
    sample = "first %d of %d maps of the workload per step (%dx%d, N=%d), fwd+bwd via autograd" % (b, B, size, size, N)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "batch_per_step": b, "size": size, "N": N},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = its eager-PyTorch CPU path restated in oracle/reference_port.py (bit-exact to the "
                    "reference in fp32, tests/test_oracle_golden.py); the reference itself is Python source that does "
                    "not exist on the GPU box"}
    args.out.emit(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
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
        nfl = B * 12 * HW
        pin = [lib.svbrdf_b200_ctx_pinned(ctx, w) for w in range(3)]
        ctypes.memmove(pin[0], host_in.data_ptr(), nfl * 4)
        ctypes.memmove(pin[1], host_tg.data_ptr(), nfl * 4)
        lossf = ctypes.c_float(0.0)
        e2e_steps = max(3, min(steps, 20))

        def e2e_step(i):
            rec = records[i % len(records)]
            _cabi.check(lib.svbrdf_b200_rendering_loss_host(ctx, pin[0], pin[1], B, rec.data_ptr(), N,
                                                            ctypes.byref(lossf), pin[2]))
        for i in range(3):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        # informational: the same call when the gradient is left on the device (what a training loop does with it)
        def e2e_step_nograd(i):
            rec = records[i % len(records)]
            _cabi.check(lib.svbrdf_b200_rendering_loss_host(ctx, pin[0], pin[1], B, rec.data_ptr(), N, ctypes.byref(lossf), None))
        for i in range(2):
            e2e_step_nograd(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step_nograd(i)
        torch.cuda.synchronize()
        dt_ng = time.perf_counter() - t0
        # parity of the two entry points on the same inputs (bitwise: same kernels, same coordinates)
        step(0)
        torch.cuda.synchronize()
        rec0 = records[0]
        _cabi.check(lib.svbrdf_b200_rendering_loss_host(ctx, pin[0], pin[1], B, rec0.data_ptr(), N, ctypes.byref(lossf), pin[2]))
        same = abs(lossf.value - float(loss.item())) <= 1e-7 * abs(float(loss.item()))
        lib.svbrdf_b200_ctx_destroy(ctx)
        os.sched_setaffinity(0, all_cpus)           # the CPU baseline below uses every host thread again
        e2e = {"value": world * evals_per_step / (dt / e2e_steps) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": 2 * nfl * 4, "d2h_bytes_per_step": nfl * 4 + 4, "steps": e2e_steps,
               "ms_per_step": dt / e2e_steps * 1e3, "entry": "svbrdf_b200_rendering_loss_host (pinned host maps -> loss + grad on host)",
               "matches_device_entry": bool(same), "host_placement": numa_note,
               "loss_only_variant": {"value": evals_per_step / (dt_ng / e2e_steps) / 1e9, "unit": UNIT, "ms_per_step": dt_ng / e2e_steps * 1e3,
                                     "d2h_bytes_per_step": 4, "note": "per rank; forward only (no gradient is computed or downloaded): "
                                     "the upload alone, i.e. the PCIe floor of this entry point"}}

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
    kernel_ms = ms_per_step if world == 1 else statistics.median(per_step)   # step = fused kernel + 1-CTA finalize (~2 %)
    ach_tflops = FLOP_PER_EVAL * evals_per_step / (kernel_ms * 1e-3) / 1e12
    ach_gbs = BYTES_PER_PIXEL * B * HW / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(args.workload)
    except Exception:
        pass
    roofline = {"bound": "fp32", "kernel": "loss_kernel<BWD=1,MIXED=0>", "achieved": ach_tflops, "peak": nominal_fp32,
                "unit": "TFLOP/s", "frac": ach_tflops / nominal_fp32, "traffic": traffic,
                "peak_source": "nominal %d SM x 128 lanes x 2 x %.0f MHz (no FP32 figure in MEASURED_PEAKS.json)" % (sms, sm_max_mhz),
                "peak_measured_ffma": probes.get("ffma"), "frac_of_measured_ffma": ach_tflops / probes["ffma"] if probes.get("ffma") else None,
                "flop_per_eval": FLOP_PER_EVAL, "evals_per_launch": evals_per_step, "kernel_ms": kernel_ms,
                "G_evals_per_s_kernel": evals_per_step / (kernel_ms * 1e-3) / 1e9,
                "G_evals_per_s_at_100pct": nominal_fp32 * 1e12 / FLOP_PER_EVAL / 1e9}
    # the bound that actually binds: operand fetch from the two register-file banks of each scheduler
    warp_records = B * HW * N / 64.0
    rf_ms = RF_CYCLES_PER_WARP_RECORD * warp_records / (sms * 4) / (sm_max_mhz * 1e6) * 1e3
    roofline_rf = {"bound": "register-file operand bandwidth (model)", "cycles_per_warp_record": RF_CYCLES_PER_WARP_RECORD,
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
        val, sec, b, threads = time_cpu_port(args.workload, 15, 1, budget_s=40.0)
        cpu = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d of %d maps of the workload, %dx%d, N=%d, 1 warm-up + 15 timed fwd+bwd passes (%.2f s each)"
                         % (b, B, size, size, N, sec)}

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
            "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_register_file": roofline_rf, "fp32_probes_tflops": probes,
            "accurate_variant": accurate,
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
    args = ap.parse_args()
    args.out = OnlyJsonOnStdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
