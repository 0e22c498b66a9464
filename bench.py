#!/usr/bin/env python
"""bench.py -- headline benchmark of the sketch-similarity hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): count-min sketch update of synthetic Zipf(1.1)
(user, item, pref) events over 10^7 items into one depth-4 x width-2^20 sketch; one step = one
pass of K1 over one HBM-resident batch of 10^9 events (12 B/event read: item key + pref).
Prints ONE JSON line (see DESIGN.md "Measurement" for every field).

N > 1 (launched by torchrun, one rank per GPU): every rank streams its own 10^9-event slice
into a private replica sketch (weak scaling, no data-path collective) and one NCCL all-reduce
of the 32 MiB counter array per step produces the single global sketch.

--impl reference times the reference's CPU algorithm (oracle/ C port, all host threads) on a
bounded sample of the same event stream; the Java itself cannot run here (no JVM in the image).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240002           # 20240001 + config number (config 2 -> ...0003 is 1-based; fixed here)
SKETCH_SEED = 42
DEPTH, WIDTH = 4, 1 << 20
ITEMS, USERS, ZIPF_S = 10_000_000, 1_000_000, 1.1
ALGO_BYTES_PER_EVENT = 20 + DEPTH * 16   # 20 B event + d x (8 B read + 8 B write)
METRIC = "sketch_updates_per_sec"
UNIT = "events/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=float(j["hbm_gbs"]), bf16=float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                    bf16_burst=float(j["bf16_tflops"]), source="measured")
    return dict(hbm=6650.0, bf16=1400.0, bf16_burst=1590.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def _physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------
def _cpu_update_rate(item, pref, threads: int):
    """events/s of the reference algorithm (oracle C port) on the given events."""
    import oracle as orc
    a, b = orc.hash_params(SKETCH_SEED, DEPTH)
    bank = np.zeros((1, DEPTH, WIDTH))
    t0 = time.perf_counter()
    orc.bank_update(bank, DEPTH, WIDTH, a, b, None, item, pref, nthreads=threads)
    dt = time.perf_counter() - t0
    return item.shape[0] / dt, dt, bank


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as orc
    from mahout_b200 import synth
    threads = orc.max_threads()
    cdf = synth.zipf_cdf(ITEMS, ZIPF_S)
    _, item, pref = synth.events_numpy(SEED, 0, 1 << 22, USERS, cdf)
    rate, _, _ = _cpu_update_rate(item, pref, threads)           # calibration
    per_step = int(min(max(rate * 4.0, 1 << 25), 1 << 27))       # ~4 s of CPU work per step
    _, item, pref = synth.events_numpy(SEED, 0, per_step, USERS, cdf)
    for w in range(args.warmup):
        m = min(per_step, 1 << 22)
        _cpu_update_rate(item[:m], pref[:m], threads)
    t = 0.0
    for s in range(args.steps):
        _, dt, _ = _cpu_update_rate(item, pref, threads)
        t += dt
    value = per_step * args.steps / t
    sample = (f"first {per_step} events of the config-2 stream (Zipf({ZIPF_S}) over {ITEMS} items) per step, "
              f"d={DEPTH}, W=2^20")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": _config(args.gpus, per_step),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference Java cannot run (no JVM in image): oracle/ C port of HashFunction.hash + "
                "DoubleCountMinSketch.update, OpenMP over event slices with private sketches",
    }
    print(json.dumps(line), flush=True)


def _config(n_gpus: int, events_per_step: int):
    return {"workload": "configs[1]: count-min sketch update, synthetic Zipf(1.1) (user,item,pref) events over "
                        "1e7 items into depth=4 x width=2^20 sketch",
            "events_per_step_per_gpu": events_per_step, "depth": DEPTH, "width": WIDTH,
            "items": ITEMS, "zipf_s": ZIPF_S, "sketch_seed": SKETCH_SEED, "event_seed": SEED,
            "l2": "inputs_exceed_l2 (12 GB of events per step per GPU; the 32 MiB counter array is L2-resident state)",
            "parallelism": f"replica sketches x{n_gpus} + all-reduce(int64 sum)" if n_gpus > 1 else "single GPU"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import mahout_b200 as mb
    from mahout_b200 import _native as N
    from mahout_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    peaks = _peaks()

    ctx = mb.Context(local)
    # one non-default stream carries everything: the library's kernels, NCCL and the timing events
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    n = int(args.events)
    cdf = synth.zipf_cdf(ITEMS, ZIPF_S)
    cdf_dev = torch.from_numpy(cdf).to(dev)
    # resident batch: each rank owns events [rank*n, (rank+1)*n) of the stream
    _, item, pref = synth.events_device(ctx, SEED, rank * n, n, USERS, cdf_dev, None, want_user=False)
    bank = mb.SketchBank(1, WIDTH, DEPTH, SKETCH_SEED, 1, ctx)
    cptr, cells = bank.counters_ptr()
    from mahout_b200.sketch import _as_tensor
    counters = _as_tensor(cptr, cells, local)
    global_sketch = torch.empty_like(counters) if world > 1 else None

    def step():
        bank.update(None, item, pref)
        if world > 1:
            global_sketch.copy_(counters)
            dist.all_reduce(global_sketch, op=dist.ReduceOp.SUM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    bank.check()
    barrier()
    ctx.set_profiling(True)
    ctx.reset_profile()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(_physical_gpu_index(local))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    k_ms, k_n = ctx.kernel_time(N.K_UPDATE)
    ctx.set_profiling(False)
    launches = ctx.launch_count() - launches0
    bank.check()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * n * args.steps / (ms * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers (H2D inside, sketch read back) ----------
    e2e_n = int(min(n, args.e2e_events))
    hk = torch.empty(e2e_n, dtype=torch.int64).pin_memory()
    hp = torch.empty(e2e_n, dtype=torch.float32).pin_memory()
    hk.copy_(item[:e2e_n])
    hp.copy_(pref[:e2e_n])
    hout = torch.empty(DEPTH * WIDTH, dtype=torch.float64).pin_memory()
    hk_np, hp_np = hk.numpy(), hp.numpy()
    ebank = mb.SketchBank(1, WIDTH, DEPTH, SKETCH_SEED, 1, ctx)

    def e2e_step():
        N.check(N.lib().mb200_bank_update(ebank.handle, None, C.c_void_p(hk_np.ctypes.data),
                                          C.c_void_p(hp_np.ctypes.data), e2e_n, N.MEM_HOST), ctx.handle)
        N.check(N.lib().mb200_bank_read(ebank.handle, 0, 1, C.c_void_p(hout.data_ptr()), N.MEM_HOST),
                ctx.handle)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * e2e_n * e2e_steps / e2e_s

    # ---- parity + cpu_baseline on rank 0 ---------------------------------------------------------
    cpu = None
    parity = None
    if rank == 0:
        import oracle as orc
        threads = orc.max_threads()
        cal = 1 << 22
        rate, _, _ = _cpu_update_rate(item[:cal].cpu().numpy(), pref[:cal].cpu().numpy(), threads)
        sample = int(min(max(rate * 12.0, 1 << 25), 1 << 28, n))
        cpu_rate, cpu_dt, cpu_bank = _cpu_update_rate(item[:sample].cpu().numpy(),
                                                      pref[:sample].cpu().numpy(), threads)
        pbank = mb.SketchBank(1, WIDTH, DEPTH, SKETCH_SEED, 1, ctx)
        pbank.update(None, item[:sample], pref[:sample])
        got = pbank.read()
        parity = {"sketch_bit_exact": bool(got.tobytes() == cpu_bank.tobytes()), "events_checked": sample}
        pbank.close()
        cpu = {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {sample} events of rank 0's stream, {cpu_dt:.1f} s; oracle/ C port "
                         "(the Java reference cannot run: no JVM in the image)"}

    if rank == 0:
        kern_s = (k_ms / max(k_n, 1)) * 1e-3
        achieved = ALGO_BYTES_PER_EVENT * n / kern_s / 1e9
        physical = 12 * n / kern_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64 fixed-point counters (== reference f64, exact)",
            "data": "synthetic", "config": _config(world, n),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 12 * e2e_n,
                    "d2h_bytes_per_step": 8 * DEPTH * WIDTH,
                    "sample": f"{e2e_n} events/step from pinned host memory through mb200_bank_update(MEM_HOST) "
                              "+ mb200_bank_read of the whole sketch"},
            "roofline": {"bound": "hbm", "kernel": "k_update_single", "achieved": achieved, "peak": peaks["hbm"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm"], "traffic": None,
                         "peak_source": peaks["source"], "algorithmic_bytes_per_event": ALGO_BYTES_PER_EVENT,
                         "kernel_ms_per_launch": kern_s * 1e3, "launches_timed": int(k_n),
                         "physical_event_read_GBps": physical,
                         "atomic_updates_per_s": DEPTH * n / kern_s},
            "cpu_baseline": cpu, "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events", type=float, default=1e9, help="events per step per GPU")
    ap.add_argument("--e2e-events", type=float, default=float(1 << 27))
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
