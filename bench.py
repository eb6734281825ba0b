#!/usr/bin/env python
"""bench.py -- lazy-Gramian MVM throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c1|c3|c4]

A "step" is one pass of the hot path: one `mul!(b, K, a)` with K = gramian(EQ(), x), d = 3, n = 2^20, Float64
(BASELINE.json configs[1]).  Rows of K are sharded as contiguous blocks over the ranks (one process per GPU under
torchrun), x and a are replicated, and every step ends with an NCCL all-gather of b (the chained-MVM form).

  value        whole-job kernel-pair evaluations per second, operands resident in HBM (device pointers, CUDA events)
  e2e          the same metric through the public host API: gramian(k, X) (uploads X), mul_(b, G, a) with pinned HOST
               buffers (uploads a, downloads b) -- every copy inside the timed region
  roofline     FP64 FMA-pipe roofline of the dominant kernel: algorithmic flops (46 flops = 23
               issue slots per EQ pair at d = 3) / measured kernel time, against the DFMA peak measured in this run
               by cf_peak_probe (MEASURED_PEAKS.json carries no FP64 figure)
  cpu_baseline the CPU restatement of the reference loop (oracle/, OpenMP over rows like Julia's @threads) on a
               bounded row sample of the same workload
`--impl reference` times that CPU restatement alone (Julia is not installed, so the reference binary cannot run).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SLOTS = {"c1": 35, "c2": 23, "c3": 111, "c4": 97, "c5": 45, "x1": 80, "x2": 81, "x3": 61, "x4": 45, "x5": 110, "x6": 125}  # FP64 issue slots per pair/block


def workload(name):
    import covfn_b200 as cf

    if name == "c2":
        return dict(kernel=cf.EQ(), kname="EQ", d=3, n=1 << 20, nrhs=1, gradient=False,
                    desc="EQ Gramian MVM, d=3, n=2^20, Float64 (BASELINE.json configs[1])")
    if name == "c1":
        return dict(kernel=cf.MaternP(2), kname="MaternP(2)", d=3, n=16384, nrhs=1, gradient=False,
                    desc="MaternP(2) Gramian MVM, d=3, n=16384, Float64 (BASELINE.json configs[0])")
    if name == "c3":
        return dict(kernel=0.5 * cf.RQ(2) + cf.Dot() ** 2, kname="1/2*RQ(2)+Dot()^2", d=32, n=262144, nrhs=64,
                    gradient=False, desc="1/2*RQ(2)+Dot()^2 multi-RHS mul!(B,K,A), d=32, n=262144, 64 columns, Float64 (configs[2])")
    if name == "c4":
        return dict(kernel=cf.EQ(), kname="GradientKernel(EQ)", d=16, n=65536, nrhs=1, gradient=True,
                    desc="GradientKernel(EQ) MVM, d=16, n=65536, Float64 (configs[3])")
    if name == "x1":  # not a BASELINE config: composite-kernel MVM used to measure the interpreter vs run-time specialisation
        return dict(kernel=0.5 * cf.EQ() + cf.MaternP(2) * cf.RQ(2), kname="1/2*EQ+MaternP(2)*RQ(2)", d=3, n=262144, nrhs=1,
                    gradient=False, desc="composite-kernel Gramian MVM, d=3, n=262144, Float64 (auxiliary)")
    if name == "x2":  # auxiliary: high-dimensional single-RHS MVM (README.md:369-395 shape at larger n)
        return dict(kernel=cf.EQ(), kname="EQ", d=32, n=131072, nrhs=1, gradient=False,
                    desc="EQ Gramian MVM, d=32, n=131072, Float64 (auxiliary)")
    if name == "x3":
        return dict(kernel=cf.MaternP(2), kname="MaternP(2)", d=16, n=131072, nrhs=1, gradient=False,
                    desc="MaternP(2) Gramian MVM, d=16, n=131072, Float64 (auxiliary)")
    if name == "x4":  # auxiliary: the MVM inside BASELINE config 5 (CG), at a quarter of its n
        return dict(kernel=cf.MaternP(2), kname="MaternP(2)", d=8, n=131072, nrhs=1, gradient=False,
                    desc="MaternP(2) Gramian MVM, d=8, n=131072, Float64 (auxiliary; config 5's operator)")
    if name == "x5":  # auxiliary: gradient operator of a non-EQ kernel (generic jets)
        return dict(kernel=cf.MaternP(2), kname="GradientKernel(MaternP(2))", d=16, n=32768, nrhs=1, gradient=True,
                    desc="GradientKernel(MaternP(2)) MVM, d=16, n=32768, Float64 (auxiliary)")
    if name == "x6":  # auxiliary: gradient operator of a composite kernel (generic 8-wide jets)
        return dict(kernel=cf.EQ() + 0.5 * cf.RQ(2), kname="GradientKernel(EQ+1/2*RQ(2))", d=16, n=32768, nrhs=1, gradient=True,
                    desc="GradientKernel(EQ + 1/2 RQ(2)) MVM, d=16, n=32768, Float64 (auxiliary)")
    raise SystemExit(f"unknown config {name}")


def make_inputs(w, seed=0xC0F00002):
    rng = np.random.Generator(np.random.Philox(seed))
    d, n = w["d"], w["n"]
    scale = 1.0 if d <= 3 else 1.0 / np.sqrt(d)  # scale X by 1/sqrt(d) for C3/C4
    X = rng.standard_normal((n, d)) * scale      # (n, d) row-major == d x n column-major
    blk = d if w["gradient"] else 1
    if w["nrhs"] == 1:
        a = rng.standard_normal(n * blk)
    else:
        a = np.asfortranarray(rng.standard_normal((n * blk, w["nrhs"])))
    return X, a


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons of one GPU every 100 ms while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_port_rate(w, X, a, target_s=12.0, threads=None):
    """pairs/s of the oracle's restatement of the reference loop on a bounded row sample; returns (rate, rows, secs, threads)"""
    from oracle import oracle as O

    O.build()
    O.set_num_threads(threads or os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    prog = w["kernel"].program()
    n = w["n"]

    def run(rows):
        t0 = time.perf_counter()
        if w["gradient"]:
            O.gradient_mul(prog, X, a, rows=(0, rows))
        elif w["nrhs"] == 1:
            O.mul_vec(prog, X, a, rows=(0, rows))
        else:
            O.mul_mat(prog, X, a, rows=(0, rows))  # reference loop order: entry re-evaluated per RHS column
        return time.perf_counter() - t0

    nt = O.num_threads()
    rows = max(nt, 8)
    t = run(rows)  # calibration (also warms the threads)
    rows = int(min(n, max(rows, rows * target_s / max(t, 1e-3))))
    rows = max(nt, (rows // nt) * nt)
    t = run(rows)
    return rows * float(n) / t, rows, t, nt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = workload(args.config)
    n, d = w["n"], w["d"]
    pairs_per_step = float(n) * float(n)
    unit = "kernel-pair evaluations/s"
    metric = "Gramian MVM kernel-evals/s"
    cfg = {"workload": w["desc"], "kernel": w["kname"], "d": d, "n": n, "nrhs": w["nrhs"],
           "sharding": f"contiguous row blocks over {world} rank(s), x and a replicated, all-gather of b per step" if world > 1
           else "single GPU", "l2": "256 MiB L2 flush (memset) between timed steps, inside the timed region"}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        X, a = make_inputs(w)
        rate0, rows, _, nt = cpu_port_rate(w, X, a, target_s=6.0)
        from oracle import oracle as O

        prog = w["kernel"].program()
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            if w["gradient"]:
                O.gradient_mul(prog, X, a, rows=(0, rows))
            elif w["nrhs"] == 1:
                O.mul_vec(prog, X, a, rows=(0, rows))
            else:
                O.mul_mat(prog, X, a, rows=(0, rows))
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
        t = float(np.mean(times))
        value = rows * float(n) / t
        sample = f"rows 0..{rows} of the n={n} row MVM ({rows * float(n):.3g} pairs per step), all {n} columns"
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3 * (n / rows), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": unit, "cores": nt, "kind": "port", "sample": sample,
                             "note": "C/OpenMP restatement of reference src/gramian.jl:78-87 (Julia is not installed; "
                                     "the reference binary cannot run here). ms_per_step is extrapolated to all n rows."},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------------------------------------ our arm
    import torch

    import covfn_b200 as cf

    if not torch.cuda.is_available() or cf.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    X, a_host = make_inputs(w)
    blk = d if w["gradient"] else 1
    k = cf.GradientKernel(w["kernel"]) if w["gradient"] else w["kernel"]
    r0, r1 = n * rank // world, n * (rank + 1) // world
    G = cf.gramian(k, X.T).set_row_range(r0, r1)
    G.handle()
    nrhs = w["nrhs"]
    a_dev = torch.from_numpy(np.ascontiguousarray(a_host.T if nrhs > 1 else a_host)).to(dev)  # column-major m x nrhs
    b_full = torch.empty((nrhs, n * blk) if nrhs > 1 else (n * blk,), dtype=torch.float64, device=dev)
    b_loc = torch.empty((nrhs, (r1 - r0) * blk) if nrhs > 1 else ((r1 - r0) * blk,), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        flush.zero_()
        G.mul_device(b_loc.data_ptr(), a_dev.data_ptr(), nrhs=nrhs, ldy=(r1 - r0) * blk, ldx=n * blk, stream=stream.cuda_stream)
        if dist is not None:
            if nrhs == 1:
                dist.all_gather_into_tensor(b_full, b_loc)
            else:
                parts = [torch.empty_like(b_loc) for _ in range(world)]
                dist.all_gather(parts, b_loc)
        return b_loc

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    sampler.stop_flag = True
    total_ms = e0.elapsed_time(e1)
    if dist is not None:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = pairs_per_step / (ms_per_step * 1e-3)

    # dominant kernel alone (CUDA events recorded by the library on the launch stream, around the kernel + its reduction)
    for _ in range(args.steps):
        G.mul_device(b_loc.data_ptr(), a_dev.data_ptr(), nrhs=nrhs, ldy=(r1 - r0) * blk, ldx=n * blk)  # library stream, blocking
        ms, launches = G.last_timing()
        kern_ms.append(ms)
    kernel_ms = float(np.mean(kern_ms))
    launches_per_step = launches

    # opt-in symmetric variant (each unordered pair evaluated once; NOT the headline: see include/covfn_b200.h CF_OPT_SYMMETRIC)
    sym = None
    if world == 1 and nrhs == 1 and not w["gradient"]:
        G.set_symmetric(True)
        ts = []
        for _ in range(1 + min(args.steps, 3)):
            G.mul_device(b_loc.data_ptr(), a_dev.data_ptr())
            ts.append(G.last_timing()[0])
        G.set_symmetric(False)
        sym_ms = float(np.mean(ts[1:]))
        sym = {"ms_per_step": sym_ms, "mvm_equivalent_pairs_per_s": pairs_per_step / (sym_ms * 1e-3),
               "evaluated_pairs_per_s": 0.5 * pairs_per_step / (sym_ms * 1e-3),
               "note": "cf_gramian_set_option(CF_OPT_SYMMETRIC): K = K^T, every unordered pair evaluated once and used for b_i and "
                       "b_j; column half accumulated with fp64 atomics (not bit-reproducible), off by default, not used for `value`"}

    # end to end through the public host API, pinned host buffers, X uploaded every step
    a_pin = torch.from_numpy(np.ascontiguousarray(a_host.T if nrhs > 1 else a_host)).pin_memory()
    b_pin = torch.empty_like(b_loc, device="cpu").pin_memory()
    a_np = a_pin.numpy().T if nrhs > 1 else a_pin.numpy()
    b_np = b_pin.numpy().T if nrhs > 1 else b_pin.numpy()
    XT = X.T  # d x n, column-major (columns are points): the layout of a Julia Matrix passed to gramian(k, X)

    def step_e2e():
        t_a = time.perf_counter()
        Ge = cf.gramian(k, XT).set_row_range(r0, r1)  # create: uploads X (reference: gramian(k, x) is O(1) lazy)
        Ge.handle()
        t_b = time.perf_counter()
        cf.mul_(b_np, Ge, a_np)
        t_c = time.perf_counter()
        Ge.close()
        if os.environ.get("CF_BENCH_DEBUG"):
            print(f"[e2e] create {1e3 * (t_b - t_a):.1f} ms, mul_ {1e3 * (t_c - t_b):.1f} ms, close {1e3 * (time.perf_counter() - t_c):.1f} ms",
                  file=sys.stderr)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 3))
    for _ in range(e2e_steps):
        step_e2e()
    t_loop = time.perf_counter()
    barrier()
    if os.environ.get("CF_BENCH_DEBUG"):
        print(f"[e2e] loop {1e3 * (t_loop - t0):.1f} ms, barrier {1e3 * (time.perf_counter() - t_loop):.1f} ms", file=sys.stderr)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = pairs_per_step / e2e_s
    h2d = X.nbytes + a_host.nbytes
    d2h = b_loc.numel() * 8

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel against the DFMA peak measured now
    peak_lane_ops, _ = cf.peak_probe("dfma", 1 << 15)
    slots = SLOTS[args.config]
    my_pairs = float(r1 - r0) * float(n)
    achieved_tflops = 2.0 * slots * my_pairs / (kernel_ms * 1e-3) / 1e12
    peak_tflops = 2.0 * peak_lane_ops / 1e12
    alg_bytes = X.nbytes + a_host.nbytes + (r1 - r0) * blk * 8 * nrhs
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {
        "bound": "fp64_fma_pipe", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved_tflops / peak_tflops,
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one ncu --set full capture at this workload
        # (profiles/r1_ncu_mvm_eq_c2_n1048576.md): 33.90 MB + 0.76 MB per launch; only known for the 1-GPU c2 shape
        "traffic": 34653184 if (args.config == "c2" and world == 1) else None,
        "kernel_ms": kernel_ms, "flops_per_pair": 2 * slots,
        # `frac` uses SURVEY.md 8d's fixed reference instruction sequence (exp = 16 slots, distance = 2 d), so a kernel with a
        # cheaper exp (9 FP64 instructions here) or a tensor-core distance can exceed 1; the counter-level view is in profiles/
        "note": "reference-slot accounting (implementation independent); executed FP64 instructions per pair for c2: 15 of the "
                "23 slots -> the FP64 pipe itself is at frac * 15 / 23 of the probe peak (ncu: 78.8 % pipe-active)",
        "fp64_instr_frac": (achieved_tflops / peak_tflops) * 15.0 / 23.0 if args.config == "c2" else None,
        "peak_source": "cf_peak_probe DFMA microbenchmark in this run (MEASURED_PEAKS.json has no FP64 entry); "
                       "nominal 64 DFMA/clk/SM x 148 SMs x 1.965 GHz = 37.2 TFLOP/s",
        "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"},
    }
    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg, "clocks": sampler.result(),
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3,
                "note": "gramian(k, X) handle creation (packs and uploads X) + mul_(b, G, a) with pinned host a, b, every step"},
        "gpu_launches": int(launches_per_step) * args.steps,
        "roofline": roofline,
        "pct_of_fp64_peak": 100.0 * achieved_tflops / peak_tflops,
    }
    if sym is not None:
        out["symmetric_variant"] = sym
    if world == 1 and not args.no_cpu_baseline:
        rate, rows, secs, nt = cpu_port_rate(w, X, a_host, target_s=12.0)
        out["cpu_baseline"] = {
            "value": rate, "unit": unit, "cores": nt, "kind": "port",
            "sample": f"rows 0..{rows} of the same n={n} MVM ({rows * float(n):.3g} pairs, {secs:.1f} s)",
            "note": "C/OpenMP restatement of reference src/gramian.jl:78-87 (oracle/); Julia absent, reference binary not runnable; "
                    "README.md:37-38 publishes 4.59e8 pairs/s for MaternP(2), d=3, n=16384 on unstated hardware",
        }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
