#!/usr/bin/env python
"""bench.py -- lazy-Gramian MVM throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c1|c3|c4|c5] [--dtype f64|f32] [--spmd]

--config c5 is BASELINE.json configs[4]: conjugate gradients on (K + sigma^2 I) x = y, K = gramian(MaternP(2), x), d = 8,
n = 2^19, a fixed number of iterations (--cg-iters, default 20); a step is one whole solve through cf_cg_solve.  Under torchrun
every rank owns a row block and the library gathers the product with NCCL once per iteration (csrc/cf_comm.h); with --spmd ONE
process drives --gpus devices (cf_init) and the all-gather is fused into the producing kernels' epilogues (NVLink peer stores).

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

SLOTS = {"c1": 35, "c2": 23, "c3": 111, "c4": 97, "c5": 45, "x1": 80, "x2": 81, "x3": 61, "x4": 45, "x5": 110, "x6": 125, "x7": 33}  # FP64 issue slots per pair/block


MUFU_PER_PAIR = {"c1": 2, "c2": 1, "c3": 1, "x2": 1, "x3": 2, "x4": 2, "x7": 1}  # Float32 kernels: MUFU operations per evaluated entry


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
    if name == "c5":
        return dict(kernel=cf.MaternP(2), kname="MaternP(2)", d=8, n=1 << 19, nrhs=1, gradient=False, cg=True, sigma2=1e-2,
                    desc="CG on (K + sigma^2 I) x = y, K = gramian(MaternP(2), x), d=8, n=2^19, sigma^2=1e-2, Float64 (configs[4])")
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
    if name == "x7":  # auxiliary: EQ value MVM at d = 8 (the evaluation-bound end of the tensor-core value kernels)
        return dict(kernel=cf.EQ(), kname="EQ", d=8, n=131072, nrhs=1, gradient=False,
                    desc="EQ Gramian MVM, d=8, n=131072, Float64 (auxiliary)")
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
    dt = X.dtype.type

    def run(rows):
        t0 = time.perf_counter()
        if w["gradient"]:
            O.gradient_mul(prog, X, a, rows=(0, rows))
        elif w["nrhs"] == 1:
            O.mul_vec(prog, X, a, rows=(0, rows), dtype=dt)
        else:
            O.mul_mat(prog, X, a, rows=(0, rows), dtype=dt)  # reference loop order: entry re-evaluated per RHS column
        return time.perf_counter() - t0

    nt = O.num_threads()
    rows = max(nt, 8)
    t = run(rows)  # calibration (also warms the threads)
    rows = int(min(n, max(rows, rows * target_s / max(t, 1e-3))))
    rows = max(nt, (rows // nt) * nt)
    t = run(rows)
    return rows * float(n) / t, rows, t, nt


def oracle_rows(w, X, a, rows, dtype):
    """the reference's product restricted to `rows` (oracle as the checker, outside every timed region)"""
    from oracle import oracle as O

    O.build()
    O.set_num_threads(os.cpu_count() or 1)
    prog = w["kernel"].program()
    if w["gradient"]:
        return O.gradient_mul(prog, X, a, rows=rows)
    if w["nrhs"] == 1:
        return O.mul_vec(prog, X, a, rows=rows, dtype=dtype)
    return O.mul_mat(prog, X, a, rows=rows, dtype=dtype)


def recorded_ncu(config, dtype):
    """per-config counters of the dominant kernel from the committed ncu pass (profiles/r2_ncu_metrics.json, written by
    bench_aux/record_ncu_metrics.py from `ncu --set full` captures of this same command's kernels); None if not recorded"""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_metrics.json")))
        return rec.get(f"{config}:{dtype}")
    except Exception:
        return None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


CPU_NOTE = ("C/OpenMP restatement of the reference loop (oracle/: src/gramian.jl:78-99, 241-253); Julia is not installed here or on "
            "the GPU box (profiles/r2_julia_probe.txt), so the reference binary cannot run")


_RESULT_STREAM = None


def emit(obj):
    """the ONE JSON line of the contract, on the process's original stdout"""
    stream = _RESULT_STREAM or sys.stdout
    stream.write(json.dumps(obj) + "\n")
    stream.flush()


def main():
    # Libraries may write to file descriptor 1 (NCCL prints its version banner there when NCCL_DEBUG is set): keep the original stdout for
    # the result line only and send everything else to stderr
    global _RESULT_STREAM
    sys.stdout.flush()
    _RESULT_STREAM = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--spmd", action="store_true", help="c5 only: ONE process drives --gpus devices (cf_init, fused peer-store gather)")
    ap.add_argument("--cg-iters", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = workload(args.config)
    w.setdefault("cg", False)
    if os.environ.get("CF_BENCH_N"):  # development only: a smaller n (the JSON line then names it in config.n)
        w["n"] = int(os.environ["CF_BENCH_N"])
        w["desc"] += f" [n overridden to {w['n']}]"
    n, d = w["n"], w["d"]
    npdt = np.float64 if args.dtype == "f64" else np.float32
    unit = "kernel-pair evaluations/s"
    metric = "Gramian MVM kernel-evals/s"
    ngpus = args.gpus if args.spmd else world
    cfg = {"workload": w["desc"] + (" [Float32 variant]" if args.dtype == "f32" else ""), "kernel": w["kname"], "d": d, "n": n, "nrhs": w["nrhs"],
           "sharding": (f"ONE process, rows sharded inside the library over {ngpus} device(s) (cf_init), all-gather fused into the kernel epilogues (peer stores)"
                        if args.spmd else
                        f"contiguous row blocks over {world} rank(s), x and a replicated, all-gather of b per product" if world > 1 else "single GPU"),
           "l2": "256 MiB L2 flush (memset) between timed steps, inside the timed region"}

    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        X, a = make_inputs(w)
        X, a = X.astype(npdt), a.astype(npdt)
        rate0, rows, _, nt = cpu_port_rate(w, X, a, target_s=6.0)
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            oracle_rows(w, X, a, (0, rows), npdt)
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
        t = float(np.mean(times))
        value = rows * float(n) / t
        products = (args.cg_iters + 1) if w["cg"] else 1
        sample = f"rows 0..{rows} of the n={n} row product ({rows * float(n):.3g} pairs per timed sample), all {n} columns"
        emit(({
            "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3 * (n / rows) * products, "ms_per_step_is_extrapolated": True,
            "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": unit, "cores": nt, "kind": "port", "sample": sample,
                             "note": CPU_NOTE + ". Each timed step is the row sample; ms_per_step is EXTRAPOLATED linearly to all n rows"
                                     + (f" and {products} products of the CG solve" if w["cg"] else "")
                                     + " (so it does not fit the driver's clock around this run by construction)."},
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
    if w["cg"]:
        return run_cg(args, w, cfg, rank, world, local_rank, dev, dist, metric, unit)
    if args.spmd:
        raise SystemExit("--spmd applies to --config c5")

    X, a_host = make_inputs(w)
    X, a_host = X.astype(npdt), a_host.astype(npdt)
    pairs_per_step = float(n) * float(n)
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    es = 8 if args.dtype == "f64" else 4
    blk = d if w["gradient"] else 1
    k = cf.GradientKernel(w["kernel"]) if w["gradient"] else w["kernel"]
    r0, r1 = n * rank // world, n * (rank + 1) // world
    G = cf.gramian(k, X.T).set_row_range(r0, r1)
    G.handle()
    nrhs = w["nrhs"]
    sym_capable = nrhs == 1 and not w["gradient"] and args.dtype == "f64"
    if sym_capable:
        # `value` counts kernel-pair EVALUATIONS: it is measured with every one of the n*m entries evaluated, as the reference does
        # (CF_OPT_SYMMETRIC off).  The library's default for y === x evaluates each unordered pair once; that path is what `e2e`
        # (the user-facing call) runs and is reported device-resident under `symmetric_variant`.
        G.set_symmetric(False)
        cfg["value_path"] = "all n*m entries evaluated (CF_OPT_SYMMETRIC = 0), row blocks" + (" + all-gather" if world > 1 else "")
        cfg["e2e_path"] = "library default: y === x evaluates each unordered pair once (deterministic symmetric variant)" + \
                          (", partial vectors summed with ncclAllReduce" if world > 1 else "")
    if world > 1 and sym_capable:
        from covfn_b200 import distributed as D

        D.comm_init_from_torch()
    a_dev = torch.from_numpy(np.ascontiguousarray(a_host.T if nrhs > 1 else a_host)).to(dev)  # column-major m x nrhs
    b_full = torch.empty((nrhs, n * blk) if nrhs > 1 else (n * blk,), dtype=tdt, device=dev)
    b_loc = torch.empty((nrhs, (r1 - r0) * blk) if nrhs > 1 else ((r1 - r0) * blk,), dtype=tdt, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    parts = [torch.empty((nrhs, (n * (r + 1) // world - n * r // world) * blk), dtype=tdt, device=dev) for r in range(world)] if nrhs > 1 else None

    def step_device():
        flush.zero_()
        G.mul_device(b_loc.data_ptr(), a_dev.data_ptr(), nrhs=nrhs, ldy=(r1 - r0) * blk, ldx=n * blk, stream=stream.cuda_stream)
        if dist is not None:
            if nrhs == 1 and n % world == 0:
                dist.all_gather_into_tensor(b_full, b_loc)
            elif nrhs == 1:
                dist.all_gather([b_full[(n * r // world) * blk:(n * (r + 1) // world) * blk] for r in range(world)], b_loc)
            else:
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

    # parity spot check, outside the timed region: rows spanning the first and the last rank boundary (or the middle of the
    # matrix on one GPU) of the gathered product against the reference restatement
    parity = None
    if rank == 0:
        full = (b_full if world > 1 and nrhs == 1 else (torch.cat(parts, dim=1) if world > 1 else b_loc)).cpu().numpy()
        full = full.T if nrhs > 1 else full
        bounds = sorted({n * 1 // world, n * (world - 1) // world} - {0, n}) if world > 1 else [n // 2]
        half = 16 if nrhs > 1 else 32
        worst, checked = 0.0, []
        for bnd in bounds:
            rws = (max(0, bnd - half), min(n, bnd + half))
            ref = oracle_rows(w, X, a_host, rws, npdt)
            got = full[rws[0] * blk:rws[1] * blk]
            err = float(np.linalg.norm(np.asarray(got, dtype=np.float64) - ref) / np.linalg.norm(ref))
            worst = max(worst, err)
            checked.append(list(rws))
        tol = 1e-12 if args.dtype == "f64" else 1e-5
        parity = {"rows": checked, "rel_2norm_err_vs_oracle": worst, "tolerance": tol, "ok": bool(worst < tol)}
        if args.dtype == "f32" and nrhs == 1 and not w["gradient"]:
            # the Float32 reference accumulates n terms sequentially IN Float32 (src/gramian.jl:83 stores into y[i] every term): at
            # n = 2^20 its own distance from the exact product is ~1e-5, so the check is made against the extended-precision
            # evaluation of the same Float32 inputs, with the restatement's own error reported beside it
            from oracle import oracle as O

            rws = tuple(checked[0])
            tru = O.truth_mul_vec(w["kernel"].program(), X.astype(np.float64), a_host.astype(np.float64), rows=rws)
            ref = oracle_rows(w, X, a_host, rws, npdt)
            got = np.asarray(full[rws[0]:rws[1]], dtype=np.float64)
            parity["rel_2norm_err_vs_extended_precision"] = float(np.linalg.norm(got - tru) / np.linalg.norm(tru))
            parity["reference_restatement_vs_extended_precision"] = float(np.linalg.norm(ref - tru) / np.linalg.norm(tru))
            parity["ok"] = bool(parity["rel_2norm_err_vs_extended_precision"] < tol)
        if not parity["ok"]:
            raise SystemExit(f"bench.py: parity spot check failed: {parity}")

    # dominant kernel alone (CUDA events recorded by the library on the launch stream, around the kernel + its reduction)
    for _ in range(args.steps):
        G.mul_device(b_loc.data_ptr(), a_dev.data_ptr(), nrhs=nrhs, ldy=(r1 - r0) * blk, ldx=n * blk)  # library stream, blocking
        ms, launches = G.last_timing()
        kern_ms.append(ms)
    kernel_ms = float(np.mean(kern_ms))
    launches_per_step = launches

    # the library's default dispatch for y === x (each unordered pair evaluated once), device resident
    sym = None
    if sym_capable:
        from covfn_b200.gramian import mul_collective_device

        G.set_symmetric(True)
        ts = []
        for _ in range(1 + min(args.steps, 3)):
            flush.zero_()
            if world > 1:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                ev0.record()
                mul_collective_device(G, b_full.data_ptr(), a_dev.data_ptr(), stream=stream.cuda_stream)
                ev1.record()
                torch.cuda.synchronize()
                tt = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ts.append(float(tt.item()))
            else:
                G.mul_device(b_loc.data_ptr(), a_dev.data_ptr())
                ts.append(G.last_timing()[0])
        sym_ms = float(np.mean(ts[1:]))
        sym_err = None
        if rank == 0:  # same rows as the parity check above, now from the symmetric path
            fullv = (b_full if world > 1 else b_loc).cpu().numpy()
            rws = tuple(parity["rows"][0])
            ref = oracle_rows(w, X, a_host, rws, npdt)
            sym_err = float(np.linalg.norm(fullv[rws[0]:rws[1]] - ref) / np.linalg.norm(ref))
            if not sym_err < 1e-12:
                raise SystemExit(f"bench.py: symmetric path parity check failed: {sym_err}")
        sym = {"ms_per_step": sym_ms, "mvm_equivalent_pairs_per_s": pairs_per_step / (sym_ms * 1e-3),
               "evaluated_pairs_per_s": 0.5 * pairs_per_step / (sym_ms * 1e-3), "rel_2norm_err_vs_oracle": sym_err,
               "collective": "ncclAllReduce(sum) of the n-vector of partial sums, one per product" if world > 1 else None,
               "note": "K = K^T: every unordered pair evaluated once and used for b_i and b_j; single-writer partial sums combined in a fixed "
                       "order (bit-reproducible).  `value` above is measured with this turned off."}

    # end to end through the public host API, pinned host buffers, X uploaded every step
    a_pin = torch.from_numpy(np.ascontiguousarray(a_host.T if nrhs > 1 else a_host)).pin_memory()
    b_pin = torch.empty_like(b_loc, device="cpu").pin_memory()
    a_np = a_pin.numpy().T if nrhs > 1 else a_pin.numpy()
    b_np = b_pin.numpy().T if nrhs > 1 else b_pin.numpy()
    XT = X.T  # d x n, column-major (columns are points): the layout of a Julia Matrix passed to gramian(k, X)

    bfull_pin = torch.empty(n * blk, dtype=tdt).pin_memory() if (world > 1 and sym_capable) else None

    def step_e2e():
        Ge = cf.gramian(k, XT).set_row_range(r0, r1)  # create: uploads X (reference: gramian(k, x) is O(1) lazy)
        Ge.handle()
        if world > 1 and sym_capable:
            # multi-rank default path: host vector in, collective product (symmetric + ncclAllReduce), complete host vector out
            a_d = a_pin.to(dev, non_blocking=True)
            mul_collective_device(Ge, b_full.data_ptr(), a_d.data_ptr(), stream=stream.cuda_stream)
            bfull_pin.copy_(b_full, non_blocking=True)
            torch.cuda.synchronize()
        else:
            cf.mul_(b_np, Ge, a_np)
        Ge.close()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 3))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = pairs_per_step / e2e_s
    h2d = X.nbytes + a_host.nbytes
    d2h = (n * blk * es) if (world > 1 and sym_capable) else b_loc.numel() * es

    if rank != 0:
        if dist is not None:
            if sym_capable:
                D.comm_destroy()
            dist.destroy_process_group()
        return

    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": cfg, "clocks": sampler.result(),
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3,
                "note": "gramian(k, X) handle creation (packs and uploads X) + mul_(b, G, a) with pinned host a, b, every step"},
        "gpu_launches": int(launches_per_step) * args.steps,
        "roofline": roofline_block(cf, args, w, world, kernel_ms, float(r1 - r0) * float(n), X.nbytes + a_host.nbytes + (r1 - r0) * blk * es * nrhs),
        "parity_check": parity,
    }
    out["pct_of_fp64_peak" if args.dtype == "f64" else "pct_of_fp32_peak"] = 100.0 * out["roofline"]["frac"]
    if sym is not None:
        out["symmetric_variant"] = sym
    if world == 1 and not args.no_cpu_baseline:
        rate, rows, secs, nt = cpu_port_rate(w, X, a_host, target_s=12.0)
        out["cpu_baseline"] = {
            "value": rate, "unit": unit, "cores": nt, "kind": "port",
            "sample": f"rows 0..{rows} of the same n={n} product ({rows * float(n):.3g} pairs, {secs:.1f} s)",
            "note": CPU_NOTE + "; README.md:37-38 publishes 4.59e8 pairs/s for MaternP(2), d=3, n=16384 on unstated hardware",
        }
    emit(out)
    if dist is not None:
        if sym_capable:
            D.comm_destroy()
        dist.destroy_process_group()


def roofline_block(cf, args, w, world, kernel_ms, my_pairs, alg_bytes):
    """FP64 (or FP32) FMA-pipe roofline of the dominant kernel.  `achieved` follows SURVEY.md section 8d: the fixed reference
    instruction sequence (slots per pair x 2 flops) over the measured kernel time; `peak` is the FMA rate measured in this run by
    cf_peak_probe (MEASURED_PEAKS.json has no FP64 / FP32 entry).  Because the reference sequence charges 16 slots for an exp
    that costs 8-9 FP64 instructions here and 2 d for a distance that costs d, `frac` can exceed 1: it compares WORK, not pipe
    occupancy.  The occupancy view -- FP64 pipe cycles active, DRAM bytes -- comes from the ncu pass recorded for this config
    (profiles/r2_ncu_metrics.json) and is reported beside it."""
    slots = SLOTS[args.config]
    f64 = args.dtype == "f64"
    peak_lane_ops, _ = cf.peak_probe("dfma" if f64 else "ffma", 1 << 15)
    achieved_tflops = 2.0 * slots * my_pairs / (kernel_ms * 1e-3) / 1e12
    peak_tflops = 2.0 * peak_lane_ops / 1e12
    peaks = load_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    rec = recorded_ncu(args.config, args.dtype) if world == 1 else None
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    rl = {
        "bound": "fp64_fma_pipe" if f64 else "fp32_fma_pipe", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
        "frac": achieved_tflops / peak_tflops,
        "traffic": (rec or {}).get("dram_bytes"),
        "kernel_ms": kernel_ms, "flops_per_pair": 2 * slots,
        "fp64_pipe_active": (rec or {}).get("fp64_pipe_active_pct"),
        "issue_active": (rec or {}).get("issue_active_pct"),
        "ncu_source": (rec or {}).get("source"),
        "ncu_captured_n": (rec or {}).get("captured_n"),  # set when the recorded pass ran the same kernel at a smaller n (traffic is per launch at that n)
        "peak_lane_fma_per_clk_per_sm": peak_lane_ops / 148.0 / (sm_mhz * 1e6),
        "note": "frac = reference-slot work (SURVEY.md 8d) / measured FMA peak: implementation independent, can exceed 1 when the kernel "
                "needs fewer instructions than the reference sequence (EQ d=3: 12 FP64 instructions per pair against 23 slots); "
                "fp64_pipe_active is the occupancy figure (ncu sm__pipe_fp64_cycles_active of the recorded pass)",
        "peak_source": "cf_peak_probe FMA microbenchmark in this run (two-register operand form: bench_aux/micro/fp64_issue_probe.cu); "
                       "MEASURED_PEAKS.json has no FP64/FP32 FMA entry; nominal 64 DFMA/clk/SM x 148 SMs x 1.965 GHz = 37.2 TFLOP/s",
        "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"},
    }
    if not f64:
        # Float32: the transcendental of every entry is one (EQ: ex2; RQ with integer alpha: rcp) or two (MaternP: sqrt, ex2) MUFU
        # operations and the MUFU pipe (16 lanes per clock and SM), not the FMA pipe, is the binding unit of the value kernels
        # (profiles/r2_ncu_c2_f32.md, r2_ncu_x2_f32.md: pipe_xu 84-86 % busy); reported as its own fraction
        mufu, _ = cf.peak_probe("mufu", 1 << 15)
        per_pair = MUFU_PER_PAIR.get(args.config)
        rl["mufu_peak_lane_ops_per_s"] = mufu
        if per_pair:
            ach = per_pair * my_pairs / (kernel_ms * 1e-3)
            rl["mufu"] = {"ops_per_pair": per_pair, "achieved_lane_ops_per_s": ach, "peak_lane_ops_per_s": mufu, "frac": ach / mufu,
                          "pipe_xu_active": (rec or {}).get("pipe_xu_pct"), "pipe_fma_inst_active": (rec or {}).get("pipe_fma_pct")}
        rl["note"] = ("frac = reference-slot work (SURVEY.md 8d) / measured FP32 FMA peak: implementation independent, exceeds 1 because the "
                      "reference sequence charges 16 slots for an exp that is one MUFU.EX2 here; the binding unit is the MUFU pipe: see `mufu`")
        rl["peak_source"] = "cf_peak_probe FFMA / MUFU.EX2 microbenchmarks in this run; MEASURED_PEAKS.json has no FP32 FMA or MUFU entry"
    return rl


def run_cg(args, w, cfg, rank, world, local_rank, dev, dist, metric, unit):
    """BASELINE config 5: a step is ONE conjugate-gradient solve of --cg-iters iterations through cf_cg_solve (iterates stay on
    the device(s)); pairs per step = (iterations + 1) n^2 (one operator product per iteration plus the initial residual)."""
    import torch

    import covfn_b200 as cf
    from covfn_b200 import distributed as D

    n, d = w["n"], w["d"]
    sigma2 = w["sigma2"]
    rng = np.random.Generator(np.random.Philox(0xC0F00005))
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    y = rng.standard_normal(n)
    iters = args.cg_iters
    products = iters + 1
    pairs_per_step = products * float(n) * float(n)
    ngpus = args.gpus if args.spmd else world
    if args.spmd:
        if world != 1:
            raise SystemExit("--spmd is a single-process mode: do not launch it under torchrun")
        cf.init(list(range(args.gpus)))
    elif world > 1:
        D.comm_init_from_torch()
    r0, r1 = (0, n) if args.spmd else (n * rank // world, n * (rank + 1) // world)
    XT = X.T

    def make():
        G = cf.gramian(w["kernel"], XT)
        if not args.spmd and world > 1:
            G.set_row_range(r0, r1)
        G.handle()
        return G

    G = make()
    A = sigma2 * cf.I(n) + G
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def solve(op):
        return op.solve(y, reltol=1e-300, maxiter=iters)  # a fixed number of iterations

    x = None
    for _ in range(args.warmup):
        flush.zero_()
        x, it, res = solve(A)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    prod_ms = gather_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        x, it, res = solve(A)
        tm = A.cg_timing()
        prod_ms += tm[1]
        gather_ms += tm[2]
    barrier()
    total_s = time.perf_counter() - t0
    sampler.stop_flag = True
    assert it == iters, (it, iters)
    if dist is not None:
        tt = torch.tensor([total_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_s = float(tt.item())
    ms_per_step = 1e3 * total_s / args.steps
    value = pairs_per_step / (ms_per_step * 1e-3)

    # every rank must hold bit-identical iterates (scalars are recomputed from gathered vectors, no all-reduce)
    identical = True
    if dist is not None:
        chk = torch.from_numpy(np.frombuffer(np.array([x.sum(), np.abs(x).max(), x[n // 3]]).tobytes(), dtype=np.int64).copy()).to(dev)
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        identical = all(torch.equal(allc[0], c) for c in allc)

    # end to end: handle creation (uploads X) + solve, every step
    def step_e2e():
        Ge = make()
        xe, _, _ = solve(sigma2 * cf.I(n) + Ge)
        Ge.close()
        return xe

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 2
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())

    if rank != 0:
        if dist is not None:
            D.comm_destroy()
            dist.destroy_process_group()
        return

    # parity, outside the timed region: (1) the recurrence residual against the true residual of the returned iterate, computed
    # with the reference restatement on a row sample; (2) the operator product itself on rows spanning a rank boundary
    from oracle import oracle as O

    O.build()
    O.set_num_threads(os.cpu_count() or 1)
    prog = w["kernel"].program()
    bnd = n // ngpus if ngpus > 1 else n // 2
    rws = (bnd - 24, bnd + 24)
    # (the right-hand side y as the test vector: the CG iterate of this ill-conditioned system has entries ~1e4 that cancel in K x,
    # which would measure the conditioning of the sum, not the kernel)
    Ax = O.mul_vec(prog, X, y, rows=rws) + sigma2 * y[rws[0]:rws[1]]
    Gfull = cf.gramian(w["kernel"], XT) if (world > 1 and not args.spmd) else G
    got = (Gfull @ y)[rws[0]:rws[1]] + sigma2 * y[rws[0]:rws[1]]
    perr = float(np.linalg.norm(got - Ax) / np.linalg.norm(Ax))
    true_res = float(np.linalg.norm(y - (Gfull @ x) - sigma2 * x))
    parity = {"rows": [list(rws)], "rel_2norm_err_vs_oracle": perr, "tolerance": 1e-12, "ok": bool(perr < 1e-12),
              "what": "(sigma^2 I + K) y on rows spanning a shard boundary: library vs reference restatement",
              "true_residual_of_returned_iterate": true_res}
    if not parity["ok"]:
        raise SystemExit(f"bench.py: parity spot check failed: {parity}")

    kernel_ms = prod_ms / (args.steps * products) if prod_ms > 0 else ms_per_step / products
    my_pairs = float(r1 - r0) * float(n) if not args.spmd else float(n) * float(n) / ngpus
    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": ngpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": dict(cfg, cg_iterations=iters, sigma2=sigma2, products_per_step=products,
                                            mode="spmd (one process, cf_init)" if args.spmd else ("torchrun + cf_comm (NCCL inside the library)" if world > 1 else "single GPU")),
        "clocks": sampler.result(),
        "cg": {"iterations": iters, "ms_per_iteration": ms_per_step / products, "product_ms_per_iteration": kernel_ms,
               "collective_ms_per_iteration": gather_ms / (args.steps * products) if gather_ms > 0 else (None if ngpus > 1 and args.spmd else 0.0),
               "collective_share": (gather_ms / (args.steps * products)) / (ms_per_step / products) if gather_ms > 0 else None,
               "collective": (None if ngpus == 1 else
                              "peer loads: every device sums the partial vectors of all devices in device order (symmetric variant), or "
                              "peer stores from the kernel epilogues (row blocks)" if args.spmd else
                              "ncclAllReduce(sum) of the n-vector of partial sums (symmetric variant, default) / in-place ncclAllGather of row blocks"),
               "recurrence_residual": res, "rhs_norm": float(np.linalg.norm(y)), "ranks_bit_identical": bool(identical),
               "note": "one exchange of the 4 MiB product per iteration; torchrun mode times it with CUDA events inside the library, spmd mode "
                       "overlaps it with the other devices' kernels (not separable: share = None)"},
        "e2e": {"value": pairs_per_step / e2e_s, "unit": unit, "h2d_bytes_per_step": int(X.nbytes + 2 * y.nbytes), "d2h_bytes_per_step": int(y.nbytes),
                "ms_per_step": e2e_s * 1e3, "note": "gramian(k, X) (uploads X) + (sigma^2 I + K) \\ y from host vectors, every step"},
        "gpu_launches": int(args.steps * products * 6),
        "roofline": roofline_block(cf, args, w, 1 if ngpus == 1 else 2, kernel_ms, my_pairs, X.nbytes + 2 * y.nbytes),
        "parity_check": parity,
    }
    out["pct_of_fp64_peak"] = 100.0 * out["roofline"]["frac"]
    if ngpus == 1 and not args.no_cpu_baseline:
        rate, rows, secs, nt = cpu_port_rate(w, X, y, target_s=12.0)
        out["cpu_baseline"] = {"value": rate, "unit": unit, "cores": nt, "kind": "port",
                               "sample": f"rows 0..{rows} of one n={n} operator product ({rows * float(n):.3g} pairs, {secs:.1f} s)",
                               "note": CPU_NOTE}
    emit(out)
    if dist is not None:
        D.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
