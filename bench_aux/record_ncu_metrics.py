"""Record, per bench configuration, the ncu counters of the dominant kernel (one `ncu --set full --clock-control none` capture each)
and write profiles-ready artefacts: gpurun_out/r2_ncu_<tag>.ncu-rep, gpurun_out/r2_ncu_<tag>.md and gpurun_out/r2_ncu_metrics.json,
the file bench.py reads for `roofline.traffic`, `roofline.fp64_pipe_active` and `roofline.issue_active`.
    python bench_aux/record_ncu_metrics.py            (on the GPU box, one GPU)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

# key in r2_ncu_metrics.json -> (run_one arguments, kernel regex, environment)
RUNS = {
    "c2:f64": (["--config", "c2"], "gram_mvm_eq_kernel", {"COVFN_SYMMETRIC": "0"}),
    "c2:f64:sym": (["--config", "c2"], "gram_mvm_sym_kernel", {"COVFN_SYMMETRIC": "1"}),
    "c2:f64:k1": (["--config", "c2"], "gram_mvm_kernel", {"COVFN_SYMMETRIC": "0", "COVFN_MVM_SCALAR": "1"}),
    "c1:f64": (["--config", "c1"], "gram_mvm_eq_kernel", {}),  # K1m: the MaternP form of the scaled-domain kernel
    "c1:f64:k1": (["--config", "c1"], "gram_mvm_kernel", {"COVFN_MVM_SCALAR": "1"}),
    "c3:f64": (["--config", "c3"], "gram_mm_dmma_kernel", {}),
    "c4:f64": (["--config", "c4"], "grad_mvm_dmma_kernel", {}),
    "c5:f64": (["--config", "x4"], "gram_mvm_sym_kernel", {}),  # config 5's operator at n = 131072 (a sixteenth of the pairs: shorter capture)
    "c5:f64:k1": (["--config", "x4"], "gram_mvm_sym_kernel", {"COVFN_MVM_SCALAR": "1"}),
    "c2:f32": (["--config", "c2", "--dtype", "f32", "--n", "262144"], "gram_mvm_f32p_kernel", {}),  # (a sixteenth of the pairs: same rate, shorter capture)
    "c2:f32:k1": (["--config", "c2", "--dtype", "f32"], "gram_mvm_kernel", {"COVFN_MVM_SCALAR": "1"}),
    "x2:f32": (["--config", "x2", "--dtype", "f32"], "gram_mvm_tc5_kernel", {}),
    "x2:f32:legacy": (["--config", "x2", "--dtype", "f32"], "gram_mvm_tf32_kernel", {"COVFN_MVM_LEGACY": "1"}),
    "x3:f32": (["--config", "x3", "--dtype", "f32"], "gram_mvm_tc5_kernel", {}),
    "c3:f32": (["--config", "c3", "--dtype", "f32"], "gram_mm_tf32", {}),
}
only = sys.argv[1:]
metrics = {}
for key, (argv, kregex, env) in RUNS.items():
    if only and key not in only:
        continue
    tag = key.replace(":", "_")
    rep = os.path.join(OUT, f"r2_ncu_{tag}")
    e = dict(os.environ)
    e.update(env)
    cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{kregex}", "-c", "1", "-f", "-o", rep,
           sys.executable, os.path.join(ROOT, "bench_aux", "run_one.py"), "--reps", "1"] + argv
    r = subprocess.run(cmd, env=e, capture_output=True, text=True)
    if not os.path.exists(rep + ".ncu-rep"):
        print(key, "capture failed", r.stderr[-400:])
        continue
    raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))

    def num(name, scale_units=True):
        try:
            v = float(d[name].replace(",", ""))
        except Exception:
            return None
        if scale_units:
            un = u.get(name, "")
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(un, 1.0)
        return v

    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    dur = num("gpu__time_duration.sum", False)
    dur_ms = dur * {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u.get("gpu__time_duration.sum", "ms"), 1.0) if dur is not None else None
    metrics[key] = {
        "kernel": d.get("Kernel Name", "")[:120], "duration_ms": dur_ms,
        "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", False),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "pipe_xu_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", False),    # MUFU
        "pipe_fma_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", False),
        "pipe_alu_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", False),
        "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active", False),
        "inst_executed": num("smsp__inst_executed.sum", False),
        "registers_per_thread": num("launch__registers_per_thread", False),
        "dram_bytes": (rd + wr) if rd is not None and wr is not None else None,
        "dram_bytes_read": rd, "dram_bytes_write": wr,
        "source": f"profiles/r2_ncu_{tag}.md",
    }
    subprocess.run([sys.executable, os.path.join(ROOT, "bench_aux", "ncu_summary.py"), rep + ".ncu-rep", os.path.join(OUT, f"r2_ncu_{tag}.md"),
                    f"{key}: {kregex} ({' '.join(argv)}; env {env})"], capture_output=True)
    print(key, json.dumps(metrics[key]))
path = os.path.join(OUT, "r2_ncu_metrics.json")
old = {}
if os.path.exists(path):
    old = json.load(open(path))
old.update(metrics)
json.dump(old, open(path, "w"), indent=1)
print("wrote", path)
