"""Build profiles/r2_ncu_metrics.json (read by bench.py: roofline.traffic / fp64_pipe_active / issue_active) and the per-capture
markdown summaries from the .ncu-rep files that bench_aux/record_ncu_metrics.py captured on the GPU box.
    python bench_aux/ncu_metrics_from_reps.py gpurun_out profiles"""
import csv
import io
import json
import os
import subprocess
import sys

src, dst = sys.argv[1], sys.argv[2]
KEYS = {"r2_ncu_c2_f64": "c2:f64", "r2_ncu_c2_f64_sym": "c2:f64:sym", "r2_ncu_c2_f64_k1": "c2:f64:k1", "r2_ncu_c1_f64": "c1:f64",
        "r2_ncu_c3_f64": "c3:f64", "r2_ncu_c4_f64": "c4:f64", "r2_ncu_c5_f64": "c5:f64", "r2_ncu_c2_f32": "c2:f32", "r2_ncu_c3_f32": "c3:f32",
        "r2_ncu_c3_f32_tc5": "c3:f32"}
path = os.path.join(dst, "r2_ncu_metrics.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for stem, key in KEYS.items():
    rep = os.path.join(src, stem + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))

    def num(name, scale=False):
        try:
            v = float(d[name].replace(",", ""))
        except Exception:
            return None
        if scale:
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u.get(name, ""), 1.0)
        return v

    rd, wr = num("dram__bytes_read.sum", True), num("dram__bytes_write.sum", True)
    out[key] = {
        "kernel": d.get("Kernel Name", "")[:120], "grid": d.get("Grid Size"), "duration_ms": num("gpu__time_duration.sum", True),
        "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "inst_executed": num("smsp__inst_executed.sum"), "registers_per_thread": num("launch__registers_per_thread"),
        "dram_bytes": (rd + wr) if rd is not None and wr is not None else None, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "source": f"profiles/{stem}.md",
    }
    md = os.path.join(src, stem + ".md")
    if os.path.exists(md):
        open(os.path.join(dst, stem + ".md"), "w").write(open(md).read().replace("gpurun_out/", "gpurun_out/ (scratch) "))
    print(key, json.dumps(out[key]))
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
