"""Summarise an .ncu-rep (ncu --set full) into a small markdown file for profiles/.
    python bench_aux/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_xxx.md "title" """
import csv
import io
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
with open(out, "w") as f:
    f.write(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none), one section per captured launch\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write(f"\n## {d.get('Kernel Name', '')[:110]}\n\ngrid {d.get('Grid Size')} block {d.get('Block Size')}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in d and d[k] != "":
                f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
        f.write("\nwarp stall reasons (warps stalled per issued instruction, > 0.05):\n\n")
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[h])
                except ValueError:
                    continue
                if v > 0.05:
                    name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
                    f.write(f"- {name}: {v:.3f}\n")
print("wrote", out)
