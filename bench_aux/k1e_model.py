"""Offline evaluation of gram_mvm_eq.cuh variants with the measured issue model (bench_aux/sass_cost.py): compiles
bench_aux/micro/k1e_variants.cu with the given macro flags to SASS (no GPU needed), finds the innermost loop of the kernel and prints its
modelled cycles per pair.     python bench_aux/k1e_model.py [-DVR=8 -DVNT=128 -DCF_EQ_UNROLL=2 ...]
The model predicted the shipped K1e to 1 % (31.1 cycles per pair -> 1.20e12 pairs/s; measured 1.21e12) and K1m to 6 %."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
flags = sys.argv[1:] or ["-DVR=8", "-DVNT=128"]
rows = 8
for f in flags:
    if f.startswith("-DVR="):
        rows = int(f[5:])
with tempfile.TemporaryDirectory() as td:
    cubin = os.path.join(td, "k.cubin")
    subprocess.check_call(["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "covariancefunctions.jl_b200", "csrc"),
                           "-Xcudafe", "--diag_suppress=177", "-cubin", "-o", cubin] + flags + [os.path.join(ROOT, "bench_aux", "micro", "k1e_variants.cu")])
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", cubin], capture_output=True, text=True).stdout
ins = []
infn = False
for l in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        infn = "gram_mvm_eq_kernel" in m.group(1)
        continue
    if not infn:
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
cands = []
for a, txt in ins:
    t = txt.split()
    op = t[1] if t[0].startswith("@") else t[0]
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < a:
            body = [x for x in ins if int(m.group(1), 16) <= x[0] <= a]
            nf = sum(1 for x in body if re.match(r"(@!?U?P\d )?(DFMA|DMUL|DADD)", x[1]))
            if nf >= 10 * rows:
                cands.append((len(body), nf, body))
cands.sort(key=lambda x: x[0])
ln, nf, body = cands[0]
path = os.path.join(tempfile.gettempdir(), "k1e_loop.txt")
open(path, "w").write("\n".join(x[1] + " ;" for x in body))
pairs = nf / 12.0  # 12 FP64 instructions per pair in every variant of this kernel
regs = re.findall(r"gram_mvm_eq_kernel.*?\n.*?REG:(\d+)", res)
print("flags", " ".join(flags), "| registers", regs[:1], "| loop", ln, "instructions,", nf, "FP64 =", pairs, "pairs")
print(subprocess.run([sys.executable, os.path.join(ROOT, "bench_aux", "sass_cost.py"), path, str(pairs)], capture_output=True, text=True).stdout.strip())
