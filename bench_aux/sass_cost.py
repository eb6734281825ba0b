"""Estimate the issue cost of a SASS loop on a B200 sub-partition with the model measured by bench_aux/micro/fp64_issue_probe.cu:
an FP64 instruction costs max(2, number of DISTINCT 64-bit source REGISTERS actually fetched) cycles (operand-reuse hits, uniform
registers, constants and immediates are free; a register named in two slots is fetched once), every other instruction 1 cycle.
    python bench_aux/sass_cost.py loop.txt [pairs_per_iteration]"""
import re
import sys

lines = [l.split("/*")[0].strip() for l in open(sys.argv[1]) if l.strip()]
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cache = {}
fp = other = 0
cost = 0
hist = {}
for l in lines:
    t = l.replace(";", "").split(None, 1)
    if not t:
        continue
    if t[0].startswith("@"):
        t = t[1].split(None, 1)
    op = t[0]
    ops = [o.strip() for o in t[1].split(",")] if len(t) > 1 else []
    srcs = ops[1:]
    if op.split(".")[0] in ("DFMA", "DADD", "DMUL"):
        fetched = set()  # DISTINCT registers fetched: a register named in two operand slots is read once (fma(x, x, y): 2.02 cycles, measured)
        for slot, o in enumerate(srcs):
            m = re.match(r"[-|]*R(\d+)", o)
            if not m:
                continue
            reg = m.group(1)
            if cache.get(slot) == reg:
                pass  # reuse hit
            else:
                fetched.add(reg)
            if "reuse" in o:
                cache[slot] = reg
            else:
                cache.pop(slot, None)
        reads = len(fetched)
        c = max(2, reads)
        hist[c] = hist.get(c, 0) + 1
        cost += c
        fp += 1
    else:
        for slot, o in enumerate(srcs):
            if "reuse" in o:
                m = re.match(r"[-|~]*R(\d+)", o)
                if m:
                    cache[slot] = m.group(1)
            else:
                cache.pop(slot, None)
        other += 1
        cost += 1
print(f"{fp} FP64 ({hist}), {other} other; model cycles {cost} -> {cost / pairs:.2f} per pair; FP64 pipe active {2 * fp / cost:.1%}")
