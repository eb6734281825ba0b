"""Print the measured pipe peaks of the current GPU (DFMA / FFMA / MUFU lane-instructions per second)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import covfn_b200 as cf  # noqa: E402

out = {}
for kind in ("dfma", "ffma", "mufu"):
    ops, ms = cf.peak_probe(kind, 1 << 15)
    out[kind] = {"lane_ops_per_s": ops, "ms": ms, "per_clk_per_sm_at_1965MHz": ops / 148 / 1.965e9}
print(json.dumps(out))
