"""Measures the integer-ALU roofline denominators on this box (hx_int_peak) and prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyperex_b200 as hx  # noqa: E402

NAMES = {0: "lop3", 1: "iadd3", 2: "shf", 3: "imad", 4: "lop3+imad 1:1", 5: "imad.hi", 6: "lop3+iadd3 1:1",
         7: "lop3+imad 2:1", 8: "imad.wide(+xor)", 9: "popc(+or)", 10: "myers_step_v0", 11: "myers_step_v1", 12: "myers_step_v2", 13: "myers_step_v3", 14: "myers_step_wide", 15: "myers_step_popc_score"}
out = {}
for v, name in NAMES.items():
    try:
        best = max(hx.int_peak(v, 3000) for _ in range(3))
        out[name] = round(best, 1)
    except Exception as e:
        out[name] = str(e)
print(json.dumps({"unit": "1e9 lane-ops/s (variants 10+: 1e9 word-steps/s)", "peaks": out}))
