"""Print the headline fields and the per-shape GEMM table of bench.py JSON lines (files given as arguments)."""
import json
import sys

for f in sys.argv[1:]:
    lines = [x for x in open(f) if x.startswith("{")]
    if not lines:
        print(f, "no JSON line")
        continue
    d = json.loads(lines[-1])
    print("==", f)
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
    print("e2e", d.get("e2e"))
    r = d.get("roofline", {})
    print({k: v for k, v in r.items() if k not in ("per_shape", "peak_note", "traffic_note", "kernel")})
    for k, v in r.get("per_shape", {}).items():
        print("  ", k, {a: round(b, 3) if isinstance(b, float) else b for a, b in v.items()})
