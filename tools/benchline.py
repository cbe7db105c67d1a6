"""Compact view of a bench.py JSON line read from stdin: tools/benchline.py [label]"""
import json
import sys

label = sys.argv[1] if len(sys.argv) > 1 else ""
lines = [l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")]
if not lines:
    print(label, "NO JSON LINE")
    sys.exit(1)
d = json.loads(lines[-1])
e2e = d.get("e2e") or {}
print(label, "ms/step=%.4f" % d["ms_per_step"], "value=%.4g" % d["value"], "frac=%.4f" % d.get("roofline", {}).get("frac", 0),
      "steps(min/med/max)=%.4f/%.4f/%.4f" % tuple(d["step_ms_rank0"][k] for k in ("min", "median", "max")),
      "e2e_ms=%s" % (("%.1f" % e2e["ms_per_step"]) if e2e else "-"), "n_gpus=%d" % d["n_gpus"], "clocks=%s" % (d.get("clocks") or {}).get("sm_mhz"))
