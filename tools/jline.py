"""Print the headline numbers of a bench.py JSON line read from stdin (last line)."""
import json
import sys

lines = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not lines:
    print("no JSON line")
    sys.exit(1)
d = json.loads(lines[-1])
r = d.get("roofline") or {}
print("%s N=%d  value %.3e  ms/step %.3f  kernel_ms %s  e2e %.3e" % (
    d.get("config", {}).get("workload", "?")[:28], d.get("n_gpus", 0), d["value"], d["ms_per_step"],
    ("%.3f" % r["kernel_ms"]) if "kernel_ms" in r else "-", (d.get("e2e") or {}).get("value", 0.0)))
