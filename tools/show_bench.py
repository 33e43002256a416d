"""Pretty-print a bench.py JSON line read from stdin or a file (development helper)."""
import json
import sys

txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
line = [l for l in txt.splitlines() if l.startswith("{")][-1]
d = json.loads(line)
print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.1f}  "
      f"launches/step {d.get('gpu_launches_per_step')}  clocks {d.get('clocks')}")
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f"  {k:14s} {v['ms_per_step']:8.3f} ms  {v['achieved']:8.1f} {v['unit']:8s} frac {v['frac']:.3f}  n={v['launches']}")
print("  conv_total", d.get("conv_total"))
print("  roofline", d.get("roofline"))
print("  cpu_baseline", d.get("cpu_baseline"))
