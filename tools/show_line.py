"""prints the headline fields of a bench.py JSON line: python tools/show_line.py file.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])   # (the C3 run prints the insert-size histogram first)
print("value %.0f %s  ms_per_step %.1f  kernel_ms %s  parity %s  k2 frac %s  cpu_baseline %s" % (
    d["value"], d["unit"], d["ms_per_step"], {k: round(v, 1) for k, v in (d.get("kernel_ms") or {}).items()},
    (d.get("parity") or {}).get("identical"), (d.get("roofline_k2") or {}).get("frac"),
    (d.get("cpu_baseline") or {}).get("value")))
