#!/bin/bash
# K2 device time of the bench's one-worker pass for several CTA counts of the 16-lane kernel
# (SMB_SW_CTAS_PER_SM, default 6): tools/k2_grid_sweep.sh 4 5 6 8
for g in "$@"; do
  SMB_SW_CTAS_PER_SM=$g timeout 200 python bench.py --no-cpu-baseline --no-cli --parity 0 --steps 3 2>/dev/null > /tmp/k2_$g.json
  python - "$g" <<'PY'
import json, sys
g = sys.argv[1]
d = json.load(open("/tmp/k2_%s.json" % g))
print("CTAs per SM", g, "e2e %.0f" % d["value"], "k2 %.2f ms" % d["kernel_ms"]["k2_sw_score"], "%.0f GCUPS" % d["sw_gcups"], flush=True)
PY
done
