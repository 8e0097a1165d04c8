#!/bin/bash
# e2e of the C2 bench for several device batch sizes (SMALT_B200_BATCH, reads per combined batch)
for b in "$@"; do
  SMALT_B200_BATCH=$b timeout 200 python bench.py --no-cpu-baseline --no-cli --no-device-pass --parity 0 --steps 6 2>/dev/null > /tmp/bs_$b.json
  python - "$b" <<'PY'
import json, sys
b = sys.argv[1]
d = json.load(open("/tmp/bs_%s.json" % b))
print("batch", b, "e2e %.0f reads/s" % d["value"], "%.1f ms per step" % d["ms_per_step"], flush=True)
PY
done
