python - <<'P'
import sys, os
sys.path.insert(0, "tools")
import paired_check as pc
tmp = "/tmp/pp"; os.makedirs(tmp, exist_ok=True)
pc.make(tmp, 20000, 4, 4)
P
SMB_PLAN_DEBUG=1 smalt_b200/bin/smalt_b200 map -r 7 -n 1 -i 600 -j 200 -o /tmp/pp/o.sam /tmp/pp/idx /tmp/pp/r1.fq /tmp/pp/r2.fq 2>&1 | grep "K3 plan" | head -16
