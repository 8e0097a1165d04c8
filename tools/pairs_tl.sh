set +e
python - <<'P'
import sys, os
sys.path.insert(0, "tools")
import paired_check as pc
tmp = "/tmp/pp"; os.makedirs(tmp, exist_ok=True)
pc.make(tmp, 200000, 20, 4)
P
for cfg in "16 512" "32 512" "64 512" "32 1024" "48 2048"; do
set -- $cfg
SMALT_B200_BLOCK=$2 SMALT_B200_TIMING=1 smalt_b200/bin/smalt_b200 map -r 7 -n $1 -O -i 600 -j 200 -o /tmp/pp/o.sam /tmp/pp/idx /tmp/pp/r1.fq /tmp/pp/r2.fq 2>&1 | grep -v "^#" > /tmp/pp/tl.txt
echo "== n=$1 block=$2"
grep -v "block\|worker context" /tmp/pp/tl.txt
python - <<'P'
import re
t=[float(m.group(1)) for m in re.finditer(r"done at ([0-9.]+) s", open("/tmp/pp/tl.txt").read())]
t.sort()
n=len(t)
print("blocks",n,"first done %.2f last %.2f"%(t[0],t[-1]),"steady: blocks %d..%d in %.2f s"%(n//4,n-1,t[-1]-t[n//4]))
P
done
