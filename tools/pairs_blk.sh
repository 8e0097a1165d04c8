set +e
python - <<'P'
import sys, os
sys.path.insert(0, "tools")
import paired_check as pc
tmp = "/tmp/pp"; os.makedirs(tmp, exist_ok=True)
pc.make(tmp, 200000, 20, 4)
P
for b in 512 1024 4096; do
echo "block=$b n=16"
time env SMALT_B200_BLOCK=$b smalt_b200/bin/smalt_b200 map -r 7 -n 16 -O -i 600 -j 200 -o /tmp/pp/o.sam /tmp/pp/idx /tmp/pp/r1.fq /tmp/pp/r2.fq 2>&1 | grep -v "^#" | tail -1
done
echo "ref n=16"; time oracle/_ref/smalt map -r 7 -n 16 -O -i 600 -j 200 -o /tmp/pp/r.sam /tmp/pp/idx /tmp/pp/r1.fq /tmp/pp/r2.fq 2>&1 | tail -1
