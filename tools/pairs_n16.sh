set +e
python - <<'P'
import sys, os
sys.path.insert(0, "tools")
import paired_check as pc
tmp = "/tmp/pp"; os.makedirs(tmp, exist_ok=True)
pc.make(tmp, 200000, 20, 4)
P
for n in 4 8 16; do
echo "n=$n"; time env SMALT_B200_TIMING=1 SMALT_B200_STATS=/tmp/pp/st.json smalt_b200/bin/smalt_b200 map -r 7 -n $n -O -i 600 -j 200 -o /tmp/pp/o.sam /tmp/pp/idx /tmp/pp/r1.fq /tmp/pp/r2.fq 2>&1 | grep -v "^#" | tail -12
head -2 /tmp/pp/st.json
done
