set +e
python - <<'P'
import sys, os
sys.path.insert(0, "tools")
import paired_check as pc
tmp = "/tmp/pp"; os.makedirs(tmp, exist_ok=True)
pc.make(tmp, 200000, 20, 4)
P
mkdir -p gpurun_out
SMALT_B200_PROF_CALLERS=1 SMALT_B200_PROF=gpurun_out/prof_pairs16.txt smalt_b200/bin/smalt_b200 map -r 7 -n 16 -O -i 600 -j 200 -o /tmp/pp/o.sam /tmp/pp/idx /tmp/pp/r1.fq /tmp/pp/r2.fq 2>&1 | tail -2
sort -n -r gpurun_out/prof_pairs16.txt | head -25
