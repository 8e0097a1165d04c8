# diagnostic: sampling profile of the fiber path (SMALT_B200_PROF)
python - <<'P'
import sys, os
sys.path.insert(0, "tools")
import paired_check as pc
tmp = "/tmp/pp"; os.makedirs(tmp, exist_ok=True)
pref, f1, f2 = pc.make(tmp, 20000, 4, 4)
print(pref, f1, f2)
P
mkdir -p gpurun_out; SMALT_B200_PROF=gpurun_out/prof_pairs.txt SMALT_B200_STATS=/tmp/pp/st.json smalt_b200/bin/smalt_b200 map -r 7 -n 1 -i 600 -j 200 -o /tmp/pp/o.sam /tmp/pp/idx /tmp/pp/r1.fq /tmp/pp/r2.fq 2>&1 | tail -3
cat /tmp/pp/st.json | tail -1
sort -n -r gpurun_out/prof_pairs.txt | head -5
