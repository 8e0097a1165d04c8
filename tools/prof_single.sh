set +e
mkdir -p gpurun_out /tmp/ps
python - <<'P'
import sys, os
sys.path.insert(0, ".")
import bench, numpy as np
g = bench.make_genome()
reads,_,_,_ = bench.simulate_reads(g, 1000000, seed=43)
pref, fq, ix = bench.write_workload_files("/tmp/ps", g, reads)
print(pref, fq)
P
SMALT_B200_PROF=gpurun_out/prof_single.txt smalt_b200/bin/smalt_b200 map -n 16 -O -o /tmp/ps/o.sam /tmp/ps/c2 /tmp/ps/reads.fq 2>&1 | grep -v "block|^#" | tail -12
sort -n -r gpurun_out/prof_single.txt | head -3
