# CPU-time profile of the host side of the paired path (C3-shaped, 300 k pairs): tools/prof_host_pairs.sh [workers]
set +e
mkdir -p gpurun_out /tmp/php
N=${1:-32}
python - <<'P'
import sys
sys.path.insert(0, ".")
import bench
cfg = dict(bench.CONFIGS["c3"])
wl = bench.Workload("/tmp/php", cfg, 1500000)
print(wl.pref, wl.files)
P
ls -la /tmp/php | head
SMALT_B200_PROF=gpurun_out/prof_pairs_flat.txt smalt_b200/bin/smalt_b200 map -n $N -O -i 600 -j 200 -o /dev/null /tmp/php/idx /tmp/php/reads_1.fq /tmp/php/reads_2.fq > /dev/null 2> gpurun_out/prof_pairs.err
SMALT_B200_PROF_CALLERS=1 SMALT_B200_PROF=gpurun_out/prof_pairs_callers.txt smalt_b200/bin/smalt_b200 map -n $N -O -i 600 -j 200 -o /dev/null /tmp/php/idx /tmp/php/reads_1.fq /tmp/php/reads_2.fq > /dev/null 2>&1
tail -3 gpurun_out/prof_pairs.err
