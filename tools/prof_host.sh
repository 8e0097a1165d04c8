# CPU-time profile of the host side of `smalt_b200 map` on the C2 workload (sampling profiler built into the driver,
# SMALT_B200_PROF): top functions by samples, flat and attributed to the innermost frame inside the driver's libraries
set +e
mkdir -p gpurun_out /tmp/ph
python - <<'P'
import sys
sys.path.insert(0, ".")
import bench
wl = bench.Workload("/tmp/ph", bench.CONFIGS["c2"], 1000000)
print(wl.pref, wl.files[0])
P
N=${1:-14}
CORES=${2:-16}
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do cat /tmp/ph/reads.fq; done > /tmp/ph/reads6.fq
SMALT_B200_PROF=gpurun_out/prof_host_flat.txt taskset -c 0-$((CORES-1)) smalt_b200/bin/smalt_b200 map -n $N -O -o /dev/null /tmp/ph/idx /tmp/ph/reads6.fq > /dev/null 2>&1
SMALT_B200_PROF_CALLERS=1 SMALT_B200_PROF=gpurun_out/prof_host_callers.txt taskset -c 0-$((CORES-1)) smalt_b200/bin/smalt_b200 map -n $N -O -o /dev/null /tmp/ph/idx /tmp/ph/reads6.fq > /dev/null 2>&1
SMALT_B200_PROF_CALLERS=2 SMALT_B200_PROF=gpurun_out/prof_host_callers2.txt taskset -c 0-$((CORES-1)) smalt_b200/bin/smalt_b200 map -n $N -O -o /dev/null /tmp/ph/idx /tmp/ph/reads6.fq > /dev/null 2>&1
python tools/prof_top.py gpurun_out/prof_host_flat.txt gpurun_out/prof_host_callers.txt gpurun_out/prof_host_callers2.txt
