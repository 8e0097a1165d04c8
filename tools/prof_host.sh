# CPU-time profile of the host side of `smalt_b200 map` on the C2 workload (sampling profiler built into the driver,
# SMALT_B200_PROF): top functions by samples
set +e
mkdir -p gpurun_out /tmp/ph
python - <<'P'
import sys
sys.path.insert(0, ".")
import bench
wl = bench.Workload("/tmp/ph", bench.CONFIGS["c2"], 1000000)
print(wl.pref, wl.files[0])
P
SMALT_B200_PROF=gpurun_out/prof_host.txt smalt_b200/bin/smalt_b200 map -n 14 -O -o /dev/null /tmp/ph/idx /tmp/ph/reads.fq > /dev/null 2>&1
python - <<'P'
import collections
tot = collections.Counter()
for ln in open("gpurun_out/prof_host.txt"):
    f = ln.split()
    if len(f) < 4: continue
    tot[(f[1].split("/")[-1], f[3])] += int(f[0])
n = sum(tot.values())
print("samples", n)
for (lib, fn), c in tot.most_common(45):
    print("%6.2f %%  %-28s %s" % (100.0 * c / n, lib, fn))
P
