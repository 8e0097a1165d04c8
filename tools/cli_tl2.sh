set +e
mkdir -p /tmp/ps
python - <<'P'
import sys
sys.path.insert(0, ".")
import bench
g = bench.make_genome()
reads,_,_,_ = bench.simulate_reads(g, 1000000, seed=43)
bench.write_workload_files("/tmp/ps", g, reads)
P
for rep in 1 2 3; do
for pg in 0 1; do
echo "== rep $rep pregrow=$pg"
{ time env SMALT_B200_NOPREGROW=$((1-pg)) SMALT_B200_TIMING=1 smalt_b200/bin/smalt_b200 map -n 16 -O -o /tmp/ps/o.sam /tmp/ps/c2 /tmp/ps/reads.fq > /tmp/ps/tl.txt 2>&1 ; } 2> /tmp/ps/time.txt; tr "\n" " " < /tmp/ps/time.txt; echo
grep "CUDA init\|fastmap set-up\|wall " /tmp/ps/tl.txt
python - <<'P'
import re
t=sorted(float(m.group(1)) for m in re.finditer(r"block \d+ \(\d+ reads\) [0-9.]+ s, done at ([0-9.]+) s", open("/tmp/ps/tl.txt").read()))
n=len(t); print("blocks",n,"first done %.2f, 25%% %.2f, last %.2f"%(t[0],t[n//4],t[-1]))
P
sleep 2
done
done
