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
for n in 32 16 32; do
echo "== n=$n"
SMALT_B200_TIMING=1 smalt_b200/bin/smalt_b200 map -n $n -O -o /tmp/ps/o.sam /tmp/ps/c2 /tmp/ps/reads.fq > /tmp/ps/tl.txt 2>&1
grep -v "^#" /tmp/ps/tl.txt | grep -v "block [0-9]* (" | head -50
grep "block [0-9]* (" /tmp/ps/tl.txt | sort -t' ' -k9 -n | head -5
done
