# ncu captures of the round's kernels (one GPU): launch list + --set full of one launch per kernel
set +e
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --reads 100000 --threads 1 --no-cli --no-cpu-baseline --no-paired"
$B > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s3.csv $B > gpurun_out/ncu_l.log 2>&1
for k in band_pack_kernel sw_score2_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_${k}_s3 $B > gpurun_out/ncu_$k.log 2>&1
done
# band_wide_kernel: a paired run (pass 4 tasks of the rescue searches)
python - <<'P'
import sys, os
sys.path.insert(0, "tools")
import paired_check as pc
os.makedirs("/tmp/pp", exist_ok=True)
pc.make("/tmp/pp", 20000, 4, 4)
P
ncu --set full --clock-control none --import-source on -k regex:band_wide_kernel -s 2 -c 1 -f -o gpurun_out/prof_band_wide_kernel_s3 smalt_b200/bin/smalt_b200 map -r 7 -n 1 -i 600 -j 200 -o /tmp/pp/o.sam /tmp/pp/idx /tmp/pp/r1.fq /tmp/pp/r2.fq > gpurun_out/ncu_band_wide.log 2>&1
