# ncu captures of the round's kernels (one GPU): launch list + --set full of one launch per kernel.
# The bench's device pass maps blocks of 32000 reads on one stream; -s skips the warm-up launches.
set +e
mkdir -p gpurun_out
TAG=${1:-r1c}
B="python bench.py --steps 1 --warmup 1 --reads 100000 --threads 1 --no-cli --no-cpu-baseline --no-paired"
$B > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_l.log 2>&1
for k in band_pack_kernel sw_score2_kernel seed_warp_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${k}_$TAG $B > gpurun_out/ncu_$k.log 2>&1
done
