# ncu --set full capture of one band_pack_kernel launch (32000-read block of the bench workload)
set +e
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --reads 100000 --threads 1 --no-cli --no-cpu-baseline --no-paired"
$B > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:band_pack_kernel -s 3 -c 1 -f -o gpurun_out/prof_band_pack_s3b $B > gpurun_out/ncu_k3b.log 2>&1
tail -3 gpurun_out/ncu_k3b.log
