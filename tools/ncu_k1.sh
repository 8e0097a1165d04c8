# ncu --set full capture of one seed_warp_kernel launch (32000-read block of the bench workload)
set +e
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --reads 100000 --threads 1 --no-cli --no-cpu-baseline --no-paired"
ncu --set full --clock-control none --import-source on -k regex:seed_warp_kernel -s 2 -c 1 -f -o gpurun_out/prof_seed_warp_s3b $B > gpurun_out/ncu_k1b.log 2>&1
tail -3 gpurun_out/ncu_k1b.log
