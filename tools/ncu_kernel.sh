# one --set full capture of one launch of a kernel of the one-worker device pass (C2, 32000-read launches)
# usage: bash tools/ncu_kernel.sh <kernel regex> <tag> [skip launches] [device block]
set +e
mkdir -p gpurun_out
K=$1; TAG=$2; SKIP=${3:-3}; DB=${4:-32000}
B="python bench.py --steps 1 --warmup 1 --reads 100000 --threads 1 --device-block $DB --no-cli --no-cpu-baseline --parity 0"
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/prof_${TAG} $B > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
