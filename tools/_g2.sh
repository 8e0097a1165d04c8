set +e
bash tools/ncu_launches.sh r2c c2 100000 32000 | tail -25
timeout 900 python bench.py --config c5 --steps 2 --warmup 1 --no-cli > gpurun_out/r2_c5d.json 2> gpurun_out/r2_c5d.err; tail -c 200 gpurun_out/r2_c5d.err
timeout 1500 python bench.py --config c3 --steps 2 --warmup 1 --no-cli > gpurun_out/r2_c3d.json 2> gpurun_out/r2_c3d.err; tail -c 200 gpurun_out/r2_c3d.err
python - <<P
import json
for f in ("r2_c5d","r2_c3d"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d.get("device_value"), d.get("kernel_ms"), d["parity"].get("identical"), d.get("cpu_baseline",{}).get("value"), d["run"]["host_workers_per_gpu"])
    except Exception as e: print(f, "ERR", e)
P
