set +e
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c2d.json 2> gpurun_out/r2_c2d.err; tail -c 300 gpurun_out/r2_c2d.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_c2d_ref.json 2> gpurun_out/r2_c2d_ref.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2_c2d.json").read().strip().splitlines()[-1])
print(d["value"], d.get("device_value"), d["kernel_ms"], d["run"]["host_workers_per_gpu"], d["run"]["host_cores"], d["parity"]["identical"], d["e2e_cli"], d["cpu_baseline"]["value"])
r=json.loads(open("gpurun_out/r2_c2d_ref.json").read().strip().splitlines()[-1])
print(r["value"], r.get("impl"))
P
