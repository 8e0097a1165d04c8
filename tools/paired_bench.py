"""The paired-end section of bench.py alone (in-process, all workers): python tools/paired_bench.py [steps]"""
import os, sys, tempfile, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cores = bench.host_threads()
with tempfile.TemporaryDirectory() as tmp:
    r = bench.run_paired(tmp, 2 * cores, cores, 250000, 1000, steps, 2)
    print(json.dumps({k: r[k] for k in ("e2e", "kernel_ms")}))
