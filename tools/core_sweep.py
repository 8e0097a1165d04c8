"""e2e of one rank confined to a few host cores (the 8-GPU box has 4 cores per GPU):
python tools/core_sweep.py "cores:workers[:nospin[:block]]" ...   e.g. 4:4 4:3 4:6:1 16:14 16:16:1 4:4:0:1024"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 2 and sys.argv[1] == "--child":
    import bench
    from smalt_b200.mapper import Mapper
    tmpdir, workers = sys.argv[2], int(sys.argv[3])
    wl = bench.Workload(tmpdir, bench.CONFIGS["c2"], 1_000_000)
    text = wl.texts[0]
    m = Mapper(wl.pref, workers)
    steps = []
    for _ in range(2):
        t0 = time.perf_counter()
        m.map_fastq_nocopy(text)
        steps.append(time.perf_counter() - t0)
    c0 = time.process_time()
    t0 = time.perf_counter()
    for _ in range(3):
        t1 = time.perf_counter()
        m.map_fastq_nocopy(text)
        steps.append(time.perf_counter() - t1)
    dt = (time.perf_counter() - t0) / 3
    cpu = (time.process_time() - c0) / 3
    m.close()
    print("RESULT %.1f ms  %.2f M reads/s  cpu %.2f core-s per step  steps(ms) %s" % (
        1e3 * dt, wl.nreads / dt / 1e6, cpu, " ".join("%.0f" % (1e3 * x) for x in steps)), flush=True)
    sys.exit(0)

tmp = tempfile.TemporaryDirectory()
import bench
bench.Workload(tmp.name, bench.CONFIGS["c2"], 1_000_000)   # builds index and reads once
for spec in sys.argv[1:]:
    f = spec.split(":")
    cores, workers = int(f[0]), int(f[1])
    env = dict(os.environ)
    if len(f) > 2 and f[2] == "1":
        env["SMALT_B200_NOSPIN"] = "1"
    if len(f) > 3 and f[3] != "0":
        env["SMALT_B200_BLOCK"] = f[3]
    for kv in f[4:]:                      # further fields: NAME=value environment settings
        k, v = kv.split("=")
        env[k] = v
    cmd = ["taskset", "-c", "0-%d" % (cores - 1), sys.executable, __file__, "--child", tmp.name, str(workers)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    print(spec, res[0] if res else "FAILED " + out.stderr[-300:], flush=True)
