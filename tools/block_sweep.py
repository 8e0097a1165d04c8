"""e2e (in-process, all workers) against the driver's block size and worker count:
python tools/block_sweep.py [blocks...]   (SWEEP_WORKERS="16 32 48" to choose the worker counts)"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from smalt_b200.mapper import Mapper

blocks = [int(x) for x in sys.argv[1:]] or [0, 4096, 8192, 16384]
cores = bench.host_threads()
workers_list = [int(x) for x in os.environ.get("SWEEP_WORKERS", "%d %d %d" % (cores, 2 * cores, 3 * cores)).split()]
tmp = tempfile.TemporaryDirectory()
wl = bench.Workload(tmp.name, bench.CONFIGS["c2"], 1_000_000)
text = wl.texts[0]
for workers in workers_list:
    for b in blocks:
        if b:
            os.environ["SMALT_B200_BLOCK"] = str(b)
        else:
            os.environ.pop("SMALT_B200_BLOCK", None)
        m = Mapper(wl.pref, workers)
        for _ in range(2):
            m.map_fastq_nocopy(text)
        t0 = time.perf_counter()
        for _ in range(3):
            m.map_fastq_nocopy(text)
        dt = (time.perf_counter() - t0) / 3
        st = m.stats.as_dict()
        m.close()
        print("workers %d block %5d: %.1f ms  %.2f M reads/s  kernel ms %.0f/%.0f/%.0f" % (
            workers, b, 1e3 * dt, wl.nreads / dt / 1e6, st["k1_ms"], st["k2_ms"], st["k3_ms"]), flush=True)
