"""e2e (in-process, all workers) against the driver's block size: python tools/block_sweep.py [blocks...]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from smalt_b200.mapper import Mapper

blocks = [int(x) for x in sys.argv[1:]] or [0, 2048, 3072, 4096, 6144, 8192, 12288]
n = 1_000_000
cores = bench.host_threads()
genome = bench.make_genome()
reads, _, _, _ = bench.simulate_reads(genome, n, seed=43)
text = bench.fastq_text(reads)
tmp = tempfile.TemporaryDirectory()
pref, ix = bench.write_index_files(tmp.name, genome)
for workers in (2 * cores, 3 * cores):
    for b in blocks:
        if b:
            os.environ["SMALT_B200_BLOCK"] = str(b)
        else:
            os.environ.pop("SMALT_B200_BLOCK", None)
        m = Mapper(pref, workers)
        for _ in range(2):
            m.map_fastq_nocopy(text)
        t0 = time.perf_counter()
        for _ in range(3):
            m.map_fastq_nocopy(text)
        dt = (time.perf_counter() - t0) / 3
        m.close()
        print("workers %d block %5d: %.1f ms  %.2f M reads/s" % (workers, b, 1e3 * dt, n / dt / 1e6), flush=True)
