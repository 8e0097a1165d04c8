"""multi-GPU check (run under gpurun --gpus N): torchrun mapreads on N ranks == reference SAM"""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
ngpu = int(sys.argv[2]) if len(sys.argv) > 2 else 2
genome = bench.make_genome()
reads, _, _, _ = bench.simulate_reads(genome, n, seed=43)
with tempfile.TemporaryDirectory() as tmp:
    pref, fq, ix = bench.write_workload_files(tmp, genome, reads)
    out = os.path.join(tmp, "multi.sam")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ngpu), "--master-addr", "127.0.0.1",
           "--master-port", "29533", "-m", "smalt_b200.mapreads", "-r", "7", "-o", out, pref, fq]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=bench.ROOT)
    print("mapreads rc", r.returncode, "\n".join(l for l in r.stderr.splitlines() if not l.startswith("#"))[:4000] if r.returncode else "")
    ref = os.path.join(bench.ROOT, "oracle", "_ref", "smalt")
    rout = os.path.join(tmp, "ref.sam")
    subprocess.run([ref, "map", "-r", "7", "-n", "16", "-O", "-o", rout, pref, fq], check=True, capture_output=True)
    a = [l for l in open(out).read().splitlines() if not l.startswith("@PG")]
    b = [l for l in open(rout).read().splitlines() if not l.startswith("@PG")]
    d = [(x, y) for x, y in zip(a, b) if x != y]
    print("lines", len(a), len(b), "differing", len(d))
    for x, y in d[:3]:
        print(x[:120]); print(y[:120])
