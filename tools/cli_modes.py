"""whole-program timing of `smalt_b200 map` on the C2 workload with the stage timers on (SMALT_B200_TIMING):
python tools/cli_modes.py [reads] [ENV=VALUE ...]   (one run per given environment, plus the default)"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
modes = [{}] + [dict([a.split("=", 1)]) for a in sys.argv[2:]]
cfg = bench.CONFIGS["c2"]
with tempfile.TemporaryDirectory() as tmp:
    wl = bench.Workload(tmp, cfg, n)
    for mode in modes:
        t0 = time.time()
        r = subprocess.run([bench.SMALT_B200, "map", "-n", "32", "-O", "-o", "/dev/null", wl.pref] + wl.files,
                           capture_output=True, text=True, env=dict(os.environ, SMALT_B200_TIMING="1", **mode))
        dt = time.time() - t0
        print("cli %s: rc %d %.2f s %.0f reads/s" % (mode, r.returncode, dt, n / dt))
        lines = [l for l in r.stderr.splitlines() if "timing" in l]
        keep = [l for l in lines if "batch slot" in l or "device batches" in l or "fastmap set-up" in l]
        print("\n".join(keep[:80]))
        print("...")
        print("\n".join(lines[-4:]))
