#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native SMALT hot path.

Metric (BASELINE.json): mapped reads/s (and SW GCUPS) next to the reference's own CPU `smalt`
timed on this box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c1|c3|c4|c5] [--reads R]
  python bench.py --impl reference ...      # the reference CPU arm (oracle/_ref/smalt map)

Workloads (genome = uniform ACGT from numpy PCG64 seeded with the config number,
reads from the reference's own misc/simread with the seeds fixed there):
  c2 (default, the configuration the metric is quoted on): 5 Mb genome, 1 M single-end 150 bp reads per
      GPU, 2 % error, smalt index -k 13 -s 6
  c1  1 Mb, 10 k x 100 bp (the reference's CPU-runnable case)
  c3  100 Mb (4 x 25 Mb), pairs of 2 x 150 bp, insert 400, `smalt sample` + `map -g`
  c4  3.1 Gb (24 sequences), k = 20 s = 13, pairs of 2 x 150 bp
  c5  100 Mb, reads of 5-10 kb with 12 % indel-dominated error
c1/c3/c4/c5 are builder-run lines (profiles/), the driver runs c2.

One "step" = one pass of the mapping path over one batch of reads per GPU: FASTQ text in host memory ->
SAM text in host memory through the in-process driver (include/smalt_b200_map.h).

  value = e2e  reads/s of that call, wall clock over exactly K steps between barriers (host buffers in,
         host buffers out; every H2D/D2H copy and all host stages inside), max over ranks - the headline;
  device_value  reads / sum of the device times of all kernels of a step (CUDA events on the launching
         stream, a pass with ONE host worker so that no two streams overlap): explains the e2e figure;
  roofline / roofline_k1 / _k2 / _k3  per kernel, against the measured integer-issue peak (DP kernels;
         no tensor/HBM bound applies) or the measured HBM bandwidth (seed lookup);
  parity  SAM of a prefix of THIS run's reads from a one-worker run of the product, compared byte for
         byte with the reference's own `smalt map` on the same prefix (per rank when N > 1).

Multi-GPU: one process per GPU (torchrun), reads sharded by rank, index + reference replicated, no
collective on the data path (weak scaling); host cores are split between ranks; the per-rank SAM texts are
merged in input order by offset writes into one shared file (smalt_b200/shard.py), inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
SMALT_REF = os.path.join(REF_DIR, "smalt")
SIMREAD = os.path.join(REF_DIR, "simread")
SMALT_B200 = os.path.join(ROOT, "smalt_b200", "bin", "smalt_b200")

# seeds and simread arguments of SURVEY.md 8(d)
CONFIGS = {
    "c1": dict(name="C1", seqs=[1_000_000], gseed=1, k=13, s=6, qlen=100, units=10_000, err=1.0, insert=0, std=0,
               simseed=42, paired=False),
    "c2": dict(name="C2", seqs=[5_000_000], gseed=2, k=13, s=6, qlen=150, units=1_000_000, err=2.0, insert=0, std=0,
               simseed=43, paired=False),
    "c3": dict(name="C3", seqs=[25_000_000] * 4, gseed=3, k=13, s=6, qlen=150, units=1_000_000, err=2.0, insert=400,
               std=0.1, simseed=44, paired=True),
    "c4": dict(name="C4", seqs=[129_166_667] * 24, gseed=4, k=20, s=13, qlen=150, units=500_000, err=2.0, insert=400,
               std=0.1, simseed=45, paired=True),
    "c5": dict(name="C5", seqs=[25_000_000] * 4, gseed=3, k=13, s=6, qlen=0, units=512, err=12.0, insert=0, std=0,
               simseed=46, paired=False, long_reads=(5000, 10000)),
}
# integer instructions per DP cell of the recurrences as restated for two 16-bit lanes per register
# ("rooflines"): K2 5.5 ALU-pipe instructions per cell pair (PRMT over the row's score table, 3 x
# VIADDMNMX, 1.5 x VIMNMX3; H - gap_init is an IMAD on the FMA pipe; rounds 1 and 2a counted 6.5 with the LOP3 of
# the selector form, K2_OPS_PER_CELL_R1 keeps that yardstick beside the new one).  K3: the ALGORITHMIC minimum of the
# restricted recurrence on the ALU pipe per packed cell pair - h (VIADD), m = max(E, F) (VIMNMX), the
# comparison h <= m as a mask (3), t (VIADDMNMX.RELU + LOP3), E', F' (2 x VIADDMNMX), H' (VIMNMX), the running
# maximum key (VIMNMX), the direction code (4) = 16 per pair = 8 per cell; everything above that
# (staging, backtrace, band masks, lanes outside the band) is overhead the fraction exposes.
K2_OPS_PER_CELL = 2.75
K2_OPS_PER_CELL_R1 = 3.25
K3_OPS_PER_CELL = 8.0


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
def make_genome(cfg):
    rng = np.random.default_rng(cfg["gseed"])
    return [rng.integers(0, 4, n, dtype=np.uint8) for n in cfg["seqs"]]


def write_index(tmp, cfg, seqs, device=0, tag="idx"):
    """.smi/.sma of the genome (byte-identical to `smalt index -k K -s S`, tests/test_indexer.py,
    test_gpu_index_build.py).  Genomes beyond 20 Mb are indexed on the GPU (smb_index_build)."""
    from smalt_b200 import indexer
    from smalt_b200.seqpack import pack3
    pref = os.path.join(tmp, tag)
    total = sum(len(s) for s in seqs)
    words = pack3(np.concatenate(list(seqs) + [np.array([7], np.uint8)]))
    check = None
    if total > 20_000_000:
        from smalt_b200.capi import Context
        ctx = Context(device)
        try:
            ix = indexer.build_index_gpu(ctx, seqs, cfg["k"], cfg["s"], words=words)
        finally:
            ctx.close()
        check = index_invariants(ix, seqs, cfg["k"], cfg["s"])
    else:
        ix = indexer.build_index(seqs, cfg["k"], cfg["s"])
    indexer.write_smi(pref, ix)
    indexer.write_sma(pref, ["chr%d" % (i + 1) for i in range(len(seqs))], seqs, words=words)
    info = {"nwords": int(ix["nwords"]), "nkeys": int(ix["nkeys"]), "npos": int(ix["npos"]), "typ": int(ix["typ"])}
    if check:
        info["check"] = check
    return pref, info


def index_invariants(ix, seqs, k, nskip, nsample=5000):
    """a GPU-built index is too large for the host builder: the size-independent checks of smalt_b200/indexcheck.py
    (pinned against the host builder in tests/test_indexcheck.py); raises if one fails"""
    from smalt_b200 import indexcheck
    what = []
    if int(ix["npos"]) <= 100_000_000:   # (the array-wide passes take ~16 bytes per position of host memory)
        indexcheck.check_structure(ix)
        what.append("offset arrays monotone and complete, words / positions in order")
    n = indexcheck.check_samples(ix, seqs, k, nskip, nsample=nsample)
    what.append("%d sampled grid positions found under their words" % n)
    return "; ".join(what)


def simread(pref, cfg, n, seed, out, name_prefix):
    """reads of the reference's own simulator (misc/simread.c:738-755; drand48 seeded -> deterministic)
    -> list of FASTQ files (two for pairs)"""
    cmd = [SIMREAD, pref, str(cfg["qlen"]), str(n), "%g" % cfg["err"], "y", str(cfg["insert"]), "%g" % cfg["std"],
           str(seed), name_prefix, out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("simread failed: " + r.stdout.decode()[-300:])
    if cfg["paired"]:
        return [out + "_1.fq", out + "_2.fq"]
    return [out + ".fq"]


def simulate_long_reads(seqs, n, seed, lo, hi, err):
    """C5: misc/simread only makes fixed lengths and substitution-dominated errors, so the
    long reads come from this seeded simulator: length uniform in [lo, hi], `err` % errors of which 80 %
    are indels of 1-3 bases, random strand -> FASTQ text"""
    rng = np.random.default_rng(seed)
    let = np.frombuffer(b"ACGT", np.uint8)
    out = []
    p = err / 100.0
    for i in range(n):
        L = int(rng.integers(lo, hi + 1))
        s = seqs[int(rng.integers(0, len(seqs)))]
        st = int(rng.integers(0, len(s) - 2 * L))
        ev = rng.random(L)
        is_del = ev < 0.4 * p
        is_ins = (ev >= 0.4 * p) & (ev < 0.8 * p)
        is_sub = (ev >= 0.8 * p) & (ev < p)
        step = np.ones(L, np.int64)
        step[is_del] += rng.integers(1, 4, int(is_del.sum()))
        step[is_ins] = 0
        step[0] = 0
        idx = st + np.cumsum(step)
        rd = s[np.minimum(idx, len(s) - 1)].copy()
        rnd = rng.integers(0, 4, L).astype(np.uint8)
        rd[is_ins] = rnd[is_ins]
        rd[is_sub] = (rd[is_sub] + 1 + rnd[is_sub] % 3) & 3
        if rng.integers(0, 2):
            rd = 3 - rd[::-1]
        out.append(b"@lr%d\n" % i + let[rd].tobytes() + b"\n+\n" + b"5" * L + b"\n")
    return b"".join(out)


def nth_record(text, n):
    """byte offset behind the n-th 4-line record of a FASTQ text"""
    p = 0
    for _ in range(4 * n):
        p = text.find(b"\n", p) + 1
        if p == 0:
            return len(text)
    return p


class Workload:
    """index files + the reads of one rank (texts in memory and files on disk)"""

    def __init__(self, tmp, cfg, units, rank=0, device=0):
        self.cfg, self.units, self.tmp = cfg, units, tmp
        t0 = time.time()
        seqs = make_genome(cfg)
        self.pref, self.ixinfo = write_index(tmp, cfg, seqs, device)
        self.t_index = time.time() - t0
        t0 = time.time()
        seed = cfg["simseed"] + 1000 * rank       # rank r maps its own reads (weak scaling)
        if cfg.get("long_reads"):
            lo, hi = cfg["long_reads"]
            text = simulate_long_reads(seqs, units, seed, lo, hi, cfg["err"])
            f = os.path.join(tmp, "reads.fq")
            with open(f, "wb") as fh:
                fh.write(text)
            self.files, self.texts = [f], [text]
            self.generator = "own seeded simulator (lengths %d-%d, %g %% errors, 80 %% of them indels)" % (lo, hi, cfg["err"])
        else:
            if not os.path.exists(SIMREAD):
                raise RuntimeError("oracle/_ref/simread not built")
            self.files = simread(self.pref, cfg, units, seed, os.path.join(tmp, "reads"), "r%d" % rank)
            self.texts = [open(f, "rb").read() for f in self.files]
            self.generator = "misc/simread %d %d %g y %d %g %d" % (cfg["qlen"], units, cfg["err"], cfg["insert"],
                                                                  cfg["std"], seed)
        self.t_reads = time.time() - t0
        self.nreads = units * (2 if cfg["paired"] else 1)
        del seqs

    def prefix_files(self, n, tag):
        """the first n units as files"""
        out = []
        for i, t in enumerate(self.texts):
            f = os.path.join(self.tmp, "%s_%d.fq" % (tag, i + 1))
            with open(f, "wb") as fh:
                fh.write(t[:nth_record(t, n)])
            out.append(f)
        return out


def workload_config(cfg, units, world=1):
    what = "%s: %s synthetic genome (uniform ACGT, numpy PCG64 seed %d)" % (
        cfg["name"], " + ".join("%d" % n for n in sorted(set(cfg["seqs"]))) + (" x %d" % len(cfg["seqs"]) if len(cfg["seqs"]) > 1 else "") + " bp",
        cfg["gseed"])
    if cfg.get("long_reads"):
        reads = "%d reads of %d-%d bp per GPU, %g%% indel-dominated error" % ((units,) + cfg["long_reads"] + (cfg["err"],))
    elif cfg["paired"]:
        reads = "%d pairs of 2 x %d bp per GPU, insert %d, %g%% error (misc/simread seed %d + 1000 x rank)" % (
            units, cfg["qlen"], cfg["insert"], cfg["err"], cfg["simseed"])
    else:
        reads = "%d single-end %d bp reads per GPU, %g%% error (misc/simread seed %d + 1000 x rank)" % (
            units, cfg["qlen"], cfg["err"], cfg["simseed"])
    return {"workload": "%s, %s, smalt index -k %d -s %d" % (what, reads, cfg["k"], cfg["s"]),
            "reads_per_gpu": units * (2 if cfg["paired"] else 1), "gpus": world,
            "l2": "inputs larger than L2 (FASTQ text, task lists and outputs of a step are > 126 MB)"}


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def count_mapped(data):
    mapped = total = 0
    for ln in data.split(b"\n"):
        if not ln or ln[:1] == b"@":
            continue
        total += 1
        if not int(ln.split(b"\t", 2)[1]) & 4:
            mapped += 1
    return mapped, total


def run_program(exe, args, env=None):
    t0 = time.time()
    r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    dt = time.time() - t0
    if r.returncode != 0:
        return None, "%s failed: %s" % (os.path.basename(exe), r.stderr.decode()[-300:])
    return dt, None


def scaled_config(args):
    cfg = dict(CONFIGS[args.config])
    if args.genome_scale != 1.0:
        cfg["seqs"] = [max(1000, int(n * args.genome_scale)) for n in cfg["seqs"]]
    return cfg


def host_threads():
    try:
        return max(1, min(len(os.sched_getaffinity(0)), 64))
    except AttributeError:
        return max(1, min(os.cpu_count() or 1, 64))


def dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def sam_records(path):
    with open(path, "rb") as f:
        return [ln for ln in f.read().split(b"\n") if ln and ln[:1] != b"@"]


def map_options(cfg, hist=None):
    if not cfg["paired"]:
        return []
    return ["-g", hist] if hist else ["-i", "600", "-j", "200"]


def parity_check(wl, n, device, hist):
    """SAM of the first n units: one-worker `smalt_b200 map -r 7` (this product, this GPU) against the
    reference's own `smalt map -r 7` (fixed seed: by default the reference draws among equally good
    placements with a calendar-seeded drand48) - every record, byte for byte"""
    if not os.path.exists(SMALT_REF) or not os.path.exists(SMALT_B200):
        return {"checked": False, "why": "oracle/_ref/smalt or smalt_b200/bin/smalt_b200 not built"}
    n = min(n, wl.units)
    files = wl.prefix_files(n, "par")
    opts = ["map", "-r", "7"] + map_options(wl.cfg, hist)
    ref_out, own_out = os.path.join(wl.tmp, "par_ref.sam"), os.path.join(wl.tmp, "par_own.sam")
    dt_r, err = run_program(SMALT_REF, opts + ["-o", ref_out, wl.pref] + files)
    if dt_r is None:
        return {"checked": False, "why": err}
    env = dict(os.environ, SMALT_B200_DEVICE=str(device))
    dt_o, err = run_program(SMALT_B200, opts + ["-o", own_out, wl.pref] + files, env=env)
    if dt_o is None:
        return {"checked": False, "why": err}
    a, b = sam_records(ref_out), sam_records(own_out)
    ndiff = sum(1 for x, y in zip(a, b) if x != y) + abs(len(a) - len(b))
    return {"checked": True, "units": n, "sam_records": len(a), "differing_records": ndiff, "identical": ndiff == 0,
            "how": "first %d %s of this rank's input, `map -r 7` one worker, all records compared with the "
                   "reference's own program" % (n, "pairs" if wl.cfg["paired"] else "reads")}


def reference_rate(wl, n, threads, hist, tag="ref"):
    """whole `oracle/_ref/smalt map -n threads -O` program on the first n units -> (reads/s, seconds)"""
    files = wl.prefix_files(n, tag) if n < wl.units else wl.files
    out = os.path.join(wl.tmp, tag + ".sam")
    dt, err = run_program(SMALT_REF, ["map", "-n", str(threads), "-O"] + map_options(wl.cfg, hist) + ["-o", out, wl.pref] + files)
    if dt is None:
        return None, err, None
    return n * (2 if wl.cfg["paired"] else 1) / dt, dt, out


def sample_histogram(wl, exe, threads, device=None, skip=100):
    """insert-size estimation: `smalt sample` on every skip-th pair (pairs 0, skip, 2 skip, ... - the pairs
    insIsInSample picks with a sampling interval of `skip`, insert.c:215-218) -> histogram file for `map -g`
    (smalt.c:838-878, :1288, :1397; insert.c).  The sampled pairs are cut out of the input first, so that the
    program does not read the other 99 %."""
    from smalt_b200.shard import every_nth_record
    out = os.path.join(wl.tmp, "insert_%s.hist" % os.path.basename(exe))
    sub = []
    for i, t in enumerate(wl.texts):
        f = os.path.join(wl.tmp, "sampled_%d.fq" % (i + 1))
        if not os.path.exists(f):
            with open(f, "wb") as fh:
                fh.write(every_nth_record(t, skip))
        sub.append(f)
    env = dict(os.environ, SMALT_B200_DEVICE=str(device)) if device is not None else None
    dt, err = run_program(exe, ["sample", "-u", "1", "-n", str(threads), "-o", out, wl.pref] + sub, env=env)
    return (out, dt, None) if dt is not None else (None, None, err)


# --------------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    """the reference's own CPU `smalt map` on this box's host cores (rank 0 only), same workload and
    read count as the product's arm"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not os.path.exists(SMALT_REF):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/smalt not built"}))
        return
    cfg = scaled_config(args)
    units = args.reads or cfg["units"]
    nsample = min(args.ref_sample or units, units)
    cores = host_threads()
    with tempfile.TemporaryDirectory() as tmp:
        try:
            wl = Workload(tmp, cfg, units)
        except Exception as exc:
            print(json.dumps({"impl": "reference", "unavailable": repr(exc)[:200]}))
            return
        hist = None
        if cfg["paired"]:
            hist, _, err = sample_histogram(wl, SMALT_REF, cores)
            if hist is None:
                print(json.dumps({"impl": "reference", "unavailable": err}))
                return
        # the reference does not scale past ~16 PROC threads (slower at -n 32 than at -n 16):
        # one untimed run per candidate thread count, the timed steps use the faster
        cands = sorted({cores, min(cores, 16)})
        trial = {}
        for t in cands:
            v, dt, _ = reference_rate(wl, nsample, t, hist)
            if v is None:
                print(json.dumps({"impl": "reference", "unavailable": dt}))
                return
            trial[t] = v
        threads = max(trial, key=trial.get)
        rates, out = [], None
        for it in range(max(0, args.warmup - len(cands)) + args.steps):
            v, dt, out = reference_rate(wl, nsample, threads, hist)
            if it >= max(0, args.warmup - len(cands)):
                rates.append(v)
        mapped, total = count_mapped(open(out, "rb").read())
    v = float(np.mean(rates))
    nreads = nsample * (2 if cfg["paired"] else 1)
    sample = ("%s %d %s of the workload per step; whole `smalt map -n %d -O` program (index load, FASTQ parsing, "
              "SAM output); reads/s at the thread counts tried: %s"
              % ("all" if nsample == units else "first", nsample, "pairs" if cfg["paired"] else "reads", threads,
                 ", ".join("-n %d: %.0f" % (t, r) for t, r in sorted(trial.items()))))
    print(json.dumps({
        "impl": "reference", "metric": "mapped reads/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * nreads / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/i16 (SSE2)", "data": "synthetic",
        "config": workload_config(cfg, units, args.gpus),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference", "sample": sample,
                         "host_cores": cores, "mapped_fraction": mapped / max(total, 1)},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------------------------------
# the product's arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="reads (pairs) per GPU per step (0: the config's size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=0, help="reference arm: units per step (0: all, same job)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="cpu_baseline: units of the sample (0: config default)")
    ap.add_argument("--threads", type=int, default=0, help="host worker threads per GPU (0: chosen from the cores per GPU)")
    ap.add_argument("--device-block", type=int, default=32000, help="reads per launch of the device-time pass")
    ap.add_argument("--parity", type=int, default=20000, help="units of the SAM parity check (0: none)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true")
    ap.add_argument("--no-device-pass", action="store_true")
    ap.add_argument("--genome-scale", type=float, default=1.0, help="scale the sequence lengths of the config (trial runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank, world, local, dist = dist_setup()
    from smalt_b200.capi import Context
    from smalt_b200.mapper import Mapper
    from smalt_b200 import shard

    cfg = scaled_config(args)
    units = args.reads or cfg["units"]
    cores = host_threads()
    # host worker threads per GPU: the paired and long-read paths keep a device stream per worker (long reads: twice as
    # many workers as cores hide the device latency); the single-end path (workers never wait for the device, two polling device
    # threads that sleep between polls) wants a worker per core on 4 cores, one core less from 6 cores, and no more
    # than 14 workers (they saturate one GPU; tools/core_sweep.py: 4 cores 4 > 5 > 6 workers, 8 cores 7 > 8, 16 cores
    # 14 ~ 13 > 16)
    per_rank = cores / world
    if cfg["paired"]:
        threads = args.threads or max(1, int(per_rank))         # blocks of 4096 pairs: one worker per core (C3: 16 > 24 > 32 > 12)
    elif cfg.get("long_reads"):
        threads = args.threads or max(1, int(round(2.0 * per_rank)))
    else:
        threads = args.threads or min(14, max(2, int(per_rank) - (1 if per_rank >= 6 else 0)))
    tmpdir = tempfile.TemporaryDirectory()
    tmp = tmpdir.name
    wl = Workload(tmp, cfg, units, rank, local)
    nreads = wl.nreads
    paired = cfg["paired"]

    def barrier():
        if dist is not None:
            dist.barrier()

    # insert-size estimation (paired configs): `sample` on this rank's pairs; the histograms of the ranks
    # (every 100th pair; with several ranks the sampled pairs of all ranks are merged on the host and sampled
    # once - the only cross-read state of the reference, smalt.c:838-878)
    hist, t_sample = None, None
    if paired:
        t0 = time.time()
        if dist is None:
            hist, _, err = sample_histogram(wl, SMALT_B200, max(1, cores // world), device=local)
            if hist is None:
                raise RuntimeError("sample failed: " + err)
        else:
            hist = shard.sample_insert_sizes(dist, SMALT_B200, wl.pref, wl.texts[0], wl.texts[1], 100,
                                             "/dev/shm/smalt_b200_bench_%s.hist" % os.environ.get("MASTER_PORT", "0"), cores,
                                             env=dict(os.environ, SMALT_B200_DEVICE=str(local)))
        t_sample = time.time() - t0
    opts = map_options(cfg, hist)

    def run_step(m, copy=False):
        if paired:
            return m.map_fastq_pairs(wl.texts[0], wl.texts[1], copy=copy)
        return m.map_fastq(wl.texts[0]) if copy else m.map_fastq_nocopy(wl.texts[0])

    # ---- pass 1: device-resident kernel times, ONE host worker (no overlapping streams), large launches ----
    s1 = None
    if not args.no_device_pass:
        os.environ["SMALT_B200_BLOCK"] = str(args.device_block)
        m = Mapper(wl.pref, 1, options=opts, paired=paired)
        run_step(m)   # warm-up of this mapper (buffers, kernels)
        run_step(m)
        s1 = m.stats.as_dict()
        m.close()
        del os.environ["SMALT_B200_BLOCK"]

    # ---- pass 2: e2e through the in-process driver, all host workers; per-rank SAM merged in input order ----
    m = Mapper(wl.pref, threads, options=opts, paired=paired)
    merged = os.path.join(os.environ.get("SMALT_B200_MERGE_DIR", "/dev/shm"), "smalt_b200_bench_%s.sam" % os.environ.get("MASTER_PORT", "0"))
    for _ in range(args.warmup):
        if dist is None:
            run_step(m)
        else:   # the merge is part of a step: its file is written once before the timed region
            sam = m.map_fastq_view(wl.texts[0], wl.texts[1] if paired else None)
            shard.merge_to_file(dist, sam, merged, final_size=False)
            del sam
    c0 = m.stats.as_dict()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    sam_bytes = merged_bytes = 0
    t_map = t_merge = 0.0
    for _ in range(args.steps):
        if dist is None:
            sam_bytes = run_step(m)
        else:
            ta = time.perf_counter()
            sam = m.map_fastq_view(wl.texts[0], wl.texts[1] if paired else None)   # the mapper's own buffer, no copy
            tb = time.perf_counter()
            sam_bytes = len(sam)
            merged_bytes = shard.merge_to_file(dist, sam, merged, final_size=False)   # copies at the ranks' offsets, ends with a barrier
            del sam
            t_map += tb - ta
            t_merge += time.perf_counter() - tb
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    c1 = m.stats.as_dict()
    sam = run_step(m, copy=True) if units <= 50_000 else None
    m.close()
    if dist is not None and rank == 0 and os.path.exists(merged):
        os.unlink(merged)
    mapped_fraction = None
    if sam is not None:
        mp, tot = count_mapped(sam)
        mapped_fraction = mp / max(tot, 1)

    # ---- parity: SAM of a prefix against the reference, on every rank ----
    parity = parity_check(wl, args.parity, local, hist) if args.parity else {"checked": False, "why": "--parity 0"}

    wall_ms = 1e3 * wall / args.steps
    dev_ms = (s1["k1_ms"] + s1["k2_ms"] + s1["k3_ms"]) if s1 else 0.0
    par_ok = 1.0 if parity.get("identical") else 0.0
    par_checked = 1.0 if parity.get("checked") else 0.0
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, wall_ms, -par_ok, -par_checked], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
        parity["all_ranks_identical"] = bool(-float(t[2]) > 0.5)
        parity["all_ranks_checked"] = bool(-float(t[3]) > 0.5)
    if rank != 0:
        return

    e2e = world * nreads / (wall_ms * 1e-3)
    launches = (c1["gpu_launches"] - c0["gpu_launches"]) // args.steps
    h2d = (c1["h2d_bytes"] - c0["h2d_bytes"]) // args.steps
    d2h = (c1["d2h_bytes"] - c0["d2h_bytes"]) // args.steps
    line = {
        "metric": "mapped reads/sec", "value": e2e, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i16x2 DPX (K2, K3), u32/u64 (K1)", "data": "synthetic",
        "config": workload_config(cfg, units, world),
        "run": {"host_workers_per_gpu": threads, "host_cores": cores, "reads_generator": wl.generator,
                "index_s": wl.t_index, "reads_s": wl.t_reads, "index": wl.ixinfo},
        "timing": "value = e2e: reads of all ranks / wall clock of K calls of smbm_map_fastq%s (FASTQ text in host "
                  "memory -> SAM text in host memory%s), barriers on both sides, max over ranks; device_value: reads "
                  "/ sum of CUDA-event kernel times of a step (one host worker = one stream, inputs resident, %d reads "
                  "per launch)" % ("_pairs" if paired else "", ", per-rank SAM merged into one file in input order" if world > 1 else "",
                                   args.device_block),
        "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": wall_ms, "sam_bytes_per_step": int(sam_bytes), "merged_sam_bytes_per_step": int(merged_bytes),
                "rank0_map_ms_per_step": 1e3 * t_map / args.steps, "rank0_merge_ms_per_step": 1e3 * t_merge / args.steps,
                "host_stage_wall_s": {k: v for k, v in c1["host_stage_s"].items() if v},   # (results.*: only with SMALT_B200_TIMING)
                "host_stage_cpu_s": c1["host_cpu_s"],
                # who formatted CIGAR / NM of the SAM records of the timed steps: the device's output stage
                # (csrc/cigar.cu; default up to 8 host workers, SMALT_B200_DEVCIGAR=1 forces it) or diffstr.c on the host
                "cigar_records_device_per_step": int(c1["cigar_dev"]),   # (of the last timed step)
                "cigar_records_host_per_step": int(c1["cigar_host"])},
        "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
    }
    if mapped_fraction is not None:
        line["mapped_fraction"] = mapped_fraction
    if t_sample is not None:
        line["sample_s"] = t_sample
    if s1:
        ctx = Context(local)
        # giga thread-instructions/s: VIADDMNMX, VIMNMX3, IADD+IMNMX pairs (ops), VIADDMNMX.S16x2, VIMNMX3.S16x2
        peaks = ctx.int_peak()
        ctx.close()
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            peak_src = "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
        k2_gcups = s1["k2_cells"] / (s1["k2_ms"] * 1e-3) / 1e9 if s1["k2_ms"] else 0.0
        k3_gcups = s1["k3_cells"] / (s1["k3_ms"] * 1e-3) / 1e9 if s1["k3_ms"] else 0.0
        k2_peak, k3_peak = peaks[3] / K2_OPS_PER_CELL, peaks[3] / K3_OPS_PER_CELL
        # K1 algorithmic bytes: per lookup 8 (idx pair) + 4*ceil(log2(bucket+1)) (wordidx probes) + 8
        # (posidx pair); lookups = 2 strands x (qlen - k + 1) per read
        qlen_mean = (sum(len(t) for t in wl.texts) / max(1, nreads) - 8) / 2 if cfg.get("long_reads") else cfg["qlen"]
        nlook = 2 * nreads * max(1.0, qlen_mean - cfg["k"] + 1)
        bucket = max(1.0, wl.ixinfo["nwords"] / max(1, wl.ixinfo["nkeys"])) if wl.ixinfo["typ"] else 1.0
        probes = float(np.ceil(np.log2(bucket + 1))) if wl.ixinfo["typ"] else 0.0
        k1_bytes = nlook * (8 + 4 * probes + 8)
        k1_seed_ms = s1["k1_ms"] - s1["cand_ms"]
        k1_gbs = k1_bytes / (k1_seed_ms * 1e-3) / 1e9 if k1_seed_ms > 0 else 0.0
        k1_only = s1["k1_ms"] - s1["cand_ms"]
        kernel_ms = {"k1_seed_hits": k1_only, "candidates_replay": s1["cand_ms"], "k2_sw_score": s1["k2_ms"],
                     "k3_band_align": s1["k3_ms"]}
        longr = bool(cfg.get("long_reads"))
        roof_k3 = {"kernel": "K3 banded DP + backtrace (%s)" % ("band_long_kernel<16|32>, one CTA per task, TMA-staged windows" if longr
                                                               else "band_pack_kernel for short reads"), "bound": "alu", "achieved": k3_gcups,
                   "peak": k3_peak, "unit": "GCUPS", "frac": k3_gcups / k3_peak if k3_peak else None,
                   "traffic": 1143.7e6 if longr else 13.1e6,
                   "note": "integer-issue bound (no tensor/HBM bound applies to this DP): peak = measured VIADDMNMX.S16x2 issue "
                           "rate %.0f G thread-instr/s / %.1f ALU-pipe instructions per cell, the algorithmic minimum of the "
                           "restricted recurrence with directions (16 per packed cell pair, bench.py K3_OPS_PER_CELL); staging, "
                           "backtrace and lanes outside the band count against the fraction" % (peaks[3], K3_OPS_PER_CELL)}
        roof_k2 = {"kernel": "K2 SW score (%s)" % ("sw_long2_kernel, 512-column blocks, strips in HBM" if longr
                                                    else "sw_score2_kernel, 2 tasks per 16 lanes"), "bound": "alu", "achieved": k2_gcups,
                   "peak": k2_peak, "unit": "GCUPS", "frac": k2_gcups / k2_peak if k2_peak else None,
                   "traffic": 9.68e9 if longr else 11.8e6,
                   "frac_round1_definition": k2_gcups * K2_OPS_PER_CELL_R1 / peaks[3] if peaks[3] else None,
                   "note": "DPX issue bound: peak = measured VIADDMNMX.S16x2 issue rate %.0f G thread-instr/s / %.2f "
                           "ALU-pipe instructions per cell (5.5 per packed cell pair: PRMT, 3 x VIADDMNMX, 1.5 x VIMNMX3; "
                           "frac_round1_definition keeps the 6.5-instruction yardstick of the earlier kernel)"
                           % (peaks[3], K2_OPS_PER_CELL)}
        roof_k1 = {"kernel": "K1 seed tables + hit lists (seed_warp_kernel, hits_warp_kernel)", "bound": "hbm", "achieved": k1_gbs, "peak": hbm_peak,
                   "unit": "GB/s", "frac": k1_gbs / hbm_peak, "traffic": 37.3e6, "peak_source": peak_src,
                   "note": "algorithmic bytes of the index probes only (%.0f B per lookup, %.1f wordidx probes); dependent "
                           "4-byte loads, latency bound; the time includes the hit lists" % (8 + 4 * probes + 8, probes)}
        roof_cand = {"kernel": "candidate selection, task lists, score replay (block.cu)", "bound": "latency", "achieved": None,
                     "peak": None, "unit": None, "frac": None, "traffic": None,
                     "note": "integer bookkeeping on per-read lists of a few dozen hits (one lane per hit list): no roofline applies; "
                             "%.1f ns per read" % (1e6 * s1["cand_ms"] / max(1, nreads))}
        dominant = max(kernel_ms, key=kernel_ms.get)
        line.update({
            "device_value": world * nreads / (dev_ms * 1e-3) if dev_ms else None, "device_ms_per_step": dev_ms,
            "kernel_ms": kernel_ms, "sw_gcups": k2_gcups, "band_gcups": k3_gcups,
            "tasks_per_step": {"k2_tasks": s1["k2_tasks"], "k2_cells": s1["k2_cells"], "k3_tasks": s1["k3_tasks"],
                               "k3_cells": s1["k3_cells"]},
            "int_peaks_ginstr": {"viaddmnmx": peaks[0], "vimnmx3": peaks[1], "iadd_imnmx_ops": peaks[2],
                                 "viaddmnmx_s16x2": peaks[3], "vimnmx3_s16x2": peaks[4]},
            "roofline": {"k3_band_align": roof_k3, "k2_sw_score": roof_k2, "k1_seed_hits": roof_k1,
                         "candidates_replay": roof_cand}[dominant],
            "roofline_k2": roof_k2, "roofline_k3": roof_k3, "roofline_k1": roof_k1})
    if world == 1 and not args.no_cli:
        # the same job as a whole program, like the reference arm runs it (process start, CUDA start-up, index
        # load, file I/O included); two runs, the faster (a program started while the driver tears down the
        # contexts of the passes above waits seconds in CUDA start-up)
        runs, err = [], None
        for _ in range(2):
            time.sleep(2.0)
            dt, err = run_program(SMALT_B200, ["map", "-n", str(threads), "-O"] + opts + ["-o", os.path.join(tmp, "cli.sam"), wl.pref] + wl.files)
            if not dt:
                break
            runs.append(dt)
        dt = min(runs) if runs else None
        line["e2e_cli"] = ({"value": nreads / dt, "unit": "reads/s", "seconds": dt, "runs_seconds": runs,
                            "what": "whole `smalt_b200 map -n %d -O` program on the same reads" % threads}
                           if dt else {"value": None, "unavailable": err})
    if world == 1 and not args.no_cpu_baseline:
        ns = min(args.cpu_sample or units, units)
        if cfg.get("long_reads"):
            ns = min(ns, args.cpu_sample or 64)
        if os.path.exists(SMALT_REF):
            ref_hist = hist
            t_best, best = None, None
            for t in sorted({cores, min(cores, 16)}):
                v, dt, _ = reference_rate(wl, ns, t, ref_hist, tag="cpu")
                if v is not None and (best is None or v > best):
                    best, t_best, dt_best = v, t, dt
            if best is not None:
                line["cpu_baseline"] = {"value": best, "unit": "reads/s", "cores": t_best, "host_cores": cores, "kind": "reference",
                                        "sample": "%s %d %s of this workload, whole `oracle/_ref/smalt map -n %d -O` "
                                                  "program (%.1f s)" % ("all" if ns == units else "first", ns,
                                                                        "pairs" if paired else "reads", t_best, dt_best)}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": cores, "kind": "reference",
                                        "sample": "unavailable: " + str(dt)}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": cores, "kind": "reference",
                                    "sample": "unavailable: oracle/_ref/smalt not built"}
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        if "torch.distributed" in sys.modules:
            import torch.distributed as _d
            if _d.is_available() and _d.is_initialized():
                _d.destroy_process_group()
