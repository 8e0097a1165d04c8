#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native SMALT hot path.

Metric (BASELINE.json): mapped reads/s (and SW GCUPS) on config C2 - 5 Mb synthetic genome,
1 M single-end 150 bp reads, smalt index -k 13 -s 6 - next to the reference's own CPU `smalt`
timed on this box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R]
  python bench.py --impl reference ...      # the reference CPU arm (oracle/_ref/smalt map)

One "step" = one pass of the mapping path over one batch of R reads per GPU (default 1 M):
FASTQ text in host memory -> SAM text in host memory through the in-process driver
(include/smalt_b200_map.h: the reference's unmodified candidate selection / results / SAM
writer around K1 seed lookup + hit lists, K2 SW score, K3 banded DP + backtrace on the GPU).

  e2e    reads/s of that call, wall clock (host buffers in, host buffers out; every H2D/D2H
         copy and all host stages inside the timed region) - the headline;
  value  reads/s with inputs resident in HBM: reads / sum of the device times of all kernels
         of a step (CUDA events on the launching stream, measured in a pass with ONE host
         worker so that no two streams overlap);
  roofline / roofline_k2 / roofline_k1: per kernel, against the measured integer-issue peak
         (DP kernels; no tensor/HBM bound applies) or the measured HBM bandwidth (seed lookup).

Multi-GPU: one process per GPU (torchrun), reads sharded by rank, index + reference
replicated, no collective on the data path (weak scaling); host cores are split between ranks.
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

GENOME_LEN = 5_000_000
READ_LEN = 150
K, NSKIP = 13, 6
ERR = 0.02
# integer instructions per DP cell of the recurrences as restated for two 16-bit lanes per
# register ("rooflines"): K2 6.5 ALU-pipe instructions per cell pair (LOP3, PRMT, 3 x VIADDMNMX,
# 1.5 x VIMNMX3; the eighth, H - gap_init, is an IMAD on the FMA pipe); K3: ALU-pipe instructions per
# iteration of the DP loop in the SASS (62, loop overhead included) / 4 cells (two packed cell pairs)
K2_OPS_PER_CELL = 3.25
K3_OPS_PER_CELL = 15.5


def make_genome(seed=2, n=GENOME_LEN):
    return np.random.default_rng(seed).integers(0, 4, n).astype(np.uint8)


def simulate_reads(genome, n, seed, qlen=READ_LEN, err=ERR):
    """Seeded, vectorised read simulator: substitutions (80 % of errors) and short indels
    (20 %), random strand.  -> reads[n, qlen] codes, pos[n], strand[n], span[n]"""
    rng = np.random.default_rng(seed)
    G = len(genome)
    pos = rng.integers(0, G - qlen - 64, n)
    ev = rng.random((n, qlen))
    p_indel = err * 0.2
    is_del = ev < p_indel / 2
    is_ins = (ev >= p_indel / 2) & (ev < p_indel)
    is_sub = (ev >= p_indel) & (ev < p_indel + err * 0.8)
    step = np.ones((n, qlen), np.int64)
    step[is_del] += rng.integers(1, 4, int(is_del.sum()))
    step[is_ins] = 0
    step[:, 0] = 0
    idx = pos[:, None] + np.cumsum(step, axis=1)
    reads = genome[idx]
    rnd = rng.integers(0, 4, (n, qlen)).astype(np.uint8)
    reads[is_ins] = rnd[is_ins]
    reads[is_sub] = (reads[is_sub] + 1 + rnd[is_sub] % 3) & 3
    span = idx[:, -1] - pos + 1
    strand = rng.integers(0, 2, n).astype(np.uint8)
    rc = strand == 1
    reads[rc] = 3 - reads[rc][:, ::-1]
    return np.ascontiguousarray(reads), pos, strand, span


def fastq_text(reads, first=0):
    """4-line FASTQ text of the code matrix reads[n, qlen] (names r<first+i>, quality 'I')."""
    n, qlen = reads.shape
    let = np.frombuffer(b"ACGT", np.uint8)
    names = np.char.add("@r", np.arange(first, first + n).astype(str)).astype("S")
    w = names.dtype.itemsize
    rec = np.full((n, w + 1 + qlen + 3 + qlen + 1), ord("\n"), np.uint8)
    nm = np.frombuffer(names.tobytes(), np.uint8).reshape(n, w)
    rec[:, :w] = nm                       # padded with NULs, removed below
    rec[:, w + 1:w + 1 + qlen] = let[reads]
    rec[:, w + 2 + qlen] = ord("+")
    rec[:, w + 4 + qlen:w + 4 + 2 * qlen] = ord("I")
    flat = rec.reshape(-1)
    return flat[flat != 0].tobytes()


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


def simulate_at(genome, pos, rc, seed, qlen=READ_LEN, err=ERR):
    """like simulate_reads, but at given positions / strands (rc[i]: read i is the reverse complement of
    genome[pos[i] ...]) -> reads[n, qlen]"""
    rng = np.random.default_rng(seed)
    n = len(pos)
    ev = rng.random((n, qlen))
    p_indel = err * 0.2
    is_del = ev < p_indel / 2
    is_ins = (ev >= p_indel / 2) & (ev < p_indel)
    is_sub = (ev >= p_indel) & (ev < p_indel + err * 0.8)
    step = np.ones((n, qlen), np.int64)
    step[is_del] += rng.integers(1, 4, int(is_del.sum()))
    step[is_ins] = 0
    step[:, 0] = 0
    idx = pos[:, None] + np.cumsum(step, axis=1)
    reads = genome[np.minimum(idx, len(genome) - 1)]
    rnd = rng.integers(0, 4, (n, qlen)).astype(np.uint8)
    reads[is_ins] = rnd[is_ins]
    reads[is_sub] = (reads[is_sub] + 1 + rnd[is_sub] % 3) & 3
    reads[rc] = 3 - reads[rc][:, ::-1]
    return np.ascontiguousarray(reads)


def paired_workload(tmp, npairs, nseq=4, seqlen=5_000_000, seed=3):
    """C3 scaled down (BASELINE.json configs[2]: 100 Mb genome, 5 M pairs): nseq x seqlen bases, pairs of
    2 x 150 bp from fragments of 400 +- 40 bases (forward/reverse), 2 % error, 2 % of the mates random"""
    from smalt_b200 import indexer
    rng = np.random.default_rng(seed)
    seqs = [rng.integers(0, 4, seqlen).astype(np.uint8) for _ in range(nseq)]
    genome = np.concatenate(seqs)
    pref = os.path.join(tmp, "c3")
    indexer.write_smi(pref, indexer.build_index(seqs, K, NSKIP))
    indexer.write_sma(pref, ["chr%d" % (i + 1) for i in range(nseq)], seqs)
    ins = np.clip(rng.normal(400, 40, npairs), 200, 600).astype(np.int64)
    chrom = rng.integers(0, nseq, npairs)
    start = chrom * seqlen + (rng.random(npairs) * (seqlen - 700)).astype(np.int64)
    flip = rng.integers(0, 2, npairs).astype(bool)      # which mate is the forward one
    r_fwd = simulate_at(genome, start, np.zeros(npairs, bool), seed + 1)
    r_rev = simulate_at(genome, start + ins - READ_LEN - 8, np.ones(npairs, bool), seed + 2)
    junk = rng.random(npairs) < 0.02
    r_rev[junk] = rng.integers(0, 4, (int(junk.sum()), READ_LEN)).astype(np.uint8)
    r1 = np.where(flip[:, None], r_rev, r_fwd)
    r2 = np.where(flip[:, None], r_fwd, r_rev)
    return pref, fastq_text(r1), fastq_text(r2)


def run_paired(tmp, threads, cores, npairs, ref_pairs, steps, warmup):
    """paired-end throughput through the in-process driver (smbm_map_fastq_pairs) + the reference's CPU
    `smalt map` on a sample of the same pairs"""
    from smalt_b200.mapper import Mapper
    pref, t1, t2 = paired_workload(tmp, npairs)
    opts = ["-i", "600", "-j", "200"]
    m = Mapper(pref, threads, options=opts, paired=True)
    for _ in range(warmup):
        m.map_fastq_pairs(t1, t2, copy=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        nb = m.map_fastq_pairs(t1, t2, copy=False)
    wall = (time.perf_counter() - t0) / steps
    st = m.stats.as_dict()
    sam = m.map_fastq_pairs(t1[:_nth_record(t1, 20000)], t2[:_nth_record(t2, 20000)])
    m.close()
    proper = sum(1 for ln in sam.split(b"\n") if ln and ln[:1] != b"@" and int(ln.split(b"\t", 2)[1]) & 2)
    out = {"workload": "C3 scaled to 4 x 5 Mb: %d pairs of 2 x %d bp per step, fragments 400 +- 40, %.0f%% error, 2%% random "
                       "mates, smalt index -k %d -s %d, map -i 600 -j 200" % (npairs, READ_LEN, ERR * 100, K, NSKIP),
           "e2e": {"value": 2 * npairs / wall, "unit": "reads/s", "ms_per_step": 1e3 * wall, "sam_bytes_per_step": int(nb)},
           "kernel_ms": {"k1": st["k1_ms"], "k2": st["k2_ms"], "k3": st["k3_ms"]},
           "proper_pair_fraction_sample": proper / 40000.0, "host_workers": threads}
    smalt = os.path.join(ROOT, "oracle", "_ref", "smalt")
    if os.path.exists(smalt):
        f1, f2 = os.path.join(tmp, "p1.fq"), os.path.join(tmp, "p2.fq")
        with open(f1, "wb") as f:
            f.write(t1[:_nth_record(t1, ref_pairs)])
        with open(f2, "wb") as f:
            f.write(t2[:_nth_record(t2, ref_pairs)])
        t0 = time.time()
        r = subprocess.run([smalt, "map", "-n", str(cores), "-O"] + opts + ["-o", os.path.join(tmp, "pref.sam"), pref, f1, f2],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        dt = time.time() - t0
        out["cpu_baseline"] = ({"value": 2 * ref_pairs / dt, "unit": "reads/s", "cores": cores, "kind": "reference",
                                "sample": "first %d pairs, whole `oracle/_ref/smalt map -n %d -O -i 600 -j 200` program "
                                          "(%.1f s)" % (ref_pairs, cores, dt)}
                               if r.returncode == 0 else {"value": None, "sample": "unavailable: reference failed"})
    return out


def _nth_record(text, n):
    """byte offset behind the n-th 4-line record of a FASTQ text"""
    p = 0
    for _ in range(4 * n):
        p = text.find(b"\n", p) + 1
        if p == 0:
            return len(text)
    return p


def write_index_files(tmp, genome):
    """index files (own builder, byte-identical to `smalt index -k 13 -s 6`)"""
    from smalt_b200 import indexer
    pref = os.path.join(tmp, "c2")
    ix = indexer.build_index([genome], K, NSKIP)
    indexer.write_smi(pref, ix)
    indexer.write_sma(pref, ["chr1"], [genome])
    return pref, ix


def write_workload_files(tmp, genome, reads):
    pref, ix = write_index_files(tmp, genome)
    fq = os.path.join(tmp, "reads.fq")
    with open(fq, "wb") as f:
        f.write(fastq_text(reads))
    return pref, fq, ix


def count_mapped(sam_path_or_bytes):
    data = sam_path_or_bytes if isinstance(sam_path_or_bytes, bytes) else open(sam_path_or_bytes, "rb").read()
    mapped = total = 0
    for ln in data.split(b"\n"):
        if not ln or ln[:1] == b"@":
            continue
        total += 1
        if not int(ln.split(b"\t", 2)[1]) & 4:
            mapped += 1
    return mapped, total


def run_cli(exe, threads, pref, fq, out, env=None):
    cmd = [exe, "map", "-n", str(threads), "-O", "-o", out, pref, fq]
    t0 = time.time()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    dt = time.time() - t0
    if r.returncode != 0:
        return None, "%s failed: %s" % (os.path.basename(exe), r.stderr.decode()[-200:])
    return dt, None


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


def workload_config(nreads, world=1):
    return {"workload": "C2: 5 Mb synthetic genome (uniform ACGT, seed 2), %d single-end %d bp reads per GPU, "
                        "%.0f%% error, smalt index -k %d -s %d" % (nreads, READ_LEN, ERR * 100, K, NSKIP),
            "reads_per_gpu": nreads, "gpus": world,
            "l2": "inputs larger than L2 (FASTQ text, task lists and outputs of a step are > 126 MB)"}


def run_reference_arm(args):
    """the reference's own CPU `smalt map` on this box's host cores (rank 0 only)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    smalt = os.path.join(ROOT, "oracle", "_ref", "smalt")
    if not os.path.exists(smalt):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/smalt not built"}))
        return
    genome = make_genome()
    nsample = args.ref_sample
    reads, _, _, _ = simulate_reads(genome, nsample, seed=43)
    threads = host_threads()
    rates, mapped = [], 0
    with tempfile.TemporaryDirectory() as tmp:
        pref, fq, _ = write_workload_files(tmp, genome, reads)
        out = os.path.join(tmp, "out.sam")
        for it in range(args.warmup + args.steps):
            dt, err = run_cli(smalt, threads, pref, fq, out)
            if dt is None:
                print(json.dumps({"impl": "reference", "unavailable": err}))
                return
            if it >= args.warmup:
                rates.append(nsample / dt)
        mapped, _ = count_mapped(out)
    v = float(np.mean(rates))
    sample = ("first %d reads of the C2 read set per step; whole `smalt map -n %d -O` program (index load, FASTQ "
              "parsing, SAM output)" % (nsample, threads))
    print(json.dumps({
        "impl": "reference", "metric": "mapped reads/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * nsample / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/i16 (SSE2)", "data": "synthetic",
        "config": workload_config(nsample),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference", "sample": sample,
                         "mapped_fraction": mapped / nsample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=250_000)
    ap.add_argument("--cpu-sample", type=int, default=250_000)
    ap.add_argument("--threads", type=int, default=0, help="host worker threads per GPU (0: 2 x cores / GPUs)")
    ap.add_argument("--device-block", type=int, default=32000, help="reads per launch of the device-time pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true")
    ap.add_argument("--no-paired", action="store_true")
    ap.add_argument("--pairs", type=int, default=250_000, help="pairs per step of the paired-end section")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank, world, local, dist = dist_setup()
    from smalt_b200.capi import Context
    from smalt_b200.mapper import Mapper

    n = args.reads
    cores = host_threads()
    threads = args.threads or max(1, int(round(2.0 * cores / world)))
    genome = make_genome()
    # reads are sharded by rank: rank r maps reads [r*n, (r+1)*n) of the job
    reads, pos, strand, span = simulate_reads(genome, n, seed=43 + 1000 * rank)
    text = fastq_text(reads, first=rank * n)
    tmpdir = tempfile.TemporaryDirectory()
    tmp = tmpdir.name
    pref, ix = write_index_files(tmp, genome)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- pass 1: device-resident kernel times, ONE host worker (no overlapping streams) ----
    # (blocks of DEVICE_BLOCK reads per launch: with one stream nothing else fills the tail of a launch,
    # which the e2e pass does with the launches of its other workers' streams)
    os.environ["SMALT_B200_BLOCK"] = str(args.device_block)
    m = Mapper(pref, 1)
    m.map_fastq_nocopy(fastq_text(reads[:max(1, n // 8)]))   # warm-up of this mapper
    m.map_fastq_nocopy(text)
    s1 = m.stats.as_dict()
    m.close()
    del os.environ["SMALT_B200_BLOCK"]
    dev_ms = s1["k1_ms"] + s1["k2_ms"] + s1["k3_ms"]

    # ---- pass 2: e2e through the in-process driver, all host workers ----
    m = Mapper(pref, threads)
    for _ in range(args.warmup):
        m.map_fastq_nocopy(text)
    c0 = m.stats.as_dict()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    sam_bytes = 0
    for _ in range(args.steps):
        sam_bytes = m.map_fastq_nocopy(text)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    c1 = m.stats.as_dict()
    sam = m.map_fastq(fastq_text(reads[:max(1, n // 50)]))
    mapped, total = count_mapped(sam)
    m.close()

    wall_ms = 1e3 * wall / args.steps
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
    if rank != 0:
        return

    value = world * n / (dev_ms * 1e-3)
    e2e = world * n / (wall_ms * 1e-3)
    launches = (c1["gpu_launches"] - c0["gpu_launches"]) // args.steps
    h2d = (c1["h2d_bytes"] - c0["h2d_bytes"]) // args.steps
    d2h = (c1["d2h_bytes"] - c0["d2h_bytes"]) // args.steps
    k2_gcups = s1["k2_cells"] / (s1["k2_ms"] * 1e-3) / 1e9
    k3_gcups = s1["k3_cells"] / (s1["k3_ms"] * 1e-3) / 1e9
    ctx = Context(local)
    # giga thread-instructions/s: VIADDMNMX, VIMNMX3, IADD+IMNMX pairs (ops), VIADDMNMX.S16x2, VIMNMX3.S16x2
    peaks = ctx.int_peak()
    ctx.close()
    k2_peak = peaks[3] / K2_OPS_PER_CELL
    k3_peak = peaks[3] / K3_OPS_PER_CELL
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    # K1 algorithmic bytes: per lookup 8 (idx pair) + 4*ceil(log2(bucket+1)) (wordidx probes) +
    # 8 (posidx pair); per hit 4 (pos) + 8 (sqdat) + 16 (sort)
    nlook = 2 * n * (READ_LEN - K + 1)
    bucket = max(1.0, ix["nwords"] / ix["nkeys"])
    k1_bytes = nlook * (8 + 4 * np.ceil(np.log2(bucket + 1)) + 8)
    k1_gbs = k1_bytes / (s1["k1_ms"] * 1e-3) / 1e9
    kernel_ms = {"k1_seed_hits": s1["k1_ms"], "k2_sw_score": s1["k2_ms"], "k3_band_align": s1["k3_ms"]}
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch (one block of 32000 reads) from the
    # ncu --set full captures under profiles/ (r1c_ncu_full_k1/k2/k3_raw_selected.csv) - K2 and K3 are
    # bound by the integer ALU pipe (sm__pipe_alu_cycles_active 93 % / 72 %), not by memory
    roof_k3 = {"kernel": "band_pack_kernel (K3: banded DP + backtrace, 4 tasks per warp)", "bound": "alu",
               "achieved": k3_gcups, "peak": k3_peak, "unit": "GCUPS", "frac": k3_gcups / k3_peak if k3_peak else None,
               "traffic": 13.1e6,
               "note": "integer-issue bound (no tensor/HBM bound applies to this DP): peak = measured VIADDMNMX.S16x2 "
                       "issue rate %.0f G thread-instr/s / %.1f ALU-pipe instructions per cell (62 per DP-loop iteration "
                       "of two packed cell pairs in the SASS); cells include staging and backtrace time" % (peaks[3], K3_OPS_PER_CELL)}
    roof_k2 = {"kernel": "sw_score2_kernel (K2: SW score, 2 tasks per warp)", "bound": "alu", "achieved": k2_gcups,
               "peak": k2_peak, "unit": "GCUPS", "frac": k2_gcups / k2_peak if k2_peak else None, "traffic": 11.8e6,
               "note": "DPX issue bound: peak = measured VIADDMNMX.S16x2 issue rate %.0f G thread-instr/s / %.1f "
                       "ALU-pipe instructions per cell (6.5 per packed cell pair)" % (peaks[3], K2_OPS_PER_CELL)}
    roof_k1 = {"kernel": "seed_kernel + hits_kernel (K1)", "bound": "hbm", "achieved": k1_gbs, "peak": hbm_peak,
               "unit": "GB/s", "frac": k1_gbs / hbm_peak, "traffic": 37.3e6, "peak_source": peak_src,
               "note": "dependent 4-byte index probes (latency bound); the 5 Mb index (11 MB) is L2 resident"}
    dominant = max(kernel_ms, key=kernel_ms.get)
    line = {
        "metric": "mapped reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i16x2 DPX (K2, K3), u32/u64 (K1)", "data": "synthetic",
        "config": dict(workload_config(n, world), host_workers_per_gpu=threads, host_cores=cores),
        "timing": "value: reads / sum of CUDA-event kernel times of a step (one host worker = one stream, "
                  "inputs resident, %d reads per launch); " % args.device_block +
                  "e2e: wall clock of smbm_map_fastq (FASTQ text in host memory -> SAM text in host memory), "
                  "max over ranks",
        "device_ms_per_step": dev_ms, "kernel_ms": kernel_ms, "sw_gcups": k2_gcups, "band_gcups": k3_gcups,
        "tasks_per_step": {"k2_tasks": s1["k2_tasks"], "k2_cells": s1["k2_cells"], "k3_tasks": s1["k3_tasks"],
                           "k3_cells": s1["k3_cells"]},
        "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": wall_ms, "sam_bytes_per_step": int(sam_bytes),
                "host_stage_wall_s": c1["host_stage_s"], "host_stage_cpu_s": c1["host_cpu_s"]},
        "gpu_launches": int(launches), "clocks": clocks,
        "int_peaks_ginstr": {"viaddmnmx": peaks[0], "vimnmx3": peaks[1], "iadd_imnmx_ops": peaks[2],
                             "viaddmnmx_s16x2": peaks[3], "vimnmx3_s16x2": peaks[4]},
        "roofline": {"k3_band_align": roof_k3, "k2_sw_score": roof_k2, "k1_seed_hits": roof_k1}[dominant],
        "roofline_k2": roof_k2, "roofline_k3": roof_k3, "roofline_k1": roof_k1,
        "mapped_fraction": mapped / max(total, 1),
    }
    if world == 1 and not args.no_cli:
        # the same job as a whole program, like the reference arm runs it (process start, CUDA
        # start-up, index load, file I/O included)
        fq = os.path.join(tmp, "reads.fq")
        with open(fq, "wb") as f:
            f.write(text)
        exe = os.path.join(ROOT, "smalt_b200", "bin", "smalt_b200")
        # (the driver is still tearing down the contexts of the passes above when this process's mappers
        # are closed: a program started right then waits seconds in CUDA start-up; two runs, the faster)
        runs = []
        for _ in range(2):
            time.sleep(2.0)
            dt, err = run_cli(exe, cores, pref, fq, os.path.join(tmp, "cli.sam"))
            if not dt:
                break
            runs.append(dt)
        dt = min(runs) if runs else None
        line["e2e_cli"] = ({"value": n / dt, "unit": "reads/s", "seconds": dt, "runs_seconds": runs,
                            "what": "whole `smalt_b200 map -n %d -O` program on the same %d reads" % (cores, n)}
                           if dt else {"value": None, "unavailable": err})
    if world == 1 and not args.no_cpu_baseline:
        ns = min(args.cpu_sample, n)
        smalt = os.path.join(ROOT, "oracle", "_ref", "smalt")
        fq = os.path.join(tmp, "sample.fq")
        with open(fq, "wb") as f:
            f.write(fastq_text(reads[:ns]))
        dt, err = run_cli(smalt, cores, pref, fq, os.path.join(tmp, "ref.sam")) if os.path.exists(smalt) else \
            (None, "oracle/_ref/smalt not built")
        if dt:
            line["cpu_baseline"] = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "reference",
                                    "sample": "first %d reads of this workload, whole `oracle/_ref/smalt map -n %d "
                                              "-O` program (%.1f s)" % (ns, cores, dt)}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": cores, "kind": "reference",
                                    "sample": "unavailable: " + err}
    if world == 1 and not args.no_paired:
        # configs[2] of BASELINE.json (paired-end, insert sizes) scaled to one GPU and a few seconds: not the
        # headline metric, reported next to it
        try:
            time.sleep(2.0)   # (let the driver finish tearing down the whole-program runs above)
            line["paired"] = run_paired(tmp, threads, cores, args.pairs, min(50_000, args.pairs), 3, 3)
        except Exception as exc:   # the headline line must not be lost
            line["paired"] = {"unavailable": repr(exc)[:300]}
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        if "torch.distributed" in sys.modules:
            import torch.distributed as _d
            if _d.is_available() and _d.is_initialized():
                _d.destroy_process_group()
