#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native SMALT hot path.

Metric (BASELINE.json): mapped reads/s (and SW GCUPS) on config C2 - 5 Mb synthetic
genome, 1 M single-end 150 bp reads, k=13 s=6 - next to the reference's own CPU `smalt`
timed on this box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R]
  python bench.py --impl reference ...      # the reference CPU arm (oracle/_ref/smalt)

One "step" = one pass of the hot path over one batch of R reads (default 1 M per GPU):
K1 seed lookup + ranking + hit lists, K2 SW score, K3 banded DP + backtrace.
`value` is reads/s with inputs resident in HBM (sum of CUDA-event kernel times);
`e2e` is reads/s through the C ABI with host buffers (H2D of reads/tasks and D2H of every
result inside the timed region, wall clock between device synchronisations).
Multi-GPU: one process per GPU (torchrun), reads sharded, index + reference replicated,
no collective on the data path (weak scaling).
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
WINDOW_PAD = 21            # reference window = read + band margins (150 x 171)
K2_DECOYS, K3_DECOYS = 3.42, 1.55   # reference-measured calls/read beyond the true locus
OPS_PER_CELL = 7.5         # algorithmic integer ops per DP cell of K2


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


def plan_tasks(reads_pos, strand, n, seed, qlen=READ_LEN, G=GENOME_LEN):
    """PLACEHOLDER task planner (round 1): the true-locus window of every read plus random
    decoy windows at the per-read rates the reference executes on this workload.  It is
    replaced by the real candidate selection once that stage is built."""
    from smalt_b200.capi import BAND_TASK_DTYPE, SW_TASK_DTYPE
    rng = np.random.default_rng(seed)
    wl = qlen + WINDOW_PAD
    wstart = np.clip(reads_pos - 10, 0, G - wl)
    nd2 = rng.poisson(K2_DECOYS, n)
    nd3 = rng.poisson(K3_DECOYS, n)
    def build(nd):
        owner = np.concatenate([np.arange(n), np.repeat(np.arange(n), nd)])
        ws = np.concatenate([wstart, rng.integers(0, G - wl, int(nd.sum()))])
        st = np.concatenate([strand, rng.integers(0, 2, int(nd.sum())).astype(np.uint8)])
        return owner, ws, st
    o2, w2, s2 = build(nd2)
    sw = np.zeros(len(o2), SW_TASK_DTYPE)
    sw["read_off"] = o2.astype(np.uint64) * qlen
    sw["ref_off"] = w2
    sw["read_len"] = qlen
    sw["ref_len"] = wl
    sw["flags"] = 2 | s2
    o3, w3, s3 = build(nd3)
    bt = np.zeros(len(o3), BAND_TASK_DTYPE)
    bt["read_off"] = o3.astype(np.uint64) * qlen
    bt["ref_off"] = w3
    bt["read_len"] = qlen
    bt["ref_len"] = wl
    bt["flags"] = 2 | s3
    off = np.concatenate([reads_pos - wstart, np.full(len(o3) - n, 10)])
    bt["l_edge"] = -off - 9
    bt["r_edge"] = -off + 9
    bt["p_left"], bt["p_right"] = 0, qlen - 1
    bt["u_left"], bt["u_right"] = 0, wl - 1
    bt["minscore"], bt["minscorlen"] = 50, 30
    return sw, bt


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


def write_workload_files(tmp, genome, reads):
    """index files (own builder, byte-identical to `smalt index`) + FASTQ for the CPU arm"""
    from smalt_b200 import indexer
    pref = os.path.join(tmp, "c2")
    ix = indexer.build_index([genome], K, NSKIP)
    indexer.write_smi(pref, ix)
    indexer.write_sma(pref, ["chr1"], [genome])
    let = np.frombuffer(b"ACGT", np.uint8)
    fq = os.path.join(tmp, "reads.fq")
    qual = "I" * reads.shape[1]
    with open(fq, "w") as f:
        for i in range(len(reads)):
            f.write("@r%d\n%s\n+\n%s\n" % (i, let[reads[i]].tobytes().decode(), qual))
    return pref, fq, ix


def cpu_reference_rate(genome, reads, threads):
    """reads/s of the UNMODIFIED reference (`oracle/_ref/smalt map -n threads -O`) on `reads`."""
    smalt = os.path.join(ROOT, "oracle", "_ref", "smalt")
    if not os.path.exists(smalt):
        return None, "oracle/_ref/smalt not built"
    with tempfile.TemporaryDirectory() as tmp:
        pref, fq, _ = write_workload_files(tmp, genome, reads)
        cmd = [smalt, "map", "-n", str(threads), "-O", "-o", os.path.join(tmp, "out.sam"), pref, fq]
        t0 = time.time()
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        dt = time.time() - t0
        if r.returncode != 0:
            return None, "smalt map failed: " + r.stderr.decode()[-200:]
        mapped = 0
        with open(os.path.join(tmp, "out.sam")) as f:
            for ln in f:
                if ln[0] != "@" and not (int(ln.split("\t", 2)[1]) & 4):
                    mapped += 1
    return dict(rate=len(reads) / dt, seconds=dt, mapped=mapped), None


def host_threads():
    try:
        return max(1, min(len(os.sched_getaffinity(0)), 64))
    except AttributeError:
        return max(1, min(os.cpu_count() or 1, 64))


def dist_setup(n_gpus):
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


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    genome = make_genome()
    nsample = args.ref_sample
    reads, _, _, _ = simulate_reads(genome, nsample, seed=43)
    threads = host_threads()
    rates, mapped = [], 0
    for it in range(args.warmup + args.steps):
        res, err = cpu_reference_rate(genome, reads, threads)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": err}))
            return
        if it >= args.warmup:
            rates.append(res["rate"])
            mapped = res["mapped"]
    v = float(np.mean(rates))
    sample = "first %d reads of the C2 read set per step, smalt map -n %d -O incl. index load" % (nsample, threads)
    print(json.dumps({
        "impl": "reference", "metric": "mapped reads/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * nsample / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/i16 (SSE2)", "data": "synthetic",
        "config": workload_config(nsample),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference", "sample": sample,
                         "mapped_fraction": mapped / nsample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(nreads):
    return {"workload": "C2: 5 Mb synthetic genome (uniform ACGT, seed 2), %d single-end %d bp reads, %.0f%% error, "
                        "smalt index -k %d -s %d" % (nreads, READ_LEN, ERR * 100, K, NSKIP),
            "reads_per_gpu": nreads, "l2": "inputs larger than L2 (reads + task lists + outputs > 126 MB)",
            "tasks": "PLACEHOLDER planner: true-locus window + decoys at the reference's measured call rates "
                     "(K2 4.42/read, K3 2.55/read); K1 runs on every read x strand"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=100_000)
    ap.add_argument("--cpu-sample", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank, world, local, dist = dist_setup(args.gpus)
    import smalt_b200
    from smalt_b200 import indexer
    from smalt_b200.capi import HIT_REQ_DTYPE
    from smalt_b200.seqpack import pack3

    n = args.reads
    genome = make_genome()
    reads, pos, strand, span = simulate_reads(genome, n, seed=43 + 1000 * rank)
    ix = indexer.as_loaded(indexer.build_index([genome], K, NSKIP))
    ctx = smalt_b200.Context(local)
    ctx.index_upload(ix)
    words = pack3(np.concatenate([genome, np.array([7], np.uint8)]))
    ctx.refseq_upload(words, len(genome) + 1, np.array([0, len(genome)], np.uint64))
    sw, bt = plan_tasks(pos, strand, n, seed=7 + rank)
    arena = reads.reshape(-1)
    read_off = np.arange(n, dtype=np.uint64) * READ_LEN
    read_len = np.full(n, READ_LEN, np.uint32)
    req = np.zeros(2 * n, HIT_REQ_DTYPE)
    req["lo"], req["hi"] = 0, len(genome)
    req["read"] = np.repeat(np.arange(n, dtype=np.uint32), 2)
    req["strand"] = np.tile(np.array([0, 1], np.uint8), n)
    req["nhit_max"], req["use_short"] = 10000, 1
    k2_cells = float((sw["read_len"].astype(np.float64) * sw["ref_len"]).sum())

    def barrier():
        if dist is not None:
            dist.barrier()

    def step():
        """one pass of the hot path through the C ABI with host buffers"""
        kms = {}
        ctx.arena_upload(arena)
        info, _ = ctx.seed_batch(read_off, read_len, None, 10000, 16384, 0, full=False)
        kms["k1_seed"] = ctx.last_kernel_ms
        sq, first, herr = ctx.hits_batch(req, max_hits=48 * n)
        kms["k1_hits"] = ctx.last_kernel_ms
        scores, serr = ctx.sw_score(sw)
        kms["k2"] = ctx.last_kernel_ms
        res, rfirst, diff, berr, cells = ctx.band_align(bt, max_results=2 * len(bt), max_diff=24 * len(bt))
        kms["k3"] = ctx.last_kernel_ms
        d2h = info.nbytes + sq.nbytes + first.nbytes + scores.nbytes + serr.nbytes + res.nbytes + diff.nbytes + \
            rfirst.nbytes + berr.nbytes
        h2d = arena.nbytes + read_off.nbytes + read_len.nbytes + req.nbytes + sw.nbytes + bt.nbytes
        out = dict(kms=kms, cells=cells, nhits=len(sq), h2d=h2d, d2h=d2h, nres=len(res),
                   mapped=int((scores[:n] >= 50).sum()))
        return out

    for _ in range(args.warmup):
        step()
    launches0 = ctx.total_kernel_launches
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    outs = [step() for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.total_kernel_launches - launches0

    kms_tot = {k: float(np.mean([o["kms"][k] for o in outs])) for k in outs[0]["kms"]}
    dev_ms = sum(kms_tot.values())           # device time of the kernels of one step (CUDA events)
    wall_ms = 1e3 * wall / args.steps
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0]), float(t[1])
    if rank != 0:
        return
    o = outs[-1]
    value = world * n / (dev_ms * 1e-3)
    e2e = world * n / (wall_ms * 1e-3)
    k2_gcups = k2_cells / (kms_tot["k2"] * 1e-3) / 1e9
    k3_gcups = o["cells"] / (kms_tot["k3"] * 1e-3) / 1e9
    peaks = ctx.int_peak()
    peak_gcups = peaks[0] / OPS_PER_CELL
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, peak_src = 6650.0, "fallback"
    # K1 algorithmic bytes: per lookup 8 (idx pair) + 4*ceil(log2(bucket+1)) + 8 (posidx pair);
    # per hit 4 (pos) + 8 (sqdat) + 16 (sort)
    nlook = 2 * n * (READ_LEN - K + 1)
    bucket = max(1.0, ix["nwords"] / ix["nkeys"])
    k1_bytes = nlook * (8 + 4 * np.ceil(np.log2(bucket + 1)) + 8) + o["nhits"] * 28.0
    k1_gbs = k1_bytes / ((kms_tot["k1_seed"] + kms_tot["k1_hits"]) * 1e-3) / 1e9
    line = {
        "metric": "mapped reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32 DPX (K2), i16x2 (K3), u32/u64 (K1)", "data": "synthetic",
        "config": workload_config(n),
        "timing": "value: CUDA events around the kernels of a step (inputs resident); e2e: wall clock between "
                  "synchronisations incl. H2D/D2H through the C ABI; max over ranks",
        "kernel_ms": kms_tot, "sw_gcups": k2_gcups, "band_gcups": k3_gcups,
        "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": int(o["h2d"]),
                "d2h_bytes_per_step": int(o["d2h"]), "ms_per_step": wall_ms},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "sw_score_kernel (K2)", "bound": "alu", "achieved": k2_gcups, "peak": peak_gcups,
                     "unit": "GCUPS", "frac": k2_gcups / peak_gcups if peak_gcups else None, "traffic": None,
                     "note": "no tensor/HBM bound applies: integer DPX issue bound; peak = measured VIADDMNMX "
                             "rate %.0f Gop/s / %.1f ops per cell" % (peaks[0], OPS_PER_CELL)},
        "roofline_hbm": {"kernel": "seed_kernel + hits_kernel (K1)", "bound": "hbm", "achieved": k1_gbs,
                         "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak, "traffic": None,
                         "peak_source": peak_src,
                         "note": "latency bound dependent 4-byte loads; the 5 Mb index is L2 resident"},
        "mapped_fraction": o["mapped"] / n,
    }
    if not args.no_cpu_baseline and world == 1:
        ns = min(args.cpu_sample, n)
        threads = host_threads()
        res, err = cpu_reference_rate(genome, reads[:ns], threads)
        if res is not None:
            line["cpu_baseline"] = {"value": res["rate"], "unit": "reads/s", "cores": threads, "kind": "reference",
                                    "sample": "first %d reads of this workload, oracle/_ref/smalt map -n %d -O "
                                              "(whole program incl. index load, %.1f s)" % (ns, threads, res["seconds"])}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": threads, "kind": "reference",
                                    "sample": "unavailable: " + err}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
