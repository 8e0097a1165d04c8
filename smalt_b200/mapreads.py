"""`python -m smalt_b200.mapreads` - multi-GPU `smalt map` of a FASTQ file: one process per
GPU (launch with torchrun), reads sharded by rank, SAM merged in input order by offset writes into the output file.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m smalt_b200.mapreads \
      [-n threads_per_gpu] [-i max_insert -j min_insert] -o out.sam <index_prefix> <reads.fq> [<mates.fq>]
With a mate file the pairs are sharded (both files cut at the same record numbers) and mapped
with the paired-end path (rmapPair).
"""
import argparse
import os
import sys


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", type=int, default=0, help="host worker threads per GPU (0: 2 x cores / GPUs)")
    ap.add_argument("-o", required=True)
    ap.add_argument("-r", type=int, default=None, help="seed of the draw among equally good hits (smalt map -r)")
    ap.add_argument("--backend", default=None, help="torch.distributed backend (default nccl; gloo for tests)")
    ap.add_argument("-i", type=int, default=None, help="maximum insert size (smalt map -i)")
    ap.add_argument("-j", type=int, default=None, help="minimum insert size (smalt map -j)")
    ap.add_argument("index")
    ap.add_argument("reads")
    ap.add_argument("mates", nargs="?", default=None)
    args = ap.parse_args(argv)
    from .mapper import Mapper
    from .shard import merge_to_file, pair_shard_of, shard_of
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        backend = args.backend or "nccl"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    text = open(args.reads, "rb").read()
    cores = len(os.sched_getaffinity(0))
    opts = [] if args.r is None else ["-r", str(args.r)]
    for flag, val in (("-i", args.i), ("-j", args.j)):
        if val is not None:
            opts += [flag, str(val)]
    nthreads = args.n or max(1, int(2 * cores / world))
    if args.mates is None:
        mine = shard_of(text, rank, world)
        m = Mapper(args.index, nthreads, opts)
        header = m.sam_header() if rank == 0 else b""
        sam = m.map_fastq(mine)
    else:
        mine1, mine2 = pair_shard_of(text, open(args.mates, "rb").read(), rank, world)
        m = Mapper(args.index, nthreads, opts, paired=True)
        header = m.sam_header() if rank == 0 else b""
        sam = m.map_fastq_pairs(mine1, mine2) if mine1 else b""
    m.close()
    if dist is None:
        with open(args.o, "wb") as f:
            f.write(header)
            f.write(sam)
    else:
        if rank == 0 and os.path.exists(args.o):
            os.unlink(args.o)
        dist.barrier()
        merge_to_file(dist, sam, args.o, header=header)   # every rank writes its records at its offset
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
