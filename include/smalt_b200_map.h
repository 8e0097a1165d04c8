/* smalt_b200_map.h - in-process C ABI of the smalt_b200 mapping driver (libsmalt_b200_map.so).
 *
 * The library is the `smalt_b200 map` program without its main(): the reference's own driver
 * code (option parsing menu.c, index loading hashidx.c/sequence.c, candidate selection
 * segment.c, result post-processing results.c, SAM formatting report.c - compiled from the
 * reference tree, unchanged) around the B200 hot path of smalt_b200.h.  It replaces, for a
 * caller that already holds the reads in memory, the command line
 *     smalt map -n <nthreads> -O [options] <index_prefix> <reads.fq>     (smalt.c:1482 main,
 *     smalt.c:1316 mapReads)
 * and returns the same SAM records the reference prints (everything but the @-header lines,
 * which smbm_sam_header gives separately).
 *
 * One mapper per process (the reference driver keeps global state, threads.c) - multi-GPU
 * runs use one process per GPU (device = LOCAL_RANK or SMALT_B200_DEVICE).
 * All functions return 0 or an error code of smalt_b200.h / the reference's elib.h.
 */
#ifndef SMALT_B200_MAP_H
#define SMALT_B200_MAP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smbm_mapper smbm_mapper;

typedef struct {
  uint64_t n_reads;          /* reads mapped by the last smbm_map_fastq call */
  double wall_s;             /* its wall-clock time */
  double k1_ms, k2_ms, k3_ms;/* device time of the kernels (CUDA events, summed over the worker streams) */
  uint64_t k2_tasks, k2_cells, k3_tasks, k3_cells;
  uint64_t gpu_launches;     /* kernels launched by this process so far */
  uint64_t h2d_bytes, d2h_bytes; /* bytes copied host->device / device->host by this process so far */
  double host_stage_s[12];   /* summed over worker threads: staging, seed, hits, candidates, score,
				replay, align, results, parse, and inside results: add, sort+filter, emit */
  double host_cpu_s[8];      /* thread CPU seconds of the first eight stages (no waiting for the GPU) */
  double cand_ms;            /* the part of k1_ms spent in candidate selection / task lists / score replay */
  uint64_t cigar_dev, cigar_host; /* SAM records of the last call whose CIGAR / NM came from the device's
				     output stage / were formatted by the reference's diffstr.c on the host */
} smbm_stats;

/* Loads <index_prefix>.smi/.sma (hashTableRead hashidx.c:1257, seqSetReadBinFil sequence.c:2521),
 * uploads them to the GPU and starts `nthreads` worker threads' worth of state.
 * options: further `smalt map` command-line options as separate strings, e.g. {"-S","match=2"}. */
int smbm_open(smbm_mapper **m, const char *index_prefix, int nthreads, int noptions,
	      const char *const *options);

/* Maps the reads of a FASTQ (4-line) or FASTA text buffer; *sam points to the SAM records in
 * input order (owned by the mapper, valid until the next call). */
int smbm_map_fastq(smbm_mapper *m, const char *fastq, size_t nbytes, const char **sam, size_t *sam_len,
		   smbm_stats *stats);

/* Paired-end reads (rmapPair, rmap.c:1744): a mapper opened with smbm_open_paired (the options
 * carry the insert size range, e.g. {"-i","600","-j","200"}) maps record i of `fastq` with record
 * i of `fastq_mates` (two plain 4-line FASTQ texts with the same number of records) and returns
 * both SAM records of every pair, in input order. */
int smbm_open_paired(smbm_mapper **m, const char *index_prefix, int nthreads, int noptions,
		     const char *const *options);
int smbm_map_fastq_pairs(smbm_mapper *m, const char *fastq, size_t nbytes, const char *fastq_mates,
			 size_t nbytes_mates, const char **sam, size_t *sam_len, smbm_stats *stats);

/* the @HD/@SQ/@PG header lines the reference writes (report.c writeSAMHeaderf); free with smbm_free */
int smbm_sam_header(smbm_mapper *m, char **text, size_t *len);
void smbm_free(void *p);

int smbm_close(smbm_mapper *m);

/* Host-only helper (no GPU, no mapper): the byte offsets at which the block-parallel pipeline
 * would cut `text` into blocks of about chunk_bytes - every offset is the start of a record
 * (4-line FASTQ or FASTA).  Returns ERRCODE_FASTA (6) if the text is not whole 4-line FASTQ
 * records.  Used by the tests and by the multi-GPU read sharding. */
int smbm_split_blocks(const char *text, size_t nbytes, size_t chunk_bytes, size_t *starts,
		      size_t max_starts, size_t *nstarts, size_t *nrecords);

#ifdef __cplusplus
}
#endif
#endif
