/* rmap_wave.h - batched single-end mapping (rmap_wave.c) */
#ifndef SMALT_B200_RMAP_WAVE_H
#define SMALT_B200_RMAP_WAVE_H
#include "rmap.h"
typedef struct RmapWave_ RmapWave;
typedef int (RMAPWAVE_EMITF)(void *user, int i, const ResultSet *rsltp);
RmapWave *rmapWaveCreate(const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
			 const ScoreMatrix *scormtxp);
void rmapWaveDelete(RmapWave *w);
/* SMB_CIGAR_* flags (include/smalt_b200.h) for the blocks of single-end reads that follow: the device also
 * emits the CIGAR text and the NM edit distance of every alignment (0: off) */
void rmapWaveSetCigarMode(int flags);
/* ms: device time of K1, K2(+K2'), K3; counts: reads, K2 tasks, K2 cells, K3 tasks, K3 cells */
void rmapWaveGetStats(const RmapWave *w, double ms[3], uint64_t counts[5]);
/* host wall-clock seconds per stage: staging, seed, hits, candidates, score, replay, align, results */
void rmapWaveGetWall(const RmapWave *w, double wall[11]);
/* the part of ms[0] spent in candidate selection, task lists and the score replay (block.cu) */
double rmapWaveGetCandMs(const RmapWave *w);
/* thread CPU seconds of the eight stages */
void rmapWaveGetCpu(const RmapWave *w, double cpu[8]);
/* Maps reads[0..n) (SEQCOD_MANGLED) like n calls of rmapSingle (rmap.c:1648) would and calls
 * emitf(user, i, result set) for i = 0..n-1 in order.  Returns ERRCODE_ARGINVAL if the flag
 * combination is not handled by the wave path (the caller then uses rmapSingle per read). */
int rmapSingleWave(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, int n, SeqFastq **reads,
		   const uint32_t *min_cover_arr, int ktuple_maxhit, int min_swatscor,
		   int min_swatscor_below_max, unsigned char min_basqval, short target_depth, short max_depth,
		   RMAPFLG_t rmapflg, const ScoreMatrix *scormtxp, const ResultFilter *rsfp,
		   const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
		   RMAPWAVE_EMITF *emitf, void *user);
/* Device batches shared by several worker threads (rmap_wave.c, "combined device batches"): nslots batches
 * in flight, a batch closes at target_reads reads or as soon as a device thread is free; spin: the device
 * threads poll instead of sleeping (worth a core each when there are many). */
typedef struct WaveCombiner_ WaveCombiner;
WaveCombiner *waveCombinerCreate(const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
				 const ScoreMatrix *scormtxp, int nslots, int target_reads, int spin);
void waveCombinerDelete(WaveCombiner *wc);
/* the waves of the batch slots (for the statistics) -> number of slots; counts: batches run, reads in them */
void waveCombinerSetRunning(WaveCombiner *wc, int n);
int waveCombinerSlots(const WaveCombiner *wc, RmapWave **waves, uint64_t counts[2]);
/* The three stages of a block that goes through the combiner (see rmap_wave.c): a worker delivers its reads
 * (waveCombinerDeliver, 1 = no slot free right now), a device thread runs closed batches (waveCombinerRunNext ->
 * slot or -1 after waveCombinerFlush(stop)), a worker turns its share of a finished batch into results. */
typedef struct { int slot, first, n; } WaveTicket;
int waveCombinerTakes(const WaveCombiner *wc, int n, SeqFastq **reads, RMAPFLG_t rmapflg, const HashTable *htp);
int waveCombinerDeliver(WaveCombiner *wc, int n, SeqFastq **reads, const uint32_t *min_cover_arr, int min_swatscor,
			const ScoreMatrix *scormtxp, const HashTable *htp, WaveTicket *tk);
void waveCombinerFlush(WaveCombiner *wc, int stop);
void waveCombinerRestart(WaveCombiner *wc);
int waveCombinerRunNext(ErrMsg *errmsgp, WaveCombiner *wc, int ktuple_maxhit, int min_swatscor_below_max,
			unsigned char min_basqval, short target_depth, short max_depth, RMAPFLG_t rmapflg,
			const ScoreMatrix *scormtxp, SeqFastq *any_read, const SeqSet *ssp, int *errcode);
int waveCombinerResults(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, WaveCombiner *wc, const WaveTicket *tk, SeqFastq **reads,
			const uint32_t *min_cover_arr, int min_swatscor, short max_depth, RMAPFLG_t rmapflg,
			const ScoreMatrix *scormtxp, const ResultFilter *rsfp, const SeqSet *ssp, const SeqCodec *codecp,
			RMAPWAVE_EMITF *emitf, void *user);
/* Paired reads: reads[2p] / reads[2p+1] = read and mate of pair p (mincov likewise).  On return
 * status[p] is RMAPPAIR_DONE (finish with rmapPairWaveFinish) or RMAPPAIR_FALLBACK (map the pair
 * with the reference's rmapPair).  ERRCODE_ARGINVAL: flag combination not handled here. */
enum { RMAPPAIR_DONE = 0, RMAPPAIR_FALLBACK = 1 };
int rmapPairWave(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, int npairs, SeqFastq **reads, const uint32_t *mincov,
		 int d_min, int d_max, RSLTPAIRLIB_t pairlibcode, int ktuple_maxhit, int min_swatscor,
		 unsigned char min_basqval, short target_depth, short max_depth, RMAPFLG_t rmapflg,
		 const ScoreMatrix *scormtxp, const HashTable *htp, const SeqSet *ssp, const SeqCodec *codecp,
		 unsigned char *status);
int rmapPairWaveFinish(ErrMsg *errmsgp, RMap *rmp, RmapWave *w, int p, int d_min, int d_max,
		       RSLTPAIRLIB_t pairlibcode, const ResultFilter *rsfp, SeqFastq *readp, SeqFastq *matep,
		       const ResultSet **rsltp, const ResultSet **rslt_matep, const ResultPairs **pairp,
		       RSLTPAIRFLG_t *pairflg);
/* pairs seen, pairs left to rmapPair, pairs with a third pass, pairs with a fourth pass */
void rmapWaveGetPairStats(const RmapWave *w, uint64_t counts[4]);
#endif
