// api_block.cu - smb_block_run / smb_block_fetch: a block of reads resident on the device from the
// seed tables to the alignments (include/smalt_b200.h, "resident block").  Host work per block:
// sizing buffers from a few counters read back at four points, launch grids from class histograms.
#include "ctx.h"
#include "block.cuh"
#include "cigar.cuh"
#include <cmath>

namespace {

// what smb_block_fetch needs from the last smb_block_run
struct Pending {
  bool valid = false, multipass = false;
  int njobs = 0;
  size_t ncand = 0, nk3 = 0, nres = 0, ndiff = 0;
  // fast path: device pointers for the gather
  smb_ali_result *d_res = nullptr;
  uint32_t *d_nres = nullptr, *d_dused = nullptr, *d_first = nullptr;
  int32_t *d_errs = nullptr;
  uint64_t *d_diff_off = nullptr;
  unsigned long long *d_diff_first = nullptr;
  smb_block_read *d_rd = nullptr;
  smb_block_cand *d_k3c = nullptr;
  int max_res = 4;
  // debug access
  BlockArgs args{};
  // multi-pass path: results assembled on the host
  std::vector<smb_ali_result> v_res;
  std::vector<uint32_t> v_first;
  std::vector<uint8_t> v_diff;
  std::vector<int32_t> v_errs;
  // output stage (SMB_CIGAR_ON)
  bool cigar = false;
  int cigar_flags = 0;
  size_t ncig = 0;                             // text bytes
  unsigned long long *d_cig_first = nullptr;   // fast path: per task, offset of its first alignment's text
  CigarArgs cig{};                             // multi-pass path: dense results uploaded again
  unsigned long long *d_cig_off = nullptr;
};

inline size_t al256(size_t n) { return (n + 255) & ~(size_t)255; }

struct Carver {   // carves aligned arrays out of one device buffer
  char *base;
  size_t off = 0;
  explicit Carver(void *p) : base((char *)p) {}
  template <class T> T *take(size_t n) {
    T *r = (T *)(base ? base + off : nullptr);
    off += al256(n * sizeof(T));
    return r;
  }
};

// count pass of the output stage over dense / explicit alignments `ca` (len / nm / off / tile carved out of blk_cig
// behind `extra` bytes the caller uses for uploads); the total is copied to *h_total (valid after the next
// synchronisation)
static int cigar_count_pass(smb_ctx *ctx, CigarArgs *ca, unsigned long long **d_off_out, size_t extra,
                            unsigned long long *h_total, cudaStream_t st, int *nl) {
  const size_t nscan = (size_t)ca->n + 1;
  Carver c(nullptr);
  c.off = al256(extra);
  c.take<uint32_t>(nscan); c.take<int32_t>(nscan); c.take<unsigned long long>(nscan + 1);
  c.take<unsigned long long>((size_t)compact_tiles((int)nscan) + 2);
  CU(ctx->blk_cig.ensure(c.off));
  Carver d(ctx->blk_cig.p);
  d.off = al256(extra);
  ca->len = d.take<uint32_t>(nscan);
  ca->nm = d.take<int32_t>(nscan);
  unsigned long long *d_off = d.take<unsigned long long>(nscan + 1);
  unsigned long long *d_tile = d.take<unsigned long long>((size_t)compact_tiles((int)nscan) + 2);
  CU(launch_cigar_count(*ca, d_off, d_tile, st, nl));
  CU(d2h(h_total, d_off + (nscan - 1), 8, st));
  *d_off_out = d_off;
  return SMB_OK;
}

// fill pass + one copy of the blob {first[n + 1], nm[n], text[ntext]} to the host
static int cigar_fill_blob(smb_ctx *ctx, CigarArgs ca, const unsigned long long *d_off, size_t ntext, void *blob,
                           cudaStream_t st, int *nl) {
  const size_t n = (size_t)ca.n, head = (2 * n + 1) * 4;
  CU(ctx->blk_cigtext.ensure(head + ntext + 64));
  char *d_blob = ctx->blk_cigtext.as<char>();
  ca.off = d_off;
  ca.first_out = (uint32_t *)d_blob;
  ca.text = d_blob + head;
  CU(launch_cigar_fill(ca, st, nl));
  CU(cudaMemcpyAsync(d_blob + (n + 1) * 4, ca.nm, n * 4, cudaMemcpyDeviceToDevice, st));
  CU(d2h(blob, d_blob, head + ntext, st));
  return SMB_OK;
}

}  // namespace

struct BlockState { Pending p; };

void block_state_free(smb_ctx *ctx) {
  delete ctx->blk;
  ctx->blk = nullptr;
}

// Device time of a block: spans of GPU work between two events on the stream, none of them across a
// host synchronisation (the gaps in which the device waits for the host are not kernel time).
enum { SPAN_HITS = 0, SPAN_CAND = 1, SPAN_K2 = 2, SPAN_K3 = 3, SPAN_KINDS = 4 };
struct Spans {
  smb_ctx *ctx;
  int n = 0;
  int kind[SMB_BLK_EVENTS / 2];
  explicit Spans(smb_ctx *c) : ctx(c) {}
  cudaError_t begin(int k) {
    if (n >= SMB_BLK_EVENTS / 2) return cudaSuccess;   // (more spans than events: the rest goes untimed)
    kind[n] = k;
    return cudaEventRecord(ctx->blk_ev[2 * n], ctx->stream);
  }
  cudaError_t end() {
    if (n >= SMB_BLK_EVENTS / 2) return cudaSuccess;
    return cudaEventRecord(ctx->blk_ev[2 * n++ + 1], ctx->stream);
  }
  void sum(float ms[SPAN_KINDS]) const {   // after a synchronisation of the stream
    for (int k = 0; k < SPAN_KINDS; ++k) ms[k] = 0.f;
    for (int i = 0; i < n; ++i) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, ctx->blk_ev[2 * i], ctx->blk_ev[2 * i + 1]) == cudaSuccess) ms[kind[i]] += t;
      else cudaGetLastError();
    }
  }
};

extern "C" {

int smb_block_run(smb_ctx *ctx, const smb_block_params *prm, const smb_block_job *jobs, int njobs,
                  const smb_block_ival *ivals, int nivals, smb_block_sizes *sizes) {
  if (!ctx || !prm || !sizes || njobs < 0 || nivals < 0 || (njobs && !jobs) || (nivals && !ivals)) return SMB_ERR_ARG;
  memset(sizes, 0, sizeof *sizes);
  ctx->last_ms = 0.f;
  ctx->last_launches = 0;
  if (!ctx->blk) ctx->blk = new (std::nothrow) BlockState();
  if (!ctx->blk) return SMB_ERRCODE_NOMEM;
  Pending &P = ctx->blk->p;
  P.valid = false;
  if (!njobs) { P = Pending(); P.valid = true; P.cigar = (prm->cigar & SMB_CIGAR_ON) != 0; return SMB_OK; }
  if (!ctx->seed_nreads) return fail(ctx, SMB_ERR_STATE, "smb_seed_batch() first");
  if (!ctx->src.packed || !ctx->d_seq_offs || ctx->seq_offs.size() < 2)
    return fail(ctx, SMB_ERR_STATE, "smb_refseq_upload() with sequence offsets first");
  const int nseq = (int)ctx->seq_offs.size() - 1;
  const Scoring &sc = ctx->sc;
  if (sc.match - sc.mismatch < 1 || sc.gap_ext <= 0 || sc.mismatch >= 0 || sc.match - (-sc.gap_init) < 1)
    return fail(ctx, SMB_ERRCODE_ASSERT, "penalties outside the range mapSingleRead accepts (rmap.c:1266, :639)");
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  for (cudaEvent_t &e : ctx->blk_ev)
    if (!e) CU(cudaEventCreate(&e));

  // requests of every job: 2 strands x (sequences | intervals)
  const size_t nj = (size_t)njobs;
  const size_t stage_need = al256((nj + 1) * 4) + al256(sizeof(BlockCounters)) + 1024;
  CU(ctx->stage.ensure(stage_need));
  uint32_t *h_job_req = ctx->stage.as<uint32_t>();
  BlockCounters *h_cnt = (BlockCounters *)((char *)ctx->stage.p + al256((nj + 1) * 4));
  unsigned long long *h_tot = (unsigned long long *)(h_cnt + 1);   // a few totals read back
  uint64_t nreq64 = 0;
  uint32_t maxlen = 0;
  int max_lists = 0;   // hit lists of the job with the most: group size of the candidate kernel
  for (int j = 0; j < njobs; ++j) {
    if (jobs[j].seed_read >= (uint32_t)ctx->seed_nreads) return fail(ctx, SMB_ERR_ARG, "job %d: read %u out of range", j, jobs[j].seed_read);
    if (jobs[j].niv >= 0 && (uint64_t)jobs[j].iv_first + (uint64_t)jobs[j].niv > (uint64_t)nivals)
      return fail(ctx, SMB_ERR_ARG, "job %d: intervals out of range", j);
    h_job_req[j] = (uint32_t)nreq64;
    nreq64 += 2ull * (uint64_t)(jobs[j].niv < 0 ? nseq : jobs[j].niv);
    max_lists = std::max(max_lists, 2 * (jobs[j].niv < 0 ? nseq : jobs[j].niv));
    const uint32_t l = ctx->seed_len[jobs[j].seed_read];
    if (l > maxlen) maxlen = l;
  }
  if (nreq64 > 0x7fffffffull) return fail(ctx, SMB_ERR_ARG, "too many hit-list requests in one block");
  h_job_req[njobs] = (uint32_t)nreq64;
  const int nreq = (int)nreq64;
  const size_t nr = (size_t)nreq;
  for (int k = 0; k < nivals; ++k)
    if (ivals[k].seqidx < 0 || ivals[k].seqidx >= nseq) return fail(ctx, SMB_ERR_ARG, "interval %d: sequence out of range", k);

  // ---- per-job arrays ----
  const int jtiles = compact_tiles(njobs > nreq ? njobs : (nreq > 0 ? nreq : 1)) + 1;
  {
    Carver c(nullptr);
    c.take<smb_block_job>(nj); c.take<uint32_t>(nj + 1); c.take<smb_block_ival>((size_t)nivals + 1);
    c.take<uint32_t>(nj); c.take<uint32_t>(nj); c.take<uint32_t>(2 * nj); c.take<smb_block_read>(nj);
    c.take<unsigned long long>(nj + 1); c.take<unsigned long long>(nj + 1); c.take<unsigned long long>((size_t)jtiles);
    c.take<BlockCounters>(1); c.take<int>(64);
    CU(ctx->blk_jobs.ensure(c.off));
  }
  BlockArgs a{};
  Carver cj(ctx->blk_jobs.p);
  smb_block_job *d_jobs = cj.take<smb_block_job>(nj);
  uint32_t *d_job_req = cj.take<uint32_t>(nj + 1);
  smb_block_ival *d_ivals = cj.take<smb_block_ival>((size_t)nivals + 1);
  a.n_sort = cj.take<uint32_t>(nj);
  a.nk3 = cj.take<uint32_t>(nj);
  a.cover_deficit = cj.take<uint32_t>(2 * nj);
  a.rd = cj.take<smb_block_read>(nj);
  unsigned long long *d_cand_first = cj.take<unsigned long long>(nj + 1);
  unsigned long long *d_k3_first = cj.take<unsigned long long>(nj + 1);
  unsigned long long *d_tile = cj.take<unsigned long long>((size_t)jtiles);
  a.cnt = cj.take<BlockCounters>(1);
  int *d_k2_tickets = cj.take<int>(64);
  a.jobs = d_jobs; a.njobs = njobs; a.ivals = d_ivals; a.prm = *prm;
  a.ktup = ctx->seed_ix.wordlen; a.nskip = ctx->seed_ix.nskip; a.nseq = nseq;
  a.match = sc.match; a.mismatch = sc.mismatch; a.gap_init = sc.gap_init; a.gap_ext = sc.gap_ext;
  a.seq_offs = ctx->d_seq_offs;
  a.seed = ctx->seed_args;
  a.job_req = d_job_req;
  a.cand_first = d_cand_first;
  a.k3_first = d_k3_first;
  CU(h2d(d_jobs, jobs, nj * sizeof(smb_block_job), st));
  CU(h2d(d_job_req, h_job_req, (nj + 1) * 4, st));
  if (nivals) CU(h2d(d_ivals, ivals, (size_t)nivals * sizeof(smb_block_ival), st));
  CU(cudaMemsetAsync(a.cnt, 0, sizeof(BlockCounters), st));

  // ---- hit lists (as smb_hits_batch, lists stay on the device) ----
  uint32_t nhits_alloc;
  {
    const double ql = (double)ctx->seed_maxlen;
    double target = ql > 1 ? ql * log(ql) * 32.0 : 0.0;
    if (target > 2147483647.0) target = 2147483647.0;
    if (target < 8192.0) target = 8192.0;
    const size_t t = (size_t)target;
    nhits_alloc = 16384;
    if (t > nhits_alloc) nhits_alloc = (uint32_t)((t + 16383) / 16384 * 16384);
  }
  {
    Carver c(nullptr);
    c.take<smb_hit_req>(nr + 1); c.take<uint64_t>(nr + 2); c.take<uint32_t>(nr + 1); c.take<uint32_t>(nr + 1);
    c.take<int32_t>(nr + 1); c.take<int32_t>(nr + 1); c.take<uint32_t>(nr + 1); c.take<uint8_t>(nr + 1);
    c.take<unsigned long long>(2 * nj + 1);
    CU(ctx->hit_meta.ensure(c.off));
  }
  Carver ch(ctx->hit_meta.p);
  smb_hit_req *d_req = ch.take<smb_hit_req>(nr + 1);
  uint64_t *d_off = ch.take<uint64_t>(nr + 2);
  uint32_t *d_count = ch.take<uint32_t>(nr + 1);
  uint32_t *d_used = ch.take<uint32_t>(nr + 1);
  int32_t *d_rerrs = ch.take<int32_t>(nr + 1);
  int32_t *d_rseq = ch.take<int32_t>(nr + 1);
  a.req_ncand = ch.take<uint32_t>(nr + 1);
  a.req_skip = ch.take<uint8_t>(nr + 1);
  // with several reference sequences most requests are empty: mark them from the seeds' positions first
  a.seqmask = (nseq >= 8 && nseq <= 64 && !ctx->seed_args.tab) ? ch.take<unsigned long long>(2 * nj + 1) : nullptr;
  a.ix = ctx->seed_ix;
  a.req = d_req; a.req_seqidx = d_rseq; a.hit_off = d_off; a.req_err = d_rerrs;
  ctx->hit_qmask_valid = false;
  HitArgs ha{};
  ha.seed = ctx->seed_args; ha.req = d_req; ha.nreq = nreq; ha.nhits_alloc = nhits_alloc;
  ha.count = d_count; ha.maxhit_used = d_used; ha.errs = d_rerrs; ha.offset = d_off; ha.sqdat = nullptr;
  ha.list_qmask = nullptr; ha.qmask_off = nullptr;
  ha.req_skip = a.req_skip;
  int nl = 0;
  Spans sp(ctx);
  CU(sp.begin(SPAN_HITS));
  CU(launch_block_seqmask(a, st, &nl));
  CU(launch_block_reqs(a, st, &nl));
  uint64_t total_hits = 0;
  if (nreq > 0) {
    CU(launch_hits(ctx->seed_ix, ha, false, st, &nl));
    CU(launch_scan_counts(d_count, nreq, (unsigned long long *)d_off, d_tile, st, &nl));
    CU(sp.end());
    CU(d2h(h_tot, d_off + nreq, 8, st));
    CU(ctx_sync(ctx));                                                          // sync 1: total hits
    total_hits = h_tot[0];
  } else {
    CU(cudaMemsetAsync(d_off, 0, 16, st));
    CU(sp.end());
  }
  CU(ctx->hit_data.ensure((size_t)(total_hits + 1) * 8));
  ha.sqdat = ctx->hit_data.as<uint64_t>();
  a.sqdat = ha.sqdat;

  // ---- candidate selection ----
  const size_t H = (size_t)total_hits + 1;
  a.mask_words = (maxlen + 31u) / 32u + 1u;
  const size_t mask_total = a.mask_words > 8u ? nj * 32u * a.mask_words : 64;   // (CW_MASKW words per lane live in shared memory)
  {
    Carver c(nullptr);
    c.take<uint64_t>(H); c.take<int32_t>(H); c.take<uint32_t>(H); c.take<int32_t>(H); c.take<uint32_t>(H);
    c.take<SegCand>(H); c.take<uint32_t>(H); c.take<uint32_t>(H); c.take<uint32_t>(mask_total);
    CU(ctx->blk_scr.ensure(c.off));
  }
  Carver cs(ctx->blk_scr.p);
  a.sd_sqo = cs.take<uint64_t>(H); a.sd_len = cs.take<int32_t>(H); a.sg_ix = cs.take<uint32_t>(H);
  a.sg_nseed = cs.take<int32_t>(H); a.sg_cover = cs.take<uint32_t>(H); a.cand = cs.take<SegCand>(H);
  a.sort_key = cs.take<uint32_t>(H); a.sort_idx = cs.take<uint32_t>(H); a.mask = cs.take<uint32_t>(mask_total);
  CU(sp.begin(SPAN_HITS));
  if (nreq > 0 && total_hits) CU(launch_hits(ctx->seed_ix, ha, true, st, &nl));
  CU(sp.end());
  CU(sp.begin(SPAN_CAND));
  CU(launch_block_cands(a, max_lists, (double)total_hits / (double)njobs, st, &nl));
  CU(launch_scan_counts(a.n_sort, njobs, d_cand_first, d_tile, st, &nl));
  CU(sp.end());
  CU(d2h(h_tot, d_cand_first + njobs, 8, st));
  CU(d2h(h_cnt, a.cnt, sizeof(BlockCounters), st));
  CU(ctx_sync(ctx));                                                            // sync 2: candidates, K2 classes
  const size_t ncand = (size_t)h_tot[0];
  const size_t NC = ncand + 1;
  unsigned int nsw = 0;
  for (int b = 0; b < BLK_K2_BINS; ++b) a.k2_start[b] = 0;
  for (int b = 1; b <= 16; ++b) { a.k2_start[b] = nsw; nsw += h_cnt->k2_hist[b]; }
  a.k2_start[18] = nsw; nsw += h_cnt->k2_hist[18];   // long reads, paired
  const unsigned int nbf = h_cnt->k2_hist[17];
  if ((size_t)nsw + nbf != ncand) return fail(ctx, SMB_ERRCODE_ASSERT, "block: candidate classes do not add up");
  {
    Carver c(nullptr);
    c.take<DCand>(NC); c.take<uint32_t>(NC); c.take<smb_sw_task>(NC); c.take<int32_t>(NC); c.take<int32_t>(NC);
    c.take<int32_t>(NC); c.take<uint8_t>(NC); c.take<int>(NC); c.take<int>(NC); c.take<smb_band_task>(NC);
    CU(ctx->blk_cand.ensure(c.off));
  }
  Carver cc(ctx->blk_cand.p);
  a.dc = cc.take<DCand>(NC); a.dc_job = cc.take<uint32_t>(NC); a.swt = cc.take<smb_sw_task>(NC);
  a.score = cc.take<int32_t>(NC); a.serr = cc.take<int32_t>(NC); a.k3rank = cc.take<int32_t>(NC);
  a.k3cls = cc.take<uint8_t>(NC); a.k2_order = cc.take<int>(NC); a.bf_order = cc.take<int>(NC);
  a.bft = cc.take<smb_band_task>(NC);
  CU(sp.begin(SPAN_CAND));
  CU(launch_block_emit_k2(a, ncand, st, &nl));
  CU(sp.end());

  // ---- K2 (and K2' for the candidates outside the SIMD predicate) ----
  if (nsw) {
    SwPlan plan;
    for (int c = 0; c <= 16; ++c) { plan.count[c] = c ? (int)h_cnt->k2_hist[c] : 0; plan.start[c] = (int)a.k2_start[c]; }
    plan.count[SW_SLOT_LONG2] = (int)h_cnt->k2_hist[18];
    plan.start[SW_SLOT_LONG2] = (int)a.k2_start[18];
    plan.max_grid = ctx->sm_count * 8;
    plan.bstride = (h_cnt->max_rlen_multi + 31u) & ~31u;
    plan.strip_bytes = (size_t)plan.max_grid * 4u * 2u * plan.bstride * sizeof(int2);   // SW_WARPS = 4 (sw_score.cu)
    CU(ctx->scratch.ensure(plan.strip_bytes + 256));
    CU(sp.begin(SPAN_K2));
    CU(launch_sw_score(sc, ctx->src, a.swt, plan, d_k2_tickets, a.k2_order, ctx->scratch.p, a.score, a.serr, st, &nl));
    CU(sp.end());
  }
  // K2' of a list of candidates (indices in h_bf): planned on the host like smb_band_score_batch
  auto run_band_fast = [&](const std::vector<int> &h_bf) -> int {
    const int n = (int)h_bf.size();
    if (!n) return SMB_OK;
    std::vector<smb_band_task> all(ncand), sub((size_t)n);
    CU(d2h(all.data(), a.bft, ncand * sizeof(smb_band_task), st));
    CU(ctx_sync(ctx));
    for (int i = 0; i < n; ++i) sub[(size_t)i] = all[(size_t)h_bf[(size_t)i]];
    BandPlan plan;
    plan_band(sub.data(), n, false, sc, plan);
    std::vector<int> order((size_t)n);
    for (int i = 0; i < n; ++i) order[(size_t)i] = h_bf[(size_t)plan.order[(size_t)i]];
    const size_t gring_words = band_gring_words(plan);
    CU(ctx->scratch.ensure(gring_words * sizeof(uint32_t) + 512));
    CU(h2d(a.bf_order, order.data(), (size_t)n * sizeof(int), st));
    CU(ctx_sync(ctx));   // `order` goes out of scope
    unsigned long long *d_cells = &a.cnt->k2_cells_ref;   // (cells of K2' are not reported; any 8-byte slot)
    BandOut bo{nullptr, nullptr, nullptr, a.serr, d_cells, nullptr};
    CU(sp.begin(SPAN_K2));
    CU(launch_band(sc, ctx->src, a.bft, plan, a.bf_order, false, a.score, bo, 0, nullptr, nullptr, nullptr, nullptr,
                   ctx->scratch.as<uint32_t>(), ctx->ticket.as<int>(), ctx->sm_count, st, &nl));
    CU(sp.end());
    return SMB_OK;
  };
  if (nbf) {
    std::vector<int> h_bf((size_t)nbf);
    CU(d2h(h_bf.data(), a.bf_order, (size_t)nbf * sizeof(int), st));
    CU(ctx_sync(ctx));
    const int rc = run_band_fast(h_bf);
    if (rc) return rc;
    CU(cudaMemsetAsync(&a.cnt->k2_cells_ref, 0, 8, st));
  }

  // ---- replay of the sequential score bookkeeping, K3 task list ----
  size_t nk3 = 0;
  for (int round = 0;; ++round) {
    CU(sp.begin(SPAN_CAND));
    CU(launch_block_replay(a, st, &nl));
    CU(launch_scan_counts(a.nk3, njobs, d_k3_first, d_tile, st, &nl));
    CU(sp.end());
    CU(d2h(h_tot, d_k3_first + njobs, 8, st));
    CU(d2h(h_cnt, a.cnt, sizeof(BlockCounters), st));
    CU(ctx_sync(ctx));                                                          // sync 3: K3 classes and sizes
    nk3 = (size_t)h_tot[0];
    if (!h_cnt->n_exceed) break;
    if (round) return fail(ctx, SMB_ERRCODE_ASSERT, "block: SWATEXCEED after the banded fallback");
    // ERRCODE_SWATEXCEED -> banded fast variant (rmap.c:730-744), then the replay again
    CU(cudaMemsetAsync(&a.cnt->bf_cursor, 0, sizeof(unsigned int), st));
    CU(launch_block_exceed(a, ncand, st, &nl));
    CU(d2h(h_cnt, a.cnt, sizeof(BlockCounters), st));
    CU(ctx_sync(ctx));
    std::vector<int> h_bf((size_t)h_cnt->bf_cursor);
    if (!h_bf.empty()) {
      CU(d2h(h_bf.data(), a.bf_order, h_bf.size() * sizeof(int), st));
      CU(ctx_sync(ctx));
      const int rc = run_band_fast(h_bf);
      if (rc) return rc;
    }
    BlockCounters z = *h_cnt;   // counters of the replay start again; the K2 histogram is history
    memset(z.k3_hist, 0, sizeof z.k3_hist);
    z.pack_maxrows = z.pack_maxread = z.n_exceed = 0;
    z.dir_words = z.diff_bytes = z.k2_cells_ref = z.k2_tasks_ref = 0;
    *h_cnt = z;
    CU(h2d(a.cnt, h_cnt, sizeof(BlockCounters), st));
    CU(ctx_sync(ctx));
  }
  sizes->nhits = total_hits;
  sizes->ncand = ncand;
  sizes->nk2 = nsw;
  sizes->nk2_band = nbf;
  sizes->nk3 = nk3;
  sizes->k2_cells = h_cnt->k2_cells;
  sizes->k2_cells_ref = h_cnt->k2_cells_ref;
  sizes->k2_tasks_ref = h_cnt->k2_tasks_ref;

  P = Pending();
  P.cigar = (prm->cigar & SMB_CIGAR_ON) != 0;
  P.cigar_flags = prm->cigar;
  P.njobs = njobs;
  P.ncand = ncand;
  P.nk3 = nk3;
  P.d_rd = a.rd;
  P.args = a;
  const int max_res = 4;
  P.max_res = max_res;
  float ms_k3 = 0.f;
  if (nk3) {
    if (nk3 > 0x7fffffffull) return fail(ctx, SMB_ERR_ARG, "too many alignment tasks in one block");
    const int n3 = (int)nk3;
    BandPlan plan;
    unsigned int pos = 0;
    for (int b = 0; b < BLK_K3_BINS; ++b) { a.k3_start[b] = pos; pos += h_cnt->k3_hist[b]; }
    if (pos != nk3) return fail(ctx, SMB_ERRCODE_ASSERT, "block: alignment classes do not add up");
    for (int b = 0; b < BAND_CLS_THREAD_END; ++b)
      if (h_cnt->k3_hist[b]) plan.classes.push_back(BandPlan::Class{32 << b, (int)a.k3_start[b], (int)h_cnt->k3_hist[b]});
    plan.long16_start = (int)a.k3_start[BAND_CLS_LONG16]; plan.long16_count = (int)h_cnt->k3_hist[BAND_CLS_LONG16];
    plan.long32_start = (int)a.k3_start[BAND_CLS_LONG32]; plan.long32_count = (int)h_cnt->k3_hist[BAND_CLS_LONG32];
    plan.wide_start = (int)a.k3_start[BAND_CLS_WIDE]; plan.wide_count = (int)h_cnt->k3_hist[BAND_CLS_WIDE];
    plan.pack_start = (int)a.k3_start[BAND_CLS_PACK]; plan.pack_count = (int)h_cnt->k3_hist[BAND_CLS_PACK];
    plan.half_start = (int)a.k3_start[BAND_CLS_HALF]; plan.half_count = (int)h_cnt->k3_hist[BAND_CLS_HALF];
    plan.warp_start = (int)a.k3_start[BAND_CLS_WARP]; plan.warp_count = (int)h_cnt->k3_hist[BAND_CLS_WARP];
    plan.pack_maxrows = (int)h_cnt->pack_maxrows; plan.pack_maxread = (int)h_cnt->pack_maxread;
    const size_t N3 = nk3 + 1;
    const int ntiles = compact_tiles(n3);
    {
      Carver c(nullptr);
      c.take<smb_band_task>(N3); c.take<smb_block_cand>(N3); c.take<uint32_t>(N3); c.take<uint32_t>(N3); c.take<uint32_t>(N3);
      c.take<int>(N3); c.take<uint64_t>(N3 + 1); c.take<uint64_t>(N3 + 1); c.take<unsigned long long>((size_t)ntiles + 1);
      CU(ctx->blk_k3.ensure(c.off));
    }
    Carver c3(ctx->blk_k3.p);
    a.bat = c3.take<smb_band_task>(N3); a.k3c = c3.take<smb_block_cand>(N3); a.dir_words_arr = c3.take<uint32_t>(N3);
    a.diff_cap = c3.take<uint32_t>(N3); a.diff_stride = c3.take<uint32_t>(N3); a.k3_order = c3.take<int>(N3);
    uint64_t *d_dir_off = c3.take<uint64_t>(N3 + 1);
    uint64_t *d_diff_off = c3.take<uint64_t>(N3 + 1);
    unsigned long long *d_tile3 = c3.take<unsigned long long>((size_t)ntiles + 1);
    P.d_k3c = a.k3c;
    P.args = a;
    const bool too_big = h_cnt->dir_words > ((uint64_t)1 << 30);   // long-read sized: chunked multi-pass path
    const size_t res_bytes = (size_t)n3 * max_res * sizeof(smb_ali_result);
    CU(sp.begin(SPAN_CAND));
    CU(launch_block_emit_k3(a, ncand, st, &nl));
    CU(sp.end());
    bool multipass = too_big;
    if (!too_big) {
      const size_t outa = res_bytes + (size_t)n3 * (2 * sizeof(uint32_t) + sizeof(int32_t)) + 64;
      CU(ctx->out_b.ensure(outa));
      CU(ctx->dirs.ensure((size_t)(h_cnt->dir_words + 4) * sizeof(uint32_t)));
      CU(ctx->diff.ensure((size_t)h_cnt->diff_bytes + 64));
      CU(sp.begin(SPAN_CAND));
      CU(launch_scan_counts(a.dir_words_arr, n3, (unsigned long long *)d_dir_off, d_tile3, st, &nl));
      CU(launch_scan_counts(a.diff_stride, n3, (unsigned long long *)d_diff_off, d_tile3, st, &nl));
      CU(sp.end());
      char *ob = ctx->out_b.as<char>();
      smb_ali_result *d_res = (smb_ali_result *)ob;
      uint32_t *d_nres = (uint32_t *)(ob + res_bytes);
      uint32_t *d_dused = d_nres + n3;
      int32_t *d_errs = (int32_t *)(d_dused + n3);
      unsigned long long *d_cells = (unsigned long long *)(((uintptr_t)(d_errs + n3) + 15) & ~(uintptr_t)15);
      CU(cudaMemsetAsync(d_cells, 0, sizeof(unsigned long long), st));
      BandOut bo{d_res, d_nres, ctx->diff.as<uint8_t>(), d_errs, d_cells, d_dused};
      const size_t gring_words = band_gring_words(plan);
      CU(ctx->scratch.ensure(gring_words * sizeof(uint32_t) + 512));
      const size_t cmp_bytes = (size_t)ntiles * 2 * 8 + 64 + (size_t)(n3 + 1) * 4 + 64 + (size_t)n3 * 8 + 64;
      CU(ctx->cmp.ensure(cmp_bytes));
      char *cb = ctx->cmp.as<char>();
      unsigned long long *d_tile_res = (unsigned long long *)cb;
      unsigned long long *d_tile_diff = d_tile_res + ntiles;
      CompactTotals *d_tot = (CompactTotals *)(d_tile_diff + ntiles);
      unsigned long long *d_diff_first = (unsigned long long *)(((uintptr_t)(d_tot + 1) + 63) & ~(uintptr_t)63);
      uint32_t *d_first = (uint32_t *)(d_diff_first + n3);
      CU(sp.begin(SPAN_K3));
      CU(launch_band(sc, ctx->src, a.bat, plan, a.k3_order, true, nullptr, bo, max_res, d_dir_off, ctx->dirs.as<uint32_t>(),
                     d_diff_off, a.diff_cap, ctx->scratch.as<uint32_t>(), ctx->ticket.as<int>(), ctx->sm_count, st, &nl,
                     &ctx->side));
      uint32_t *d_cig_task = nullptr;
      unsigned long long *d_tile_cig = nullptr, *d_cig_first = nullptr;
      if (P.cigar) {   // output stage: text bytes per task, scanned with the result and DiffStr counts
        Carver c(nullptr);
        c.take<uint32_t>((size_t)n3 + 1); c.take<unsigned long long>((size_t)ntiles + 1); c.take<unsigned long long>((size_t)n3 + 1);
        CU(ctx->blk_cig.ensure(c.off));
        Carver d(ctx->blk_cig.p);
        d_cig_task = d.take<uint32_t>((size_t)n3 + 1);
        d_tile_cig = d.take<unsigned long long>((size_t)ntiles + 1);
        d_cig_first = d.take<unsigned long long>((size_t)n3 + 1);
        CigarSlots cs{d_res, d_nres, d_diff_off, ctx->diff.as<uint8_t>(), a.bat, n3, max_res, (int)prm->cigar};
        CU(launch_cigar_task_count(cs, d_cig_task, st, &nl));
      }
      CU(launch_compact_scan(d_nres, d_dused, d_errs, n3, d_tile_res, d_tile_diff, d_tot, d_first, d_diff_first, st, &nl,
                             d_cig_task, d_tile_cig, d_cig_first));
      CU(sp.end());
      CompactTotals *h_ct = (CompactTotals *)(h_tot + 2);
      unsigned long long *h_cells = (unsigned long long *)(h_ct + 1);
      CU(d2h(h_ct, d_tot, sizeof(CompactTotals), st));
      CU(d2h(h_cells, d_cells, sizeof(unsigned long long), st));
      CU(ctx_sync(ctx));                                                        // sync 4: result sizes
      if (h_ct->capacity_flag || h_ct->ndiff > 0xffffffffull) multipass = true;
      else {
        P.nres = (size_t)h_ct->nresults;
        P.ndiff = (size_t)h_ct->ndiff;
        P.d_res = d_res; P.d_nres = d_nres; P.d_dused = d_dused; P.d_first = d_first; P.d_errs = d_errs;
        P.d_diff_off = d_diff_off; P.d_diff_first = d_diff_first;
        sizes->k3_cells = *h_cells;
        if (P.cigar) { P.ncig = (size_t)h_ct->ncig; P.d_cig_first = d_cig_first; }
      }
    }
    if (multipass) {
      std::vector<smb_band_task> h_bat(nk3);
      CU(d2h(h_bat.data(), a.bat, nk3 * sizeof(smb_band_task), st));
      CU(ctx_sync(ctx));
      uint64_t cells = 0;
      const float keep_ms = ctx->last_ms;
      ctx->last_ms = 0.f;
      const int rc = band_align_multipass(ctx, h_bat.data(), n3, P.v_res, P.v_first, P.v_diff, P.v_errs, &cells);
      if (rc) return rc;
      ms_k3 += ctx->last_ms;
      ctx->last_ms = keep_ms;
      P.multipass = true;
      P.nres = P.v_res.size();
      P.ndiff = P.v_diff.size();
      sizes->k3_cells = cells;
      if (P.cigar && P.nres) {   // output stage on the dense results the multi-pass path assembled on the host
        const size_t rb = al256(P.nres * sizeof(smb_ali_result)), extra = rb + al256(P.ndiff + 1);
        {   // size the buffer first so that the uploads land in the final allocation
          Carver c(nullptr);
          c.off = al256(extra);
          c.take<uint32_t>(P.nres + 1); c.take<int32_t>(P.nres + 1); c.take<unsigned long long>(P.nres + 2);
          c.take<unsigned long long>((size_t)compact_tiles((int)P.nres + 1) + 2);
          CU(ctx->blk_cig.ensure(c.off));
        }
        smb_ali_result *d_mres = (smb_ali_result *)ctx->blk_cig.p;
        uint8_t *d_mdiff = (uint8_t *)ctx->blk_cig.p + rb;
        CU(h2d(d_mres, P.v_res.data(), P.nres * sizeof(smb_ali_result), st));
        if (P.ndiff) CU(h2d(d_mdiff, P.v_diff.data(), P.ndiff, st));
        CigarArgs ca{};
        ca.res = d_mres; ca.diff = d_mdiff; ca.tasks = a.bat; ca.n = (int)P.nres; ca.flags = prm->cigar;
        unsigned long long *h_cig = h_tot + 8;
        const int rc2 = cigar_count_pass(ctx, &ca, &P.d_cig_off, extra, h_cig, st, &nl);
        if (rc2) return rc2;
        CU(ctx_sync(ctx));
        P.cig = ca;
        P.ncig = (size_t)*h_cig;
      }
    }
  }
  sizes->nresults = P.nres;
  sizes->ndiffbytes = P.ndiff;
  sizes->ncigarbytes = P.ncig;
  {
    float ms[SPAN_KINDS];
    sp.sum(ms);
    sizes->ms_hits = ms[SPAN_HITS];
    sizes->ms_cand = ms[SPAN_CAND];
    sizes->ms_k2 = ms[SPAN_K2];
    sizes->ms_k3 = ms[SPAN_K3] + ms_k3;
  }
  sizes->launches = nl;
  ctx->last_ms = sizes->ms_hits + sizes->ms_cand + sizes->ms_k2 + sizes->ms_k3;
  ctx->last_launches += nl;
  ctx->total_launches += nl;
  g_launches += nl;
  P.valid = true;
  return SMB_OK;
}

static int block_fetch_impl(smb_ctx *ctx, smb_block_read *reads, smb_block_cand *cands, int32_t *errs, uint32_t *first_result,
                            smb_ali_result *results, uint8_t *diffstr, bool want_cigar, void *cigar_blob) {
  if (!ctx) return SMB_ERR_ARG;
  if (!ctx->blk || !ctx->blk->p.valid) return fail(ctx, SMB_ERR_STATE, "smb_block_run() first");
  Pending &P = ctx->blk->p;
  if (!P.njobs) {
    if (want_cigar && cigar_blob) *(uint32_t *)cigar_blob = 0;
    return SMB_OK;
  }
  if (!reads || (P.nk3 && (!cands || !errs || !first_result)) || (P.nres && !results) || (P.ndiff && !diffstr))
    return SMB_ERR_ARG;
  if (want_cigar && !P.cigar) return fail(ctx, SMB_ERR_STATE, "smb_block_run() without SMB_CIGAR_ON");
  if (want_cigar && !cigar_blob) return SMB_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  int nl = 0;
  CU(d2h(reads, P.d_rd, (size_t)P.njobs * sizeof(smb_block_read), st));
  if (P.nk3) {
    CU(d2h(cands, P.d_k3c, P.nk3 * sizeof(smb_block_cand), st));
    if (P.multipass) {
      memcpy(errs, P.v_errs.data(), P.nk3 * sizeof(int32_t));
      memcpy(first_result, P.v_first.data(), (P.nk3 + 1) * sizeof(uint32_t));
      if (P.nres) memcpy(results, P.v_res.data(), P.nres * sizeof(smb_ali_result));
      if (P.ndiff) memcpy(diffstr, P.v_diff.data(), P.ndiff);
    } else {
      // dense arrays reuse the direction-strip buffer (no longer needed)
      const size_t dense_bytes = P.nres * sizeof(smb_ali_result) + P.ndiff + 64;
      CU(ctx->dirs.ensure(dense_bytes));
      smb_ali_result *d_dense = ctx->dirs.as<smb_ali_result>();
      uint8_t *d_ddiff = (uint8_t *)(d_dense + P.nres);
      GatherCigar gc{};
      const size_t head = (2 * P.nres + 1) * 4;
      if (want_cigar && P.cigar && P.nres) {   // output stage: the gather also writes the blob {first, nm, text}
        CU(ctx->blk_cigtext.ensure(head + P.ncig + 64));
        char *d_blob = ctx->blk_cigtext.as<char>();
        gc.tasks = P.args.bat; gc.cig_first = P.d_cig_first; gc.first_out = (uint32_t *)d_blob;
        gc.nm = (int32_t *)(d_blob + (P.nres + 1) * 4); gc.text = d_blob + head; gc.flags = P.cigar_flags;
        gc.nres_total = (uint32_t)P.nres; gc.ncig_total = P.ncig;
      }
      CU(launch_compact_gather(P.d_res, P.d_nres, ctx->diff.as<uint8_t>(), P.d_diff_off, P.d_dused, (int)P.nk3, P.max_res,
                               P.d_first, P.d_diff_first, d_dense, d_ddiff, st, &nl, gc.text ? &gc : nullptr));
      if (gc.text) CU(d2h(cigar_blob, gc.first_out, head + P.ncig, st));
      if (P.nres) CU(d2h(results, d_dense, P.nres * sizeof(smb_ali_result), st));
      if (P.ndiff) CU(d2h(diffstr, d_ddiff, P.ndiff, st));
      CU(d2h(first_result, P.d_first, (P.nk3 + 1) * sizeof(uint32_t), st));
      CU(d2h(errs, P.d_errs, P.nk3 * sizeof(int32_t), st));
    }
  }
  if (want_cigar) {
    if (!P.nres || !P.nk3) *(uint32_t *)cigar_blob = 0;
    else if (P.multipass) {
      const int rc = cigar_fill_blob(ctx, P.cig, P.d_cig_off, P.ncig, cigar_blob, st, &nl);
      if (rc) return rc;
    }
  }
  CU(ctx_sync(ctx));
  ctx->last_launches += nl;
  ctx->total_launches += nl;
  g_launches += nl;
  return SMB_OK;
}

int smb_block_fetch(smb_ctx *ctx, smb_block_read *reads, smb_block_cand *cands, int32_t *errs, uint32_t *first_result,
                    smb_ali_result *results, uint8_t *diffstr) {
  return block_fetch_impl(ctx, reads, cands, errs, first_result, results, diffstr, false, nullptr);
}

int smb_block_fetch_cigar(smb_ctx *ctx, smb_block_read *reads, smb_block_cand *cands, int32_t *errs, uint32_t *first_result,
                          smb_ali_result *results, uint8_t *diffstr, void *cigar_blob) {
  return block_fetch_impl(ctx, reads, cands, errs, first_result, results, diffstr, true, cigar_blob);
}

// the output stage for explicit alignment strings
int smb_cigar_batch(smb_ctx *ctx, const uint8_t *diffstr, size_t ndiffbytes, const uint32_t *diff_off,
                    const uint32_t *clip_start, const uint32_t *clip_end, int n, int flags, void *cigar_blob,
                    size_t max_text, size_t *ntext) {
  if (!ctx || n < 0 || !cigar_blob || !ntext) return SMB_ERR_ARG;
  *ntext = 0;
  *(uint32_t *)cigar_blob = 0;
  if (!n) return SMB_OK;
  if (!diffstr || !ndiffbytes || !diff_off || !clip_start || !clip_end) return SMB_ERR_ARG;
  if (diffstr[ndiffbytes - 1]) return fail(ctx, SMB_ERR_ARG, "alignment strings: the last byte must be the terminator");
  for (int i = 0; i < n; ++i)
    if (diff_off[i] >= ndiffbytes) return fail(ctx, SMB_ERR_ARG, "alignment string %d starts outside the buffer", i);
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  int nl = 0;
  // this call reuses the buffers the output stage of a block keeps between smb_block_run and smb_block_fetch_cigar:
  // a block that is still pending loses its text (its fetch then fails with SMB_ERR_STATE instead of reading them)
  if (ctx->blk && ctx->blk->p.valid) ctx->blk->p.cigar = false;
  const size_t N = (size_t)n, ab = al256(N * 4), extra = al256(ndiffbytes) + 3 * ab;
  {
    Carver c(nullptr);
    c.off = al256(extra);
    c.take<uint32_t>(N + 1); c.take<int32_t>(N + 1); c.take<unsigned long long>(N + 2);
    c.take<unsigned long long>((size_t)compact_tiles(n + 1) + 2);
    CU(ctx->blk_cig.ensure(c.off));
  }
  uint8_t *d_diff = ctx->blk_cig.as<uint8_t>();
  uint32_t *d_xoff = (uint32_t *)(d_diff + al256(ndiffbytes)), *d_cs = (uint32_t *)((char *)d_xoff + ab),
           *d_ce = (uint32_t *)((char *)d_cs + ab);
  CU(h2d(d_diff, diffstr, ndiffbytes, st));
  CU(h2d(d_xoff, diff_off, N * 4, st));
  CU(h2d(d_cs, clip_start, N * 4, st));
  CU(h2d(d_ce, clip_end, N * 4, st));
  CigarArgs ca{};
  ca.diff = d_diff; ca.x_off = d_xoff; ca.x_cs = d_cs; ca.x_ce = d_ce; ca.n = n; ca.flags = flags;
  CU(ctx->stage.ensure(64));
  unsigned long long *h_total = ctx->stage.as<unsigned long long>();
  unsigned long long *d_off = nullptr;
  int rc = cigar_count_pass(ctx, &ca, &d_off, extra, h_total, st, &nl);
  if (rc) return rc;
  CU(ctx_sync(ctx));
  *ntext = (size_t)*h_total;
  if (*ntext > max_text) return SMB_ERR_CAPACITY;
  rc = cigar_fill_blob(ctx, ca, d_off, *ntext, cigar_blob, st, &nl);
  if (rc) return rc;
  CU(ctx_sync(ctx));
  ctx->last_launches = nl;
  ctx->total_launches += nl;
  g_launches += nl;
  return SMB_OK;
}

int smb_block_debug_cands(smb_ctx *ctx, uint64_t *cand_first, smb_block_cand *cands, uint32_t *cover, uint32_t *qs_qe,
                          size_t max_cands) {
  if (!ctx || !cand_first) return SMB_ERR_ARG;
  if (!ctx->blk || !ctx->blk->p.valid) return fail(ctx, SMB_ERR_STATE, "smb_block_run() first");
  const Pending &P = ctx->blk->p;
  if (!P.njobs) return SMB_OK;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  CU(d2h(cand_first, P.args.cand_first, ((size_t)P.njobs + 1) * sizeof(uint64_t), st));
  CU(ctx_sync(ctx));
  if (P.ncand > max_cands) return fail(ctx, SMB_ERR_CAPACITY, "need room for %zu candidates", P.ncand);
  if (!P.ncand) return SMB_OK;
  std::vector<DCand> dc(P.ncand);
  std::vector<int32_t> score(P.ncand);
  CU(d2h(dc.data(), P.args.dc, P.ncand * sizeof(DCand), st));
  CU(d2h(score.data(), P.args.score, P.ncand * sizeof(int32_t), st));
  CU(ctx_sync(ctx));
  for (size_t i = 0; i < P.ncand; ++i) {
    if (cands) {
      smb_block_cand &o = cands[i];
      memset(&o, 0, sizeof o);
      o.rs = dc[i].rs; o.sqidx = dc[i].sqidx; o.swscor = score[i]; o.reflen = dc[i].reflen;
      o.band_l = dc[i].band_l; o.band_r = dc[i].band_r; o.reverse = dc[i].rev;
    }
    if (cover) cover[i] = dc[i].cover;
    if (qs_qe) { qs_qe[2 * i] = dc[i].qs; qs_qe[2 * i + 1] = dc[i].qe; }
  }
  return SMB_OK;
}

}  // extern "C"
