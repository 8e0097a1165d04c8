// api.cu - the C ABI of include/smalt_b200.h: context, device buffers, batch entry points.
#include "ctx.h"
#if defined(__linux__)
#include <sys/prctl.h>
#endif
#include "block.cuh"
#include <cmath>
#include <new>

std::atomic<unsigned long long> g_h2d_bytes{0}, g_d2h_bytes{0};
std::atomic<long long> g_launches{0};

static void make_scoring(Scoring &sc, int match, int mismatch, int gapopen, int gapext) {
  sc.match = match;
  sc.mismatch = mismatch;
  sc.gap_init = -gapopen;
  sc.gap_ext = -gapext;
  for (int a = 0; a < 8; ++a)
    for (int b = 0; b < 8; ++b) {
      int s;
      if (a >= 6 || b >= 6 || a == 5 || b == 5) s = 0;  // N / outside the alphabet (score.c:158-161)
      else if (a == 4 || b == 4) s = mismatch - match;  // X (score.c:162-163)
      else s = (a == b) ? match : mismatch;
      sc.S[a * 8 + b] = (signed char)s;
    }
}

extern "C" {

const char *smb_version(void) { return "smalt-b200 0.1 (sm_100a)"; }

int smb_device_warmup(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return SMB_ERR_NODEVICE;
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) return SMB_ERR_CUDA;
  if (warm_sw() != cudaSuccess || warm_band() != cudaSuccess || warm_band_warp() != cudaSuccess || warm_band_wide() != cudaSuccess || warm_band_long() != cudaSuccess || warm_band_pack() != cudaSuccess ||
      warm_seed() != cudaSuccess ||
      warm_compact() != cudaSuccess || warm_block() != cudaSuccess)
    return SMB_ERR_CUDA;
  return SMB_OK;
}

int smb_ctx_create(smb_ctx **out, int device) {
  if (!out) return SMB_ERR_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return SMB_ERR_NODEVICE;
  smb_ctx *ctx = new (std::nothrow) smb_ctx();
  if (!ctx) return SMB_ERRCODE_NOMEM;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_done, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_spin, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->side.stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->side.fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->side.join, cudaEventDisableTiming) != cudaSuccess) {
    delete ctx;
    return SMB_ERR_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  make_scoring(ctx->sc, 1, -2, -4, -3);
  if (ctx->ticket.ensure(256) != cudaSuccess) {
    smb_ctx_destroy(ctx);
    return SMB_ERR_CUDA;
  }
  *out = ctx;
  return SMB_OK;
}

void smb_ctx_destroy(smb_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  DevBuf *bufs[] = {&ctx->arena, &ctx->packed, &ctx->tasks, &ctx->out_a, &ctx->out_b,
                    &ctx->scratch, &ctx->dirs, &ctx->diff, &ctx->offs, &ctx->index, &ctx->qualbuf,
                    &ctx->seed_meta, &ctx->seed_u32, &ctx->seed_u8, &ctx->hit_meta, &ctx->hit_data, &ctx->cmp, &ctx->ticket, &ctx->hit_qmask, &ctx->aux_index,
                    &ctx->seq_offs_buf, &ctx->blk_jobs, &ctx->blk_scr, &ctx->blk_cand, &ctx->blk_k3, &ctx->blk_cig,
                    &ctx->blk_cigtext};
  for (DevBuf *b : bufs) b->release();
  for (cudaEvent_t &e : ctx->blk_ev) if (e) cudaEventDestroy(e);
  block_state_free(ctx);
  if (ctx->built.block) cudaFree(ctx->built.block);
  ctx->stage.release();
  if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
  if (ctx->ev_spin) cudaEventDestroy(ctx->ev_spin);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->side.fork) cudaEventDestroy(ctx->side.fork);
  if (ctx->side.join) cudaEventDestroy(ctx->side.join);
  if (ctx->side.stream) cudaStreamDestroy(ctx->side.stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *smb_last_error(const smb_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }
float smb_last_kernel_ms(const smb_ctx *ctx) { return ctx ? ctx->last_ms : 0.f; }
int smb_last_kernel_launches(const smb_ctx *ctx) { return ctx ? ctx->last_launches : 0; }
long long smb_total_kernel_launches(const smb_ctx *ctx) { return ctx ? ctx->total_launches : 0; }
void smb_process_counters(unsigned long long *launches, unsigned long long *h2d_bytes, unsigned long long *d2h_bytes) {
  if (launches) *launches = (unsigned long long)g_launches.load();
  if (h2d_bytes) *h2d_bytes = g_h2d_bytes.load();
  if (d2h_bytes) *d2h_bytes = g_d2h_bytes.load();
}

void *smb_host_alloc(size_t nbytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void smb_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int smb_set_scoring(smb_ctx *ctx, int match, int mismatch, int gapopen, int gapext) {
  if (!ctx) return SMB_ERR_ARG;
  // scoreSetPenalty ranges (score.c:90-112)
  if (match < 0 || match > 127 || mismatch > 0 || mismatch < -127 || gapopen > 0 || gapopen < -127 ||
      gapext > 0 || gapext < -127 || mismatch - match < -127)
    return fail(ctx, SMB_ERRCODE_SWATEXCEED, "penalty out of range");
  make_scoring(ctx->sc, match, mismatch, gapopen, gapext);
  return SMB_OK;
}

int smb_ctx_set_spin(smb_ctx *ctx, int spin) {
  if (!ctx) return SMB_ERR_ARG;
  ctx->spin = spin < 0 ? 0 : (spin >= 2 ? 2 : spin);
  if (spin > 2) ctx->poll_ns = (long)spin * 1000L;   // 3 and more: the sleep between polls in microseconds
  if (ctx->spin == 2) {
#if defined(__linux__)
    prctl(PR_SET_TIMERSLACK, 1000UL, 0UL, 0UL, 0UL);   // of the calling thread: sleeps of tens of microseconds mean what they say
#endif
  }
  return SMB_OK;
}

int smb_ctx_share_index(smb_ctx *dst, const smb_ctx *src) {
  if (!dst || !src || dst->device != src->device) return SMB_ERR_ARG;
  dst->ix = src->ix;
  dst->have_index = src->have_index;
  dst->src.packed = src->src.packed;
  dst->src.packed_nbases = src->src.packed_nbases;
  dst->seq_offs = src->seq_offs;
  dst->d_seq_offs = src->d_seq_offs;
  return SMB_OK;
}

int smb_int_peak(smb_ctx *ctx, double gops[5]) {
  if (!ctx || !gops) return SMB_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  CU(ctx->scratch.ensure((size_t)ctx->sm_count * 8 * 256 * sizeof(int)));
  const int iters = 4096;
  for (int mode = 0; mode < 5; ++mode) {
    double ops = 0, best = 0;
    for (int rep = 0; rep < 4; ++rep) {
      CU(cudaEventRecord(ctx->ev0, st));
      CU(run_int_peak(mode, ctx->sm_count, ctx->scratch.as<int>(), iters, st, &ops));
      CU(cudaEventRecord(ctx->ev1, st));
      CU(ctx_sync(ctx));
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
      ++ctx->total_launches;
      if (rep > 0 && ms > 0.f && ops / ms * 1e-6 > best) best = ops / ms * 1e-6;
    }
    gops[mode] = best;
  }
  return SMB_OK;
}

int smb_arena_upload(smb_ctx *ctx, const uint8_t *codes, size_t nbytes) {
  if (!ctx || (!codes && nbytes)) return SMB_ERR_ARG;
  cudaSetDevice(ctx->device);
  CU(ctx->arena.ensure(nbytes + 16));
  CU(h2d(ctx->arena.p, codes, nbytes, ctx->stream));
  CU(ctx_sync(ctx));
  ctx->arena_bytes = nbytes;
  ctx->src.arena = ctx->arena.as<uint8_t>();
  return SMB_OK;
}

int smb_refseq_upload(smb_ctx *ctx, const uint32_t *words, size_t nwords, uint64_t nbases,
                      const uint64_t *seq_offs, int nseq) {
  if (!ctx || !words || nwords * 10u < nbases) return SMB_ERR_ARG;
  cudaSetDevice(ctx->device);
  CU(ctx->packed.ensure(nwords * sizeof(uint32_t) + 1040));   // (slack: the K3 long-read kernel stages whole 512-byte chunks)
  CU(h2d(ctx->packed.p, words, nwords * sizeof(uint32_t), ctx->stream));
  CU(ctx_sync(ctx));
  ctx->src.packed = ctx->packed.as<uint32_t>();
  ctx->src.packed_nbases = nbases;
  ctx->seq_offs.clear();
  ctx->d_seq_offs = nullptr;
  if (seq_offs && nseq > 0) {
    ctx->seq_offs.assign(seq_offs, seq_offs + nseq + 1);
    CU(ctx->seq_offs_buf.ensure(((size_t)nseq + 1) * sizeof(uint64_t)));
    CU(h2d(ctx->seq_offs_buf.p, seq_offs, ((size_t)nseq + 1) * sizeof(uint64_t), ctx->stream));
    CU(ctx_sync(ctx));
    ctx->d_seq_offs = ctx->seq_offs_buf.as<uint64_t>();
  }
  return SMB_OK;
}

static int check_seq_ranges(smb_ctx *ctx, uint64_t read_off, uint32_t read_len, uint64_t ref_off,
                            uint32_t ref_len, uint32_t flags, int i) {
  if (read_off > ctx->arena_bytes || read_len > ctx->arena_bytes - read_off)
    return fail(ctx, SMB_ERR_ARG, "task %d: read [%llu,+%u) outside the arena (%zu bytes)", i,
                (unsigned long long)read_off, read_len, ctx->arena_bytes);
  if (flags & SMB_TASK_REF_PACKED) {
    if (!ctx->src.packed) return fail(ctx, SMB_ERR_STATE, "task %d: no packed reference uploaded", i);
    if (ref_off > ctx->src.packed_nbases || ref_len > ctx->src.packed_nbases - ref_off)
      return fail(ctx, SMB_ERR_ARG, "task %d: window outside the packed reference", i);
  } else if (ref_off > ctx->arena_bytes || ref_len > ctx->arena_bytes - ref_off) {
    return fail(ctx, SMB_ERR_ARG, "task %d: window outside the arena", i);
  }
  return SMB_OK;
}

int smb_sw_score_batch(smb_ctx *ctx, const smb_sw_task *tasks, int ntasks, int32_t *scores,
                       int32_t *errs) {
  if (!ctx || ntasks < 0 || (ntasks && (!tasks || !scores || !errs))) return SMB_ERR_ARG;
  ctx->last_ms = 0.f;
  ctx->last_launches = 0;
  if (!ntasks) return SMB_OK;
  if (!ctx->src.arena) return fail(ctx, SMB_ERR_STATE, "smb_arena_upload() first");
  for (int i = 0; i < ntasks; ++i) {
    int rcode = check_seq_ranges(ctx, tasks[i].read_off, tasks[i].read_len, tasks[i].ref_off,
                                 tasks[i].ref_len, tasks[i].flags, i);
    if (rcode) return rcode;
  }
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  CU(ctx->tasks.ensure((size_t)ntasks * sizeof(smb_sw_task)));
  CU(ctx->out_a.ensure((size_t)ntasks * 2 * sizeof(int32_t)));
  CU(h2d(ctx->tasks.p, tasks, (size_t)ntasks * sizeof(smb_sw_task), st));
  int32_t *d_scores = ctx->out_a.as<int32_t>(), *d_errs = d_scores + ntasks;
  int nl = 0;
  SwPlan plan;
  plan_sw(tasks, ntasks, ctx->sm_count, ctx->sc, plan);   // host planning happens before the timed events
  const size_t off_order = 256, off_strip = (off_order + (size_t)ntasks * sizeof(int) + 255) & ~(size_t)255;
  CU(ctx->scratch.ensure(off_strip + plan.strip_bytes + 256));
  char *sb = ctx->scratch.as<char>();
  CU(h2d(sb + off_order, plan.order.data(), (size_t)ntasks * sizeof(int), st));
  CU(cudaEventRecord(ctx->ev0, st));
  CU(launch_sw_score(ctx->sc, ctx->src, ctx->tasks.as<smb_sw_task>(), plan, (int *)sb, (const int *)(sb + off_order),
                     sb + off_strip, d_scores, d_errs, st, &nl));
  CU(cudaEventRecord(ctx->ev1, st));
  CU(d2h(scores, d_scores, (size_t)ntasks * sizeof(int32_t), st));
  CU(d2h(errs, d_errs, (size_t)ntasks * sizeof(int32_t), st));
  CU(ctx_sync(ctx));
  CU(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
  ctx->last_launches = nl;
  ctx->total_launches += nl;
  g_launches += nl;
  return SMB_OK;
}

int smb_band_score_batch(smb_ctx *ctx, const smb_band_task *tasks, int ntasks, int32_t *scores,
                         int32_t *errs) {
  if (!ctx || ntasks < 0 || (ntasks && (!tasks || !scores || !errs))) return SMB_ERR_ARG;
  ctx->last_ms = 0.f;
  ctx->last_launches = 0;
  if (!ntasks) return SMB_OK;
  if (!ctx->src.arena) return fail(ctx, SMB_ERR_STATE, "smb_arena_upload() first");
  for (int i = 0; i < ntasks; ++i) {
    int rcode = check_seq_ranges(ctx, tasks[i].read_off, tasks[i].read_len, tasks[i].ref_off,
                                 tasks[i].ref_len, tasks[i].flags, i);
    if (rcode) return rcode;
  }
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  CU(ctx->tasks.ensure((size_t)ntasks * sizeof(smb_band_task)));
  CU(ctx->out_a.ensure((size_t)ntasks * 2 * sizeof(int32_t) + 64));
  CU(h2d(ctx->tasks.p, tasks, (size_t)ntasks * sizeof(smb_band_task), st));
  int32_t *d_scores = ctx->out_a.as<int32_t>(), *d_errs = d_scores + ntasks;
  unsigned long long *d_cells = (unsigned long long *)(((uintptr_t)(d_errs + ntasks) + 15) & ~(uintptr_t)15);
  CU(cudaMemsetAsync(d_cells, 0, sizeof(unsigned long long), st));
  BandOut bo{nullptr, nullptr, nullptr, d_errs, d_cells, nullptr};
  int nl = 0;
  BandPlan plan;
  plan_band(tasks, ntasks, false, ctx->sc, plan);
  const size_t gring_words = band_gring_words(plan);
  CU(ctx->scratch.ensure((size_t)ntasks * sizeof(int) + gring_words * sizeof(uint32_t) + 512));
  int *d_order = ctx->scratch.as<int>();
  uint32_t *d_gring = (uint32_t *)(ctx->scratch.as<char>() + (((size_t)ntasks * sizeof(int) + 255) & ~(size_t)255));
  CU(h2d(d_order, plan.order.data(), (size_t)ntasks * sizeof(int), st));
  CU(cudaEventRecord(ctx->ev0, st));
  CU(launch_band(ctx->sc, ctx->src, ctx->tasks.as<smb_band_task>(), plan, d_order, false, d_scores, bo, 0,
                 nullptr, nullptr, nullptr, nullptr, d_gring, ctx->ticket.as<int>(), ctx->sm_count, st, &nl));
  CU(cudaEventRecord(ctx->ev1, st));
  CU(d2h(scores, d_scores, (size_t)ntasks * sizeof(int32_t), st));
  CU(d2h(errs, d_errs, (size_t)ntasks * sizeof(int32_t), st));
  CU(ctx_sync(ctx));
  CU(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
  ctx->last_launches = nl;
  ctx->total_launches += nl;
  g_launches += nl;
  return SMB_OK;
}

// One pass of K3 over `idx` (indices into tasks) with the given per-task capacities.
static int band_align_pass(smb_ctx *ctx, const smb_band_task *tasks, const std::vector<int> &idx,
                           int max_res, int diff_scale, std::vector<smb_ali_result> &h_res,
                           std::vector<uint32_t> &h_nres, std::vector<int32_t> &h_errs,
                           std::vector<uint8_t> &h_diff, std::vector<uint64_t> &diff_off,
                           unsigned long long *cells, float *ms, int *nlaunch) {
  const int n = (int)idx.size();
  std::vector<smb_band_task> sub((size_t)n);
  std::vector<uint64_t> dir_off((size_t)n + 1, 0);
  std::vector<uint32_t> diff_cap((size_t)n);
  diff_off.assign((size_t)n + 1, 0);
  for (int i = 0; i < n; ++i) {
    const smb_band_task &t = tasks[idx[(size_t)i]];
    sub[(size_t)i] = t;
    const uint64_t words = band_dir_words(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right,
                                          (int)t.ref_len);
    dir_off[(size_t)i + 1] = dir_off[(size_t)i] + words;
    const uint32_t cap = (uint32_t)diff_scale * (t.read_len + t.ref_len + 64u);
    diff_cap[(size_t)i] = cap;
    diff_off[(size_t)i + 1] = diff_off[(size_t)i] + cap + (t.read_len + t.ref_len + 8u);
  }
  cudaStream_t st = ctx->stream;
  const size_t res_bytes = (size_t)n * max_res * sizeof(smb_ali_result);
  const size_t outa = res_bytes + (size_t)n * (sizeof(uint32_t) + sizeof(int32_t)) + 64;
  CU(ctx->tasks.ensure((size_t)n * sizeof(smb_band_task)));
  CU(ctx->out_b.ensure(outa));
  CU(ctx->dirs.ensure(dir_off[(size_t)n] * sizeof(uint32_t)));
  CU(ctx->diff.ensure(diff_off[(size_t)n]));
  CU(ctx->offs.ensure((size_t)(n + 1) * 2 * sizeof(uint64_t) + (size_t)n * sizeof(uint32_t)));
  uint64_t *d_dir_off = ctx->offs.as<uint64_t>();
  uint64_t *d_diff_off = d_dir_off + (n + 1);
  uint32_t *d_diff_cap = (uint32_t *)(d_diff_off + (n + 1));
  CU(h2d(ctx->tasks.p, sub.data(), (size_t)n * sizeof(smb_band_task), st));
  CU(h2d(d_dir_off, dir_off.data(), (size_t)(n + 1) * sizeof(uint64_t), st));
  CU(h2d(d_diff_off, diff_off.data(), (size_t)(n + 1) * sizeof(uint64_t), st));
  CU(h2d(d_diff_cap, diff_cap.data(), (size_t)n * sizeof(uint32_t), st));
  char *ob = ctx->out_b.as<char>();
  smb_ali_result *d_res = (smb_ali_result *)ob;
  uint32_t *d_nres = (uint32_t *)(ob + res_bytes);
  int32_t *d_errs = (int32_t *)(d_nres + n);
  unsigned long long *d_cells = (unsigned long long *)(((uintptr_t)(d_errs + n) + 15) & ~(uintptr_t)15);
  CU(cudaMemsetAsync(d_cells, 0, sizeof(unsigned long long), st));
  BandOut bo{d_res, d_nres, ctx->diff.as<uint8_t>(), d_errs, d_cells, nullptr};
  BandPlan plan;
  plan_band(sub.data(), n, true, ctx->sc, plan);
  const size_t gring_words = band_gring_words(plan);
  CU(ctx->scratch.ensure((size_t)n * sizeof(int) + gring_words * sizeof(uint32_t) + 512));
  int *d_order = ctx->scratch.as<int>();
  uint32_t *d_gring = (uint32_t *)(ctx->scratch.as<char>() + (((size_t)n * sizeof(int) + 255) & ~(size_t)255));
  CU(h2d(d_order, plan.order.data(), (size_t)n * sizeof(int), st));
  CU(cudaEventRecord(ctx->ev0, st));
  CU(launch_band(ctx->sc, ctx->src, ctx->tasks.as<smb_band_task>(), plan, d_order, true, nullptr, bo, max_res,
                 d_dir_off, ctx->dirs.as<uint32_t>(), d_diff_off, d_diff_cap, d_gring, ctx->ticket.as<int>(), ctx->sm_count, st, nlaunch, &ctx->side));
  CU(cudaEventRecord(ctx->ev1, st));
  h_res.resize((size_t)n * max_res);
  h_nres.resize((size_t)n);
  h_errs.resize((size_t)n);
  h_diff.resize(diff_off[(size_t)n]);
  unsigned long long c = 0;
  CU(d2h(h_res.data(), d_res, res_bytes, st));
  CU(d2h(h_nres.data(), d_nres, (size_t)n * sizeof(uint32_t), st));
  CU(d2h(h_errs.data(), d_errs, (size_t)n * sizeof(int32_t), st));
  CU(d2h(h_diff.data(), ctx->diff.p, diff_off[(size_t)n], st));
  CU(d2h(&c, d_cells, sizeof c, st));
  CU(ctx_sync(ctx));
  float m = 0.f;
  CU(cudaEventElapsedTime(&m, ctx->ev0, ctx->ev1));
  *ms += m;
  *cells += c;
  return SMB_OK;
}

// Fast path of smb_band_align_batch: one kernel pass with the default slot capacities, dense
// output assembled on the device (compact.cu).  Returns 1 (and leaves the outputs untouched)
// when a task ran out of slot capacity - the caller then takes the multi-pass path.
static int band_align_fast(smb_ctx *ctx, const smb_band_task *tasks, int ntasks, smb_ali_result *results,
                           size_t max_results, size_t *nresults, uint32_t *first_result, uint8_t *diffstr,
                           size_t max_diffbytes, size_t *ndiffbytes, int32_t *errs, uint64_t *ncells) {
  const int n = ntasks, max_res = 4;
  cudaStream_t st = ctx->stream;
  // host-side geometry: direction strip and DiffStr slot of every task
  const size_t stage_bytes = (size_t)(n + 1) * 2 * sizeof(uint64_t) + (size_t)n * 2 * sizeof(uint32_t) + 64;
  const size_t stage_tail = (stage_bytes + 63) & ~(size_t)63;  // totals read back behind the arrays
  CU(ctx->stage.ensure(stage_tail + 64));
  uint64_t *dir_off = ctx->stage.as<uint64_t>();
  uint64_t *diff_off = dir_off + (n + 1);
  uint32_t *diff_cap = (uint32_t *)(diff_off + (n + 1));
  int *h_order = (int *)(diff_cap + n);
  dir_off[0] = diff_off[0] = 0;
  for (int i = 0; i < n; ++i) {
    const smb_band_task &t = tasks[i];
    const uint64_t words = band_dir_words(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right,
                                          (int)t.ref_len);
    dir_off[i + 1] = dir_off[i] + words;
    // final DiffStr area: one byte per aligned column at most, plus slack for several results
    const uint32_t cap = t.read_len + t.ref_len + 64u;
    diff_cap[i] = cap;
    diff_off[i + 1] = diff_off[i] + cap + (t.read_len + t.ref_len + 8u);
  }
  if (dir_off[n] > ((uint64_t)1 << 30)) return 1;  // long-read sized batch: chunked multi-pass path
  BandPlan plan;
  plan_band(tasks, n, true, ctx->sc, plan);
  memcpy(h_order, plan.order.data(), (size_t)n * sizeof(int));

  const size_t res_bytes = (size_t)n * max_res * sizeof(smb_ali_result);
  const size_t outa = res_bytes + (size_t)n * (2 * sizeof(uint32_t) + sizeof(int32_t)) + 64;
  CU(ctx->tasks.ensure((size_t)n * sizeof(smb_band_task)));
  CU(ctx->out_b.ensure(outa));
  CU(ctx->dirs.ensure(dir_off[n] * sizeof(uint32_t)));
  CU(ctx->diff.ensure(diff_off[n]));
  CU(ctx->offs.ensure(stage_bytes));
  CU(h2d(ctx->tasks.p, tasks, (size_t)n * sizeof(smb_band_task), st));
  CU(h2d(ctx->offs.p, ctx->stage.p, stage_bytes, st));
  uint64_t *d_dir_off = ctx->offs.as<uint64_t>();
  uint64_t *d_diff_off = d_dir_off + (n + 1);
  uint32_t *d_diff_cap = (uint32_t *)(d_diff_off + (n + 1));
  int *d_order = (int *)(d_diff_cap + n);
  char *ob = ctx->out_b.as<char>();
  smb_ali_result *d_res = (smb_ali_result *)ob;
  uint32_t *d_nres = (uint32_t *)(ob + res_bytes);
  uint32_t *d_dused = d_nres + n;
  int32_t *d_errs = (int32_t *)(d_dused + n);
  unsigned long long *d_cells = (unsigned long long *)(((uintptr_t)(d_errs + n) + 15) & ~(uintptr_t)15);
  CU(cudaMemsetAsync(d_cells, 0, sizeof(unsigned long long), st));
  BandOut bo{d_res, d_nres, ctx->diff.as<uint8_t>(), d_errs, d_cells, d_dused};
  const size_t gring_words = band_gring_words(plan);
  CU(ctx->scratch.ensure(gring_words * sizeof(uint32_t) + 512));
  // compaction scratch: tile sums, totals, per-task offsets
  const int ntiles = compact_tiles(n);
  const size_t cmp_bytes = (size_t)ntiles * 2 * 8 + 64 + (size_t)(n + 1) * 4 + 64 + (size_t)n * 8 + 64;
  CU(ctx->cmp.ensure(cmp_bytes));
  char *cb = ctx->cmp.as<char>();
  unsigned long long *d_tile_res = (unsigned long long *)cb;
  unsigned long long *d_tile_diff = d_tile_res + ntiles;
  CompactTotals *d_tot = (CompactTotals *)(d_tile_diff + ntiles);
  unsigned long long *d_diff_first = (unsigned long long *)(((uintptr_t)(d_tot + 1) + 63) & ~(uintptr_t)63);
  uint32_t *d_first = (uint32_t *)(d_diff_first + n);
  int nl = 0;
  CU(cudaEventRecord(ctx->ev0, st));
  CU(launch_band(ctx->sc, ctx->src, ctx->tasks.as<smb_band_task>(), plan, d_order, true, nullptr, bo, max_res,
                 d_dir_off, ctx->dirs.as<uint32_t>(), d_diff_off, d_diff_cap, ctx->scratch.as<uint32_t>(), ctx->ticket.as<int>(),
                 ctx->sm_count, st, &nl, &ctx->side));
  CU(launch_compact_scan(d_nres, d_dused, d_errs, n, d_tile_res, d_tile_diff, d_tot, d_first, d_diff_first, st, &nl));
  CU(cudaEventRecord(ctx->ev1, st));
  CompactTotals *h_tot = (CompactTotals *)((char *)ctx->stage.p + stage_tail);
  unsigned long long *h_cells = (unsigned long long *)(h_tot + 1);
  CU(d2h(h_tot, d_tot, sizeof(CompactTotals), st));
  CU(d2h(h_cells, d_cells, sizeof(unsigned long long), st));
  CU(ctx_sync(ctx));
  float ms0 = 0.f, ms1 = 0.f;
  CU(cudaEventElapsedTime(&ms0, ctx->ev0, ctx->ev1));
  ctx->last_ms += ms0;
  ctx->last_launches += nl;
  ctx->total_launches += nl;
  g_launches += nl;
  if (getenv("SMB_PLAN_DEBUG")) {
    int mxbw = 0, mxrows = 0, mxread = 0;
    for (int i = 0; i < n; ++i) {
      mxbw = std::max(mxbw, tasks[i].r_edge - tasks[i].l_edge + 1);
      mxrows = std::max(mxrows, (int)tasks[i].ref_len);
      mxread = std::max(mxread, (int)tasks[i].read_len);
    }
    fprintf(stderr, "K3 plan: n %d pack8 %d pack %d half %d warp %d wide %d thread-classes %zu (", n, plan.pack8_count, plan.pack_count, plan.half_count,
            plan.warp_count, plan.wide_count, plan.classes.size());
    for (const auto &c : plan.classes) fprintf(stderr, " wcap %d x %d", c.wcap, c.count);
    fprintf(stderr, " ) max band %d rows %d read %d  %.3f ms\n", mxbw, mxrows, mxread, ms0);
  }
  if (h_tot->capacity_flag) return 1;
  const size_t nr = (size_t)h_tot->nresults, nd = (size_t)h_tot->ndiff;
  if (ncells) *ncells = *h_cells;
  *nresults = nr;
  *ndiffbytes = nd;
  if (nr > max_results || nd > max_diffbytes || (nr && !results) || (nd && !diffstr) || nd > 0xffffffffull)
    return fail(ctx, SMB_ERR_CAPACITY, "need %zu results and %zu diffstr bytes", nr, nd);
  // dense arrays reuse the direction-strip buffer (no longer needed)
  const size_t dense_bytes = nr * sizeof(smb_ali_result) + nd + 64;
  CU(ctx->dirs.ensure(dense_bytes));
  smb_ali_result *d_dense = ctx->dirs.as<smb_ali_result>();
  uint8_t *d_ddiff = (uint8_t *)(d_dense + nr);
  nl = 0;
  CU(cudaEventRecord(ctx->ev0, st));
  CU(launch_compact_gather(d_res, d_nres, ctx->diff.as<uint8_t>(), d_diff_off, d_dused, n, max_res, d_first,
                           d_diff_first, d_dense, d_ddiff, st, &nl));
  CU(cudaEventRecord(ctx->ev1, st));
  if (nr) CU(d2h(results, d_dense, nr * sizeof(smb_ali_result), st));
  if (nd) CU(d2h(diffstr, d_ddiff, nd, st));
  CU(d2h(first_result, d_first, (size_t)(n + 1) * sizeof(uint32_t), st));
  CU(d2h(errs, d_errs, (size_t)n * sizeof(int32_t), st));
  CU(ctx_sync(ctx));
  CU(cudaEventElapsedTime(&ms1, ctx->ev0, ctx->ev1));
  ctx->last_ms += ms1;
  ctx->last_launches += nl;
  ctx->total_launches += nl;
  g_launches += nl;
  return SMB_OK;
}

}  // extern "C"

// Multi-pass path of K3 for batches in which a task ran out of its slot capacity (more than four
// results, long DiffStrs) or whose direction strips exceed one pass: growing capacities, results
// assembled on the host.  Outputs in task order.
int band_align_multipass(smb_ctx *ctx, const smb_band_task *tasks, int ntasks, std::vector<smb_ali_result> &results,
                         std::vector<uint32_t> &first_result, std::vector<uint8_t> &diffstr,
                         std::vector<int32_t> &errs, uint64_t *ncells) {
  struct TaskOut { std::vector<smb_ali_result> res; std::vector<uint8_t> diff; int32_t err = 0; };
  std::vector<TaskOut> outs((size_t)ntasks);
  std::vector<int> todo((size_t)ntasks);
  for (int i = 0; i < ntasks; ++i) todo[(size_t)i] = i;
  int max_res = 8, diff_scale = 1, nl = 0;
  float ms = 0.f;
  unsigned long long cells = 0;
  // memory-bounded chunks (direction strips can be large for long reads)
  const uint64_t DIR_WORDS_MAX = (uint64_t)1 << 30;  // 4 GiB of direction words per pass
  for (int attempt = 0; attempt < 6 && !todo.empty(); ++attempt) {
    std::vector<int> retry;
    size_t pos = 0;
    while (pos < todo.size()) {
      std::vector<int> chunk;
      uint64_t words = 0;
      while (pos < todo.size()) {
        const smb_band_task &t = tasks[todo[pos]];
        const uint64_t w = band_dir_words(t.l_edge, t.r_edge, t.p_left, t.p_right, (int)t.read_len, t.u_left, t.u_right,
                                          (int)t.ref_len);
        if (!chunk.empty() && words + w > DIR_WORDS_MAX) break;
        words += w;
        chunk.push_back(todo[pos++]);
      }
      std::vector<smb_ali_result> h_res;
      std::vector<uint32_t> h_nres;
      std::vector<int32_t> h_errs;
      std::vector<uint8_t> h_diff;
      std::vector<uint64_t> doff;
      int rcode = band_align_pass(ctx, tasks, chunk, max_res, diff_scale, h_res, h_nres, h_errs, h_diff,
                                  doff, &cells, &ms, &nl);
      if (rcode) return rcode;
      for (size_t k = 0; k < chunk.size(); ++k) {
        const int ti = chunk[k];
        TaskOut &o = outs[(size_t)ti];
        o.res.clear();
        o.diff.clear();
        o.err = h_errs[k];
        if (o.err == SMB_ERR_CAPACITY && attempt < 5) { retry.push_back(ti); continue; }
        for (uint32_t r = 0; r < h_nres[k]; ++r) {
          smb_ali_result rr = h_res[k * (size_t)max_res + r];
          const uint8_t *src = h_diff.data() + doff[k] + rr.diff_off;
          rr.diff_off = (uint32_t)o.diff.size();
          rr.task = (uint32_t)ti;
          o.diff.insert(o.diff.end(), src, src + rr.diff_len);
          o.res.push_back(rr);
        }
      }
    }
    todo.swap(retry);
    max_res *= 8;
    diff_scale *= 4;
  }
  ctx->last_ms += ms;
  ctx->last_launches += nl;
  ctx->total_launches += nl;
  g_launches += nl;
  if (ncells) *ncells = cells;
  results.clear();
  diffstr.clear();
  first_result.assign((size_t)ntasks + 1, 0);
  errs.assign((size_t)ntasks, 0);
  for (int i = 0; i < ntasks; ++i) {
    TaskOut &o = outs[(size_t)i];
    first_result[(size_t)i] = (uint32_t)results.size();
    errs[(size_t)i] = o.err;
    const uint32_t nd = (uint32_t)diffstr.size();
    for (smb_ali_result rr : o.res) {
      rr.diff_off += nd;
      results.push_back(rr);
    }
    diffstr.insert(diffstr.end(), o.diff.begin(), o.diff.end());
  }
  first_result[(size_t)ntasks] = (uint32_t)results.size();
  return SMB_OK;
}

extern "C" {

int smb_band_align_batch(smb_ctx *ctx, const smb_band_task *tasks, int ntasks, smb_ali_result *results,
                         size_t max_results, size_t *nresults, uint32_t *first_result, uint8_t *diffstr,
                         size_t max_diffbytes, size_t *ndiffbytes, int32_t *errs, uint64_t *ncells) {
  if (!ctx || ntasks < 0 || !nresults || !ndiffbytes || (ntasks && (!tasks || !first_result || !errs)))
    return SMB_ERR_ARG;
  ctx->last_ms = 0.f;
  ctx->last_launches = 0;
  *nresults = 0;
  *ndiffbytes = 0;
  if (ncells) *ncells = 0;
  if (first_result) first_result[0] = 0;
  if (!ntasks) return SMB_OK;
  if (!ctx->src.arena) return fail(ctx, SMB_ERR_STATE, "smb_arena_upload() first");
  for (int i = 0; i < ntasks; ++i) {
    int rcode = check_seq_ranges(ctx, tasks[i].read_off, tasks[i].read_len, tasks[i].ref_off,
                                 tasks[i].ref_len, tasks[i].flags, i);
    if (rcode) return rcode;
  }
  cudaSetDevice(ctx->device);
  {
    const int rcode = band_align_fast(ctx, tasks, ntasks, results, max_results, nresults, first_result, diffstr,
                                      max_diffbytes, ndiffbytes, errs, ncells);
    if (rcode != 1) return rcode;
    *nresults = 0;
    *ndiffbytes = 0;
    if (ncells) *ncells = 0;
    first_result[0] = 0;
  }
  std::vector<smb_ali_result> v_res;
  std::vector<uint32_t> v_first;
  std::vector<uint8_t> v_diff;
  std::vector<int32_t> v_errs;
  const int rcode = band_align_multipass(ctx, tasks, ntasks, v_res, v_first, v_diff, v_errs, ncells);
  if (rcode) return rcode;
  const size_t nr = v_res.size(), nd = v_diff.size();
  *nresults = nr;
  *ndiffbytes = nd;
  if (nr > max_results || nd > max_diffbytes || (nr && !results) || (nd && !diffstr))
    return fail(ctx, SMB_ERR_CAPACITY, "need %zu results and %zu diffstr bytes", nr, nd);
  if (nr) memcpy(results, v_res.data(), nr * sizeof(smb_ali_result));
  if (nd) memcpy(diffstr, v_diff.data(), nd);
  memcpy(first_result, v_first.data(), ((size_t)ntasks + 1) * sizeof(uint32_t));
  memcpy(errs, v_errs.data(), (size_t)ntasks * sizeof(int32_t));
  return SMB_OK;
}

// ------------------------------------ K1 ------------------------------------------

int smb_index_upload(smb_ctx *ctx, int typ, int wordlen, int nskip, int nbits_key, int nbits_lo,
                     uint32_t npos, uint32_t nwords, const uint32_t *idx, const uint32_t *pos,
                     const uint32_t *wordidx, const uint32_t *posidx) {
  if (!ctx || !idx || (npos && !pos) || (typ != 0 && (!wordidx || !posidx))) return SMB_ERR_ARG;
  if (wordlen < 1 || wordlen > 31 || nskip < 1 || nskip > 32 || nbits_key > 32 || (typ != 0 && nbits_lo >= nbits_key) ||
      (typ == 0 && wordlen > 15))   // perfect hash: 4^k keys must fit 32 bits (as smb_index_build)
    return fail(ctx, SMB_ERR_ARG, "index parameters out of range (k=%d nskip=%d typ=%d)", wordlen, nskip, typ);
  cudaSetDevice(ctx->device);
  Index ix{};
  ix.typ = typ; ix.wordlen = wordlen; ix.nskip = nskip; ix.nbits_key = nbits_key; ix.nbits_lo = nbits_lo;
  ix.npos = npos; ix.nwords = nwords;
  ix.wordmask = (wordlen >= 32) ? ~0ull : ((1ull << (2 * wordlen)) - 1ull);   // hashTableCreate, hashidx.c:781-787
  if (typ == 0) {
    ix.nkeys = 1u << (2 * wordlen);
  } else {
    ix.nkeys = 1u << nbits_key;
    ix.wordmask_lo = (1ull << nbits_lo) - 1ull;
    ix.wordmask_hi = ix.wordmask & ~ix.wordmask_lo;
    ix.keymod = 1u << (nbits_key - nbits_lo);
  }
  const size_t n_idx = (size_t)ix.nkeys + 1, n_pos = npos, n_w = typ ? (size_t)nwords + 1 : 0;
  auto al = [](size_t n) { return (n + 63) & ~(size_t)63; };
  const size_t total = al(n_idx) + al(n_pos + 1) + 2 * al(n_w + 1);
  CU(ctx->index.ensure(total * sizeof(uint32_t)));
  uint32_t *base = ctx->index.as<uint32_t>();
  uint32_t *d_idx = base, *d_pos = d_idx + al(n_idx), *d_w = d_pos + al(n_pos + 1), *d_p = d_w + al(n_w + 1);
  cudaStream_t st = ctx->stream;
  CU(h2d(d_idx, idx, n_idx * 4, st));
  if (n_pos) CU(h2d(d_pos, pos, n_pos * 4, st));
  if (typ) {
    CU(h2d(d_w, wordidx, n_w * 4, st));
    CU(h2d(d_p, posidx, n_w * 4, st));
  }
  CU(ctx_sync(ctx));
  ix.idx = d_idx; ix.pos = d_pos; ix.wordidx = d_w; ix.posidx = d_p;
  ctx->ix = ix;
  ctx->have_index = true;
  return SMB_OK;
}

int smb_index_build(smb_ctx *ctx, int wordlen, int nskip, int typ, int nbits_key, int nbits_lo,
                    const smb_index_seq *seqs, int nseq, smb_index_info *info) {
  if (!ctx || !seqs || nseq < 1 || !info) return SMB_ERR_ARG;
  if (wordlen < 1 || wordlen > 31 || nskip < 1 || (typ == 0 && wordlen > 15) ||
      (typ != 0 && (nbits_key < 1 || nbits_key > 30 || nbits_lo < 0 || nbits_lo >= nbits_key || 2 * wordlen - nbits_lo > 32)))
    return fail(ctx, SMB_ERR_ARG, "index parameters out of range (k=%d nskip=%d typ=%d key bits %d/%d)", wordlen, nskip,
                typ, nbits_key, nbits_lo);
  if (!ctx->src.packed) return fail(ctx, SMB_ERR_STATE, "smb_refseq_upload() first");
  for (int i = 0; i < nseq; ++i)
    if (seqs[i].n_k && seqs[i].start + seqs[i].offs + (uint64_t)(seqs[i].n_k - 1) * nskip + wordlen > ctx->src.packed_nbases)
      return fail(ctx, SMB_ERR_ARG, "sequence %d: k-mer grid outside the uploaded reference", i);
  cudaSetDevice(ctx->device);
  if (ctx->built.block) { cudaFree(ctx->built.block); ctx->built = IndexBuildOut{}; }
  int nl = 0;
  static_assert(sizeof(smb_index_seq) == sizeof(IndexBuildSeq), "layout");
  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  CU(index_build(ctx->src.packed, (const IndexBuildSeq *)seqs, nseq, wordlen, nskip, typ, nbits_key, nbits_lo,
                 ctx->stream, &ctx->built, &nl));
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  CU(ctx_sync(ctx));
  CU(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
  ctx->built_typ = typ;
  ctx->last_launches = nl;
  ctx->total_launches += nl;
  g_launches += nl;
  info->npos = ctx->built.npos; info->nwords = ctx->built.nwords; info->nkeys = ctx->built.nkeys;
  info->kernel_ms = ctx->last_ms;
  return SMB_OK;
}

int smb_index_fetch(smb_ctx *ctx, uint32_t *idx, uint32_t *pos, uint32_t *wordidx, uint32_t *posidx) {
  if (!ctx) return SMB_ERR_ARG;
  if (!ctx->built.block) return fail(ctx, SMB_ERR_STATE, "smb_index_build() first");
  cudaSetDevice(ctx->device);
  const IndexBuildOut &b = ctx->built;
  if (idx) CU(d2h(idx, b.idx, ((size_t)b.nkeys + 1) * 4, ctx->stream));
  if (pos && b.npos) CU(d2h(pos, b.pos, (size_t)b.npos * 4, ctx->stream));
  if (ctx->built_typ != 0) {
    if (wordidx) CU(d2h(wordidx, b.wordidx, ((size_t)b.nwords + 1) * 4, ctx->stream));
    if (posidx) CU(d2h(posidx, b.posidx, ((size_t)b.nwords + 1) * 4, ctx->stream));
  }
  CU(ctx_sync(ctx));
  cudaFree(ctx->built.block);
  ctx->built = IndexBuildOut{};
  return SMB_OK;
}

static int seed_batch_impl(smb_ctx *ctx, const Index &ixt, const IndexTab *d_tab, const uint32_t *read_tab,
                           const uint64_t *read_off, const uint32_t *read_len, int nreads,
                           const uint8_t *qual, uint32_t maxhit_per_tuple, uint32_t maxhit_total, int basq_thresh,
                           int short_info, smb_seed_info *info, uint32_t *seed_posidx, uint32_t *seed_nhits, uint32_t *seed_qoffs,
                           uint32_t *sortkey, uint32_t *sidx, uint8_t *qmask) {
  if (!ctx || nreads < 0 || (nreads && (!read_off || !read_len || !info))) return SMB_ERR_ARG;
  ctx->last_ms = 0.f;
  ctx->last_launches = 0;
  if (!nreads) return SMB_OK;
  // the buffers of the previous batch may be reallocated below: its state is gone from here on and
  // the new one is published only on success
  ctx->seed_nreads = 0;
  ctx->hit_qmask_valid = false;
  if (!d_tab && !ctx->have_index) return fail(ctx, SMB_ERR_STATE, "smb_index_upload() first");
  if (!ctx->src.arena) return fail(ctx, SMB_ERR_STATE, "smb_arena_upload() first");
  if (basq_thresh < 0 || basq_thresh + 0x21 > 255) return fail(ctx, 67 /* ERRCODE_QUALVAL */, "quality threshold");
  std::vector<uint64_t> slot((size_t)nreads + 1, 0);
  for (int i = 0; i < nreads; ++i) {
    if (read_off[i] > ctx->arena_bytes || read_len[i] > ctx->arena_bytes - read_off[i])
      return fail(ctx, SMB_ERR_ARG, "read %d outside the arena", i);
    slot[(size_t)i + 1] = slot[(size_t)i] + 2ull * read_len[i];
  }
  const uint64_t nslots = slot[(size_t)nreads];
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const size_t meta_bytes = (size_t)nreads * (8 + 4 + 8 + 4) + (size_t)2 * nreads * sizeof(smb_seed_info) + 256;
  CU(ctx->seed_meta.ensure(meta_bytes));
  CU(ctx->seed_u32.ensure((size_t)(nslots + 64) * 6 * sizeof(uint32_t)));
  CU(ctx->seed_u8.ensure((size_t)(nslots + 64) * 2));
  char *mb = ctx->seed_meta.as<char>();
  uint64_t *d_off = (uint64_t *)mb;
  uint64_t *d_slot = d_off + nreads;
  smb_seed_info *d_info = (smb_seed_info *)(d_slot + nreads);
  uint32_t *d_len = (uint32_t *)(d_info + 2 * (size_t)nreads);
  uint32_t *d_rtab = d_len + nreads;
  if (d_tab) CU(h2d(d_rtab, read_tab, (size_t)nreads * 4, st));
  CU(h2d(d_off, read_off, (size_t)nreads * 8, st));
  CU(h2d(d_slot, slot.data(), (size_t)nreads * 8, st));
  CU(h2d(d_len, read_len, (size_t)nreads * 4, st));
  const uint8_t *d_qual = nullptr;
  if (qual) {
    CU(ctx->qualbuf.ensure(ctx->arena_bytes + 16));
    CU(h2d(ctx->qualbuf.p, qual, ctx->arena_bytes, st));
    d_qual = ctx->qualbuf.as<uint8_t>();
  }
  const size_t S = (size_t)nslots + 64;
  uint32_t *u = ctx->seed_u32.as<uint32_t>();
  uint8_t *b = ctx->seed_u8.as<uint8_t>();
  SeedArgs a{};
  a.tab = d_tab; a.read_tab = d_tab ? d_rtab : nullptr;
  a.read_off = d_off; a.read_len = d_len; a.slot_off = d_slot; a.qual = d_qual; a.nreads = nreads;
  a.maxhit_per_tuple = maxhit_per_tuple; a.maxhit_total = maxhit_total; a.basq_thresh = basq_thresh;
  a.is_short = short_info ? 1 : 0;
  a.maxlen = 0;
  for (int i = 0; i < nreads; ++i) if (read_len[i] > a.maxlen) a.maxlen = read_len[i];
  a.info = d_info;
  a.posidx = u; a.nhits = u + S; a.qoffs = u + 2 * S; a.sortkey = u + 3 * S; a.sidx = u + 4 * S; a.frame = u + 5 * S;
  a.qmask = b; a.qbuf = b + S;
  int nl = 0;
  CU(cudaEventRecord(ctx->ev0, st));
  CU(launch_seed(ixt, ctx->src.arena, a, st, &nl));
  CU(cudaEventRecord(ctx->ev1, st));
  CU(d2h(info, d_info, (size_t)2 * nreads * sizeof(smb_seed_info), st));
  struct { uint32_t *h; uint32_t *d; } cp[] = {{seed_posidx, a.posidx}, {seed_nhits, a.nhits}, {seed_qoffs, a.qoffs},
                                                {sortkey, a.sortkey}, {sidx, a.sidx}};
  for (auto &c : cp)
    if (c.h) CU(d2h(c.h, c.d, (size_t)nslots * 4, st));
  if (qmask) CU(d2h(qmask, a.qmask, (size_t)nslots, st));
  CU(ctx_sync(ctx));
  CU(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
  ctx->seed_nreads = nreads;
  ctx->seed_len.assign(read_len, read_len + nreads);
  ctx->seed_args = a;
  ctx->seed_ix = ixt;
  ctx->seed_maxlen = 0;
  for (int i = 0; i < nreads; ++i) if (read_len[i] > ctx->seed_maxlen) ctx->seed_maxlen = read_len[i];
  ctx->seed_slots = nslots;
  ctx->last_launches = nl;
  ctx->total_launches += nl;
  g_launches += nl;
  return SMB_OK;
}

int smb_seed_batch(smb_ctx *ctx, const uint64_t *read_off, const uint32_t *read_len, int nreads,
                   const uint8_t *qual, uint32_t maxhit_per_tuple, uint32_t maxhit_total, int basq_thresh,
                   int short_info, smb_seed_info *info, uint32_t *seed_posidx, uint32_t *seed_nhits, uint32_t *seed_qoffs,
                   uint32_t *sortkey, uint32_t *sidx, uint8_t *qmask) {
  if (!ctx) return SMB_ERR_ARG;
  return seed_batch_impl(ctx, ctx->ix, nullptr, nullptr, read_off, read_len, nreads, qual, maxhit_per_tuple, maxhit_total,
                         basq_thresh, short_info, info, seed_posidx, seed_nhits, seed_qoffs, sortkey, sidx, qmask);
}

int smb_seed_batch_tables(smb_ctx *ctx, int wordlen, int nskip, const smb_small_index *tables, int ntables,
                          const uint32_t *read_table, const uint64_t *read_off, const uint32_t *read_len, int nreads,
                          const uint8_t *qual, uint32_t maxhit_per_tuple, uint32_t maxhit_total, int basq_thresh,
                          int short_info, smb_seed_info *info) {
  if (!ctx || ntables < 0 || nreads < 0 || (nreads && (!tables || !read_table || ntables < 1))) return SMB_ERR_ARG;
  if (wordlen < 1 || wordlen > 12 || nskip < 1 || nskip > 32)
    return fail(ctx, SMB_ERR_ARG, "small index parameters out of range (k=%d nskip=%d)", wordlen, nskip);
  ctx->last_ms = 0.f;
  ctx->last_launches = 0;
  if (!nreads) return SMB_OK;
  for (int i = 0; i < nreads; ++i)
    if (read_table[i] >= (uint32_t)ntables) return fail(ctx, SMB_ERR_ARG, "read %d: table %u out of range", i, read_table[i]);
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const size_t nkeys = (size_t)1 << (2 * wordlen), n_idx = nkeys + 1;
  auto al = [](size_t n) { return (n + 63) & ~(size_t)63; };
  size_t words = al((size_t)ntables * sizeof(IndexTab) / 4 + 16);
  std::vector<size_t> off((size_t)ntables);
  for (int t = 0; t < ntables; ++t) {
    if (!tables[t].idx || (tables[t].npos && !tables[t].pos)) return SMB_ERR_ARG;
    off[(size_t)t] = words;
    words += al(n_idx) + al((size_t)tables[t].npos + 1);
  }
  CU(ctx->aux_index.ensure(words * sizeof(uint32_t)));
  uint32_t *base = ctx->aux_index.as<uint32_t>();
  std::vector<IndexTab> tab((size_t)ntables);
  for (int t = 0; t < ntables; ++t) {
    uint32_t *d_idx = base + off[(size_t)t], *d_pos = d_idx + al(n_idx);
    CU(h2d(d_idx, tables[t].idx, n_idx * 4, st));
    if (tables[t].npos) CU(h2d(d_pos, tables[t].pos, (size_t)tables[t].npos * 4, st));
    tab[(size_t)t] = IndexTab{d_idx, d_pos, tables[t].npos, 0};
  }
  CU(h2d(base, tab.data(), (size_t)ntables * sizeof(IndexTab), st));
  CU(ctx_sync(ctx));   // the host vectors above go out of scope
  Index ixt{};
  ixt.typ = 0; ixt.wordlen = wordlen; ixt.nskip = nskip;
  ixt.wordmask = (1ull << (2 * wordlen)) - 1ull;
  ixt.nkeys = (uint32_t)nkeys;
  return seed_batch_impl(ctx, ixt, (const IndexTab *)base, read_table, read_off, read_len, nreads, qual, maxhit_per_tuple,
                         maxhit_total, basq_thresh, short_info, info, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int smb_hits_batch(smb_ctx *ctx, const smb_hit_req *req, int nreq, uint32_t nhits_alloc, uint64_t *sqdat,
                   size_t max_hits, size_t *nhits_total, uint64_t *list_first, int32_t *errs) {
  if (!ctx || nreq < 0 || !nhits_total || (nreq && (!req || !list_first || !errs))) return SMB_ERR_ARG;
  ctx->last_ms = 0.f;
  ctx->last_launches = 0;
  *nhits_total = 0;
  if (list_first) list_first[0] = 0;
  if (!nreq) return SMB_OK;
  if (!ctx->seed_nreads) return fail(ctx, SMB_ERR_STATE, "smb_seed_batch() first");
  for (int i = 0; i < nreq; ++i)
    if (req[i].read >= (uint32_t)ctx->seed_nreads || req[i].strand > 1)
      return fail(ctx, SMB_ERR_ARG, "request %d: read %u / strand %u out of range", i, req[i].read, req[i].strand);
  if (!nhits_alloc) {  // initHitList / reallocHitList (hashhit.c:1262-1296, :1232-1248), starting from
                       // hashCreateHitList(HASH_MAXNHITS = 16384) (rmap.c:50, :1123)
    const double ql = (double)ctx->seed_maxlen;
    double target = ql > 1 ? ql * log(ql) * 32.0 : 0.0;
    if (target > 2147483647.0) target = 2147483647.0;
    if (target < 8192.0) target = 8192.0;
    size_t t = (size_t)target;
    nhits_alloc = 16384;
    if (t > nhits_alloc) nhits_alloc = (uint32_t)((t + 16383) / 16384 * 16384);
  }
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)nreq;
  const int ntiles = compact_tiles(nreq);
  // list masks (HITQUAL per read offset) are produced when a whole-set (mode 2) list is asked for
  bool want_qmask = false;
  for (int i = 0; i < nreq && !want_qmask; ++i) want_qmask = req[i].use_short == 2;
  ctx->hit_qmask_valid = false;
  uint64_t *d_qoff = nullptr;
  if (want_qmask) {
    ctx->hit_qmask_first.assign(n + 1, 0);
    for (size_t i = 0; i < n; ++i) ctx->hit_qmask_first[i + 1] = ctx->hit_qmask_first[i] + ctx->seed_len[req[i].read];
    CU(ctx->hit_qmask.ensure(ctx->hit_qmask_first[n] + (n + 1) * 8 + 64));
    d_qoff = (uint64_t *)(ctx->hit_qmask.as<char>() + ((ctx->hit_qmask_first[n] + 15) & ~(uint64_t)15));
    CU(h2d(d_qoff, ctx->hit_qmask_first.data(), (n + 1) * 8, ctx->stream));
  }
  CU(ctx->hit_meta.ensure(n * (sizeof(smb_hit_req) + 4 + 4 + 4) + (n + 1 + (size_t)ntiles) * 8 + 512));
  char *mb = ctx->hit_meta.as<char>();
  smb_hit_req *d_req = (smb_hit_req *)mb;
  uint64_t *d_off = (uint64_t *)(d_req + n);                 // n + 1 offsets
  unsigned long long *d_tile = (unsigned long long *)(d_off + n + 1);
  uint32_t *d_count = (uint32_t *)(d_tile + ntiles);
  uint32_t *d_used = d_count + n;
  int32_t *d_errs = (int32_t *)(d_used + n);
  CU(h2d(d_req, req, n * sizeof(smb_hit_req), st));
  HitArgs ha{};
  ha.seed = ctx->seed_args; ha.req = d_req; ha.nreq = nreq; ha.nhits_alloc = nhits_alloc;
  ha.count = d_count; ha.maxhit_used = d_used; ha.errs = d_errs; ha.offset = d_off; ha.sqdat = nullptr;
  ha.req_skip = nullptr;
  ha.list_qmask = want_qmask ? ctx->hit_qmask.as<uint8_t>() : nullptr;
  ha.qmask_off = d_qoff;
  int nl = 0;
  float ms0 = 0.f, ms1 = 0.f;
  // pass 1: list sizes, then their offsets by a device scan (only the offsets travel to the host)
  CU(cudaEventRecord(ctx->ev0, st));
  CU(launch_hits(ctx->seed_ix, ha, false, st, &nl));
  CU(launch_scan_counts(d_count, nreq, (unsigned long long *)d_off, d_tile, st, &nl));
  CU(cudaEventRecord(ctx->ev1, st));
  CU(d2h(list_first, d_off, (n + 1) * 8, st));
  CU(d2h(errs, d_errs, n * 4, st));
  CU(ctx_sync(ctx));
  CU(cudaEventElapsedTime(&ms0, ctx->ev0, ctx->ev1));
  const uint64_t total = list_first[n];
  *nhits_total = (size_t)total;
  if (total > max_hits || (total && !sqdat)) {
    ctx->last_ms = ms0;
    ctx->last_launches = nl;
    ctx->total_launches += nl;
    g_launches += nl;
    return fail(ctx, SMB_ERR_CAPACITY, "need room for %llu hits", (unsigned long long)total);
  }
  CU(ctx->hit_data.ensure((size_t)(total + 1) * 8));
  ha.sqdat = ctx->hit_data.as<uint64_t>();
  CU(cudaEventRecord(ctx->ev0, st));
  CU(launch_hits(ctx->seed_ix, ha, true, st, &nl));
  CU(cudaEventRecord(ctx->ev1, st));
  if (total) CU(d2h(sqdat, ha.sqdat, (size_t)total * 8, st));
  CU(ctx_sync(ctx));
  CU(cudaEventElapsedTime(&ms1, ctx->ev0, ctx->ev1));
  ctx->last_ms = ms0 + ms1;
  ctx->last_launches = nl;
  ctx->total_launches += nl;
  g_launches += nl;
  ctx->hit_qmask_valid = want_qmask;
  return SMB_OK;
}

int smb_hits_qmask(smb_ctx *ctx, uint8_t *qmask, size_t max_bytes, uint64_t *qmask_first) {
  if (!ctx || !qmask_first) return SMB_ERR_ARG;
  if (!ctx->hit_qmask_valid) return fail(ctx, SMB_ERR_STATE, "no whole-set (mode 2) hit lists in the last smb_hits_batch");
  const size_t n = ctx->hit_qmask_first.size() - 1, total = (size_t)ctx->hit_qmask_first[n];
  memcpy(qmask_first, ctx->hit_qmask_first.data(), (n + 1) * 8);
  if (total > max_bytes || (total && !qmask)) return fail(ctx, SMB_ERR_CAPACITY, "need %zu mask bytes", total);
  cudaSetDevice(ctx->device);
  if (total) CU(d2h(qmask, ctx->hit_qmask.p, total, ctx->stream));
  CU(ctx_sync(ctx));
  return SMB_OK;
}

}  // extern "C"
