// ctx.h - the context behind include/smalt_b200.h (private to csrc/): device buffers, stream, counters.
#pragma once
#include "common.cuh"
#include "band.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

using namespace smb;

struct DevBuf {  // grow-only device buffer
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T *as() const { return (T *)p; }
};

struct HostBuf {  // grow-only pinned host staging buffer
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T *as() const { return (T *)p; }
};

constexpr int SMB_BLK_EVENTS = 48;   // event pairs of the timed spans of a block (api_block.cu)

struct smb_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_done = nullptr;  // blocking-sync event: waiting host threads sleep instead of spinning
  cudaEvent_t ev_spin = nullptr;  // polled event (smb_ctx_set_spin)
  int spin = 0;                   // 0 block, 1 poll, 2 poll with sleeps of poll_ns (smb_ctx_set_spin)
  long poll_ns = 20000;
  BandSide side;                  // K3: stream of the small launches beside the packed kernel
  HostBuf stage;                  // pinned staging for the library's own host-side arrays
  DevBuf cmp;                     // K3 output compaction scratch
  DevBuf ticket;                  // work counters of persistent kernels
  Scoring sc;
  SeqSrc src{nullptr, nullptr, 0};
  DevBuf arena, packed, tasks, out_a, out_b, scratch, dirs, diff, offs;
  DevBuf index, qualbuf, seed_meta, seed_u32, seed_u8;
  Index ix{};
  bool have_index = false;
  // device-resident seed tables of the last smb_seed_batch (consumed by smb_hits_batch)
  int seed_nreads = 0;
  SeedArgs seed_args{};
  uint32_t seed_maxlen = 0;
  DevBuf hit_meta, hit_data, hit_qmask, aux_index;
  IndexBuildOut built{};               // arrays left on the device by smb_index_build
  int built_typ = 0;
  Index seed_ix{};                     // index (template) of the last seed batch: what smb_hits_batch reads
  std::vector<uint32_t> seed_len;      // host copy of the read lengths of the last smb_seed_batch
  std::vector<uint64_t> hit_qmask_first;  // per request of the last smb_hits_batch
  bool hit_qmask_valid = false;
  uint64_t seed_slots = 0;
  size_t arena_bytes = 0;
  std::vector<uint64_t> seq_offs;
  DevBuf seq_offs_buf;                 // device copy of seq_offs (owner context only)
  const uint64_t *d_seq_offs = nullptr;
  // resident block pipeline (api_block.cu)
  DevBuf blk_jobs, blk_scr, blk_cand, blk_k3, blk_cig, blk_cigtext;
  cudaEvent_t blk_ev[SMB_BLK_EVENTS] = {};
  struct BlockState *blk = nullptr;
  float last_ms = 0.f;
  int last_launches = 0;
  long long total_launches = 0;
  std::string err;
};

// process-wide traffic counters (all contexts): what bench.py reports as h2d/d2h bytes and launches
extern std::atomic<unsigned long long> g_h2d_bytes, g_d2h_bytes;
extern std::atomic<long long> g_launches;

static inline cudaError_t h2d(void *dst, const void *src, size_t bytes, cudaStream_t st) {
  g_h2d_bytes += bytes;
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
}
static inline cudaError_t d2h(void *dst, const void *src, size_t bytes, cudaStream_t st) {
  g_d2h_bytes += bytes;
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
}

static inline int fail(smb_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(ctx, SMB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                   \
  } while (0)

// Waits for the context's stream.  Uses a blocking-sync event so that a host worker thread
// yields its core while the GPU works (one context per host thread, more threads than cores).
static inline cudaError_t ctx_sync(smb_ctx *ctx) {
  if (ctx->spin && ctx->ev_spin) {   // the calling thread polls: no wake-up latency on a busy host (smb_ctx_set_spin)
    cudaError_t e = cudaEventRecord(ctx->ev_spin, ctx->stream);
    if (e != cudaSuccess) return e;
    if (ctx->spin == 2) {   // poll with short sleeps: the latency of a poll interval, next to no CPU time
      struct timespec ts = {0, ctx->poll_ns > 0 ? ctx->poll_ns : 20000};
      while ((e = cudaEventQuery(ctx->ev_spin)) == cudaErrorNotReady) nanosleep(&ts, nullptr);
      return e;
    }
    while ((e = cudaEventQuery(ctx->ev_spin)) == cudaErrorNotReady) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    return e;
  }
  cudaError_t e = cudaEventRecord(ctx->ev_done, ctx->stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ctx->ev_done);
}

// api.cu: K3 with growing per-task capacities, results assembled on the host (outputs in task order)
int band_align_multipass(smb_ctx *ctx, const smb_band_task *tasks, int ntasks, std::vector<smb_ali_result> &results,
                         std::vector<uint32_t> &first_result, std::vector<uint8_t> &diffstr,
                         std::vector<int32_t> &errs, uint64_t *ncells);
void block_state_free(smb_ctx *ctx);   // api_block.cu
