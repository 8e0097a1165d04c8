// compact.cu - device-side compaction of the K3 outputs.
//
// band_kernel<true> writes the results of task t into fixed slots (max_res results, a
// per-task DiffStr area).  Shipping those slots to the host costs ~1 KB per task although a
// typical task produces one 32-byte result and a DiffStr of a dozen bytes, so the dense
// arrays of the C ABI (smb_band_align_batch: results in task order, first_result[], one
// diffstr buffer) are assembled on the device: two exclusive scans (results, DiffStr bytes)
// and one gather kernel; only the dense arrays cross PCIe.
#include "common.cuh"
#include "cigar.cuh"

namespace smb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;                       // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long *sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long w = (threadIdx.x < SCAN_THREADS / 32) ? sh[threadIdx.x] : 0ull;
    for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
    if (threadIdx.x == 0) sh[0] = w;
  }
  __syncthreads();
  const unsigned long long r = sh[0];
  __syncthreads();
  return r;
}

// phase 1: per-tile sums of nres[] and dused[]; flags tasks that ran out of slot capacity
__global__ void __launch_bounds__(SCAN_THREADS)
scan_tiles(const uint32_t *__restrict__ nres, const uint32_t *__restrict__ dused, const uint32_t *__restrict__ cig,
           const int32_t *__restrict__ errs, int n, unsigned long long *__restrict__ tile_res,
           unsigned long long *__restrict__ tile_diff, unsigned long long *__restrict__ tile_cig,
           CompactTotals *__restrict__ tot) {
  __shared__ unsigned long long sh[32];
  const int base = blockIdx.x * SCAN_TILE;
  unsigned long long a = 0, b = 0, c = 0;
  bool cap = false;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int i = base + k * SCAN_THREADS + threadIdx.x;
    if (i < n) {
      a += nres[i];
      b += dused[i];
      if (cig) c += cig[i];
      cap |= errs[i] == SMB_ERR_CAPACITY;
    }
  }
  a = block_sum(a, sh);
  b = block_sum(b, sh);
  if (cig) c = block_sum(c, sh);
  if (threadIdx.x == 0) {
    tile_res[blockIdx.x] = a; tile_diff[blockIdx.x] = b;
    if (cig) tile_cig[blockIdx.x] = c;
  }
  if (cap) tot->capacity_flag = 1;
}

// phase 2: exclusive scan of the tile sums (a few thousand values: one warp, serial chunks)
__global__ void scan_top(unsigned long long *tile_res, unsigned long long *tile_diff, unsigned long long *tile_cig,
                         int ntiles, CompactTotals *tot) {
  const int lane = threadIdx.x;
  unsigned long long carry_a = 0, carry_b = 0, carry_c = 0;
  for (int base = 0; base < ntiles; base += 32) {
    const int i = base + lane;
    unsigned long long a = (i < ntiles) ? tile_res[i] : 0ull, b = (i < ntiles) ? tile_diff[i] : 0ull;
    unsigned long long c = (tile_cig && i < ntiles) ? tile_cig[i] : 0ull;
    unsigned long long ia = a, ib = b, ic = c;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
      const unsigned long long tc = __shfl_up_sync(0xffffffffu, ic, o);
      if (lane >= o) { ia += ta; ib += tb; ic += tc; }
    }
    if (i < ntiles) {
      tile_res[i] = carry_a + ia - a; tile_diff[i] = carry_b + ib - b;
      if (tile_cig) tile_cig[i] = carry_c + ic - c;
    }
    carry_a += __shfl_sync(0xffffffffu, ia, 31);
    carry_b += __shfl_sync(0xffffffffu, ib, 31);
    carry_c += __shfl_sync(0xffffffffu, ic, 31);
  }
  if (lane == 0) { tot->nresults = carry_a; tot->ndiff = carry_b; tot->ncig = carry_c; }
}

// phase 3: per-task exclusive offsets
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply(const uint32_t *__restrict__ nres, const uint32_t *__restrict__ dused, const uint32_t *__restrict__ cig, int n,
           const unsigned long long *__restrict__ tile_res, const unsigned long long *__restrict__ tile_diff,
           const unsigned long long *__restrict__ tile_cig,
           uint32_t *__restrict__ first_result, unsigned long long *__restrict__ diff_first,
           unsigned long long *__restrict__ cig_first) {
  __shared__ unsigned long long sa[SCAN_THREADS], sb[SCAN_THREADS], sc[SCAN_THREADS];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t va[SCAN_ITEMS], vb[SCAN_ITEMS], vc[SCAN_ITEMS];
  unsigned long long a = 0, b = 0, c = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int i = base + k;
    va[k] = (i < n) ? nres[i] : 0u;
    vb[k] = (i < n) ? dused[i] : 0u;
    vc[k] = (cig && i < n) ? cig[i] : 0u;
    a += va[k];
    b += vb[k];
    c += vc[k];
  }
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  sc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 1; o < SCAN_THREADS; o <<= 1) {   // Hillis-Steele inclusive scan of the thread sums
    unsigned long long ta = 0, tb = 0, tc = 0;
    if ((int)threadIdx.x >= o) { ta = sa[threadIdx.x - o]; tb = sb[threadIdx.x - o]; tc = sc[threadIdx.x - o]; }
    __syncthreads();
    sa[threadIdx.x] += ta;
    sb[threadIdx.x] += tb;
    sc[threadIdx.x] += tc;
    __syncthreads();
  }
  unsigned long long oa = tile_res[blockIdx.x] + sa[threadIdx.x] - a;
  unsigned long long ob = tile_diff[blockIdx.x] + sb[threadIdx.x] - b;
  unsigned long long oc = (cig ? tile_cig[blockIdx.x] : 0ull) + sc[threadIdx.x] - c;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int i = base + k;
    if (i < n) {
      first_result[i] = (uint32_t)oa; diff_first[i] = ob;
      if (cig) cig_first[i] = oc;
    }
    oa += va[k];
    ob += vb[k];
    oc += vc[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) first_result[n] = (uint32_t)oa;
}

// gather: results and DiffStr bytes of task t -> dense arrays
__global__ void __launch_bounds__(128)
gather_results(const smb_ali_result *__restrict__ slots, const uint32_t *__restrict__ nres,
               const uint8_t *__restrict__ diff_slots, const uint64_t *__restrict__ diff_off,
               const uint32_t *__restrict__ dused, int n, int max_res,
               const uint32_t *__restrict__ first_result, const unsigned long long *__restrict__ diff_first,
               smb_ali_result *__restrict__ out_res, uint8_t *__restrict__ out_diff, const GatherCigar cg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t nr = nres[t];
  const unsigned long long dbase = diff_first[t];
  const smb_ali_result *src = slots + (size_t)t * max_res;
  smb_ali_result *dst = out_res + first_result[t];
  const uint8_t *ds = diff_slots + diff_off[t];
  unsigned long long co = 0;
  uint32_t qlen = 0;
  if (cg.text) {   // output stage: CIGAR text, its offset and the edit distance of every alignment of the task
    co = cg.cig_first[t];
    qlen = cg.tasks[t].read_len;
    if (t == 0) cg.first_out[cg.nres_total] = (uint32_t)cg.ncig_total;
  }
  for (uint32_t k = 0; k < nr; ++k) {
    smb_ali_result r = src[k];
    if (cg.text) {
      int nm;
      cg.first_out[first_result[t] + k] = (uint32_t)co;
      co += (unsigned long long)cg_walk<true>(ds + r.diff_off, (uint32_t)r.qs, qlen - 1u - (uint32_t)r.qe, cg.flags,
                                              cg.text + co, &nm);
      cg.nm[first_result[t] + k] = nm;
    }
    r.diff_off += (uint32_t)dbase;
    r.task = (uint32_t)t;
    dst[k] = r;
  }
  uint8_t *dd = out_diff + dbase;
  const uint32_t nb = dused[t];
  for (uint32_t k = 0; k < nb; ++k) dd[k] = ds[k];
}


int compact_tiles(int n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// ---- exclusive scan of u32 counts into u64 offsets (hit-list offsets of smb_hits_batch) ----
__global__ void __launch_bounds__(SCAN_THREADS)
scan1_tiles(const uint32_t *__restrict__ in, int n, unsigned long long *__restrict__ tile) {
  __shared__ unsigned long long sh[32];
  const int base = blockIdx.x * SCAN_TILE;
  unsigned long long a = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int i = base + k * SCAN_THREADS + threadIdx.x;
    if (i < n) a += in[i];
  }
  a = block_sum(a, sh);
  if (threadIdx.x == 0) tile[blockIdx.x] = a;
}

__global__ void scan1_top(unsigned long long *tile, int ntiles) {
  const int lane = threadIdx.x;
  unsigned long long carry = 0;
  for (int base = 0; base < ntiles; base += 32) {
    const int i = base + lane;
    const unsigned long long a = (i < ntiles) ? tile[i] : 0ull;
    unsigned long long ia = a;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, ia, o);
      if (lane >= o) ia += t;
    }
    if (i < ntiles) tile[i] = carry + ia - a;
    carry += __shfl_sync(0xffffffffu, ia, 31);
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan1_apply(const uint32_t *__restrict__ in, int n, const unsigned long long *__restrict__ tile,
            unsigned long long *__restrict__ out) {
  __shared__ unsigned long long sa[SCAN_THREADS];
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  unsigned long long a = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    a += v[k];
  }
  sa[threadIdx.x] = a;
  __syncthreads();
  for (int o = 1; o < SCAN_THREADS; o <<= 1) {
    unsigned long long t = 0;
    if ((int)threadIdx.x >= o) t = sa[threadIdx.x - o];
    __syncthreads();
    sa[threadIdx.x] += t;
    __syncthreads();
  }
  unsigned long long o = tile[blockIdx.x] + sa[threadIdx.x] - a;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = o;
    o += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) out[n] = o;
}

// the same scan for short arrays in ONE launch: one CTA, every thread sums a run of consecutive items,
// the thread sums are scanned through the warps (a block of reads has a few thousand jobs / requests)
constexpr int SCAN1_SMALL_THREADS = 1024;
constexpr int SCAN1_SMALL_MAX = SCAN1_SMALL_THREADS * 32;
// one CTA: warp w scans the contiguous chunk [w * per, (w + 1) * per) with coalesced accesses (32 elements per
// step: a warp scan, the running total carried in a register); chunk totals are scanned in between
__global__ void __launch_bounds__(SCAN1_SMALL_THREADS)
scan1_small(const uint32_t *__restrict__ in, int n, unsigned long long *__restrict__ out) {
  __shared__ unsigned long long s_warp[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per = (((n + 31) / 32) + 31) & ~31;          // elements per warp, a multiple of 32
  const int b = w * per, e = min(n, b + per);
  unsigned long long a = 0;
  for (int i = b + lane; i < e; i += 32) a += in[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) s_warp[w] = a;
  __syncthreads();
  if (w == 0) {
    const unsigned long long v = s_warp[lane];
    unsigned long long iw = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, iw, o);
      if (lane >= o) iw += t;
    }
    s_warp[lane] = iw - v;   // exclusive over the warps
  }
  __syncthreads();
  unsigned long long run = s_warp[w];
  for (int i0 = b; i0 < e; i0 += 32) {
    const int i = i0 + lane;
    const unsigned long long v = i < e ? in[i] : 0u;
    unsigned long long incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (i < e) out[i] = run + incl - v;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (w == 31 && lane == 0) out[n] = run;
}

// out has n + 1 entries (out[n] = total); tile: compact_tiles(n) scratch words
cudaError_t launch_scan_counts(const uint32_t *in, int n, unsigned long long *out, unsigned long long *tile,
                               cudaStream_t st, int *nlaunch) {
  if (n <= SCAN1_SMALL_MAX) {
    scan1_small<<<1, SCAN1_SMALL_THREADS, 0, st>>>(in, n, out);
    *nlaunch += 1;
    return cudaGetLastError();
  }
  const int ntiles = compact_tiles(n);
  scan1_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(in, n, tile);
  scan1_top<<<1, 32, 0, st>>>(tile, ntiles);
  scan1_apply<<<ntiles, SCAN_THREADS, 0, st>>>(in, n, tile, out);
  *nlaunch += 3;
  return cudaGetLastError();
}


cudaError_t launch_compact_scan(const uint32_t *nres, const uint32_t *dused, const int32_t *errs, int n,
                                unsigned long long *tile_res, unsigned long long *tile_diff,
                                CompactTotals *tot, uint32_t *first_result,
                                unsigned long long *diff_first, cudaStream_t st, int *nlaunch,
                                const uint32_t *cig, unsigned long long *tile_cig, unsigned long long *cig_first) {
  const int ntiles = compact_tiles(n);
  cudaError_t e;
  if ((e = cudaMemsetAsync(tot, 0, sizeof(CompactTotals), st)) != cudaSuccess) return e;
  scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(nres, dused, cig, errs, n, tile_res, tile_diff, tile_cig, tot);
  scan_top<<<1, 32, 0, st>>>(tile_res, tile_diff, cig ? tile_cig : nullptr, ntiles, tot);
  scan_apply<<<ntiles, SCAN_THREADS, 0, st>>>(nres, dused, cig, n, tile_res, tile_diff, tile_cig, first_result, diff_first,
                                              cig_first);
  *nlaunch += 3;
  return cudaGetLastError();
}

cudaError_t launch_compact_gather(const smb_ali_result *slots, const uint32_t *nres, const uint8_t *diff_slots,
                                  const uint64_t *diff_off, const uint32_t *dused, int n, int max_res,
                                  const uint32_t *first_result, const unsigned long long *diff_first,
                                  smb_ali_result *out_res, uint8_t *out_diff, cudaStream_t st, int *nlaunch,
                                  const GatherCigar *cg) {
  GatherCigar g{};
  if (cg) g = *cg;
  gather_results<<<(n + 127) / 128, 128, 0, st>>>(slots, nres, diff_slots, diff_off, dused, n, max_res,
                                                  first_result, diff_first, out_res, out_diff, g);
  ++*nlaunch;
  return cudaGetLastError();
}

cudaError_t warm_compact() {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, scan_tiles);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, gather_results);
  return e;
}

}  // namespace smb
