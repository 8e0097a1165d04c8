// index_build.cu - hash index construction on the GPU.
//
// Builds the arrays of the reference's `struct _HashTable` (/root/reference/src/hashidx.c:105-146)
// for a whole set of reference sequences - what hashTableSetUp does in two (perfect type) or five
// (collision type) scalar passes over the set (hashidx.c:829-998, doWordsInSeq :465-531) - so that
// the `.smi` file comes out byte for byte like `smalt index`:
//   idx[nkeys+1]     first word (collision type) / first position (perfect type) of every key
//   pos[npos]        k-mer serial numbers, grouped by word, ascending inside a word
//   wordidx[nwords]  upper word bits of every distinct word, ascending inside a key  (collision type)
//   posidx[nwords+1] first position of every word
// Layout rule of the reference: words are kept in scan order inside (key, upper bits) - i.e. a
// STABLE sort of the k-mer grid by the 64-bit composite key << 32 | upper bits.
//
//   1. words_kernel    one thread per k-mer grid position: 2k-bit word from the 3-bit packed
//                      reference, key = (hash32mix(hi) % keymod) << nbits_lo | lo (hashidx.c:155-172),
//                      grid positions over a non-standard base are dropped (stream compaction by
//                      a device scan keeps the scan order)
//   2. radix sort      stable LSD radix sort of (composite, serial number), 8 bits per pass, only over the
//                      bit ranges the composite occupies: per pass a digit histogram per 4096-element tile
//                      (ib_rs_hist_kernel), one exclusive scan of the digit-major count matrix, and a
//                      scatter that ranks every element inside its tile in input order
//                      (ib_rs_scatter_kernel: warp chunks, __match_any_sync peer groups)
//   3. words / keys    word boundaries -> wordidx, posidx; histogram of keys -> idx (scans)
// Scans are ib_scan (tile sums, recursion over the tile sums, tile-local scan + offset).  No library
// primitives: every kernel launched here is in this file.
// The grid bookkeeping between sequences (offset of the first k-mer of a sequence, serial numbers
// that count skipped positions; hashidx.c:498-529) is done per sequence on the host and passed in.
#include "common.cuh"
#include <vector>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>

namespace smb {

__device__ __forceinline__ uint32_t ib_hash32mix(uint32_t a) {   // hashidx.c:163-172
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

struct IbParams {
  const uint32_t *packed;
  const IndexBuildSeq *seqs;
  const uint64_t *seq_first;   // [nseq+1] first grid position of every sequence
  int nseq, k, nskip, typ, nbits_lo;
  uint32_t keymod;
  uint64_t ngrid;
};

// composite key of every grid position (~0 = dropped) and its serial number
__global__ void __launch_bounds__(256)
ib_words_kernel(const IbParams p, unsigned long long *__restrict__ comp, uint32_t *__restrict__ serial,
                uint32_t *__restrict__ valid) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= p.ngrid) return;
  int lo = 0, hi = p.nseq;            // sequence of this grid position
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(p.seq_first + mid) <= g) lo = mid; else hi = mid;
  }
  const IndexBuildSeq s = p.seqs[lo];
  const uint64_t local = g - __ldg(p.seq_first + lo);
  const uint64_t base = s.start + s.offs + local * (uint64_t)p.nskip;
  unsigned long long w = 0;
  bool ok = true;
  for (int b = 0; b < p.k; ++b) {
    const uint32_t c = packed_base(p.packed, base + (uint64_t)b);
    ok &= c < 4u;
    w = (w << 2) | (unsigned long long)(c & 3u);
  }
  unsigned long long c64;
  if (p.typ == 0) {
    c64 = w << 32;
  } else {
    const uint32_t word_hi = (uint32_t)(w >> p.nbits_lo);
    const unsigned long long key = ((unsigned long long)(ib_hash32mix(word_hi) % p.keymod) << p.nbits_lo) +
                                   (w & ((1ull << p.nbits_lo) - 1ull));
    c64 = (key << 32) | word_hi;
  }
  comp[g] = c64;
  serial[g] = s.tup_base + (uint32_t)local;
  valid[g] = ok ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
ib_compact_kernel(const uint64_t n, const unsigned long long *__restrict__ comp, const uint32_t *__restrict__ serial,
                  const uint32_t *__restrict__ valid, const uint32_t *__restrict__ dst,
                  unsigned long long *__restrict__ ocomp, uint32_t *__restrict__ oserial) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n || !valid[g]) return;
  ocomp[dst[g]] = comp[g];
  oserial[dst[g]] = serial[g];
}

// flag[i] = 1 where a new word starts (collision type) ; key histogram of the words / positions
__global__ void __launch_bounds__(256)
ib_flag_kernel(const uint32_t npos, const unsigned long long *__restrict__ comp, uint32_t *__restrict__ flag,
               uint32_t *__restrict__ keycount, const int typ) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npos) return;
  const unsigned long long c = comp[i];
  const bool first = i == 0 || comp[i - 1] != c;
  flag[i] = first ? 1u : 0u;
  if (typ == 0 || first) atomicAdd(keycount + (uint32_t)(c >> 32), 1u);
}

__global__ void __launch_bounds__(256)
ib_words_out_kernel(const uint32_t npos, const unsigned long long *__restrict__ comp, const uint32_t *__restrict__ flag,
                    const uint32_t *__restrict__ wordno, uint32_t *__restrict__ wordidx, uint32_t *__restrict__ posidx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npos || !flag[i]) return;
  const uint32_t w = wordno[i];
  wordidx[w] = (uint32_t)comp[i];
  posidx[w] = i;
}

// ------------------------------------------------------------------------------------------------------
// exclusive prefix sum of u32 values (totals < 2^32), any length: tiles of 2048 values
// ------------------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256, SC_ITEMS = 8, SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ uint32_t ib_block_exscan(const uint32_t v, uint32_t *total) {   // exclusive over the CTA
  __shared__ uint32_t wsum[SC_THREADS / 32];
  __shared__ uint32_t wtot;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SC_THREADS / 32 ? wsum[lane] : 0u, winc = w;
#pragma unroll
    for (int o = 1; o < SC_THREADS / 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += u;
    }
    if (lane < SC_THREADS / 32) wsum[lane] = winc - w;
    if (lane == SC_THREADS / 32 - 1) wtot = winc;
  }
  __syncthreads();
  const uint32_t r = wsum[warp] + inc - v;
  if (total) *total = wtot;
  __syncthreads();   // wsum / wtot may be reused by the next call
  return r;
}

__global__ void __launch_bounds__(SC_THREADS)
ib_scan_reduce_kernel(const uint32_t *__restrict__ in, const uint64_t n, uint32_t *__restrict__ tilesum) {
  const uint64_t t0 = (uint64_t)blockIdx.x * SC_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) {
    const uint64_t i = t0 + (uint64_t)j * SC_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint32_t tot;
  (void)ib_block_exscan(s, &tot);
  if (threadIdx.x == 0) tilesum[blockIdx.x] = tot;
}

// out[i] = tileoffs[tile] + sum of in[tile start .. i)   (in == out allowed: a CTA reads its tile before it writes)
__global__ void __launch_bounds__(SC_THREADS)
ib_scan_apply_kernel(const uint32_t *in, uint32_t *out, const uint64_t n, const uint32_t *__restrict__ tileoffs) {
  const uint64_t t0 = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
  uint32_t v[SC_ITEMS], s = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) {
    v[j] = t0 + j < n ? in[t0 + j] : 0u;
    s += v[j];
  }
  uint32_t run = ib_block_exscan(s, nullptr) + (tileoffs ? tileoffs[blockIdx.x] : 0u);
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) {
    if (t0 + j < n) out[t0 + j] = run;
    run += v[j];
  }
}

static size_t ib_scan_scratch_words(uint64_t n) {   // tile sums of every recursion level
  size_t w = 0;
  while (n > (uint64_t)SC_TILE) { n = (n + SC_TILE - 1) / SC_TILE; w += (size_t)((n + 63) & ~63ull); }
  return w + 64;
}

static cudaError_t ib_scan(const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *scratch, cudaStream_t st,
                           int *nlaunch) {
  if (!n) return cudaSuccess;
  const uint64_t ntiles = (n + SC_TILE - 1) / SC_TILE;
  if (ntiles == 1) {
    ib_scan_apply_kernel<<<1, SC_THREADS, 0, st>>>(in, out, n, nullptr); ++*nlaunch;
    return cudaGetLastError();
  }
  ib_scan_reduce_kernel<<<(unsigned)ntiles, SC_THREADS, 0, st>>>(in, n, scratch); ++*nlaunch;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = ib_scan(scratch, scratch, ntiles, scratch + ((ntiles + 63) & ~63ull), st, nlaunch);
  if (e != cudaSuccess) return e;
  ib_scan_apply_kernel<<<(unsigned)ntiles, SC_THREADS, 0, st>>>(in, out, n, scratch); ++*nlaunch;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------
// stable LSD radix sort of (u64 key, u32 value), 8-bit digits, tiles of 4096 elements
// ------------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_ROUNDS = 16, RS_TILE = RS_THREADS * RS_ROUNDS;

// counts[digit * ntiles + tile] = elements of the tile with that digit
__global__ void __launch_bounds__(RS_THREADS)
ib_rs_hist_kernel(const unsigned long long *__restrict__ keys, const uint32_t n, const int shift,
                  uint32_t *__restrict__ counts, const uint32_t ntiles) {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t t0 = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const uint64_t i = t0 + (uint64_t)r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(hist + (uint32_t)((keys[i] >> shift) & 255ull), 1u);
  }
  __syncthreads();
  counts[(size_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

// base = exclusive scan of counts.  Warp w of a CTA owns elements [w * 512, (w + 1) * 512) of the tile and reads
// them 32 at a time; the rank of an element among the elements of its digit in the tile (in input order) is
//   elements of the digit in earlier warps + in earlier rounds of this warp + in lower lanes of this round.
__global__ void __launch_bounds__(RS_THREADS)
ib_rs_scatter_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals,
                     unsigned long long *__restrict__ okeys, uint32_t *__restrict__ ovals, const uint32_t n,
                     const int shift, const uint32_t *__restrict__ base, const uint32_t ntiles) {
  __shared__ uint32_t wh[RS_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) wh[w][threadIdx.x] = 0;
  __syncthreads();
  const uint64_t w0 = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)warp * (32 * RS_ROUNDS);
  const uint32_t lt = (1u << lane) - 1u;
  unsigned long long k[RS_ROUNDS];
  uint32_t v[RS_ROUNDS], rank[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const uint64_t i = w0 + (uint64_t)r * 32 + lane;
    const bool ok = i < n;
    k[r] = ok ? keys[i] : 0ull;
    v[r] = ok ? vals[i] : 0u;
    const uint32_t d = ok ? (uint32_t)((k[r] >> shift) & 255ull) : 256u;   // 256: past the end, ranked nowhere
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t old = 0;
    if (ok) old = wh[warp][d];
    rank[r] = old + __popc(peers & lt);
    __syncwarp();
    if (ok && (peers & lt) == 0u) wh[warp][d] = old + __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {   // thread = digit: where the digit's elements of every warp go
    uint32_t run = base[(size_t)threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t c = wh[w][threadIdx.x];
      wh[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    const uint64_t i = w0 + (uint64_t)r * 32 + lane;
    if (i < n) {
      const uint32_t dst = wh[warp][(uint32_t)((k[r] >> shift) & 255ull)] + rank[r];
      okeys[dst] = k[r];
      ovals[dst] = v[r];
    }
  }
}

// sorts by the bits [lo_begin, lo_end) and [hi_begin, hi_end) of the key (everything else is zero or irrelevant);
// the result is left in *keys / *vals (the pointer pairs are swapped after every pass)
static cudaError_t ib_radix_sort(unsigned long long **keys, unsigned long long **keys2, uint32_t **vals, uint32_t **vals2,
                                 uint32_t n, int lo_end, int hi_begin, int hi_end, uint32_t *counts, uint32_t *scratch,
                                 cudaStream_t st, int *nlaunch) {
  const uint32_t ntiles = (uint32_t)(((uint64_t)n + RS_TILE - 1) / RS_TILE);
  for (int part = 0; part < 2; ++part) {
    const int b0 = part ? hi_begin : 0, b1 = part ? hi_end : lo_end;
    for (int shift = b0; shift < b1; shift += 8) {
      ib_rs_hist_kernel<<<ntiles, RS_THREADS, 0, st>>>(*keys, n, shift, counts, ntiles); ++*nlaunch;
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
      e = ib_scan(counts, counts, (uint64_t)256 * ntiles, scratch, st, nlaunch);
      if (e != cudaSuccess) return e;
      ib_rs_scatter_kernel<<<ntiles, RS_THREADS, 0, st>>>(*keys, *vals, *keys2, *vals2, n, shift, counts, ntiles);
      ++*nlaunch;
      e = cudaGetLastError();
      if (e != cudaSuccess) return e;
      std::swap(*keys, *keys2);
      std::swap(*vals, *vals2);
    }
  }
  return cudaSuccess;
}

static double ib_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
#define IB_T(what) do { if (dbg) { cudaStreamSynchronize(st); const double t_ = ib_now(); \
    fprintf(stderr, "index_build: %-22s %8.2f ms\n", what, 1e3 * (t_ - t_dbg)); t_dbg = t_; } } while (0)
#define IB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { for (void *q_ : tofree) cudaFree(q_); return e_; } } while (0)

cudaError_t index_build(const uint32_t *d_packed, const IndexBuildSeq *h_seqs, int nseq, int k, int nskip, int typ,
                        int nbits_key, int nbits_lo, cudaStream_t st, IndexBuildOut *out, int *nlaunch) {
  const bool dbg = getenv("SMB_INDEX_DEBUG") != nullptr;
  double t_dbg = ib_now();
  std::vector<void *> tofree;
  std::vector<uint64_t> first((size_t)nseq + 1, 0);
  for (int i = 0; i < nseq; ++i) first[(size_t)i + 1] = first[(size_t)i] + h_seqs[i].n_k;
  const uint64_t ngrid = first[(size_t)nseq];
  const uint32_t nkeys = typ == 0 ? (1u << (2 * k)) : (1u << nbits_key);
  memset(out, 0, sizeof(*out));
  out->nkeys = nkeys;
  if (ngrid > 0xFFFFFFFFull) return cudaErrorInvalidValue;   // HASHPOS_MAX: serial numbers are 32 bit
  auto dalloc = [&](void **p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e == cudaSuccess) tofree.push_back(*p);
    return e;
  };
  IndexBuildSeq *d_seqs; uint64_t *d_first;
  unsigned long long *d_comp, *d_comp2; uint32_t *d_serial, *d_serial2, *d_valid, *d_dst;
  IB(dalloc((void **)&d_seqs, (size_t)nseq * sizeof(IndexBuildSeq)));
  IB(dalloc((void **)&d_first, ((size_t)nseq + 1) * 8));
  IB(dalloc((void **)&d_comp, ngrid * 8)); IB(dalloc((void **)&d_comp2, ngrid * 8));
  IB(dalloc((void **)&d_serial, ngrid * 4)); IB(dalloc((void **)&d_serial2, ngrid * 4));
  IB(dalloc((void **)&d_valid, ngrid * 4)); IB(dalloc((void **)&d_dst, (ngrid + 1) * 4));
  uint32_t *d_scratch;   // tile sums of the scans: the longest scan is over max(grid, 256 x sort tiles, keys + 1) values
  IB(dalloc((void **)&d_scratch, ib_scan_scratch_words(std::max<uint64_t>(ngrid + 256 * 2, (uint64_t)nkeys + 2)) * 4));
  IB(cudaMemcpyAsync(d_seqs, h_seqs, (size_t)nseq * sizeof(IndexBuildSeq), cudaMemcpyHostToDevice, st));
  IB(cudaMemcpyAsync(d_first, first.data(), ((size_t)nseq + 1) * 8, cudaMemcpyHostToDevice, st));
  IbParams p{d_packed, d_seqs, d_first, nseq, k, nskip, typ, nbits_lo,
             typ == 0 ? 1u : (1u << (nbits_key - nbits_lo)), ngrid};
  const unsigned gblocks = (unsigned)((ngrid + 255) / 256);
  uint32_t npos = 0;
  IB_T("allocations");
  if (ngrid) {
    ib_words_kernel<<<gblocks, 256, 0, st>>>(p, d_comp, d_serial, d_valid);
    IB(cudaGetLastError()); ++*nlaunch;
    IB_T("words kernel");
    // compaction of the grid positions over non-standard bases (keeps the scan order)
    IB(ib_scan(d_valid, d_dst, ngrid, d_scratch, st, nlaunch));
    uint32_t last_dst = 0, last_valid = 0;
    IB(cudaMemcpyAsync(&last_dst, d_dst + (ngrid - 1), 4, cudaMemcpyDeviceToHost, st));
    IB(cudaMemcpyAsync(&last_valid, d_valid + (ngrid - 1), 4, cudaMemcpyDeviceToHost, st));
    IB(cudaStreamSynchronize(st));
    npos = last_dst + last_valid;
    ib_compact_kernel<<<gblocks, 256, 0, st>>>(ngrid, d_comp, d_serial, d_valid, d_dst, d_comp2, d_serial2);
    IB(cudaGetLastError()); ++*nlaunch;
  }
  IB_T("compaction");
  // stable sort by the composite key (upper 32 bits: key, lower: upper word bits)
  if (npos) {
    // the composite occupies bits [0, 2k - nbits_lo) (upper word bits, collision type) and [32, 32 + key bits)
    const int lo_end = typ == 0 ? 0 : std::min(32, std::max(0, 2 * k - nbits_lo));
    const int hi_end = 32 + (typ == 0 ? 2 * k : nbits_key);
    uint32_t *d_counts;
    IB(dalloc((void **)&d_counts, (size_t)256 * (((size_t)npos + RS_TILE - 1) / RS_TILE) * 4));
    std::swap(d_comp, d_comp2); std::swap(d_serial, d_serial2);   // the compacted arrays are the sort's input
    IB(ib_radix_sort(&d_comp, &d_comp2, &d_serial, &d_serial2, npos, lo_end, 32, hi_end, d_counts, d_scratch, st,
                     nlaunch));
  }
  IB_T("radix sort");
  // outputs: one allocation {idx[nkeys+1], pos[npos], wordidx[nwords+1], posidx[nwords+1]}
  uint32_t *d_keycount; IB(dalloc((void **)&d_keycount, ((size_t)nkeys + 2) * 4));
  IB(cudaMemsetAsync(d_keycount, 0, ((size_t)nkeys + 2) * 4, st));
  uint32_t *d_flag = d_valid, *d_wordno = d_dst;   // reuse
  uint32_t nwords = 0;
  const unsigned pblocks = (unsigned)(((uint64_t)npos + 255) / 256);
  if (npos) {
    ib_flag_kernel<<<pblocks, 256, 0, st>>>(npos, d_comp, d_flag, d_keycount + 1, typ);
    IB(cudaGetLastError()); ++*nlaunch;
    if (typ != 0) {
      IB(ib_scan(d_flag, d_wordno, npos, d_scratch, st, nlaunch));
      uint32_t lw = 0, lf = 0;
      IB(cudaMemcpyAsync(&lw, d_wordno + (npos - 1), 4, cudaMemcpyDeviceToHost, st));
      IB(cudaMemcpyAsync(&lf, d_flag + (npos - 1), 4, cudaMemcpyDeviceToHost, st));
      IB(cudaStreamSynchronize(st));
      nwords = lw + lf;
    }
  }
  const size_t n_idx = (size_t)nkeys + 1, n_w = typ ? (size_t)nwords + 1 : 0;
  uint32_t *d_out;
  cudaError_t e = cudaMalloc((void **)&d_out, (n_idx + npos + 2 * n_w + 16) * 4);
  if (e != cudaSuccess) { for (void *q : tofree) cudaFree(q); return e; }
  tofree.push_back(d_out);
  uint32_t *o_idx = d_out, *o_pos = o_idx + n_idx, *o_widx = o_pos + npos, *o_pidx = o_widx + n_w;
  {
    // idx[i] = words (positions) of the keys below i: exclusive prefix of the key counts
    IB(ib_scan(d_keycount + 1, o_idx, n_idx, d_scratch, st, nlaunch));
  }
  if (npos) IB(cudaMemcpyAsync(o_pos, d_serial, (size_t)npos * 4, cudaMemcpyDeviceToDevice, st));
  if (typ != 0) {
    IB(cudaMemsetAsync(o_widx, 0, 2 * n_w * 4, st));
    if (npos) {
      ib_words_out_kernel<<<pblocks, 256, 0, st>>>(npos, d_comp, d_flag, d_wordno, o_widx, o_pidx);
      IB(cudaGetLastError()); ++*nlaunch;
    }
    IB(cudaMemcpyAsync(o_pidx + nwords, &npos, 4, cudaMemcpyHostToDevice, st));   // hashidx.c:984
  }
  IB(cudaStreamSynchronize(st));
  IB_T("words / keys / copies");
  for (void *q : tofree)
    if (q != (void *)d_out) cudaFree(q);   // d_out is handed to the caller
  IB_T("free");
  out->npos = npos; out->nwords = nwords;
  out->block = d_out; out->idx = o_idx; out->pos = o_pos;
  out->wordidx = typ ? o_widx : nullptr; out->posidx = typ ? o_pidx : nullptr;
  return cudaSuccess;
}

}  // namespace smb
