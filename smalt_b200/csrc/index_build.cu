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
//   2. radix sort      cub::DeviceRadixSort::SortPairs on (composite, serial number) - LSD radix
//                      sort is stable; library code, this is not the mapping hot path
//   3. words / keys    word boundaries -> wordidx, posidx; histogram of keys -> idx (scans)
// The grid bookkeeping between sequences (offset of the first k-mer of a sequence, serial numbers
// that count skipped positions; hashidx.c:498-529) is done per sequence on the host and passed in.
#include "common.cuh"
#include <cub/cub.cuh>
#include <vector>
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
    size_t tb = 0;
    IB(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_valid, d_dst, (int)ngrid, st));
    void *d_tmp; IB(dalloc(&d_tmp, tb));
    IB(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_valid, d_dst, (int)ngrid, st)); ++*nlaunch;
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
    size_t tb = 0;
    const int end_bit = 32 + (typ == 0 ? 2 * k : nbits_key);
    IB(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_comp2, d_comp, d_serial2, d_serial, (int)npos, 0, end_bit, st));
    void *d_tmp; IB(dalloc(&d_tmp, tb));
    IB(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_comp2, d_comp, d_serial2, d_serial, (int)npos, 0, end_bit, st));
    ++*nlaunch;
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
      size_t tb = 0;
      IB(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag, d_wordno, (int)npos, st));
      void *d_tmp; IB(dalloc(&d_tmp, tb));
      IB(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_flag, d_wordno, (int)npos, st)); ++*nlaunch;
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
    // idx = inclusive prefix of the key counts, shifted by one (idx[0] = 0)
    size_t tb = 0;
    IB(cub::DeviceScan::InclusiveSum(nullptr, tb, d_keycount, o_idx, (int)n_idx, st));
    void *d_tmp; IB(dalloc(&d_tmp, tb));
    IB(cub::DeviceScan::InclusiveSum(d_tmp, tb, d_keycount, o_idx, (int)n_idx, st)); ++*nlaunch;
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
