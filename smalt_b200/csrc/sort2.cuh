// sort2.cuh - the reference's two-array quicksort on the device (shared by seed.cu and block.cu)
#pragma once
#include <stdint.h>

namespace smb {

// sort2UINTarraysByQuickSort (sort.c:233-330): median-of-three quicksort, insertion sort
// below 7 elements, explicit stack, smaller partition first.  Unstable; the exchange
// sequence is reproduced so that ties end up in the reference's order.
__device__ inline int sort2(int n, uint32_t *key, uint32_t *val) {
  int lo = 0, hi = n - 1, sp = 0, i, j;
  int stack[62];
#define XC(a, b) do { uint32_t t_ = (a); (a) = (b); (b) = t_; } while (0)
  for (;;) {
    if (hi - lo < 7) {
      for (j = lo + 1; j <= hi; ++j) {
        const uint32_t k = key[j], v = val[j];
        for (i = j - 1; i >= lo && key[i] > k; --i) { key[i + 1] = key[i]; val[i + 1] = val[i]; }
        key[i + 1] = k; val[i + 1] = v;
      }
      if (!sp) return 0;
      hi = stack[sp--];
      lo = stack[sp--];
    } else {
      const int mid = (lo + hi) >> 1;
      XC(key[mid], key[lo + 1]); XC(val[mid], val[lo + 1]);
      if (key[lo] > key[hi]) { XC(key[lo], key[hi]); XC(val[lo], val[hi]); }
      if (key[lo + 1] > key[hi]) { XC(key[lo + 1], key[hi]); XC(val[lo + 1], val[hi]); }
      if (key[lo] > key[lo + 1]) { XC(key[lo], key[lo + 1]); XC(val[lo], val[lo + 1]); }
      i = lo + 1; j = hi;
      const uint32_t pk = key[lo + 1], pv = val[lo + 1];
      for (;;) {
        do ++i; while (key[i] < pk);
        do --j; while (key[j] > pk);
        if (j < i) break;
        XC(key[i], key[j]); XC(val[i], val[j]);
      }
      key[lo + 1] = key[j]; val[lo + 1] = val[j];
      key[j] = pk; val[j] = pv;
      sp += 2;
      if (sp > 60) return 34;  // ERRCODE_SORTSTACK
      if (hi - i + 1 >= j - lo) { stack[sp] = hi; stack[sp - 1] = i; hi = j - 1; }
      else { stack[sp] = j - 1; stack[sp - 1] = lo; lo = i; }
    }
  }
#undef XC
}

}  // namespace smb
