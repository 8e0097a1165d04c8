// The constants of the K2 score lookup (smalt_b200/csrc/sw2_lut.cuh, used by sw_score2_kernel and sw_long2_kernel)
// compiled for the HOST for tests/test_k2_rowtable_model.py.  Test infrastructure.
#include <cuda_runtime.h>
#include "sw2_lut.cuh"

extern "C" {
unsigned sw2h_lut_index(unsigned a, unsigned b) { return smb::sw2_lut_index(a, b); }
unsigned sw2h_tab_word(unsigned x, int match, int mismatch) { return smb::sw2_tab_word(x, match, mismatch); }
unsigned sw2h_wsel_word(unsigned a, unsigned b) { return smb::sw2_wsel_word(a, b); }
unsigned sw2h_qsel_tab(unsigned qa, unsigned qb) { return smb::sw2_qsel_tab(qa, qb); }
unsigned sw2h_qsel_masked(unsigned qa, unsigned qb) { return smb::sw2_qsel_masked(qa, qb); }
int sw2h_lut_n(void) { return smb::SW2_LUT_N; }
}
