// band_cell.cuh - the cell of the restricted banded recurrence shared by the warp-per-task K3
// kernels (band_warp.cu, band_wide.cu).
#pragma once
#define DIFFB(count, typ) ((uint8_t)((count) + ((typ) << 6)))

// One cell of the restricted recurrence (alignment.c:885-982), branch free.  `ok` = the cell
// exists (inside the band, the read segment and the row range); a cell that does not exist
// passes H = E = F = 0 on, exactly like the reference's zeroed row buffers.
#define BAND_CELL(ok, diag, ein, fin, s, Hout, Eout, Fout, best, bestr, r, dcode)                 \
  do {                                                                                             \
    const int h_ = (diag) + (s);                                                                   \
    const int m_ = __vimax3_s32((ein), (fin), 0);                                                  \
    const bool dia_ = h_ > m_;                                                                     \
    const int hn_ = max(h_, m_);                                                                   \
    int e_ = (ein) - (((ein) > 0) ? ge : 0);                                                       \
    int f_ = (fin) - (((fin) > 0) ? ge : 0);                                                       \
    const bool open_ = dia_ && h_ > gi;                                                            \
    const int t_ = open_ ? h_ - gi : (int)0x80000000;                                              \
    e_ = max(e_, t_);                                                                              \
    f_ = max(f_, t_);                                                                              \
    if ((ok) && open_ && h_ > (best)) { (best) = h_; (bestr) = (r); }                              \
    /* DIA 3, COL 1 (E >= F: whenever the maximum is positive max(E,0) >= max(F,0) <=> E >= F), */ \
    /* ROW 2, stop 0                                                                          */ \
    const uint32_t d_ = dia_ ? 3u : (m_ == 0 ? 0u : ((ein) >= (fin) ? 1u : 2u));                   \
    (Hout) = (ok) ? hn_ : 0;                                                                       \
    (Eout) = (ok) ? e_ : 0;                                                                        \
    (Fout) = (ok) ? f_ : 0;                                                                        \
    (dcode) = (ok) ? d_ : 0u;                                                                      \
  } while (0)

