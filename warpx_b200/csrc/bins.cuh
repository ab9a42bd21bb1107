// Supercell-major cell bins (see pic_bins in include/pic_b200.h).
#ifndef PIC_BINS_CUH_
#define PIC_BINS_CUH_
#include "pic_common.cuh"

namespace pic {

struct BinsView {
    const int* __restrict__ cell_start;
    int box_lo[3];
    int n[3];      // cells of the box
    int tile[3];
    int nt[3];     // supercells per direction (padded)
    int np_limit;  // particles beyond this index are not covered by the bins
};

inline BinsView make_bins(const pic_bins& b) {
    BinsView v;
    v.cell_start = b.cell_start;
    for (int d = 0; d < 3; ++d) {
        v.box_lo[d] = b.box_lo[d];
        v.n[d] = b.box_hi[d] - b.box_lo[d] + 1;
        v.tile[d] = b.tile[d];
        v.nt[d] = (v.n[d] + b.tile[d] - 1) / b.tile[d];
    }
    v.np_limit = (int)b.np_binned;
    return v;
}
inline long bins_count(const BinsView& v) {
    return (long)v.nt[0] * v.nt[1] * v.nt[2] * v.tile[0] * v.tile[1] * v.tile[2];
}

// bin id of the box-local cell (ci, cj, ck)
__host__ __device__ __forceinline__ long bin_of_cell(const BinsView& b, int ci, int cj, int ck) {
    const int ti = ci / b.tile[0], tj = cj / b.tile[1], tk = ck / b.tile[2];
    const int li = ci - ti * b.tile[0], lj = cj - tj * b.tile[1], lk = ck - tk * b.tile[2];
    const long t = ti + (long)b.nt[0] * (tj + (long)b.nt[1] * tk);
    return t * ((long)b.tile[0] * b.tile[1] * b.tile[2]) + li + b.tile[0] * (lj + b.tile[1] * lk);   // x fastest
}
__host__ __device__ __forceinline__ void tile_coords(const BinsView& b, int t, int tc[3]) {
    tc[0] = t % b.nt[0];
    tc[1] = (t / b.nt[0]) % b.nt[1];
    tc[2] = t / (b.nt[0] * b.nt[1]);
}

}  // namespace pic
#endif
