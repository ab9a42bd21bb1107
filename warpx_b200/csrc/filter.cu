// Bilinear (binomial) smoothing of one field component: dst = stencil (*) src.
//
// Replaces BilinearFilter::ComputeStencils + Filter::ApplyStencil/DoFilter as called by
// WarpX::ApplyFilterJ (reference: Source/Filter/BilinearFilter.cpp:26-88, Source/Filter/Filter.cpp:
// 37-133, Source/Parallelization/WarpXComm.cpp:1357-1374).  The reference evaluates the separable
// (1,2,1)/4 ^ npass kernel as a dense sum of eight mirrored reads per stencil entry over the grown
// box, the source zero-padded outside its allocation.  Here one thread per destination point walks
// the (2n0+1)(2n1+1)(2n2+1) distinct neighbours once, with the mirrored pairs folded into the
// weights (same values, different summation order: parity is a rounding-level tolerance).
//
// Memory bound: 16 B/point algorithmic (one read, one write, fp64); the neighbour reads of a warp
// are 32 consecutive doubles per row and are served by L1/L2.
#include "pic_common.cuh"

namespace pic {

constexpr int FILTER_MAX_PASS = 8;
struct FilterWeights { double w[3][2 * FILTER_MAX_PASS + 1]; int n[3]; };   // w[d][a + n[d]], a in [-n, n]

constexpr int FIL_BX = 64, FIL_BY = 4;

// N0,N1,N2 >= 0: compile-time half widths; -1: run-time (W.n)
template <int N0, int N1, int N2>
__global__ void __launch_bounds__(FIL_BX * FIL_BY)
filter_kernel(FabView S, FabView D, FilterWeights W) {
    const int li = blockIdx.x * FIL_BX + threadIdx.x;
    const int lj = blockIdx.y * FIL_BY + threadIdx.y;
    const int lk = blockIdx.z;
    if (li >= D.n0 || lj >= D.n1) return;
    const int n0 = N0 >= 0 ? N0 : W.n[0], n1 = N1 >= 0 ? N1 : W.n[1], n2 = N2 >= 0 ? N2 : W.n[2];
    // source-local coordinates of the destination point (the two fabs may differ in extent)
    const int si = li + D.lo0 - S.lo0, sj = lj + D.lo1 - S.lo1, sk = lk + D.lo2 - S.lo2;
    const bool interior = si >= n0 && si + n0 < S.n0 && sj >= n1 && sj + n1 < S.n1 && sk >= n2 && sk + n2 < S.n2;
    const double* __restrict__ base = S.p + (long)si + (long)sj * S.sj + (long)sk * S.sk;
    double acc = 0.0;
    if (interior) {
#pragma unroll
        for (int c = -n2; c <= n2; ++c) {
            double pl = 0.0;
#pragma unroll
            for (int b = -n1; b <= n1; ++b) {
                const double* row = base + (long)b * S.sj + (long)c * S.sk;
                double r = 0.0;
#pragma unroll
                for (int a = -n0; a <= n0; ++a) r += W.w[0][a + n0] * __ldg(row + a);
                pl += W.w[1][b + n1] * r;
            }
            acc += W.w[2][c + n2] * pl;
        }
    } else {
        for (int c = -n2; c <= n2; ++c) {
            if ((unsigned)(sk + c) >= (unsigned)S.n2) continue;             // zero padding, Filter.cpp:103-107
            double pl = 0.0;
            for (int b = -n1; b <= n1; ++b) {
                if ((unsigned)(sj + b) >= (unsigned)S.n1) continue;
                const double* row = base + (long)b * S.sj + (long)c * S.sk;
                double r = 0.0;
                for (int a = -n0; a <= n0; ++a)
                    if ((unsigned)(si + a) < (unsigned)S.n0) r += W.w[0][a + n0] * __ldg(row + a);
                pl += W.w[1][b + n1] * r;
            }
            acc += W.w[2][c + n2] * pl;
        }
    }
    D.p[(long)li + (long)lj * D.sj + (long)lk * D.sk] = acc;
}

// BilinearFilter.cpp:26-62; returns the FULL symmetric weights (element 0 un-halved again)
static void binomial_weights(int npass, double* w /* 2*npass+1 */) {
    double old_s[FILTER_MAX_PASS + 2] = {0}, new_s[FILTER_MAX_PASS + 2] = {0};
    old_s[0] = 1.0;
    int jmax = 1;
    for (int ipass = 1; ipass <= npass; ++ipass) {
        new_s[0] = 0.5 * old_s[0];
        if (1 < jmax) new_s[0] += 0.5 * old_s[1];
        for (int j = 1; j <= jmax; ++j) {
            double loc = 0.5 * old_s[j];
            loc += 0.25 * old_s[j - 1];
            if (j < jmax) loc += 0.25 * old_s[j + 1];
            new_s[j] = loc;
        }
        for (int j = 0; j <= npass; ++j) old_s[j] = new_s[j];
        jmax += 1;
    }
    for (int a = -npass; a <= npass; ++a) w[a + npass] = old_s[a < 0 ? -a : a];
}

}  // namespace pic

using namespace pic;

extern "C" int pic_apply_filter(const pic_fab* src, const pic_fab* dst, const int npass[3], void* stream) {
    if (!src || !dst || !npass || !src->p || !dst->p) return fail("pic_apply_filter: null argument");
    if (src->p == dst->p) return fail("pic_apply_filter: src and dst must be different arrays");
    FilterWeights W;
    for (int d = 0; d < 3; ++d) {
        if (npass[d] < 0 || npass[d] > FILTER_MAX_PASS) return fail("pic_apply_filter: npass out of range");
        if (src->stag[d] != dst->stag[d]) return fail("pic_apply_filter: staggering mismatch");
        W.n[d] = npass[d];
        binomial_weights(npass[d], W.w[d]);
    }
    const FabView S = make_view(*src), D = make_view(*dst);
    dim3 block(FIL_BX, FIL_BY), grid((D.n0 + FIL_BX - 1) / FIL_BX, (D.n1 + FIL_BY - 1) / FIL_BY, D.n2);
    cudaStream_t s = (cudaStream_t)stream;
    if (npass[0] == 1 && npass[1] == 1 && npass[2] == 1) filter_kernel<1, 1, 1><<<grid, block, 0, s>>>(S, D, W);
    else filter_kernel<-1, -1, -1><<<grid, block, 0, s>>>(S, D, W);
    count_launch();
    return check_launch("pic_apply_filter") ? 0 : 1;
}
