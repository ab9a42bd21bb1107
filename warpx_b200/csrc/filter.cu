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

// ---- npass = (1, 1, 1), the default of warpx.use_filter: one streaming pass, up to three components per launch ----
// A thread owns one (i, j) column of a chunk of planes and walks it along k.  Per source plane it forms the
// x-filtered value of its own row (three reads of one cache line), the rows of a CTA trade those through shared
// memory for the y sum, and the last three plane results stay in registers for the z sum: 3.75 cached reads and
// one write per point instead of 27 reads.  Same nesting and the same multiply-add forms as filter_kernel
// (sum over c of w2 * (sum over b of w1 * (sum over a of w0 * src))), so both give the same bits.
constexpr int FM_BX = 32, FM_BY = 8, FM_CHUNK = 32, FM_MAX = 3;
struct FilterMulti { FabView S[FM_MAX], D[FM_MAX]; int n; int nchunk; };

__global__ void __launch_bounds__(FM_BX * FM_BY)
filter_march_kernel(FilterMulti M, FilterWeights W) {
    PIC_STATIC_SMEM(double, rows, 2 * (FM_BY + 2) * FM_BX);
    const int f = blockIdx.z / M.nchunk, chunk = blockIdx.z % M.nchunk;
    const FabView& S = M.S[f];
    const FabView& D = M.D[f];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int li = blockIdx.x * FM_BX + tx, lj = blockIdx.y * FM_BY + ty;
    const int k0 = chunk * FM_CHUNK, k1 = min(D.n2, k0 + FM_CHUNK);
    if (k0 >= D.n2) return;                                        // whole CTA
    // source-local coordinates of the destination column (the two fabs may differ in extent)
    const int si = li + D.lo0 - S.lo0, sj = lj + D.lo1 - S.lo1, dk = D.lo2 - S.lo2;
    const double w00 = W.w[0][0], w01 = W.w[0][1], w02 = W.w[0][2];
    const double w10 = W.w[1][0], w11 = W.w[1][1], w12 = W.w[1][2];
    const double w20 = W.w[2][0], w21 = W.w[2][1], w22 = W.w[2][2];
    // x-filtered value of row (sj + b) in source plane sk; zero outside the source's allocation (Filter.cpp:103-107)
    auto xrow = [&](int b, int sk) -> double {
        const int j = sj + b;
        if ((unsigned)j >= (unsigned)S.n1 || (unsigned)sk >= (unsigned)S.n2) return 0.0;
        const double* __restrict__ row = S.p + (long)j * S.sj + (long)sk * S.sk;
        double r = 0.0;
        if ((unsigned)(si - 1) < (unsigned)S.n0) r += w00 * __ldg(row + si - 1);
        if ((unsigned)si < (unsigned)S.n0) r += w01 * __ldg(row + si);
        if ((unsigned)(si + 1) < (unsigned)S.n0) r += w02 * __ldg(row + si + 1);
        return r;
    };
    // xy-filtered value of the thread's point in source plane sk (all threads of the CTA call this together)
    int phase = 0;
    auto plane = [&](int sk) -> double {
        double* buf = rows + phase * (FM_BY + 2) * FM_BX;
        phase ^= 1;                                                // two buffers: one barrier per plane
        buf[(ty + 1) * FM_BX + tx] = xrow(0, sk);
        if (ty == 0) buf[tx] = xrow(-1, sk);
        if (ty == FM_BY - 1) buf[(FM_BY + 1) * FM_BX + tx] = xrow(1, sk);
        __syncthreads();
        double pl = 0.0;
        pl += w10 * buf[ty * FM_BX + tx];
        pl += w11 * buf[(ty + 1) * FM_BX + tx];
        pl += w12 * buf[(ty + 2) * FM_BX + tx];
        return pl;
    };
    const bool mine = li < D.n0 && lj < D.n1;
    double pm = plane(k0 - 1 + dk), p0 = plane(k0 + dk);
    for (int k = k0; k < k1; ++k) {
        const double pp = plane(k + 1 + dk);
        double acc = 0.0;
        acc += w20 * pm;
        acc += w21 * p0;
        acc += w22 * pp;
        if (mine) D.p[(long)li + (long)lj * D.sj + (long)k * D.sk] = acc;
        pm = p0; p0 = pp;
    }
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

extern "C" int pic_apply_filter(const pic_fab* src, const pic_fab* dst, const int npass[3], void* stream);

// 0: marching kernel for npass = (1,1,1) (default); 1: the direct 27-point kernel everywhere (PIC_FILTER_DIRECT=1)
static const bool g_filter_direct = [] { const char* v = getenv("PIC_FILTER_DIRECT"); return v && atoi(v) != 0; }();

extern "C" int pic_apply_filter_multi(const pic_fab* src, const pic_fab* dst, int nfab, const int npass[3], void* stream) {
    if (!src || !dst || !npass || nfab < 1) return fail("pic_apply_filter_multi: null argument");
    const bool march = npass[0] == 1 && npass[1] == 1 && npass[2] == 1 && nfab <= FM_MAX && !g_filter_direct;
    if (!march) {
        for (int f = 0; f < nfab; ++f)
            if (int rc = pic_apply_filter(&src[f], &dst[f], npass, stream)) return rc;
        return 0;
    }
    FilterWeights W;
    for (int d = 0; d < 3; ++d) { W.n[d] = 1; binomial_weights(1, W.w[d]); }
    FilterMulti M;
    M.n = nfab;
    int n0 = 0, n1 = 0, n2 = 0;
    for (int f = 0; f < nfab; ++f) {
        if (!src[f].p || !dst[f].p) return fail("pic_apply_filter_multi: null array");
        if (src[f].p == dst[f].p) return fail("pic_apply_filter_multi: src and dst must be different arrays");
        for (int d = 0; d < 3; ++d)
            if (src[f].stag[d] != dst[f].stag[d]) return fail("pic_apply_filter_multi: staggering mismatch");
        M.S[f] = make_view(src[f]); M.D[f] = make_view(dst[f]);
        n0 = max(n0, M.D[f].n0); n1 = max(n1, M.D[f].n1); n2 = max(n2, M.D[f].n2);
    }
    for (int f = nfab; f < FM_MAX; ++f) { M.S[f] = M.S[0]; M.D[f] = M.D[0]; }
    M.nchunk = (n2 + FM_CHUNK - 1) / FM_CHUNK;
    dim3 block(FM_BX, FM_BY), grid((n0 + FM_BX - 1) / FM_BX, (n1 + FM_BY - 1) / FM_BY, nfab * M.nchunk);
    filter_march_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(M, W);
    count_launch();
    return check_launch("pic_apply_filter_multi") ? 0 : 1;
}

extern "C" int pic_apply_filter(const pic_fab* src, const pic_fab* dst, const int npass[3], void* stream) {
    if (!src || !dst || !npass || !src->p || !dst->p) return fail("pic_apply_filter: null argument");
    if (src->p == dst->p) return fail("pic_apply_filter: src and dst must be different arrays");
    if (npass[0] == 1 && npass[1] == 1 && npass[2] == 1 && !g_filter_direct) return pic_apply_filter_multi(src, dst, 1, npass, stream);
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
