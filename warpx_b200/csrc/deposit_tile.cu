// Shared-memory Esirkepov deposition for cell-sorted particles (the timed path).
//
// One CTA per supercell (pic_bins.tile, tuned for 8x8x8 cells).  The CTA owns a shared-memory
// block of J covering the supercell plus the deposition halo, 3 components, fp64.
//
// The reference (CurrentDeposition.H:792-824) scatters (N+2)(N+3)^2 x 3 values per particle with
// one fp64 global atomic each.  Here the scatter is reorganised around the structure of the
// Esirkepov stencil: for one particle,
//     Jx[i][j][k] += cdsx[i] * Wx[j][k],   cdsx[i] = sum_{i'<=i} wq/(dt dy dz) (Sx_old[i'] - Sx_new[i']),
//     Wx[j][k]    = Sy_new[j] Az[k] + Sy_old[j] Bz[k],  Az = Sz_new/3 + Sz_old/6, Bz = Sz_old/3 + Sz_new/6
// (and cyclically for Jy, Jz) -- an outer product of a (N+2)-vector and a (N+3)^2 matrix.
// A warp works in two alternating phases on chunks of 32 consecutive (cell-sorted) particles:
//   phase 1 (lane = particle): positions, shape factors, prefix sums -> per-particle record in smem;
//   phase 2 (lane = stencil line): lanes own stencil lines of all three components and keep their
//       partial sums in REGISTERS while the warp walks the particles of the chunk.  All particles
//       of a run with the same stencil anchor (same new cell) accumulate into the same registers
//       -- this is the warp-segmented reduction: the segment is the run, the reduction happens in
//       registers without any atomic or shuffle.  Two lane layouts coexist:
//         * "quiet" particles (old and new position in the same cell in all three directions, the
//           overwhelming majority in a thermal plasma): the stencil shrinks to (N+1)^2 lines x N
//           prefix entries, so 32/(N+1)^2 particles are processed per warp pass;
//         * general particles: the full (N+3)^2 x (N+2) stencil, one particle per pass.
//   At the end of a run the registers are added to the shared J block (smem CAS-add; the lanes of
//   a warp hit distinct addresses, the per-component pitches of the block make them distinct banks).
//   Cells are visited along x, so the quiet layout SLIDES instead of flushing: lines are owned by
//   their absolute x (ring mapping), only the plane that leaves the window is retired
//   (40 instead of 144 shared-memory updates per cell).  At the end of the CTA the block is added
//   to global J with one pass of coalesced fp64 reductions (the halo overlaps neighbouring supercells).
// Particles whose stencil does not fit the block (drifted since the last sort) take the
// per-particle global-atomic path, so any particle order is CORRECT; sorted order is FAST.
#include "pic_common.cuh"
#include "deposit_common.cuh"
#include "bins.cuh"

namespace pic {

constexpr int DT_MARGIN_LO = 3;   // block starts 3 points below the supercell
constexpr int DT_EXTRA = 7;       // block extent = tile + 7  (cell_new in [t0-1, t0+T], window -2..+3)
constexpr int DT_CH = 32;         // particles per chunk
constexpr int DT_CHP = DT_CH + 1; // record pitch (odd: conflict-free column access)

template <int N> struct TileCfg {
    static constexpr int S = N + 3;                 // window slots per direction
    static constexpr int PN = N + 2;                // prefix entries actually deposited
    static constexpr int NB = (S * S + 31) / 32;    // b values per lane (general layout)
    static constexpr int BH = (S + NB - 1) / NB;    // lane rows
    static constexpr int NLANES = S * BH;           // active lanes, general layout
    static constexpr int QS = N + 1;                // quiet layout: slots 1..N+1
    static constexpr int QL = QS * QS;              // lines per particle
    static constexpr int NG = 32 / QL;              // particles per pass
    static constexpr int QP = N;                    // live prefix entries (slots 1..N)
    // record fields: sn/so for x,y (4*S) ; A,B for y,z (4*S) ; cds x,y,z (3*PN)
    static constexpr int F_SNX = 0, F_SOX = S, F_SNY = 2 * S, F_SOY = 3 * S;
    static constexpr int F_AY = 4 * S, F_BY = 5 * S, F_AZ = 6 * S, F_BZ = 7 * S;
    static constexpr int F_CDS = 8 * S;             // + comp*PN + i
    static constexpr int NF = 8 * S + 3 * PN;
};

// shifted (old-position) weights without dynamic indexing: slot s holds w[s-1-sh], sh in {-1,0,1}
template <int N>
__device__ __forceinline__ void place_old(double* so /*N+3*/, const double* w /*N+1*/, int sh) {
#pragma unroll
    for (int s = 0; s < N + 3; ++s) {
        const double wm = (s <= N) ? w[s] : 0.0;                    // sh = -1 -> index s
        const double w0 = (s >= 1 && s - 1 <= N) ? w[s - 1] : 0.0;  // sh =  0 -> index s-1
        const double wp = (s >= 2) ? w[s - 2] : 0.0;                // sh = +1 -> index s-2
        so[s] = (sh < 0) ? wm : ((sh == 0) ? w0 : wp);
    }
}

// one direction of the Esirkepov weights; returns i_new (leftmost index of the new stencil)
template <int N>
__device__ __forceinline__ int dir_weights(double x_new, double x_old, double* sn /*N+3*/,
                                           double* so /*N+3*/, int& sh) {
    double wn[N + 1], wo[N + 1];
    const int i_new = shape_factor<N>(wn, x_new);
    sn[0] = 0.0; sn[N + 2] = 0.0;
#pragma unroll
    for (int s = 0; s <= N; ++s) sn[s + 1] = wn[s];
    // old position: same formulas evaluated at x_old; its leftmost index relative to i_new decides
    // the slot shift (ShapeFactors.H:93-156; floor for N = 1, truncation otherwise)
    int i_old;
    if constexpr (N == 1) {
        const int i = (int)floor(x_old);
        const double d = x_old - (double)i;
        wo[0] = 1.0 - d; wo[1] = d;
        i_old = i;
    } else {
        i_old = shape_factor<N>(wo, x_old);
    }
    sh = i_old - i_new;
    place_old<N>(so, wo, sh);
    return i_new;
}

// Drifted particle: per-particle global reductions with the reference's loop nests
// (CurrentDeposition.H:792-824).  Kept out of line so that its dynamically indexed arrays do
// not push the hot path's weights into local memory.
template <int N>
__device__ __noinline__ void deposit_one_global(double xp, double yp, double zp, double wp, double uxp,
                                                double uyp, double uzp, const FabView& Jx,
                                                const FabView& Jy, const FabView& Jz, const DepositGeom& dg) {
    EsirkepovWeights<N> ew;
    ew.compute(xp, yp, zp, wp, uxp, uyp, uzp, dg);
    const int bi = dg.lo[0] + ew.i_new - 1, bj = dg.lo[1] + ew.j_new - 1, bk = dg.lo[2] + ew.k_new - 1;
    constexpr double one_third = 1.0 / 3.0, one_sixth = 1.0 / 6.0;
    for (int k = ew.dkl; k <= N + 2 - ew.dku; ++k)
        for (int j = ew.djl; j <= N + 2 - ew.dju; ++j) {
            double sd = 0.0;
            const double w2 = one_third * (ew.sy_new[j] * ew.sz_new[k] + ew.sy_old[j] * ew.sz_old[k])
                            + one_sixth * (ew.sy_new[j] * ew.sz_old[k] + ew.sy_old[j] * ew.sz_new[k]);
            for (int i = ew.dil; i <= N + 1 - ew.diu; ++i) {
                sd += ew.wqx * (ew.sx_old[i] - ew.sx_new[i]) * w2;
                atomicAdd(&Jx(bi + i, bj + j, bk + k), sd);
            }
        }
    for (int k = ew.dkl; k <= N + 2 - ew.dku; ++k)
        for (int i = ew.dil; i <= N + 2 - ew.diu; ++i) {
            double sd = 0.0;
            const double w2 = one_third * (ew.sx_new[i] * ew.sz_new[k] + ew.sx_old[i] * ew.sz_old[k])
                            + one_sixth * (ew.sx_new[i] * ew.sz_old[k] + ew.sx_old[i] * ew.sz_new[k]);
            for (int j = ew.djl; j <= N + 1 - ew.dju; ++j) {
                sd += ew.wqy * (ew.sy_old[j] - ew.sy_new[j]) * w2;
                atomicAdd(&Jy(bi + i, bj + j, bk + k), sd);
            }
        }
    for (int j = ew.djl; j <= N + 2 - ew.dju; ++j)
        for (int i = ew.dil; i <= N + 2 - ew.diu; ++i) {
            double sd = 0.0;
            const double w2 = one_third * (ew.sx_new[i] * ew.sy_new[j] + ew.sx_old[i] * ew.sy_old[j])
                            + one_sixth * (ew.sx_new[i] * ew.sy_old[j] + ew.sx_old[i] * ew.sy_new[j]);
            for (int k = ew.dkl; k <= N + 1 - ew.dku; ++k) {
                sd += ew.wqz * (ew.sz_old[k] - ew.sz_new[k]) * w2;
                atomicAdd(&Jz(bi + i, bj + j, bk + k), sd);
            }
        }
}

// Shared J block: component c is stored with its own row / plane pitch so that the (u, v) lanes
// of the quiet layout fall into distinct 8-byte banks when they update one stencil plane.
struct BlockGeom {
    int bd[3];        // extent in points
    int py[3], pz[3]; // pitches (doubles) of component c
    int off[3];       // offset (doubles) of component c
    int total;        // doubles
};

inline int round_up_residue(int v, int mod, int res) {   // smallest w >= v with w % mod == res
    int w = v - (v % mod) + res;
    return w >= v ? w : w + mod;
}

inline BlockGeom make_block_geom(const BinsView& bv) {
    BlockGeom g;
    for (int d = 0; d < 3; ++d) g.bd[d] = bv.tile[d] + DT_EXTRA;
    // Jx lanes vary along (y, z): py = 1, pz = 4 (mod 16); Jy along (x, z): pz = 4; Jz along (x, y): py = 4
    g.py[0] = round_up_residue(g.bd[0], 16, 1); g.pz[0] = round_up_residue(g.py[0] * g.bd[1], 16, 4);
    g.py[1] = g.bd[0];                          g.pz[1] = round_up_residue(g.py[1] * g.bd[1], 16, 4);
    g.py[2] = round_up_residue(g.bd[0], 16, 4); g.pz[2] = g.py[2] * g.bd[1];
    int o = 0;
    for (int c = 0; c < 3; ++c) { g.off[c] = o; o += g.pz[c] * g.bd[2]; }
    g.total = o;
    return g;
}

template <int N, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
deposit_tile_kernel(SoaView P, BinsView bins, FabView Jx, FabView Jy, FabView Jz, DepositGeom dg, BlockGeom bg) {
    using T = TileCfg<N>;
    constexpr int S = T::S, PN = T::PN, NB = T::NB, NF = T::NF, CHP = DT_CHP;
    constexpr int QS = T::QS, QL = T::QL, NG = T::NG, QP = T::QP;
    constexpr unsigned FULL = 0xffffffffu;
    PIC_DYNAMIC_SMEM(double, smem);
    const int BD0 = bg.bd[0], BD1 = bg.bd[1], BD2 = bg.bd[2];
    double* jblk = smem;                                  // 3 components, padded pitches (BlockGeom)
    double* recs = smem + bg.total;                       // [NW][NF][CHP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int n = tid; n < bg.total; n += NW * 32) jblk[n] = 0.0;

    // supercell and its particle range
    const int t = blockIdx.x;
    int tc[3];
    tile_coords(bins, t, tc);
    const long tvol = (long)bins.tile[0] * bins.tile[1] * bins.tile[2];
    const int p_begin = min(bins.cell_start[(long)t * tvol], bins.np_limit);
    const int p_end = min(bins.cell_start[(long)(t + 1) * tvol], bins.np_limit);
    // block origin in global index space (same for all three components: node-based weights)
    const int o0 = bins.box_lo[0] + tc[0] * bins.tile[0] - DT_MARGIN_LO;
    const int o1 = bins.box_lo[1] + tc[1] * bins.tile[1] - DT_MARGIN_LO;
    const int o2 = bins.box_lo[2] + tc[2] * bins.tile[2] - DT_MARGIN_LO;
    __syncthreads();

    // contiguous share of the supercell's particles for this warp, in multiples of the chunk
    const int npt = p_end - p_begin;
    const int nchunks = (npt + DT_CH - 1) / DT_CH;
    const int cpw = (nchunks + NW - 1) / NW;
    const int c_begin = warp * cpw, c_end = min(nchunks, c_begin + cpw);

    double* rec = recs + (size_t)warp * NF * CHP;

    // ---- phase-2 roles of this lane ----
    // general layout: lane (a, bh) owns lines (a, b), b = bh + nb*BH, of all three components
    const int a = lane % S, bh = lane / S;
    const bool active_g = lane < T::NLANES;
    // quiet layout: lane (g, u, v): particle slot g of the pass;
    //   Jx line (j, k) = (1+u, 1+v);  Jy line (i, k) = (1+ur, 1+v);  Jz line (i, j) = (1+ur, 1+v)
    // where ur = (u - (ax+1)) mod QS: a lane owns the Jy/Jz lines of a fixed ABSOLUTE x (ring
    // mapping), so that moving the anchor by one cell along x keeps them in place.
    const int g = lane / QL, ql = lane % QL, qu = ql % QS, qv = ql / QS;
    const bool active_q = g < NG;

    double accg[NB][3][PN];
    double accq[3][QP];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
#pragma unroll
            for (int i = 0; i < PN; ++i) accg[nb][c][i] = 0.0;
#pragma unroll
        for (int i = 0; i < QP; ++i) accq[c][i] = 0.0;
    }
    int cur = -1;
    bool dirty_g = false, dirty_q = false;

    auto jaddr = [&](int c, int x, int y, int z) -> double* {
        return jblk + bg.off[c] + x + bg.py[c] * y + bg.pz[c] * z;
    };
    auto ring = [&](int ax) -> int {               // relative x index of this lane's Jy/Jz lines
        int r = (qu - (ax + 1)) % QS;
        return r < 0 ? r + QS : r;
    };
    auto fold = [&](double v) -> double {          // sum over the particle slots of the pass
        double r = v;
#pragma unroll
        for (int gg = 1; gg < NG; ++gg) {
            const double o = __shfl_down_sync(FULL, v, gg * QL);
            if (lane + gg * QL < NG * QL) r += o;
        }
        return r;
    };
    // keys pack the anchor as ax | ay << 8 | az << 16 (block extents are < 256)
    auto flush_general = [&](int k) {
        if (!dirty_g) return;
        const int ax = k & 255, ay = (k >> 8) & 255, az = k >> 16;
        if (active_g) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                const int b = bh + nb * T::BH;
                if (b < S) {
#pragma unroll
                    for (int i = 0; i < PN; ++i) {
                        const double vx = accg[nb][0][i], vy = accg[nb][1][i], vz = accg[nb][2][i];
                        if (vx != 0.0) atomicAdd(jaddr(0, ax + i, ay + a, az + b), vx);   // line (j=a, k=b)
                        if (vy != 0.0) atomicAdd(jaddr(1, ax + a, ay + i, az + b), vy);   // line (i=a, k=b)
                        if (vz != 0.0) atomicAdd(jaddr(2, ax + a, ay + b, az + i), vz);   // line (i=a, j=b)
                        accg[nb][0][i] = 0.0; accg[nb][1][i] = 0.0; accg[nb][2][i] = 0.0;
                    }
                }
            }
        }
        dirty_g = false;
    };
    auto flush_quiet = [&](int k) {                // retire the whole quiet window
        if (!dirty_q) return;
        const int ax = k & 255, ay = (k >> 8) & 255, az = k >> 16;
        const int ur = ring(ax);
#pragma unroll
        for (int i = 0; i < QP; ++i) {
            const double vx = fold(accq[0][i]), vy = fold(accq[1][i]), vz = fold(accq[2][i]);
            if (lane < QL) {
                if (vx != 0.0) atomicAdd(jaddr(0, ax + 1 + i, ay + 1 + qu, az + 1 + qv), vx);
                if (vy != 0.0) atomicAdd(jaddr(1, ax + 1 + ur, ay + 1 + i, az + 1 + qv), vy);
                if (vz != 0.0) atomicAdd(jaddr(2, ax + 1 + ur, ay + 1 + qv, az + 1 + i), vz);
            }
            accq[0][i] = 0.0; accq[1][i] = 0.0; accq[2][i] = 0.0;
        }
        dirty_q = false;
    };
    auto slide_quiet = [&](int k) {                // anchor moves from k to k + 1 along x
        if (!dirty_q) return;
        const int ax = k & 255, ay = (k >> 8) & 255, az = k >> 16;
        const bool leaving = ring(ax) == 0;        // this lane's Jy/Jz lines sit at x = ax + 1
        // Jx: the plane x = ax + 1 (entry 0) leaves the window; the others shift down
        {
            const double vx = fold(accq[0][0]);
            if (lane < QL && vx != 0.0) atomicAdd(jaddr(0, ax + 1, ay + 1 + qu, az + 1 + qv), vx);
#pragma unroll
            for (int i = 0; i + 1 < QP; ++i) accq[0][i] = accq[0][i + 1];
            accq[0][QP - 1] = 0.0;
        }
        // Jy, Jz: the lines at x = ax + 1 leave; their lanes restart at zero for x = ax + 1 + QS
#pragma unroll
        for (int i = 0; i < QP; ++i) {
            const double vy = fold(accq[1][i]), vz = fold(accq[2][i]);
            if (lane < QL && leaving) {
                if (vy != 0.0) atomicAdd(jaddr(1, ax + 1, ay + 1 + i, az + 1 + qv), vy);
                if (vz != 0.0) atomicAdd(jaddr(2, ax + 1, ay + 1 + qv, az + 1 + i), vz);
            }
            if (leaving) { accq[1][i] = 0.0; accq[2][i] = 0.0; }
        }
    };
    auto retire = [&](int k_old, int k_new) {      // anchor changes from k_old to k_new
        if (k_old < 0) return;
        flush_general(k_old);
        if (k_new == k_old + 1 && (k_new & 255) + S <= BD0) slide_quiet(k_old);
        else flush_quiet(k_old);
    };

    // software prefetch: the particle data of chunk ch+1 is requested before chunk ch is processed
    double pf[7] = {0, 0, 0, 0, 0, 0, 0};
    auto prefetch = [&](int ch) {
        const int ip = p_begin + ch * DT_CH + lane;
        if (ch < c_end && ip < p_end) {
            pf[0] = P.x[ip]; pf[1] = P.y[ip]; pf[2] = P.z[ip]; pf[3] = P.w[ip];
            pf[4] = P.ux[ip]; pf[5] = P.uy[ip]; pf[6] = P.uz[ip];
        }
    };
    prefetch(c_begin);

    for (int ch = c_begin; ch < c_end; ++ch) {
        const int base = p_begin + ch * DT_CH;
        const int nval = min(DT_CH, p_end - base);
        const double xp = pf[0], yp = pf[1], zp = pf[2], wp = pf[3], uxp = pf[4], uyp = pf[5], uzp = pf[6];
        prefetch(ch + 1);
        // ---------------- phase 1: lane = particle ----------------
        int key = -2;
        bool quiet = false;
        if (lane < nval) {
            const double gaminv = 1.0 / sqrt(1.0 + uxp * uxp * INV_C2 + uyp * uyp * INV_C2 + uzp * uzp * INV_C2);
            const double wq = dg.q * wp;
            double pos_new[3], pos_old[3];
            deposit_coords(xp, dg.xyzmin[0], dg.tshift, uxp, gaminv, dg.dinv[0], dg.dt, pos_new[0], pos_old[0]);
            deposit_coords(yp, dg.xyzmin[1], dg.tshift, uyp, gaminv, dg.dinv[1], dg.dt, pos_new[1], pos_old[1]);
            deposit_coords(zp, dg.xyzmin[2], dg.tshift, uzp, gaminv, dg.dinv[2], dg.dt, pos_new[2], pos_old[2]);
            double sn[3][S], so[3][S];
            int inew[3], sh[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) inew[d] = dir_weights<N>(pos_new[d], pos_old[d], sn[d], so[d], sh[d]);
            // anchor in block coordinates
            const int ax = dg.lo[0] + inew[0] - 1 - o0, ay = dg.lo[1] + inew[1] - 1 - o1, az = dg.lo[2] + inew[2] - 1 - o2;
            const bool fits = ax >= 0 && ay >= 0 && az >= 0 && ax + S <= BD0 && ay + S <= BD1 && az + S <= BD2;
            if (fits) {
                key = ax | (ay << 8) | (az << 16);
                quiet = (sh[0] == 0) && (sh[1] == 0) && (sh[2] == 0);
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    rec[(T::F_SNX + s) * CHP + lane] = sn[0][s];
                    rec[(T::F_SOX + s) * CHP + lane] = so[0][s];
                    rec[(T::F_SNY + s) * CHP + lane] = sn[1][s];
                    rec[(T::F_SOY + s) * CHP + lane] = so[1][s];
                    rec[(T::F_AY + s) * CHP + lane] = (1.0 / 3.0) * sn[1][s] + (1.0 / 6.0) * so[1][s];
                    rec[(T::F_BY + s) * CHP + lane] = (1.0 / 3.0) * so[1][s] + (1.0 / 6.0) * sn[1][s];
                    rec[(T::F_AZ + s) * CHP + lane] = (1.0 / 3.0) * sn[2][s] + (1.0 / 6.0) * so[2][s];
                    rec[(T::F_BZ + s) * CHP + lane] = (1.0 / 3.0) * so[2][s] + (1.0 / 6.0) * sn[2][s];
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double wqd = wq * dg.invdtd[d];
                    // loop trimming of the reference (:777-788): entries outside [dl, N+1-du] are not deposited
                    const int dl = (sh[d] < 0) ? 0 : 1, du = (sh[d] > 0) ? 0 : 1;
                    double run = 0.0;
#pragma unroll
                    for (int i = 0; i < PN; ++i) {
                        run += wqd * (so[d][i] - sn[d][i]);
                        const bool live = (i >= dl) && (i <= N + 1 - du);
                        rec[(T::F_CDS + d * PN + i) * CHP + lane] = live ? run : 0.0;
                    }
                }
            } else {
                key = -1;
                deposit_one_global<N>(xp, yp, zp, wp, uxp, uyp, uzp, Jx, Jy, Jz, dg);
            }
        }
        __syncwarp();
        // ---------------- phase 2: lane = stencil lines ----------------
        {
            const int prev = __shfl_up_sync(FULL, key, 1);
            const bool head = (lane < nval) && (lane == 0 || key != prev);
            unsigned heads = __ballot_sync(FULL, head);
            const unsigned quietm = __ballot_sync(FULL, quiet);
            while (heads) {
                const int start = __ffs(heads) - 1;
                heads &= heads - 1;
                const int end = heads ? (__ffs(heads) - 1) : nval;
                const unsigned runm = ((end >= 32) ? FULL : ((1u << end) - 1u)) & ~((1u << start) - 1u);
                const int k = __shfl_sync(FULL, key, start);
                if (k < 0) continue;
                if (k != cur) { retire(cur, k); cur = k; }
                unsigned mq = runm & quietm, mg = runm & ~quietm;
                // ---- quiet particles: NG per pass ----
                while (mq) {
                    int pq = -1;
#pragma unroll
                    for (int gg = 0; gg < NG; ++gg) {
                        const int p = mq ? (__ffs(mq) - 1) : -1;
                        if (mq) mq &= mq - 1;
                        if (gg == g) pq = p;
                    }
                    if (active_q && pq >= 0) {
                        const int ur = ring(k & 255);
                        const double snx = rec[(T::F_SNX + 1 + ur) * CHP + pq], sox = rec[(T::F_SOX + 1 + ur) * CHP + pq];
                        const double sny = rec[(T::F_SNY + 1 + qu) * CHP + pq], soy = rec[(T::F_SOY + 1 + qu) * CHP + pq];
                        const double ay_ = rec[(T::F_AY + 1 + qv) * CHP + pq], by_ = rec[(T::F_BY + 1 + qv) * CHP + pq];
                        const double az_ = rec[(T::F_AZ + 1 + qv) * CHP + pq], bz_ = rec[(T::F_BZ + 1 + qv) * CHP + pq];
                        const double wx = sny * az_ + soy * bz_;   // Jx line (j, k) = (1+u, 1+v)
                        const double wy = snx * az_ + sox * bz_;   // Jy line (i, k) = (1+ur, 1+v)
                        const double wz = snx * ay_ + sox * by_;   // Jz line (i, j) = (1+ur, 1+v)
#pragma unroll
                        for (int i = 0; i < QP; ++i) {
                            accq[0][i] += rec[(T::F_CDS + 0 * PN + 1 + i) * CHP + pq] * wx;
                            accq[1][i] += rec[(T::F_CDS + 1 * PN + 1 + i) * CHP + pq] * wy;
                            accq[2][i] += rec[(T::F_CDS + 2 * PN + 1 + i) * CHP + pq] * wz;
                        }
                    }
                    dirty_q = true;
                }
                // ---- general particles: one per pass ----
                while (mg) {
                    const int pp = __ffs(mg) - 1;
                    mg &= mg - 1;
                    if (active_g) {
                        const double snx = rec[(T::F_SNX + a) * CHP + pp], sox = rec[(T::F_SOX + a) * CHP + pp];
                        const double sny = rec[(T::F_SNY + a) * CHP + pp], soy = rec[(T::F_SOY + a) * CHP + pp];
                        double cds[3][PN];
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int i = 0; i < PN; ++i) cds[c][i] = rec[(T::F_CDS + c * PN + i) * CHP + pp];
#pragma unroll
                        for (int nb = 0; nb < NB; ++nb) {
                            const int b = bh + nb * T::BH;
                            if (b < S) {
                                const double ay_ = rec[(T::F_AY + b) * CHP + pp], by_ = rec[(T::F_BY + b) * CHP + pp];
                                const double az_ = rec[(T::F_AZ + b) * CHP + pp], bz_ = rec[(T::F_BZ + b) * CHP + pp];
                                const double wx = sny * az_ + soy * bz_;   // Jx line (j=a, k=b)
                                const double wy = snx * az_ + sox * bz_;   // Jy line (i=a, k=b)
                                const double wz = snx * ay_ + sox * by_;   // Jz line (i=a, j=b)
#pragma unroll
                                for (int i = 0; i < PN; ++i) {
                                    accg[nb][0][i] += cds[0][i] * wx;
                                    accg[nb][1][i] += cds[1][i] * wy;
                                    accg[nb][2][i] += cds[2][i] * wz;
                                }
                            }
                        }
                    }
                    dirty_g = true;
                }
            }
        }
        __syncwarp();
    }
    if (cur >= 0) { flush_general(cur); flush_quiet(cur); }
    __syncthreads();

    // ---------------- block -> global J (coalesced along i; halo overlaps neighbours) ----------
    const int nrows = BD1 * BD2;
    for (int c = 0; c < 3; ++c) {
        const FabView& J = (c == 0) ? Jx : ((c == 1) ? Jy : Jz);
        for (int row = warp; row < nrows; row += NW) {
            const int lj = row % BD1, lk = row / BD1;
            const int gj = o1 + lj, gk = o2 + lk;
            if (gj < J.lo1 || gj >= J.lo1 + J.n1 || gk < J.lo2 || gk >= J.lo2 + J.n2) continue;
            const double* src = jblk + bg.off[c] + bg.py[c] * lj + bg.pz[c] * lk;
            for (int li = lane; li < BD0; li += 32) {
                const double v = src[li];
                const int gi = o0 + li;
                if (v != 0.0 && gi >= J.lo0 && gi < J.lo0 + J.n0) atomicAdd(&J(gi, gj, gk), v);
            }
        }
    }
}

template <int N, int NW>
static int launch_tile(SoaView P, const BinsView& bv, const pic_fab J[3], const DepositGeom& dg,
                       cudaStream_t s) {
    using T = TileCfg<N>;
    const BlockGeom bg = make_block_geom(bv);
    if (bg.bd[0] > 255 || bg.bd[1] > 255 || bg.bd[2] > 255) return fail("pic_deposit_esirkepov: supercell too large");
    const size_t smem = (size_t)(bg.total + (size_t)NW * T::NF * DT_CHP) * sizeof(double);
    auto kern = deposit_tile_kernel<N, NW>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return fail("pic_deposit_esirkepov: cannot raise dynamic shared memory limit");
        attr_done = true;
    }
    if (smem > 227 * 1024) return fail("pic_deposit_esirkepov: supercell too large for shared memory (%zu B)", smem);
    const int ntiles = bv.nt[0] * bv.nt[1] * bv.nt[2];
    kern<<<ntiles, NW * 32, smem, s>>>(P, bv, make_view(J[0]), make_view(J[1]), make_view(J[2]), dg, bg);
    count_launch();
    return check_launch("pic_deposit_esirkepov(tile)") ? 0 : 1;
}

int deposit_tile_launch(const pic_soa* p, long offset, long np, const pic_fab J[3],
                        const DepositGeom& dg, int nox, const pic_bins* bins, cudaStream_t s) {
    PIC_REQUIRE(offset == 0 && np == p->np, "pic_deposit_esirkepov: bins describe the whole tile (offset 0, np = all)");
    BinsView bv = make_bins(*bins);
    bv.np_limit = (int)(bins->np_binned < np ? bins->np_binned : np);
    SoaView P = make_soa(*p, 0);
    if (nox == 1) return launch_tile<1, 8>(P, bv, J, dg, s);
    if (nox == 2) return launch_tile<2, 8>(P, bv, J, dg, s);
    return launch_tile<3, 8>(P, bv, J, dg, s);
}

}  // namespace pic
