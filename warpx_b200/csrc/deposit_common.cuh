// Per-particle quantities of the Esirkepov deposition, shared by the global-atomic kernel and the
// shared-memory tile kernel.  Follows Source/Particles/Deposition/CurrentDeposition.H:683-788.
#ifndef PIC_DEPOSIT_COMMON_CUH_
#define PIC_DEPOSIT_COMMON_CUH_
#include "pic_common.cuh"

namespace pic {

struct DepositGeom {
    double dinv[3];
    double xyzmin[3];
    int lo[3];
    double q;
    double dt;
    double tshift;      // relative_time + 0.5*dt  (CurrentDeposition.H:725)
    double invdtd[3];   // (1/dt)*dinv.y*dinv.z, ... (:671-673)
};

template <int N>
struct EsirkepovWeights {
    double sx_new[N + 3], sx_old[N + 3];
    double sy_new[N + 3], sy_old[N + 3];
    double sz_new[N + 3], sz_old[N + 3];
    int i_new, j_new, k_new;
    int dil, diu, djl, dju, dkl, dku;
    double wqx, wqy, wqz;

    __device__ __forceinline__ void compute(double xp, double yp, double zp, double wp, double uxp,
                                            double uyp, double uzp, const DepositGeom& dg) {
        const double gaminv = 1.0 / sqrt(1.0 + uxp * uxp * INV_C2 + uyp * uyp * INV_C2 + uzp * uzp * INV_C2);
        const double wq = dg.q * wp;                                     // :691
        wqx = wq * dg.invdtd[0]; wqy = wq * dg.invdtd[1]; wqz = wq * dg.invdtd[2];
        // new and old positions in grid units (:725-736)
        double x_new, x_old, y_new, y_old, z_new, z_old;
        deposit_coords(xp, dg.xyzmin[0], dg.tshift, uxp, gaminv, dg.dinv[0], dg.dt, x_new, x_old);
        deposit_coords(yp, dg.xyzmin[1], dg.tshift, uyp, gaminv, dg.dinv[1], dg.dt, y_new, y_old);
        deposit_coords(zp, dg.xyzmin[2], dg.tshift, uzp, gaminv, dg.dinv[2], dg.dt, z_new, z_old);
#pragma unroll
        for (int n = 0; n < N + 3; ++n) {
            sx_new[n] = sx_old[n] = sy_new[n] = sy_old[n] = sz_new[n] = sz_old[n] = 0.0;
        }
        i_new = shape_factor<N>(sx_new + 1, x_new);                      // :759-773
        const int i_old = shifted_shape_factor<N>(sx_old, x_old, i_new);
        j_new = shape_factor<N>(sy_new + 1, y_new);
        const int j_old = shifted_shape_factor<N>(sy_old, y_old, j_new);
        k_new = shape_factor<N>(sz_new + 1, z_new);
        const int k_old = shifted_shape_factor<N>(sz_old, z_old, k_new);
        dil = (i_old < i_new) ? 0 : 1; diu = (i_old > i_new) ? 0 : 1;   // :777-788
        djl = (j_old < j_new) ? 0 : 1; dju = (j_old > j_new) ? 0 : 1;
        dkl = (k_old < k_new) ? 0 : 1; dku = (k_old > k_new) ? 0 : 1;
    }
};

// ---- shared by the register-run kernels (deposit_runs.cu) and the lane-per-cell kernel (deposit_cells.cu) ----
struct J3 { FabView v[3]; };

// weights of the new (wn) and old (wo) position and the slot shift of the old stencil
template <int N>
__device__ __forceinline__ int dr_dir(double x_new, double x_old, double* wn, double* wo, int& sh) {
    const int i_new = shape_factor<N>(wn, x_new);
    int i_old;
    if constexpr (N == 1) {   // order 1 uses floor for the old position (ShapeFactors.H:110)
        const int i = (int)floor(x_old);
        const double d = x_old - (double)i;
        wo[0] = 1.0 - d; wo[1] = d;
        i_old = i;
    } else {
        i_old = shape_factor<N>(wo, x_old);
    }
    sh = i_old - i_new;
    return i_new;
}

struct ParticleGeom {   // new/old positions in grid units and charge factor of one particle
    double pos_new[3], pos_old[3], wq;
};
__device__ __forceinline__ ParticleGeom particle_geom(double xp, double yp, double zp, double wp, double uxp,
                                                      double uyp, double uzp, const DepositGeom& dg) {
    ParticleGeom g;
    const double gaminv = 1.0 / sqrt(1.0 + uxp * uxp * INV_C2 + uyp * uyp * INV_C2 + uzp * uzp * INV_C2);  // :687-689
    g.wq = dg.q * wp;
    deposit_coords(xp, dg.xyzmin[0], dg.tshift, uxp, gaminv, dg.dinv[0], dg.dt, g.pos_new[0], g.pos_old[0]);   // :725-736
    deposit_coords(yp, dg.xyzmin[1], dg.tshift, uyp, gaminv, dg.dinv[1], dg.dt, g.pos_new[1], g.pos_old[1]);
    deposit_coords(zp, dg.xyzmin[2], dg.tshift, uzp, gaminv, dg.dinv[2], dg.dt, g.pos_new[2], g.pos_old[2]);
    return g;
}

// anchor (global index of slot 0 of the stencil) packed relative to the J arrays, 10 bits each
struct KeyBase { int b0, b1, b2; };
__device__ __forceinline__ int pack_key(int gx, int gy, int gz, const KeyBase& kb) {
    return (gx - kb.b0) | ((gy - kb.b1) << 10) | ((gz - kb.b2) << 20);
}

}  // namespace pic
#endif
