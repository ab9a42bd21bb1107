// Maxwell FDTD on the Yee grid: EvolveB / EvolveE, Yee and CKC stencils.
//
// Replaces FiniteDifferenceSolver::EvolveBCartesian<T_Algo> / EvolveECartesian<T_Algo>
// (reference: Source/FieldSolver/FiniteDifferenceSolver/EvolveB.cpp:122-186, EvolveE.cpp:120-216;
// stencils CartesianYeeAlgorithm.H:69-101, CartesianCKCAlgorithm.H:130-299).  The reference
// launches three lambdas over three staggered index boxes; here ONE kernel walks the cell box
// (plus the upper nodal layer) and updates all three components of a point, so each loaded
// neighbour value is reused from registers/L1 by the components that share it.
//
// Memory bound: algorithmic traffic 72 B/cell (EvolveB) and 96 B/cell (EvolveE), fp64.
// Layout: i is the contiguous index -> a warp reads 32 consecutive doubles (256 B) per row.
#include "pic_common.cuh"

namespace pic {

struct Coefs { double x[5], y[5], z[5]; };

// Box of points visited: cells [clo, chi] plus one extra layer on the high side (nodal duplicates).
struct PointBox { int lo[3]; int n[3]; };

template <int ALGO> struct Stencil;

template <> struct Stencil<PIC_SOLVER_YEE> {
    template <int D>
    static __device__ __forceinline__ double up(const FabView& F, const double* c, int i, int j, int k) {
        constexpr int di = (D == 0), dj = (D == 1), dk = (D == 2);
        return c[0] * (F.ld(i + di, j + dj, k + dk) - F.ld(i, j, k));
    }
};

// CKC: 18-point upward differences; coefficient slots as stored by the reference
// (CartesianCKCAlgorithm.H:84-101): [1]=alpha, [2],[3]=betas, [4]=gamma/d.
template <> struct Stencil<PIC_SOLVER_CKC> {
    template <int D>
    static __device__ __forceinline__ double up(const FabView& F, const double* c, int i, int j, int k) {
        const double alpha = c[1], b1 = c[2], b2 = c[3], gamma = c[4];
        if constexpr (D == 0) {
            return alpha * (F.ld(i+1,j,k) - F.ld(i,j,k))
                 + b1 * (F.ld(i+1,j+1,k) - F.ld(i,j+1,k) + F.ld(i+1,j-1,k) - F.ld(i,j-1,k))
                 + b2 * (F.ld(i+1,j,k+1) - F.ld(i,j,k+1) + F.ld(i+1,j,k-1) - F.ld(i,j,k-1))
                 + gamma * (F.ld(i+1,j+1,k+1) - F.ld(i,j+1,k+1) + F.ld(i+1,j-1,k+1) - F.ld(i,j-1,k+1)
                          + F.ld(i+1,j+1,k-1) - F.ld(i,j+1,k-1) + F.ld(i+1,j-1,k-1) - F.ld(i,j-1,k-1));
        } else if constexpr (D == 1) {
            // y: [3] = beta_yx multiplies the x neighbours, [2] = beta_yz the z neighbours
            return alpha * (F.ld(i,j+1,k) - F.ld(i,j,k))
                 + b2 * (F.ld(i+1,j+1,k) - F.ld(i+1,j,k) + F.ld(i-1,j+1,k) - F.ld(i-1,j,k))
                 + b1 * (F.ld(i,j+1,k+1) - F.ld(i,j,k+1) + F.ld(i,j+1,k-1) - F.ld(i,j,k-1))
                 + gamma * (F.ld(i+1,j+1,k+1) - F.ld(i+1,j,k+1) + F.ld(i-1,j+1,k+1) - F.ld(i-1,j,k+1)
                          + F.ld(i+1,j+1,k-1) - F.ld(i+1,j,k-1) + F.ld(i-1,j+1,k-1) - F.ld(i-1,j,k-1));
        } else {
            // z: [2] = beta_zx multiplies the x neighbours, [3] = beta_zy the y neighbours
            return alpha * (F.ld(i,j,k+1) - F.ld(i,j,k))
                 + b1 * (F.ld(i+1,j,k+1) - F.ld(i+1,j,k) + F.ld(i-1,j,k+1) - F.ld(i-1,j,k))
                 + b2 * (F.ld(i,j+1,k+1) - F.ld(i,j+1,k) + F.ld(i,j-1,k+1) - F.ld(i,j-1,k))
                 + gamma * (F.ld(i+1,j+1,k+1) - F.ld(i+1,j+1,k) + F.ld(i-1,j+1,k+1) - F.ld(i-1,j+1,k)
                          + F.ld(i+1,j-1,k+1) - F.ld(i+1,j-1,k) + F.ld(i-1,j-1,k+1) - F.ld(i-1,j-1,k));
        }
    }
};

// Downward differences are plain 2-point for both algorithms
// (CartesianYeeAlgorithm.H:88-101, CartesianCKCAlgorithm.H:170-183,223-241,286-299).
template <int D>
__device__ __forceinline__ double down(const FabView& F, const double* c, int i, int j, int k) {
    constexpr int di = (D == 0), dj = (D == 1), dk = (D == 2);
    return c[0] * (F.ld(i, j, k) - F.ld(i - di, j - dj, k - dk));
}

constexpr int FDTD_BX = 64, FDTD_BY = 4;   // 256 threads: 64 consecutive i, 4 rows of j

template <int ALGO>
__global__ void __launch_bounds__(FDTD_BX * FDTD_BY)
evolve_b_kernel(FabView Bx, FabView By, FabView Bz, FabView Ex, FabView Ey, FabView Ez,
                Coefs cf, PointBox pb, double dt) {
    const int li = blockIdx.x * FDTD_BX + threadIdx.x;
    const int lj = blockIdx.y * FDTD_BY + threadIdx.y;
    const int lk = blockIdx.z;
    if (li >= pb.n[0] || lj >= pb.n[1]) return;
    const int i = pb.lo[0] + li, j = pb.lo[1] + lj, k = pb.lo[2] + lk;
    // last layer in a direction only exists for components nodal in that direction
    const bool in_x = li < pb.n[0] - 1, in_y = lj < pb.n[1] - 1, in_z = lk < pb.n[2] - 1;
    using S = Stencil<ALGO>;
    if (in_y && in_z) {   // Bx(1,0,0)  EvolveB.cpp:168-171
        Bx(i, j, k) += dt * S::template up<2>(Ey, cf.z, i, j, k) - dt * S::template up<1>(Ez, cf.y, i, j, k);
    }
    if (in_x && in_z) {   // By(0,1,0)  :175-178
        By(i, j, k) += dt * S::template up<0>(Ez, cf.x, i, j, k) - dt * S::template up<2>(Ex, cf.z, i, j, k);
    }
    if (in_x && in_y) {   // Bz(0,0,1)  :182-185
        Bz(i, j, k) += dt * S::template up<1>(Ex, cf.y, i, j, k) - dt * S::template up<0>(Ey, cf.x, i, j, k);
    }
}

__global__ void __launch_bounds__(FDTD_BX * FDTD_BY)
evolve_e_kernel(FabView Ex, FabView Ey, FabView Ez, FabView Bx, FabView By, FabView Bz,
                FabView jx, FabView jy, FabView jz, Coefs cf, PointBox pb, double dt) {
    const int li = blockIdx.x * FDTD_BX + threadIdx.x;
    const int lj = blockIdx.y * FDTD_BY + threadIdx.y;
    const int lk = blockIdx.z;
    if (li >= pb.n[0] || lj >= pb.n[1]) return;
    const int i = pb.lo[0] + li, j = pb.lo[1] + lj, k = pb.lo[2] + lk;
    const bool in_x = li < pb.n[0] - 1, in_y = lj < pb.n[1] - 1, in_z = lk < pb.n[2] - 1;
    constexpr double c2 = C_LIGHT * C_LIGHT;
    if (in_x) {           // Ex(0,1,1)  EvolveE.cpp:185-188
        Ex(i, j, k) += c2 * dt * (-down<2>(By, cf.z, i, j, k) + down<1>(Bz, cf.y, i, j, k)
                                  - MU0 * jx.ld(i, j, k));
    }
    if (in_y) {           // Ey(1,0,1)  :201-204
        Ey(i, j, k) += c2 * dt * (-down<0>(Bz, cf.x, i, j, k) + down<2>(Bx, cf.z, i, j, k)
                                  - MU0 * jy.ld(i, j, k));
    }
    if (in_z) {           // Ez(1,1,0)  :210-213
        Ez(i, j, k) += c2 * dt * (-down<1>(Bx, cf.y, i, j, k) + down<0>(By, cf.x, i, j, k)
                                  - MU0 * jz.ld(i, j, k));
    }
}

static int point_box(const pic_fab E[3], PointBox* pb) {
    // cells of the box in direction d are the valid points of the component that is
    // cell-centred in d (Ex in x, Ey in y, Ez in z)
    for (int d = 0; d < 3; ++d) {
        pb->lo[d] = vlo(E[d], d);
        pb->n[d] = vhi(E[d], d) - vlo(E[d], d) + 2;   // cells + upper nodal layer
    }
    return 0;
}

static int check_guards(const pic_fab B[3], const pic_fab E[3], const pic_stencil* st) {
    PIC_REQUIRE(is_yee(E, B), "pic_evolve: fields must have Yee staggering (WarpX.cpp:2117-2125)");
    PIC_REQUIRE(st->algo == PIC_SOLVER_YEE || st->algo == PIC_SOLVER_CKC, "pic_evolve: unknown solver %d", st->algo);
    for (int c = 0; c < 3; ++c)
        for (int d = 0; d < 3; ++d) {
            // EvolveE reads B at -1 (CartesianYeeAlgorithm.H:61-64); CKC EvolveB reads E at +-1
            PIC_REQUIRE(B[c].ng[d] >= 1, "pic_evolve: B needs >= 1 guard cell");
            PIC_REQUIRE(st->algo == PIC_SOLVER_YEE || E[c].ng[d] >= 1, "pic_evolve: CKC needs >= 1 guard cell of E");
        }
    return 0;
}

// fdtd_bulk.cu: the Yee kernels with bulk-asynchronous staging (done = false: not applicable, take the kernels here)
int evolve_b_bulk_launch(const pic_fab B[3], const pic_fab E[3], const pic_stencil* st, const int lo[3], const int n[3],
                         double dt, cudaStream_t s, bool* done);
int evolve_e_bulk_launch(const pic_fab E[3], const pic_fab B[3], const pic_fab J[3], const pic_stencil* st, const int lo[3],
                         const int n[3], double dt, cudaStream_t s, bool* done);

}  // namespace pic

using namespace pic;

extern "C" int pic_evolve_b(const pic_fab B[3], const pic_fab E[3], const pic_stencil* st, double dt,
                            void* stream) {
    if (int rc = check_guards(B, E, st)) return rc;
    PointBox pb; point_box(E, &pb);
    Coefs cf;
    for (int n = 0; n < 5; ++n) { cf.x[n] = st->cx[n]; cf.y[n] = st->cy[n]; cf.z[n] = st->cz[n]; }
    dim3 block(FDTD_BX, FDTD_BY, 1);
    dim3 grid((pb.n[0] + FDTD_BX - 1) / FDTD_BX, (pb.n[1] + FDTD_BY - 1) / FDTD_BY, pb.n[2]);
    cudaStream_t s = (cudaStream_t)stream;
    {
        bool done = false;
        if (int rc = evolve_b_bulk_launch(B, E, st, pb.lo, pb.n, dt, s, &done)) return rc;
        if (done) return 0;
    }
    if (st->algo == PIC_SOLVER_YEE)
        evolve_b_kernel<PIC_SOLVER_YEE><<<grid, block, 0, s>>>(make_view(B[0]), make_view(B[1]), make_view(B[2]),
            make_view(E[0]), make_view(E[1]), make_view(E[2]), cf, pb, dt);
    else
        evolve_b_kernel<PIC_SOLVER_CKC><<<grid, block, 0, s>>>(make_view(B[0]), make_view(B[1]), make_view(B[2]),
            make_view(E[0]), make_view(E[1]), make_view(E[2]), cf, pb, dt);
    count_launch();
    return check_launch("pic_evolve_b") ? 0 : 1;
}

extern "C" int pic_evolve_e(const pic_fab E[3], const pic_fab B[3], const pic_fab J[3],
                            const pic_stencil* st, double dt, void* stream) {
    if (int rc = check_guards(B, E, st)) return rc;
    for (int c = 0; c < 3; ++c)
        for (int d = 0; d < 3; ++d)
            PIC_REQUIRE(J[c].stag[d] == E[c].stag[d], "pic_evolve_e: J must be staggered like E");
    PointBox pb; point_box(E, &pb);
    Coefs cf;
    for (int n = 0; n < 5; ++n) { cf.x[n] = st->cx[n]; cf.y[n] = st->cy[n]; cf.z[n] = st->cz[n]; }
    dim3 block(FDTD_BX, FDTD_BY, 1);
    dim3 grid((pb.n[0] + FDTD_BX - 1) / FDTD_BX, (pb.n[1] + FDTD_BY - 1) / FDTD_BY, pb.n[2]);
    {
        bool done = false;
        if (int rc = evolve_e_bulk_launch(E, B, J, st, pb.lo, pb.n, dt, (cudaStream_t)stream, &done)) return rc;
        if (done) return 0;
    }
    evolve_e_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(make_view(E[0]), make_view(E[1]), make_view(E[2]),
        make_view(B[0]), make_view(B[1]), make_view(B[2]), make_view(J[0]), make_view(J[1]), make_view(J[2]),
        cf, pb, dt);
    count_launch();
    return check_launch("pic_evolve_e") ? 0 : 1;
}
