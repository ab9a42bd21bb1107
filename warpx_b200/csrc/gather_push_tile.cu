// Supercell gather + push for cell-sorted particles (the timed path of PushPX / PushP).
//
// One CTA per supercell (pic_bins.tile).  The six field components are staged ONCE into a
// shared-memory block covering the supercell plus the gather halo (2 points each side: node
// weights reach c-1..c+2, cell weights c-2..c+2 for order 3), then every particle of the supercell
// gathers its 252 (order 3, Galerkin) grid values from shared memory.  Cell-sorted particles of
// the same cell are adjacent lanes reading the same addresses -> shared-memory broadcasts.
// Particles that drifted out of their supercell since the last sort are NOT gathered inline (one
// such lane would drag its whole warp through ~250 dependent global loads -- measured +40% kernel
// time per step since the last sort): they are appended to a list and a second, order-agnostic
// kernel pushes them with one thread each.  Any particle order is correct.
#include "pic_common.cuh"
#include "gather_common.cuh"
#include "bins.cuh"
#include <algorithm>

namespace pic {

constexpr int GT_HALO = 2;
constexpr int GT_THREADS = 256;
constexpr int GT_LOCAL_STRAYS = 1022;   // per-CTA list of particles that changed row inside the supercell

// Shared-memory block of the six components.  BD* > 0: compile-time extents (the 8x8x8 supercell:
// every stencil offset becomes an immediate of the LDS); BD0 == 0: run-time extents.
template <int BD0, int BD1, int BD2>
struct SmemFields {
    const double* blk;     // [6][BD2][BD1][BD0]
    int o0, o1, o2;        // global index of block element (0,0,0)
    int rb0, rb01, rbvol;  // run-time extents (generic instance only)
    struct Acc {
        const double* b; int s1, s2;
        __device__ __forceinline__ double operator()(int ix, int iy, int iz) const {
            if constexpr (BD0 > 0) return b[ix + BD0 * iy + BD0 * BD1 * iz];
            else return b[ix + s1 * iy + s2 * iz];
        }
    };
    __device__ __forceinline__ Acc at(int c, int i0, int j0, int k0) const {
        if constexpr (BD0 > 0)
            return Acc{blk + c * (BD0 * BD1 * BD2) + (i0 - o0) + BD0 * ((j0 - o1) + BD1 * (k0 - o2)), 0, 0};
        else
            return Acc{blk + c * rbvol + (i0 - o0) + rb0 * ((j0 - o1) + rb01 / rb0 * (k0 - o2)), rb0, rb01};
    }
};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gsrc) : "memory");
}

struct StrayList { int* idx; int* count; int cap; };

template <int N, int G, bool YEE, int TX, int TY, int TZ>
__global__ void __launch_bounds__(GT_THREADS, 2)
gather_push_tile_kernel(SoaView P, BinsView bins, GlobalFields gf, GatherGeom gg, double qdt2m,
                        double dt, int pusher, int push_position, EscapeView esc, StrayList stray) {
    extern __shared__ double smem[];
    constexpr bool FIXED = TX > 0;
    const int BD0 = FIXED ? TX + 2 * GT_HALO : bins.tile[0] + 2 * GT_HALO;
    const int BD1 = FIXED ? TY + 2 * GT_HALO : bins.tile[1] + 2 * GT_HALO;
    const int BD2 = FIXED ? TZ + 2 * GT_HALO : bins.tile[2] + 2 * GT_HALO;
    const int bvol = BD0 * BD1 * BD2;
    const int t = blockIdx.x;
    int tc[3];
    tile_coords(bins, t, tc);
    const long tvol = (long)bins.tile[0] * bins.tile[1] * bins.tile[2];
    const int p_begin = min(bins.cell_start[(long)t * tvol], bins.np_limit);
    const int p_end = min(bins.cell_start[(long)(t + 1) * tvol], bins.np_limit);
    if (p_begin >= p_end) return;
    const int t0 = bins.box_lo[0] + tc[0] * bins.tile[0];
    const int t1 = bins.box_lo[1] + tc[1] * bins.tile[1];
    const int t2 = bins.box_lo[2] + tc[2] * bins.tile[2];
    SmemFields<FIXED ? TX + 2 * GT_HALO : 0, FIXED ? TY + 2 * GT_HALO : 0, FIXED ? TZ + 2 * GT_HALO : 0> sf;
    sf.blk = smem; sf.o0 = t0 - GT_HALO; sf.o1 = t1 - GT_HALO; sf.o2 = t2 - GT_HALO;
    sf.rb0 = BD0; sf.rb01 = BD0 * BD1; sf.rbvol = bvol;

    // ---- stage the six sub-blocks with asynchronous 8-byte copies (LDGSTS): all ~40 copies of a
    //      thread are in flight at once, no register staging ----
    for (int n = threadIdx.x; n < 6 * bvol; n += GT_THREADS) {
        const int c = n / bvol, r = n - c * bvol;
        const int li = r % BD0, lj = (r / BD0) % BD1, lk = r / (BD0 * BD1);
        const FabView& F = gf.v[c];
        const int gi = sf.o0 + li, gj = sf.o1 + lj, gk = sf.o2 + lk;
        const bool in = gi >= F.lo0 && gi < F.lo0 + F.n0 && gj >= F.lo1 && gj < F.lo1 + F.n1 &&
                        gk >= F.lo2 && gk < F.lo2 + F.n2;
        if (in) cp_async8(smem + n, F.p + F.off(gi, gj, gk));
        else smem[n] = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // software prefetch: the next particle of this thread is requested before the current one is
    // gathered, so the HBM latency overlaps ~700 instructions of shared-memory gather + push
    double nx = 0, ny = 0, nz = 0, nux = 0, nuy = 0, nuz = 0;
    int ip = p_begin + threadIdx.x;
    if (ip < p_end) { nx = P.x[ip]; ny = P.y[ip]; nz = P.z[ip]; nux = P.ux[ip]; nuy = P.uy[ip]; nuz = P.uz[ip]; }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // Lanes of a warp hold consecutive cell-sorted particles, i.e. particles of ONE (y,z) row of
    // cells (or two at a row end): their stencil reads hit the same few shared-memory rows and
    // cost about one wavefront each.  A particle that changed row since the sort makes every LDS of
    // its warp conflict; such particles (isolated row among their neighbours) are deferred to a
    // CTA-local list and gathered after the main sweep, so the regular warps stay conflict-free.
    __shared__ int s_stray[GT_LOCAL_STRAYS];
    __shared__ int s_nstray;
    if (threadIdx.x == 0) s_nstray = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    auto finish = [&](int ipp, double xp, double yp, double zp, double ux, double uy, double uz, bool in_tile) {
        double F[6];
        if (in_tile) gather_fields<N, G, YEE>(sf, gg, xp, yp, zp, F);
        else gather_fields<N, G, YEE>(gf, gg, xp, yp, zp, F);      // global list full: inline, slow but correct
        push_particle(xp, yp, zp, ux, uy, uz, F, qdt2m, dt, pusher, push_position);
        P.ux[ipp] = ux; P.uy[ipp] = uy; P.uz[ipp] = uz;
        if (push_position) { P.x[ipp] = xp; P.y[ipp] = yp; P.z[ipp] = zp; esc.note(ipp, xp, yp, zp); }
    };
    auto cell_in_tile = [&](double xp, double yp, double zp, int& row) -> bool {
        // cell of the particle (global index) from the same coordinates the gather uses
        const int ci = gg.lo[0] + (int)((xp - gg.xyzmin[0]) * gg.dinv[0]);
        const int cj = gg.lo[1] + (int)((yp - gg.xyzmin[1]) * gg.dinv[1]);
        const int ck = gg.lo[2] + (int)((zp - gg.xyzmin[2]) * gg.dinv[2]);
        row = cj * 65536 + ck;
        return ci >= t0 && ci < t0 + bins.tile[0] && cj >= t1 && cj < t1 + bins.tile[1] &&
               ck >= t2 && ck < t2 + bins.tile[2];
    };

    for (; ip - lane < p_end; ip += GT_THREADS) {          // warp-uniform trip count
        const bool valid = ip < p_end;
        double xp = nx, yp = ny, zp = nz, ux = nux, uy = nuy, uz = nuz;
        const int in = ip + GT_THREADS;
        if (in < p_end) { nx = P.x[in]; ny = P.y[in]; nz = P.z[in]; nux = P.ux[in]; nuy = P.uy[in]; nuz = P.uz[in]; }
        int row = -1 - lane;
        const bool in_tile = valid && cell_in_tile(xp, yp, zp, row);
        const int row_prev = __shfl_up_sync(0xffffffffu, row, 1), row_next = __shfl_down_sync(0xffffffffu, row, 1);
        const bool isolated = !(lane > 0 && row == row_prev) && !(lane < 31 && row == row_next);
        bool now = valid;
        if (valid && !in_tile) {                  // left the supercell since the sort: second kernel
            const int n = atomicAdd(stray.count, 1);
            if (n < stray.cap) { stray.idx[n] = ip; now = false; }
        } else if (valid && isolated) {           // changed row inside the supercell: after the sweep
            const int n = atomicAdd(&s_nstray, 1);
            if (n < GT_LOCAL_STRAYS) { s_stray[n] = ip; now = false; }
        }
        __syncwarp();                             // one convergent pass through the 700-instruction body
        if (now) finish(ip, xp, yp, zp, ux, uy, uz, in_tile);
    }
    __syncthreads();
    const int nloc = min(s_nstray, GT_LOCAL_STRAYS);
    for (int t2l = threadIdx.x; t2l < nloc; t2l += GT_THREADS) {
        const int ipp = s_stray[t2l];
        finish(ipp, P.x[ipp], P.y[ipp], P.z[ipp], P.ux[ipp], P.uy[ipp], P.uz[ipp], true);
    }
}

// the listed strays, one thread each, fields through the read-only global path
template <int N, int G, bool YEE>
__global__ void __launch_bounds__(128)
gather_push_listed_kernel(SoaView P, StrayList stray, GlobalFields fld, GatherGeom gg, double qdt2m, double dt,
                          int pusher, int push_position, EscapeView esc) {
    const int n = min(*stray.count, stray.cap);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int ip = stray.idx[t];
        double xp = P.x[ip], yp = P.y[ip], zp = P.z[ip];
        double F[6];
        gather_fields<N, G, YEE>(fld, gg, xp, yp, zp, F);
        double ux = P.ux[ip], uy = P.uy[ip], uz = P.uz[ip];
        push_particle(xp, yp, zp, ux, uy, uz, F, qdt2m, dt, pusher, push_position);
        P.ux[ip] = ux; P.uy[ip] = uy; P.uz[ip] = uz;
        if (push_position) { P.x[ip] = xp; P.y[ip] = yp; P.z[ip] = zp; esc.note(ip, xp, yp, zp); }
    }
}

template <int N, int G>
static int launch(SoaView P, const BinsView& bv, const GlobalFields& gf, const GatherGeom& gg,
                  bool yee, double qdt2m, double dt, int pusher, int push_position, const EscapeView& esc,
                  cudaStream_t s) {
    const long bvol = (long)(bv.tile[0] + 2 * GT_HALO) * (bv.tile[1] + 2 * GT_HALO) * (bv.tile[2] + 2 * GT_HALO);
    const size_t smem = (size_t)6 * bvol * sizeof(double);
    constexpr size_t static_smem = sizeof(int) * (GT_LOCAL_STRAYS + 2);       // s_stray + s_nstray
    if (smem + static_smem > 227 * 1024) return fail("pic_gather_push: supercell too large for shared memory (%zu B)", smem);
    const int ntiles = bv.nt[0] * bv.nt[1] * bv.nt[2];
    const bool t888 = bv.tile[0] == 8 && bv.tile[1] == 8 && bv.tile[2] == 8;
    // stray list (stream-ordered scratch): a few per mille of the particles per step since the sort
    StrayList stray;
    stray.cap = (int)(bv.np_limit / 8 + 1024);
    int* scratch = nullptr;
    if (cudaMallocAsync((void**)&scratch, sizeof(int) * (size_t)(stray.cap + 1), s) != cudaSuccess)
        return fail("pic_gather_push: cannot allocate %ld B of scratch", (long)(sizeof(int) * (stray.cap + 1)));
    stray.count = scratch; stray.idx = scratch + 1;
    cudaMemsetAsync(stray.count, 0, sizeof(int), s);
#define PIC_LAUNCH(YEE_, TX_, TY_, TZ_) do { \
        auto k = gather_push_tile_kernel<N, G, YEE_, TX_, TY_, TZ_>; \
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024 - static_smem - 1024)) != cudaSuccess) \
            return fail("pic_gather_push: cannot raise the dynamic shared memory limit"); \
        k<<<ntiles, GT_THREADS, smem, s>>>(P, bv, gf, gg, qdt2m, dt, pusher, push_position, esc, stray); } while (0)
    if (yee && t888) PIC_LAUNCH(true, 8, 8, 8);       // the tuned instance: immediate LDS offsets
    else if (yee) PIC_LAUNCH(true, 0, 0, 0);
    else PIC_LAUNCH(false, 0, 0, 0);
#undef PIC_LAUNCH
    if (yee) gather_push_listed_kernel<N, G, true><<<NUM_SMS * 8, 128, 0, s>>>(P, stray, gf, gg, qdt2m, dt, pusher, push_position, esc);
    else gather_push_listed_kernel<N, G, false><<<NUM_SMS * 8, 128, 0, s>>>(P, stray, gf, gg, qdt2m, dt, pusher, push_position, esc);
    count_launch(2);
    cudaFreeAsync(scratch, s);
    return check_launch("pic_gather_push(tile)") ? 0 : 1;
}

int gather_push_tile_launch(const pic_soa* p, long offset, long np, const pic_fab E[3],
                            const pic_fab B[3], const GatherGeom& gg, double qdt2m, double dt,
                            int nox, int galerkin, int pusher, int push_position,
                            const pic_bins* bins, const EscapeView& esc, cudaStream_t s) {
    PIC_REQUIRE(offset == 0 && np == p->np, "pic_gather_push: bins describe the whole tile (offset 0, np = all)");
    BinsView bv = make_bins(*bins);
    bv.np_limit = (int)std::min<long>(bins->np_binned, np);   // the tile may have shrunk since the sort
    GlobalFields gf;
    for (int c = 0; c < 3; ++c) { gf.v[c] = make_view(E[c]); gf.v[3 + c] = make_view(B[c]); }
    SoaView P = make_soa(*p, 0);
    const bool yee = is_yee(E, B);
#define PIC_GT(N, G) return launch<N, G>(P, bv, gf, gg, yee, qdt2m, dt, pusher, push_position, esc, s)
    if (nox == 1 && galerkin) PIC_GT(1, 1);
    if (nox == 1) PIC_GT(1, 0);
    if (nox == 2 && galerkin) PIC_GT(2, 1);
    if (nox == 2) PIC_GT(2, 0);
    if (nox == 3 && galerkin) PIC_GT(3, 1);
    PIC_GT(3, 0);
#undef PIC_GT
}

}  // namespace pic
