// Yee FDTD with bulk-asynchronous (TMA engine) staging of the source field into shared memory.
//
// Same arithmetic as fdtd.cu (EvolveB.cpp:168-185 with the Yee stencil CartesianYeeAlgorithm.H:69-101; EvolveE.cpp:
// 185-213, whose downward differences are 2-point for Yee AND CKC), different data movement: a CTA owns TJ rows (j)
// of full length along x and marches through KC planes (k).  Per plane the three source components arrive as ONE
// 1-D bulk copy each (cp.async.bulk global -> shared, completion on an mbarrier): in an AMReX-shaped array the rows
// j0 .. j0+TJ of a plane are one contiguous chunk of (TJ+1) * row-length doubles.  A tensor map cannot describe these
// arrays (odd row lengths: 265 doubles, TMA tiles need 16-byte-multiple strides) but a 1-D bulk copy needs only a
// 16-byte-aligned start and size: the chunk is widened to the enclosing even element range.  Three plane slots form
// a ring: plane k and k+1 (EvolveB) / k-1 and k (EvolveE) are resident while k+2 / k+1 is in flight, so every source
// value is read from HBM/L2 once per CTA and all stencil neighbours come from shared memory; the updated component is
// a plain coalesced read-modify-write.  No thread issues a global load for the source field.
#include "pic_common.cuh"

namespace pic {

struct BulkCoefs { double cx, cy, cz; };
struct BulkBox { int lo[3]; int n[3]; };   // points visited: cells + the upper nodal layer (as fdtd.cu PointBox)

constexpr int FB_SLOTS = 3;
constexpr int FB_TX = 64, FB_TY = 4;       // 256 threads: 64 lanes along x, 4 rows
constexpr int FB_IT = 5;                   // points of a row per thread: rows up to FB_TX * FB_IT = 320 points

#ifdef PIC_SIMT_HOST      // tests/host_harness: the emulator copies synchronously, barriers are always complete
__device__ __forceinline__ void fb_mbar_init(unsigned long long*, int) {}
__device__ __forceinline__ void fb_fence_init() {}
__device__ __forceinline__ void fb_expect(unsigned long long*, unsigned) {}
__device__ __forceinline__ void fb_bulk(double* dst, const double* src, unsigned bytes, unsigned long long*) {
    for (unsigned n = 0; n < bytes / 8; ++n) dst[n] = src[n];
}
__device__ __forceinline__ void fb_wait(unsigned long long*, unsigned) {}
#else
__device__ __forceinline__ unsigned fb_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fb_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(fb_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fb_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fb_expect(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(fb_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fb_bulk(double* dst, const double* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(fb_smem(dst)), "l"(src), "r"(bytes), "r"(fb_smem(bar)) : "memory");
}
__device__ __forceinline__ void fb_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    // bounded: a copy that never completes (a bug) must end the kernel with an error, not hang the device
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(fb_smem(bar)), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();
    }
}
#endif

// One source component in the ring: where its chunk of plane p lands and how to index it.
struct BulkSrc {
    FabView F;
    long total;          // elements of the array
    __device__ __forceinline__ bool has_plane(int k) const { return k >= F.lo2 && k < F.lo2 + F.n2; }
};

// Issues the copy of rows [j_first, j_first + rows) of plane k (clamped to the array) into buf and returns the bytes the
// mbarrier will see; *shift receives the position of element (lo0, j_first, k) inside buf.  Called by ONE thread.
__device__ __forceinline__ unsigned fb_issue(const BulkSrc& S, int j_first, int rows, int k, double* buf, int* shift,
                                             unsigned long long* bar) {
    int j0 = j_first, j1 = j_first + rows;                     // clamp the rows to the allocation
    if (j0 < S.F.lo1) j0 = S.F.lo1;
    if (j1 > S.F.lo1 + S.F.n1) j1 = S.F.lo1 + S.F.n1;
    if (j1 <= j0 || !S.has_plane(k)) { *shift = 0; return 0; }
    const long eb = S.F.off(S.F.lo0, j0, k), ee = eb + (long)(j1 - j0) * S.F.sj;
    const long ab = eb & ~1L;                                  // 16-byte aligned element range [ab, ae)
    long ae = (ee + 1) & ~1L;
    // element (lo0, j_first, k) sits (eb - ab) - (j0 - j_first) * sj elements into the buffer (rows below the array
    // are never read)
    *shift = (int)(eb - ab) - (j0 - j_first) * (int)S.F.sj;
    if (ae > S.total) {                                        // the array ends on an odd element: last one by hand
        ae = ee - 1;
        buf[(ee - 1) - ab] = S.F.p[ee - 1];
    }
    const unsigned bytes = (unsigned)((ae - ab) * 8);
    if (bytes) fb_bulk(buf, S.F.p + ab, bytes, bar);
    return bytes;
}

// ------------------------------------------------------------------------------------------------------------
// EvolveB (Yee): B += dt * curl-ish of E.  Resident: E planes k and k+1, rows j0 .. j0+TJ.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FB_TX * FB_TY)
evolve_b_bulk_kernel(FabView Bx, FabView By, FabView Bz, BulkSrc Ex, BulkSrc Ey, BulkSrc Ez, BulkCoefs cf, BulkBox pb,
                     double dt, int KC, int chunk /* doubles per component slot */) {
    PIC_DYNAMIC_SMEM(double, smem);
    // [slot][component][chunk] doubles, then FB_SLOTS mbarriers, then [slot][component] shifts
    double* ring = smem;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + (size_t)FB_SLOTS * 3 * chunk);
    int* shifts = reinterpret_cast<int*>(bar + FB_SLOTS);
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * FB_TX + tx;
    const int lj0 = blockIdx.x * FB_TY;                      // first row of this CTA (box-local)
    const int lk0 = blockIdx.y * KC, lk1 = min(lk0 + KC, pb.n[2]);
    const int j_first = pb.lo[1] + lj0;
    const int rows = FB_TY + 1;                              // + the j+1 neighbour row
    if (tid == 0) {
        for (int s = 0; s < FB_SLOTS; ++s) fb_mbar_init(bar + s, 1);
        fb_fence_init();
    }
    __syncthreads();
    auto issue_plane = [&](int lk) {                         // thread 0: plane lk of all three components into slot lk % 3
        const int slot = lk % FB_SLOTS, k = pb.lo[2] + lk;
        double* b0 = ring + (size_t)slot * 3 * chunk;
        // arm first (the byte count is known before the copies are issued: same clamps as fb_issue)
        unsigned bytes = 0;
        {
            const BulkSrc* S[3] = {&Ex, &Ey, &Ez};
            for (int c = 0; c < 3; ++c) {
                int j0 = j_first, j1 = j_first + rows;
                if (j0 < S[c]->F.lo1) j0 = S[c]->F.lo1;
                if (j1 > S[c]->F.lo1 + S[c]->F.n1) j1 = S[c]->F.lo1 + S[c]->F.n1;
                if (j1 <= j0 || !S[c]->has_plane(k)) continue;
                const long eb = S[c]->F.off(S[c]->F.lo0, j0, k), ee = eb + (long)(j1 - j0) * S[c]->F.sj;
                long ae = (ee + 1) & ~1L;
                if (ae > S[c]->total) ae = ee - 1;
                bytes += (unsigned)((ae - (eb & ~1L)) * 8);
            }
        }
        fb_expect(bar + slot, bytes);
        fb_issue(Ex, j_first, rows, k, b0, shifts + slot * 3 + 0, bar + slot);
        fb_issue(Ey, j_first, rows, k, b0 + chunk, shifts + slot * 3 + 1, bar + slot);
        fb_issue(Ez, j_first, rows, k, b0 + 2 * chunk, shifts + slot * 3 + 2, bar + slot);
    };
    const int lk_last = min(lk1, pb.n[2] - 1);               // highest plane any point of this CTA reads (k+1 of the last in_z point)
    if (tid == 0) {
        issue_plane(lk0);
        if (lk0 + 1 <= lk_last) issue_plane(lk0 + 1);
    }
    const int lj = lj0 + ty;
    const bool row_ok = lj < pb.n[1];
    const bool in_y = lj < pb.n[1] - 1;
    const int j = pb.lo[1] + lj;
    // the read-modify-write values of this thread for one plane (FB_IT points of its row, three components); the
    // loads of plane lk+1 are issued before plane lk is computed (and before the wait for its source planes)
    double nbx[FB_IT], nby[FB_IT], nbz[FB_IT];
    auto load_b = [&](int lk) {
        const int k = pb.lo[2] + lk;
        const bool in_z = lk < pb.n[2] - 1;
#pragma unroll
        for (int t = 0; t < FB_IT; ++t) {
            const int li = tx + t * FB_TX, i = pb.lo[0] + li;
            const bool in = row_ok && lk < lk1 && li < pb.n[0], in_x = li < pb.n[0] - 1;
            nbx[t] = (in && in_y && in_z) ? Bx(i, j, k) : 0.0;
            nby[t] = (in && in_x && in_z) ? By(i, j, k) : 0.0;
            nbz[t] = (in && in_x && in_y) ? Bz(i, j, k) : 0.0;
        }
    };
    load_b(lk0);
    for (int lk = lk0; lk < lk1; ++lk) {
        if (tid == 0 && lk + 2 <= lk_last) issue_plane(lk + 2);     // slot (lk+2)%3 was released by the barrier below
        double vbx[FB_IT], vby[FB_IT], vbz[FB_IT];
#pragma unroll
        for (int t = 0; t < FB_IT; ++t) { vbx[t] = nbx[t]; vby[t] = nby[t]; vbz[t] = nbz[t]; }
        load_b(lk + 1);
        const int s0 = lk % FB_SLOTS, s1 = (lk + 1) % FB_SLOTS;
        fb_wait(bar + s0, ((lk - lk0) / FB_SLOTS) & 1);
        const bool in_z = lk < pb.n[2] - 1;
        if (lk + 1 <= lk_last) fb_wait(bar + s1, ((lk + 1 - lk0) / FB_SLOTS) & 1);
        __syncthreads();                                     // shifts + the hand-copied tail element are visible
        const double* e0 = ring + (size_t)s0 * 3 * chunk;
        const double* e1 = ring + (size_t)s1 * 3 * chunk;
        // pointers to element (lo0, j, k) of the row this thread works on
        const double* ex0 = e0 + shifts[s0 * 3 + 0] + ty * Ex.F.sj;
        const double* ey0 = e0 + chunk + shifts[s0 * 3 + 1] + ty * Ey.F.sj;
        const double* ez0 = e0 + 2 * chunk + shifts[s0 * 3 + 2] + ty * Ez.F.sj;
        const double* ex1 = e1 + shifts[s1 * 3 + 0] + ty * Ex.F.sj;
        const double* ey1 = e1 + chunk + shifts[s1 * 3 + 1] + ty * Ey.F.sj;
        const int k = pb.lo[2] + lk;
        if (row_ok) {
#pragma unroll
            for (int t = 0; t < FB_IT; ++t) {
                const int li = tx + t * FB_TX, i = pb.lo[0] + li;
                const bool in = li < pb.n[0], in_x = li < pb.n[0] - 1;
                const int ax = i - Ex.F.lo0, ay = i - Ey.F.lo0, az = i - Ez.F.lo0;
                if (in && in_y && in_z) {   // Bx(1,0,0)  EvolveB.cpp:168-171
                    Bx(i, j, k) = vbx[t] + (dt * (cf.cz * (ey1[ay] - ey0[ay])) - dt * (cf.cy * (ez0[az + Ez.F.sj] - ez0[az])));
                }
                if (in && in_x && in_z) {   // By(0,1,0)  :175-178
                    By(i, j, k) = vby[t] + (dt * (cf.cx * (ez0[az + 1] - ez0[az])) - dt * (cf.cz * (ex1[ax] - ex0[ax])));
                }
                if (in && in_x && in_y) {   // Bz(0,0,1)  :182-185
                    Bz(i, j, k) = vbz[t] + (dt * (cf.cy * (ex0[ax + Ex.F.sj] - ex0[ax])) - dt * (cf.cx * (ey0[ay + 1] - ey0[ay])));
                }
            }
        }
        __syncthreads();                                     // every thread is done with slot s0: it may be refilled
    }
}

// ------------------------------------------------------------------------------------------------------------
// EvolveE: E += c^2 dt (curl-ish of B - mu0 J).  Resident: B planes k-1 and k, rows j0-1 .. j0+TJ-1.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FB_TX * FB_TY)
evolve_e_bulk_kernel(FabView Ex, FabView Ey, FabView Ez, BulkSrc Bx, BulkSrc By, BulkSrc Bz, FabView jx, FabView jy,
                     FabView jz, BulkCoefs cf, BulkBox pb, double dt, int KC, int chunk) {
    PIC_DYNAMIC_SMEM(double, smem);
    double* ring = smem;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + (size_t)FB_SLOTS * 3 * chunk);
    int* shifts = reinterpret_cast<int*>(bar + FB_SLOTS);
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * FB_TX + tx;
    const int lj0 = blockIdx.x * FB_TY;
    const int lk0 = blockIdx.y * KC, lk1 = min(lk0 + KC, pb.n[2]);
    const int j_first = pb.lo[1] + lj0 - 1;                  // the j-1 neighbour row first
    const int rows = FB_TY + 1;
    if (tid == 0) {
        for (int s = 0; s < FB_SLOTS; ++s) fb_mbar_init(bar + s, 1);
        fb_fence_init();
    }
    __syncthreads();
    // ring position of box-local plane lk (planes lk0-1 .. lk1-1 are used): q = lk - (lk0 - 1)
    auto issue_plane = [&](int lk) {
        const int q = lk - (lk0 - 1), slot = q % FB_SLOTS, k = pb.lo[2] + lk;
        double* b0 = ring + (size_t)slot * 3 * chunk;
        unsigned bytes = 0;
        {
            const BulkSrc* S[3] = {&Bx, &By, &Bz};
            for (int c = 0; c < 3; ++c) {
                int j0 = j_first, j1 = j_first + rows;
                if (j0 < S[c]->F.lo1) j0 = S[c]->F.lo1;
                if (j1 > S[c]->F.lo1 + S[c]->F.n1) j1 = S[c]->F.lo1 + S[c]->F.n1;
                if (j1 <= j0 || !S[c]->has_plane(k)) continue;
                const long eb = S[c]->F.off(S[c]->F.lo0, j0, k), ee = eb + (long)(j1 - j0) * S[c]->F.sj;
                long ae = (ee + 1) & ~1L;
                if (ae > S[c]->total) ae = ee - 1;
                bytes += (unsigned)((ae - (eb & ~1L)) * 8);
            }
        }
        fb_expect(bar + slot, bytes);
        fb_issue(Bx, j_first, rows, k, b0, shifts + slot * 3 + 0, bar + slot);
        fb_issue(By, j_first, rows, k, b0 + chunk, shifts + slot * 3 + 1, bar + slot);
        fb_issue(Bz, j_first, rows, k, b0 + 2 * chunk, shifts + slot * 3 + 2, bar + slot);
    };
    if (tid == 0) {
        issue_plane(lk0 - 1);
        issue_plane(lk0);
    }
    const int lj = lj0 + ty;
    const bool row_ok = lj < pb.n[1];
    const bool in_y = lj < pb.n[1] - 1;
    constexpr double c2 = C_LIGHT * C_LIGHT;
    for (int lk = lk0; lk < lk1; ++lk) {
        if (tid == 0 && lk + 1 < lk1) issue_plane(lk + 1);
        const int q0 = lk - lk0, q1 = q0 + 1;                // ring positions of planes lk-1 and lk
        const int s0 = q0 % FB_SLOTS, s1 = q1 % FB_SLOTS;
        fb_wait(bar + s0, (q0 / FB_SLOTS) & 1);
        fb_wait(bar + s1, (q1 / FB_SLOTS) & 1);
        __syncthreads();
        const double* m0 = ring + (size_t)s0 * 3 * chunk;     // plane k-1
        const double* m1 = ring + (size_t)s1 * 3 * chunk;     // plane k
        // row (ty + 1) of the chunk is row j; row ty is j-1
        const double* bxm = m0 + shifts[s0 * 3 + 0] + (ty + 1) * Bx.F.sj;
        const double* bym = m0 + chunk + shifts[s0 * 3 + 1] + (ty + 1) * By.F.sj;
        const double* bx = m1 + shifts[s1 * 3 + 0] + (ty + 1) * Bx.F.sj;
        const double* by = m1 + chunk + shifts[s1 * 3 + 1] + (ty + 1) * By.F.sj;
        const double* bz = m1 + 2 * chunk + shifts[s1 * 3 + 2] + (ty + 1) * Bz.F.sj;
        const bool in_z = lk < pb.n[2] - 1;
        const int k = pb.lo[2] + lk, j = pb.lo[1] + lj;
        if (row_ok) {
            double vex[FB_IT], vey[FB_IT], vez[FB_IT], vjx[FB_IT], vjy[FB_IT], vjz[FB_IT];
#pragma unroll
            for (int t = 0; t < FB_IT; ++t) {      // the six global loads of every point of this thread, all in flight
                const int li = tx + t * FB_TX, i = pb.lo[0] + li;
                const bool in = li < pb.n[0], in_x = li < pb.n[0] - 1;
                vex[t] = (in && in_x) ? Ex(i, j, k) : 0.0;  vjx[t] = (in && in_x) ? jx.ld(i, j, k) : 0.0;
                vey[t] = (in && in_y) ? Ey(i, j, k) : 0.0;  vjy[t] = (in && in_y) ? jy.ld(i, j, k) : 0.0;
                vez[t] = (in && in_z) ? Ez(i, j, k) : 0.0;  vjz[t] = (in && in_z) ? jz.ld(i, j, k) : 0.0;
            }
#pragma unroll
            for (int t = 0; t < FB_IT; ++t) {
                const int li = tx + t * FB_TX, i = pb.lo[0] + li;
                const bool in = li < pb.n[0], in_x = li < pb.n[0] - 1;
                const int ax = i - Bx.F.lo0, ay = i - By.F.lo0, az = i - Bz.F.lo0;
                if (in && in_x) {           // Ex(0,1,1)  EvolveE.cpp:185-188
                    Ex(i, j, k) = vex[t] + c2 * dt * (-(cf.cz * (by[ay] - bym[ay])) + cf.cy * (bz[az] - bz[az - Bz.F.sj])
                                                      - MU0 * vjx[t]);
                }
                if (in && in_y) {           // Ey(1,0,1)  :201-204
                    Ey(i, j, k) = vey[t] + c2 * dt * (-(cf.cx * (bz[az] - bz[az - 1])) + cf.cz * (bx[ax] - bxm[ax])
                                                      - MU0 * vjy[t]);
                }
                if (in && in_z) {           // Ez(1,1,0)  :210-213
                    Ez(i, j, k) = vez[t] + c2 * dt * (-(cf.cy * (bx[ax] - bx[ax - Bx.F.sj])) + cf.cx * (by[ay] - by[ay - 1])
                                                      - MU0 * vjz[t]);
                }
            }
        }
        __syncthreads();
    }
}

long g_fdtd_bulk_launches = 0;
int g_fdtd_bulk = 1;      // pic_set_fdtd_mode bits: 1 = EvolveB staged (default), 2 = EvolveE staged; 0 = plain loads

static BulkSrc bulk_src(const pic_fab& f) {
    BulkSrc s;
    s.F = make_view(f);
    s.total = fab_size(f);
    return s;
}

// shared-memory budget: the ring holds 3 slots x 3 components x (TJ+1) rows (+2 elements of alignment slack)
static bool bulk_plan(const pic_fab src[3], const int n[3], int* chunk, size_t* smem) {
    if (n[0] > FB_TX * FB_IT) return false;                                       // longer rows: the plain kernels
    long row = 0;
    for (int c = 0; c < 3; ++c) {
        const long r = src[c].hi[0] - src[c].lo[0] + 1;
        if (r > row) row = r;
        if ((reinterpret_cast<uintptr_t>(src[c].p) & 15) != 0) return false;      // bulk copies need 16-byte aligned bases
    }
    const long ch = ((FB_TY + 1) * row + 4 + 1) & ~1L;
    *chunk = (int)ch;
    *smem = sizeof(double) * (size_t)FB_SLOTS * 3 * ch + sizeof(unsigned long long) * FB_SLOTS + sizeof(int) * FB_SLOTS * 3 + 16;
    return *smem <= 110 * 1024;       // two CTAs per SM; longer rows take the plain kernel
}

int evolve_b_bulk_launch(const pic_fab B[3], const pic_fab E[3], const pic_stencil* st, const int lo[3], const int n[3],
                         double dt, cudaStream_t s, bool* done) {
    *done = false;
    int chunk; size_t smem;
    if (!(g_fdtd_bulk & 1) || st->algo != PIC_SOLVER_YEE || !bulk_plan(E, n, &chunk, &smem)) return 0;
    BulkBox pb;
    for (int d = 0; d < 3; ++d) { pb.lo[d] = lo[d]; pb.n[d] = n[d]; }
    BulkCoefs cf{st->cx[0], st->cy[0], st->cz[0]};
    const int KC = 16;
    dim3 block(FB_TX, FB_TY, 1);
    dim3 grid((n[1] + FB_TY - 1) / FB_TY, (n[2] + KC - 1) / KC, 1);
#ifndef PIC_SIMT_HOST
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(evolve_b_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); attr = true; }
#endif
    evolve_b_bulk_kernel<<<grid, block, smem, s>>>(make_view(B[0]), make_view(B[1]), make_view(B[2]), bulk_src(E[0]),
                                                   bulk_src(E[1]), bulk_src(E[2]), cf, pb, dt, KC, chunk);
    count_launch();
    ++g_fdtd_bulk_launches;
    *done = true;
    return check_launch("pic_evolve_b(bulk)") ? 0 : 1;
}

int evolve_e_bulk_launch(const pic_fab E[3], const pic_fab B[3], const pic_fab J[3], const pic_stencil* st, const int lo[3],
                         const int n[3], double dt, cudaStream_t s, bool* done) {
    *done = false;
    int chunk; size_t smem;
    if (!(g_fdtd_bulk & 2) || !bulk_plan(B, n, &chunk, &smem)) return 0;
    BulkBox pb;
    for (int d = 0; d < 3; ++d) { pb.lo[d] = lo[d]; pb.n[d] = n[d]; }
    BulkCoefs cf{st->cx[0], st->cy[0], st->cz[0]};
    const int KC = 16;
    dim3 block(FB_TX, FB_TY, 1);
    dim3 grid((n[1] + FB_TY - 1) / FB_TY, (n[2] + KC - 1) / KC, 1);
#ifndef PIC_SIMT_HOST
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(evolve_e_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024); attr = true; }
#endif
    evolve_e_bulk_kernel<<<grid, block, smem, s>>>(make_view(E[0]), make_view(E[1]), make_view(E[2]), bulk_src(B[0]),
                                                   bulk_src(B[1]), bulk_src(B[2]), make_view(J[0]), make_view(J[1]),
                                                   make_view(J[2]), cf, pb, dt, KC, chunk);
    count_launch();
    ++g_fdtd_bulk_launches;
    *done = true;
    return check_launch("pic_evolve_e(bulk)") ? 0 : 1;
}

}  // namespace pic

extern "C" void pic_set_fdtd_mode(int mode) { pic::g_fdtd_bulk = mode & 3; }
extern "C" long pic_fdtd_bulk_launches(void) { return pic::g_fdtd_bulk_launches; }
