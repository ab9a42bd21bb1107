// Esirkepov deposition for cell-sorted particles, one LANE PER CELL (pic_set_deposit_mode(PIC_DEPOSIT_CELLS);
// replaces doEsirkepovDepositionShapeN, Source/Particles/Deposition/CurrentDeposition.H:642-907).
//
// For a particle whose old and new position lie in the cell of its bin ("quiet", > 95 % of a thermal plasma)
// the stencil is anchored at the cell: ALL particles of a cell add into the same (N+1)^2 x N nodes per
// component,
//     Jx[i][j][k] += cx[i] * (Sy_new[j] Az[k] + Sy_old[j] Bz[k]),   cx = prefix sums of wq/(dt dy dz)(Sx_old - Sx_new),
//     Az = Sz_new/3 + Sz_old/6,  Bz = Sz_old/3 + Sz_new/6            (cyclically for Jy, Jz).
// deposit_runs.cu gives every stencil LINE a lane and walks the particles: each lane re-reads the particle's
// record from shared memory (17 doubles for 15 FMA) and the kernel is bound by the shared-memory return path
// (profiles/r1k_ncu_full_256cube.txt: L1 92 %, fp64 32 %).  Here a lane owns a CELL and one component: its
// (N+1)^2 x N partial sums (48 at order 3) stay in registers while it walks the particles of its cell, and a
// particle costs it 19 doubles of shared-memory reads for 80 FMA -- the fp64 pipe becomes the bound.
//
// One CTA per supercell (8 x 8 x 8 cells), four warps with fixed roles (rotated by block so that the heavier
// role does not always land on the same SM sub-partition):
//   producer   lane = cell; slice s = the s-th particle of every cell of a group of 32 cells (4 rows along
//              x): positions, shape factors, prefix sums -> one record per lane in shared memory (double
//              buffered); particles that are not quiet in their bin's cell go to a list;
//   consumer c (c = x, y, z)   same lane = cell mapping; reads the record of its lane and accumulates component c.
// After the last slice of a group a consumer adds its sums into the CTA's shared-memory J block of its
// component with plain read-modify-writes: it is the only writer of that block, and inside one instruction the
// lanes (distinct cells, same stencil offset) touch distinct nodes.  At the end the block goes to J with one
// fp64 red.global per touched node (7 per cell instead of the 40 of deposit_runs.cu, 540 per particle in
// the reference).  The listed particles take deposit_general_kernel (deposit_runs.cu).
// Orders 1 and 3 (the stencil of every particle of a cell starts at the same node; not so at order 2).
#include "pic_common.cuh"
#include "deposit_common.cuh"
#include "bins.cuh"
#ifdef PIC_SIMT_HOST
#include <vector>
#endif

namespace pic {

constexpr unsigned DC_FULL = 0xffffffffu;

template <int N> struct CellsCfg {
    static_assert(N == 1 || N == 3, "one lane per cell: orders 1 and 3");
    static constexpr int T = 8;                   // supercell edge (pic_bins tile)
    static constexpr int QS = N + 1;              // stencil nodes per direction
    static constexpr int QP = N;                  // live prefix entries
    static constexpr int O0 = -((N - 1) / 2);     // first stencil node relative to the cell
    static constexpr int PX = 12, PY = T + QS - 1, PZ = T + QS - 1;   // J block: pitch 12 keeps rows r, r+2 in disjoint banks
    static constexpr int TS = PX * PY * PZ;       // doubles per component
    static constexpr int NCP = (QP + 1) / 2;      // double2 elements of one component's prefix sums
    // record (double2 per lane): (Sx_new,Sx_old)[QS], (Sy_new,Sy_old)[QS], (Ay,By)[QS], (Az,Bz)[QS], cds x/y/z [NCP] each
    static constexpr int F_SX = 0, F_SY = QS, F_ABY = 2 * QS, F_ABZ = 3 * QS, F_CDS = 4 * QS;
    static constexpr int NF = 4 * QS + 3 * NCP;
    static constexpr size_t smem_bytes = sizeof(double2) * 2 * NF * 32 + sizeof(double) * 3 * TS;
};

template <int N, int MINB>
__global__ void __launch_bounds__(128, MINB)
deposit_cells_kernel(SoaView P, long np_lim, BinsView bins, J3 Jp, DepositGeom dg,
                     int* __restrict__ list, int* __restrict__ list_count) {
    using T = CellsCfg<N>;
    constexpr int QS = T::QS, QP = T::QP, NF = T::NF, NCP = T::NCP, PX = T::PX, PY = T::PY, TS = T::TS;
    PIC_DYNAMIC_SMEM(double2, smem2);
    double2* rec = smem2;                                            // [2][NF][32]
    double* tile = reinterpret_cast<double*>(smem2 + 2 * NF * 32);   // [3][TS]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int role = (warp + blockIdx.x) & 3;                        // 0 producer, 1..3 consumer of component role-1
    const bool producer = role == 0;
    const int comp = role - 1;

    // supercell of this CTA
    int tc[3];
    tile_coords(bins, blockIdx.x, tc);
    const long bin0 = (long)blockIdx.x * (T::T * T::T * T::T);
    // lane -> cell of a group: 8 cells along x, rows 0,2,1,3 (rows r, r+2 share a half-warp: disjoint banks)
    const int lx = lane & 7, rsel = lane >> 3, lrow = ((rsel & 1) << 1) | (rsel >> 1);

    for (int n = threadIdx.x; n < 3 * TS; n += 128) tile[n] = 0.0;
    for (int n = threadIdx.x; n < 2 * NF * 32; n += 128) rec[n] = make_double2(0.0, 0.0);

    // ---- group walk shared by all roles: (g, s) = slice s of group g; maxn = largest cell of the group ----
    struct Walk { int g, s, maxn, p0, n; bool done; };
    auto load_group = [&](Walk& w) {
        // next group with particles
        while (true) {
            ++w.g;
            if (w.g >= 16) { w.done = true; return; }
            const int ly = ((w.g & 1) << 2) + lrow, lz = w.g >> 1;
            const long b = bin0 + lx + T::T * (ly + T::T * lz);
            int p0 = bins.cell_start[b], p1 = bins.cell_start[b + 1];
            p0 = (int)min((long)p0, np_lim); p1 = (int)min((long)p1, np_lim);
            w.p0 = p0; w.n = p1 - p0;
            int m = w.n;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) m = max(m, __shfl_xor_sync(DC_FULL, m, o));
            w.maxn = m;
            w.s = 0;
            if (m > 0) return;
        }
    };
    auto advance = [&](Walk& w) {
        if (++w.s >= w.maxn) load_group(w);
    };

    __syncthreads();

    if (producer) {
        // ======================= producer =======================
        Walk w{-1, 0, 0, 0, 0, false};
        load_group(w);
        // expected leftmost index of the new position's stencil for a particle of the lane's cell
        const int off[3] = {bins.box_lo[0] - dg.lo[0] + T::O0, bins.box_lo[1] - dg.lo[1] + T::O0, bins.box_lo[2] - dg.lo[2] + T::O0};
        double pf[7] = {0, 0, 0, 0, 0, 0, 0};
        int it = 0;
        auto produce = [&](double2* buf) {
            const bool valid = w.s < w.n;
            const long ip = (long)w.p0 + w.s;
            double xp, yp, zp, wp, uxp, uyp, uzp;
            if (w.s == 0) {
                if (valid) { pf[0] = P.x[ip]; pf[1] = P.y[ip]; pf[2] = P.z[ip]; pf[3] = P.w[ip]; pf[4] = P.ux[ip]; pf[5] = P.uy[ip]; pf[6] = P.uz[ip]; }
            }
            xp = pf[0]; yp = pf[1]; zp = pf[2]; wp = pf[3]; uxp = pf[4]; uyp = pf[5]; uzp = pf[6];
            if (w.s + 1 < w.n) {       // request the next slice of this cell
                pf[0] = P.x[ip + 1]; pf[1] = P.y[ip + 1]; pf[2] = P.z[ip + 1]; pf[3] = P.w[ip + 1];
                pf[4] = P.ux[ip + 1]; pf[5] = P.uy[ip + 1]; pf[6] = P.uz[ip + 1];
            }
            bool ok = false;
            double2* r = buf + lane;
            if (valid) {
                const ParticleGeom pg = particle_geom(xp, yp, zp, wp, uxp, uyp, uzp, dg);
                double wn[3][N + 1], wo[3][N + 1];
                int inew[3], sh[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) inew[d] = dr_dir<N>(pg.pos_new[d], pg.pos_old[d], wn[d], wo[d], sh[d]);
                const int ly = ((w.g & 1) << 2) + lrow, lz = w.g >> 1;
                const int ex = tc[0] * T::T + lx + off[0], ey = tc[1] * T::T + ly + off[1], ez = tc[2] * T::T + lz + off[2];
                ok = (sh[0] | sh[1] | sh[2]) == 0 && inew[0] == ex && inew[1] == ey && inew[2] == ez;
                if (ok) {
#pragma unroll
                    for (int s = 0; s < QS; ++s) {
                        r[(T::F_SX + s) * 32] = make_double2(wn[0][s], wo[0][s]);
                        r[(T::F_SY + s) * 32] = make_double2(wn[1][s], wo[1][s]);
                        r[(T::F_ABY + s) * 32] = make_double2((1.0 / 3.0) * wn[1][s] + (1.0 / 6.0) * wo[1][s],
                                                              (1.0 / 3.0) * wo[1][s] + (1.0 / 6.0) * wn[1][s]);
                        r[(T::F_ABZ + s) * 32] = make_double2((1.0 / 3.0) * wn[2][s] + (1.0 / 6.0) * wo[2][s],
                                                              (1.0 / 3.0) * wo[2][s] + (1.0 / 6.0) * wn[2][s]);
                    }
                    // prefix sums over the first N nodes (the sum over all N+1 vanishes and is not deposited:
                    // loop trimming of CurrentDeposition.H:777-788 with dl = du = 1)
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const double wqd = pg.wq * dg.invdtd[d];
                        double cds[2 * NCP];
                        cds[2 * NCP - 1] = 0.0;
                        double run = 0.0;
#pragma unroll
                        for (int i = 0; i < QP; ++i) {
                            run += wqd * (wo[d][i] - wn[d][i]);
                            cds[i] = run;
                        }
#pragma unroll
                        for (int m = 0; m < NCP; ++m) r[(T::F_CDS + d * NCP + m) * 32] = make_double2(cds[2 * m], cds[2 * m + 1]);
                    }
                }
            }
            if (!ok) {     // nothing to add from this lane: zero prefix sums (the weights left in the record are finite)
#pragma unroll
                for (int m = 0; m < 3 * NCP; ++m) r[(T::F_CDS + m) * 32] = make_double2(0.0, 0.0);
            }
            // particles of the bin that are not quiet in the bin's cell: list for the general kernel
            const bool listed = valid && !ok;
            const unsigned mm = __ballot_sync(DC_FULL, listed);
            if (mm) {
                int basei = 0;
                if (lane == 0) basei = atomicAdd(list_count, __popc(mm));
                basei = __shfl_sync(DC_FULL, basei, 0);
                if (listed) list[basei + __popc(mm & ((1u << lane) - 1u))] = (int)ip;
            }
        };
        if (!w.done) produce(rec);
        __syncthreads();
        while (!w.done) {           // the consumers process slice `it` while slice it+1 is produced
            advance(w);
            if (!w.done) produce(rec + (size_t)((it + 1) & 1) * NF * 32);
            __syncthreads();
            ++it;
        }
    } else {
        // ======================= consumer of component comp =======================
        Walk w{-1, 0, 0, 0, 0, false};
        load_group(w);
        // component roles: acc[a][b][p] += cds[p] * (A[a].x B[b].x + A[a].y B[b].y)
        //   Jx: a = y (Sy),  b = z (ABz), p = x     Jy: a = x (Sx), b = z (ABz), p = y     Jz: a = x (Sx), b = y (ABy), p = z
        const int fA = (comp == 0) ? T::F_SY : T::F_SX;
        const int fB = (comp == 2) ? T::F_ABY : T::F_ABZ;
        const int fC = T::F_CDS + comp * NCP;
        const int sa = (comp == 0) ? PX : 1;
        const int sb = (comp == 2) ? PX : PX * PY;
        const int sp = (comp == 0) ? 1 : ((comp == 1) ? PX : PX * PY);
        double* tl = tile + (size_t)comp * TS;
        double acc[QS][QS][QP];
#pragma unroll
        for (int a = 0; a < QS; ++a)
#pragma unroll
            for (int b = 0; b < QS; ++b)
#pragma unroll
                for (int p = 0; p < QP; ++p) acc[a][b][p] = 0.0;
        int it = 0;
        __syncthreads();             // slice 0 is in the record
        while (!w.done) {
            {
                const double2* r = rec + (size_t)(it & 1) * NF * 32 + lane;
                double cds[2 * NCP];
#pragma unroll
                for (int m = 0; m < NCP; ++m) {
                    const double2 c2 = r[(fC + m) * 32];
                    cds[2 * m] = c2.x; cds[2 * m + 1] = c2.y;
                }
                double2 B[QS];
#pragma unroll
                for (int b = 0; b < QS; ++b) B[b] = r[(fB + b) * 32];
#pragma unroll
                for (int a = 0; a < QS; ++a) {
                    const double2 A = r[(fA + a) * 32];
#pragma unroll
                    for (int b = 0; b < QS; ++b) {
                        const double wab = A.x * B[b].x + A.y * B[b].y;
#pragma unroll
                        for (int p = 0; p < QP; ++p) acc[a][b][p] += cds[p] * wab;
                    }
                }
            }
            if (w.s == w.maxn - 1) {
                // ---- last slice of the group: add the lane's sums into the CTA's J block of this component.
                // Lanes are distinct cells, so one instruction (fixed stencil offset) touches distinct nodes; two
                // offsets that differ along z never meet (the cells of a group share z), the others are ordered
                // by __syncwarp.
                const int ly = ((w.g & 1) << 2) + lrow, lz = w.g >> 1;
                double* base = tl + lx + PX * (ly + PY * lz);
                if (comp == 2) {                 // z is the prefix direction
#pragma unroll
                    for (int a = 0; a < QS; ++a)
#pragma unroll
                        for (int b = 0; b < QS; ++b) {
                            double* q = base + a * sa + b * sb;
                            double v[QP];
#pragma unroll
                            for (int p = 0; p < QP; ++p) v[p] = q[p * sp];
#pragma unroll
                            for (int p = 0; p < QP; ++p) { q[p * sp] = v[p] + acc[a][b][p]; acc[a][b][p] = 0.0; }
                            __syncwarp();
                        }
                } else {                         // z is line direction b
#pragma unroll
                    for (int a = 0; a < QS; ++a)
#pragma unroll
                        for (int p = 0; p < QP; ++p) {
                            double* q = base + a * sa + p * sp;
                            double v[QS];
#pragma unroll
                            for (int b = 0; b < QS; ++b) v[b] = q[b * sb];
#pragma unroll
                            for (int b = 0; b < QS; ++b) { q[b * sb] = v[b] + acc[a][b][p]; acc[a][b][p] = 0.0; }
                            __syncwarp();
                        }
                }
            }
            advance(w);
            __syncthreads();
            ++it;
        }
    }

    // ---- the J block goes to the arrays: one reduction per touched node ----
    __syncthreads();
    const int org[3] = {bins.box_lo[0] + tc[0] * T::T + T::O0, bins.box_lo[1] + tc[1] * T::T + T::O0,
                        bins.box_lo[2] + tc[2] * T::T + T::O0};
    for (int n = threadIdx.x; n < 3 * TS; n += 128) {
        const double v = tile[n];
        if (v == 0.0) continue;
        const int c = n / TS, r = n - c * TS;
        const int tz = r / (PX * PY), r2 = r - tz * (PX * PY), ty = r2 / PX, tx = r2 - ty * PX;
        const FabView& F = Jp.v[c];
        const int gx = org[0] + tx, gy = org[1] + ty, gz = org[2] + tz;
        if (gx < F.lo0 || gy < F.lo1 || gz < F.lo2 || gx >= F.lo0 + F.n0 || gy >= F.lo1 + F.n1 || gz >= F.lo2 + F.n2) continue;
        atomicAdd(F.p + F.off(gx, gy, gz), v);
    }
}

// deposit_runs.cu: the listed particles (full stencil)
int deposit_general_launch(SoaView P, const int* list, const int* list_count, const pic_fab J[3], const DepositGeom& dg,
                           int nox, cudaStream_t s);

template <int N>
static int launch_cells(SoaView P, long np, const pic_fab J[3], const DepositGeom& dg, const pic_bins* pb, cudaStream_t s) {
    using T = CellsCfg<N>;
    BinsView bins = make_bins(*pb);
    if (bins.tile[0] != T::T || bins.tile[1] != T::T || bins.tile[2] != T::T)
        return fail("pic_deposit_esirkepov(cells): the bins must use 8 x 8 x 8 supercells");
    const long np_lim = np < (long)bins.np_limit ? np : (long)bins.np_limit;
    const unsigned grid = (unsigned)(bins.nt[0] * bins.nt[1] * bins.nt[2]);
    J3 j3; j3.v[0] = make_view(J[0]); j3.v[1] = make_view(J[1]); j3.v[2] = make_view(J[2]);
    constexpr int MINB = 3;
    auto k = deposit_cells_kernel<N, MINB>;
#ifndef PIC_SIMT_HOST
    int* scratch = nullptr;
    if (cudaMallocAsync((void**)&scratch, sizeof(int) * (size_t)(np + 1), s) != cudaSuccess)
        return fail("pic_deposit_esirkepov: cannot allocate %ld B of scratch", (long)(sizeof(int) * (np + 1)));
    cudaMemsetAsync(scratch, 0, sizeof(int), s);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::smem_bytes);
        attr_done = true;
    }
    k<<<grid, 128, T::smem_bytes, s>>>(P, np_lim, bins, j3, dg, scratch + 1, scratch);
    count_launch();
    int rc = check_launch("pic_deposit_esirkepov(cells)") ? 0 : 1;
    if (!rc) rc = deposit_general_launch(P, scratch + 1, scratch, J, dg, N, s);
    cudaFreeAsync(scratch, s);
    return rc;
#else
    std::vector<int> scratch_h((size_t)np + 1, 0);
    int* scratch = scratch_h.data();
    k<<<grid, 128, T::smem_bytes, s>>>(P, np_lim, bins, j3, dg, scratch + 1, scratch);
    return deposit_general_launch(P, scratch + 1, scratch, J, dg, N, s);
#endif
}

// particles [0, min(np, np_binned)) through the cell kernel + the general kernel; the caller deposits the rest
int deposit_cells_launch(const pic_soa* p, long offset, long np, const pic_fab J[3], const DepositGeom& dg, int nox,
                         const pic_bins* bins, cudaStream_t s) {
    SoaView P = make_soa(*p, offset);
    if (nox == 1) return launch_cells<1>(P, np, J, dg, bins, s);
    return launch_cells<3>(P, np, J, dg, bins, s);
}

}  // namespace pic
