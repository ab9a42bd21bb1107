// C++ host driver of the PIC step: the sequence of WarpX::Evolve / OneStep_nosub for a single-level,
// periodic, explicit FDTD run on ONE rank, expressed as calls through the C ABI of this library.
//
// Mirrors (paths relative to /root/reference/Source):
//   pic_engine_evolve            <- WarpX::Evolve                          Evolve/WarpXEvolve.cpp:93-347
//   explicit_fill_boundary_eb    <- ExplicitFillBoundaryEBUpdateAux        :473-531
//   one_step_nosub               <- WarpX::OneStep_nosub                   :353-455
//   push_particles_and_deposit   <- MultiParticleContainer::Evolve         Particles/MultiParticleContainer.cpp:460-482
//                                   PhysicalParticleContainer::Evolve      Particles/PhysicalParticleContainer.cpp:1812-2095
//   sync_current                 <- WarpX::SyncCurrent / SumBoundaryJ      Parallelization/WarpXComm.cpp:1073-1240,1386-1424
//   synchronize                  <- WarpX::Synchronize                     Evolve/WarpXEvolve.cpp:64-91
//   handle_particles_at_boundaries <- WarpX::HandleParticlesAtBoundaries   :533-581
//   guard cells                  <- guardCellManager::Init                 Parallelization/GuardCellManager.cpp:62-161,310-343
//   dt                           <- WarpX::ComputeDt                       Evolve/WarpXComputeDt.cpp:56-95
// All memory is borrowed from the caller (fields, particle SoA double buffers, bins, scratch);
// every stage is an asynchronous launch on the caller's stream.  Multi-rank runs keep the same
// stage functions and interleave the neighbour exchanges from the host (warpx_b200/engine.py).
#include "pic_common.cuh"
#include <cmath>
#include <vector>

namespace pic {

struct Species {
    double q, m;
    pic_soa buf[2];        // counting sort permutes buf[cur] -> buf[1-cur]
    int cur;
    pic_bins bins;
    bool has_bins;
    void* sort_work;
    pic_escape_list esc;   // particles the last position push moved out of the domain (engine-owned)
    bool has_esc;
};

struct Engine {
    pic_geom geom;
    int box_lo[3], box_hi[3];
    int nox, galerkin, pusher, solver, sort_interval;
    double dx[3], dinv[3], dt;
    int ng_EB[3], ng_J[3], ng_FG[3], ng_FS[3];
    int use_filter = 0, npass[3] = {1, 1, 1};   // warpx.use_filter / filter_npass_each_dir
    double* filter_tmp = nullptr;               // one J-component-sized scratch array
    size_t filter_tmp_bytes = 0;
    pic_stencil st;
    pic_fab fab[9];        // Ex Ey Ez Bx By Bz jx jy jz
    std::vector<Species> species;
    bool is_synchronized = true;
    long istep = 0;
};

static void stencil_coefficients(int solver, const double dx[3], pic_stencil* st) {
    // FiniteDifferenceSolver ctor: Yee CartesianYeeAlgorithm.H:30-42, CKC CartesianCKCAlgorithm.H:31-101
    st->algo = solver;
    for (int n = 0; n < 5; ++n) st->cx[n] = st->cy[n] = st->cz[n] = 0.0;
    const double ix = 1.0 / dx[0], iy = 1.0 / dx[1], iz = 1.0 / dx[2];
    st->cx[0] = ix; st->cy[0] = iy; st->cz[0] = iz;
    if (solver != PIC_SOLVER_CKC) return;
    const double delta = fmax(ix, fmax(iy, iz));
    const double rx = (ix / delta) * (ix / delta), ry = (iy / delta) * (iy / delta), rz = (iz / delta) * (iz / delta);
    const double beta = 0.125 * (1.0 - rx * ry * rz / (ry * rz + rz * rx + rx * ry));
    const double irf = 1.0 / (ry * rz + rz * rx + rx * ry);
    const double gx = ry * rz * (0.0625 - 0.125 * ry * rz * irf);
    const double gy = rx * rz * (0.0625 - 0.125 * rx * rz * irf);
    const double gz = rx * ry * (0.0625 - 0.125 * rx * ry * irf);
    st->cx[1] = (1.0 - 2.0 * ry * beta - 2.0 * rz * beta - 4.0 * gx) * ix;
    st->cy[1] = (1.0 - 2.0 * rx * beta - 2.0 * rz * beta - 4.0 * gy) * iy;
    st->cz[1] = (1.0 - 2.0 * rx * beta - 2.0 * ry * beta - 4.0 * gz) * iz;
    st->cx[2] = ry * beta * ix; st->cx[3] = rz * beta * ix; st->cx[4] = gx * ix;
    st->cy[2] = rz * beta * iy; st->cy[3] = rx * beta * iy; st->cy[4] = gy * iy;
    st->cz[2] = rx * beta * iz; st->cz[3] = ry * beta * iz; st->cz[4] = gz * iz;
}

static double max_dt(int solver, const double dx[3]) {
    if (solver == PIC_SOLVER_YEE)
        return 1.0 / (sqrt(1.0 / (dx[0] * dx[0]) + 1.0 / (dx[1] * dx[1]) + 1.0 / (dx[2] * dx[2])) * C_LIGHT);
    return fmin(dx[0], fmin(dx[1], dx[2])) / C_LIGHT;
}

static void guard_cells(Engine& e) {
    for (int d = 0; d < 3; ++d) {
        const int ngt = e.nox;
        int ng = (ngt % 2) ? ngt + 1 : ngt;
        e.ng_J[d] = ngt + (int)ceil(C_LIGHT * 0.5 * e.dt / e.dx[d]);
        if (e.use_filter) e.ng_J[d] += e.npass[d];     // + stencil_length - 1, GuardCellManager.cpp:169-172
        e.ng_FS[d] = 1;
        ng = ng > e.ng_FS[d] ? ng : e.ng_FS[d];
        e.ng_EB[d] = ng;
        int fg = (e.nox + 1) / 2;
        fg = fg < ng ? fg : ng;
        e.ng_FG[d] = fg > e.ng_FS[d] ? fg : e.ng_FS[d];
    }
}

static void lower_corner(const Engine& e, const int ng[3], double xyzmin[3], int lo[3]) {
    // WarpX::LowerCorner of the box grown by ng (WarpX.cpp:2851-2874): prob_lo + index * dx
    for (int d = 0; d < 3; ++d) {
        lo[d] = e.box_lo[d] - ng[d];
        xyzmin[d] = e.geom.prob_lo[d] + e.dx[d] * lo[d];
    }
}

#define ENG_CALL(x) do { if (int rc_ = (x)) return rc_; } while (0)

static int fill_boundary(Engine& e, int c0, int c1, const int ng[3], void* s) {
    for (int c = c0; c < c1; ++c)
        for (int d = 0; d < 3; ++d) ENG_CALL(pic_fill_boundary_local(&e.fab[c], d, ng[d], &e.geom, s));
    return 0;
}

static int sync_current(Engine& e, void* s) {
    for (int c = 6; c < 9; ++c) {
        if (e.use_filter) {
            // WarpX::ApplyFilterJ (WarpXComm.cpp:1357-1374): filter into a temporary over the grown box, copy back
            const size_t bytes = sizeof(double) * (size_t)fab_size(e.fab[c]);
            if (bytes > e.filter_tmp_bytes) {
                if (e.filter_tmp) cudaFree(e.filter_tmp);
                if (cudaMalloc(&e.filter_tmp, bytes) != cudaSuccess) return fail("pic_engine: filter scratch allocation failed");
                e.filter_tmp_bytes = bytes;
            }
            pic_fab tmp = e.fab[c];
            tmp.p = e.filter_tmp;
            ENG_CALL(pic_apply_filter(&e.fab[c], &tmp, e.npass, s));
            cudaMemcpyAsync(e.fab[c].p, tmp.p, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)s);
        }
        // src = ng_depos_J (+ stencil_length-1 with the filter, WarpXComm.cpp:1413-1416) == ng_J either way
        for (int d = 0; d < 3; ++d) ENG_CALL(pic_sum_boundary_local(&e.fab[c], d, e.ng_J[d], &e.geom, s));
        for (int d = 0; d < 3; ++d) ENG_CALL(pic_fill_boundary_local(&e.fab[c], d, e.ng_J[d], &e.geom, s));  // all guards
    }
    return 0;
}

static int push(Engine& e, Species& sp, double dt, int push_position, void* s) {
    double xyzmin[3]; int lo[3];
    lower_corner(e, e.ng_EB, xyzmin, lo);
    const pic_soa& P = sp.buf[sp.cur];
    if (push_position && sp.has_esc) cudaMemsetAsync(sp.esc.count, 0, sizeof(int), (cudaStream_t)s);
    return pic_gather_push(&P, 0, P.np, &e.fab[0], &e.fab[3], e.dinv, xyzmin, lo, sp.q, sp.m, dt, e.nox,
                           e.galerkin, e.pusher, push_position, sp.has_bins ? &sp.bins : nullptr,
                           sp.has_esc ? &sp.esc : nullptr, s);
}

static int push_particles_and_deposit(Engine& e, void* s) {
    for (int c = 6; c < 9; ++c)                         // J.setVal(0), MultiParticleContainer.cpp:467-478
        cudaMemsetAsync(e.fab[c].p, 0, sizeof(double) * (size_t)fab_size(e.fab[c]), (cudaStream_t)s);
    double xyzmin[3]; int lo[3];
    lower_corner(e, e.ng_J, xyzmin, lo);
    for (auto& sp : e.species) {
        ENG_CALL(push(e, sp, e.dt, 1, s));
        const pic_soa& P = sp.buf[sp.cur];
        ENG_CALL(pic_deposit_esirkepov(&P, 0, P.np, &e.fab[6], e.dinv, xyzmin, lo, sp.q, e.dt, -0.5 * e.dt,
                                       e.nox, sp.has_bins ? &sp.bins : nullptr, s));
    }
    return 0;
}

static int sort_species(Engine& e, Species& sp, void* s) {
    pic_soa& in = sp.buf[sp.cur];
    pic_soa& out = sp.buf[1 - sp.cur];
    out.np = in.np;
    for (int d = 0; d < 3; ++d) { sp.bins.box_lo[d] = e.box_lo[d]; sp.bins.box_hi[d] = e.box_hi[d]; }
    ENG_CALL(pic_sort_particles_by_cell(&in, &out, &e.geom, &sp.bins, sp.sort_work, s));
    sp.bins.np_binned = in.np;
    sp.cur = 1 - sp.cur;
    sp.has_bins = true;
    return 0;
}

static int one_step(Engine& e, bool last, void* s) {
    // ---- ExplicitFillBoundaryEBUpdateAux ----
    if (e.is_synchronized) {
        ENG_CALL(fill_boundary(e, 0, 6, e.ng_EB, s));                    // ng_alloc_EB, :487-488
        for (auto& sp : e.species) ENG_CALL(push(e, sp, -0.5 * e.dt, 0, s));   // PushP(-dt/2), :492-504
        e.is_synchronized = false;
    } else {
        ENG_CALL(fill_boundary(e, 0, 6, e.ng_FG, s));                    // :515-516
    }
    // ---- OneStep_nosub ----
    ENG_CALL(push_particles_and_deposit(e, s));
    ENG_CALL(sync_current(e, s));
    ENG_CALL(pic_evolve_b(&e.fab[3], &e.fab[0], &e.st, 0.5 * e.dt, s));
    ENG_CALL(fill_boundary(e, 3, 6, e.ng_FS, s));
    ENG_CALL(pic_evolve_e(&e.fab[0], &e.fab[3], &e.fab[6], &e.st, e.dt, s));
    ENG_CALL(fill_boundary(e, 0, 3, e.ng_FS, s));
    ENG_CALL(pic_evolve_b(&e.fab[3], &e.fab[0], &e.st, 0.5 * e.dt, s));
    if (last) {                                                           // Synchronize(), :64-91
        ENG_CALL(fill_boundary(e, 0, 6, e.ng_FG, s));
        for (auto& sp : e.species) ENG_CALL(push(e, sp, 0.5 * e.dt, 0, s));
        e.is_synchronized = true;
    }
    // ---- HandleParticlesAtBoundaries ----
    const long step = e.istep++;
    for (auto& sp : e.species) {
        // amrex enforcePeriodic: only the particles the push of this step moved out of the domain
        if (sp.has_esc) ENG_CALL(pic_particles_wrap_listed(&sp.buf[sp.cur], &e.geom, &sp.esc, s));
        else ENG_CALL(pic_particles_wrap_periodic(&sp.buf[sp.cur], &e.geom, s));
        if (sp.sort_work && e.sort_interval > 0 && (step + 1) % e.sort_interval == 0) ENG_CALL(sort_species(e, sp, s));
    }
    return 0;
}

}  // namespace pic

using namespace pic;

extern "C" void* pic_engine_create(const pic_geom* geom, const int box_lo[3], const int box_hi[3], int nox,
                                   int galerkin, int pusher, int solver, double cfl, double dt,
                                   int sort_interval, int use_filter, const int filter_npass[3]) {
    Engine* e = new Engine();
    e->geom = *geom;
    e->use_filter = use_filter ? 1 : 0;
    for (int d = 0; d < 3; ++d) e->npass[d] = filter_npass ? filter_npass[d] : 1;
    for (int d = 0; d < 3; ++d) {
        e->box_lo[d] = box_lo[d]; e->box_hi[d] = box_hi[d];
        e->dx[d] = (geom->prob_hi[d] - geom->prob_lo[d]) / geom->n_cell[d];
        e->dinv[d] = 1.0 / e->dx[d];
    }
    e->nox = nox; e->galerkin = galerkin; e->pusher = pusher; e->solver = solver; e->sort_interval = sort_interval;
    e->dt = dt > 0 ? dt : cfl * max_dt(solver, e->dx);
    stencil_coefficients(solver, e->dx, &e->st);
    guard_cells(*e);
    return e;
}
extern "C" void pic_engine_destroy(void* h) {
    Engine* e = static_cast<Engine*>(h);
    if (e && e->filter_tmp) cudaFree(e->filter_tmp);
    if (e) for (auto& sp : e->species) if (sp.has_esc) cudaFree(sp.esc.count);
    delete e;
}
extern "C" double pic_engine_dt(void* h) { return static_cast<Engine*>(h)->dt; }
extern "C" void pic_engine_guards(void* h, int out[12]) {
    Engine* e = static_cast<Engine*>(h);
    for (int d = 0; d < 3; ++d) { out[d] = e->ng_EB[d]; out[3 + d] = e->ng_J[d]; out[6 + d] = e->ng_FG[d]; out[9 + d] = e->ng_FS[d]; }
}
extern "C" int pic_engine_set_fields(void* h, const pic_fab fabs[9]) {
    Engine* e = static_cast<Engine*>(h);
    for (int c = 0; c < 9; ++c) {
        e->fab[c] = fabs[c];
        const int* ng = c < 6 ? e->ng_EB : e->ng_J;
        for (int d = 0; d < 3; ++d)
            PIC_REQUIRE(fabs[c].ng[d] >= ng[d], "pic_engine_set_fields: component %d has %d guard cells, needs %d", c, fabs[c].ng[d], ng[d]);
    }
    return 0;
}
// bufA holds the particles; bufB is the sort target.  cell_start / sort_work may be NULL (no bins).
extern "C" int pic_engine_add_species(void* h, double q, double m, const pic_soa* bufA, const pic_soa* bufB,
                                      int* cell_start, const int tile[3], void* sort_work, void* stream) {
    Engine* e = static_cast<Engine*>(h);
    Species sp;
    sp.q = q; sp.m = m; sp.buf[0] = *bufA; sp.buf[1] = *bufB; sp.cur = 0; sp.has_bins = false; sp.sort_work = sort_work;
    sp.bins.cell_start = cell_start;
    for (int d = 0; d < 3; ++d) sp.bins.tile[d] = tile ? tile[d] : 8;
    sp.bins.np_binned = 0;
    // escape list: a layer of one cell next to every domain face can leave per step at most
    sp.has_esc = false;
    {
        const long cap = bufA->np / 16 + 65536;
        int* mem = nullptr;
        if (cudaMalloc(&mem, sizeof(int) * (size_t)(cap + 1)) == cudaSuccess) {
            sp.esc.count = mem; sp.esc.idx = mem + 1; sp.esc.capacity = (int)cap;
            for (int d = 0; d < 3; ++d) {
                sp.esc.lo[d] = e->geom.periodic[d] ? e->geom.prob_lo[d] : -INFINITY;
                sp.esc.hi[d] = e->geom.periodic[d] ? e->geom.prob_hi[d] : INFINITY;
            }
            sp.has_esc = true;
        } else {
            cudaGetLastError();      // no list: the engine wraps with the full sweep
        }
    }
    e->species.push_back(sp);
    if (cell_start && sort_work) return sort_species(*e, e->species.back(), stream);
    return 0;
}
// which of the two buffers currently holds species isp, and its particle count
extern "C" int pic_engine_species_buffer(void* h, int isp, long* np) {
    Engine* e = static_cast<Engine*>(h);
    if (np) *np = e->species[isp].buf[e->species[isp].cur].np;
    return e->species[isp].cur;
}
// WarpX::Evolve(numsteps): the last step synchronises u with x when synchronize_last is set
extern "C" int pic_engine_evolve(void* h, int numsteps, int synchronize_last, void* stream) {
    Engine* e = static_cast<Engine*>(h);
    for (int n = 0; n < numsteps; ++n) ENG_CALL(one_step(*e, synchronize_last && n == numsteps - 1, stream));
    return 0;
}
