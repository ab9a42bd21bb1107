"""GPU parity tests (-m gpu) of the non-periodic additions: PEC walls, the
moving window, the laser antenna, continuous plasma injection and particle boundaries -- every stage
through the C ABI against the CPU oracle, then the laser-acceleration deck through the C++ driver
against the oracle and against WarpX's own golden checksums (test_3d_laser_acceleration.json).

Tolerances: the thin kernels copy / negate / add values -> bit-exact; the laser profile uses
exp / sincos of the device -> 1e-13 of the peak momentum; loops as in test_gpu_parity.py.
(Sorts after test_gpu_parity.py.  Green on a B200: profiles/r1l_gpu_tests.txt.)"""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import rel_linf
from helpers import lower_corner, random_fields
from test_gpu_parity import Dev, _particles, _run_both, _match_particles, _sorted_device_species, box  # noqa: F401
from test_oracle import check_particle_boundaries, check_pec_particle, make_lwfa_oracle
from warpx_b200 import abi, workloads

pytestmark = pytest.mark.gpu


@pytest.fixture()
def dev(cuda):
    return Dev(cuda)


BND = {
    "pec_z": dict(field_lo=("periodic", "periodic", "pec"), field_hi=("periodic", "periodic", "pec"), periodic=(1, 1, 0)),
    "pec_xz": dict(field_lo=("pec", "periodic", "pec"), field_hi=("pec", "periodic", "pec"), periodic=(0, 1, 0)),
}


def _sync_periodic_duplicates(f, periodic):
    a = f.a
    for d in range(3):
        if periodic[d] and f.desc.stag[d]:
            ax, ng = 2 - d, f.desc.ng[d]
            n = a.shape[ax] - 2 * ng - 1
            hi, lo = [slice(None)] * 3, [slice(None)] * 3
            hi[ax], lo[ax] = ng + n, ng
            a[tuple(hi)] = a[tuple(lo)]


@pytest.mark.parametrize("case", ["pec_z", "pec_xz"])
@pytest.mark.parametrize("is_E", [1, 0])
def test_pec_field_matches_oracle(orc, dev, case, is_E):
    n, ng, ngfg = (20, 12, 33), (4, 4, 4), (2, 2, 2)
    cfg = BND[case]
    geom = abi.make_geom(n, (0, 0, 0), (1, 1, 1), periodic=cfg["periodic"])
    bnd = abi.make_boundaries(cfg["field_lo"], cfg["field_hi"])
    rng = np.random.default_rng(31)
    blo, bhi = box(n)
    F = [orc.HostFab(blo, bhi, ng, abi.YEE_STAG[c + (0 if is_E else 3)]) for c in range(3)]
    for f in F:
        f.a[...] = rng.standard_normal(f.a.shape)
    arr, tens = dev.fabs(F)
    dev.ok(dev.L.pic_apply_pec_field(arr, is_E, C.byref(geom), C.byref(bnd), abi.int3(ngfg), dev.stream))
    dev.sync()
    orc.lib().orc_apply_pec_field(orc.fab_array(F), is_E, C.byref(geom), C.byref(bnd), abi.int3(ngfg))
    for f, t in zip(F, tens):
        assert np.array_equal(t.cpu().numpy(), f.a)


@pytest.mark.parametrize("case,pbc", [("pec_z", None), ("pec_xz", None),
                                      ("pec_z", (("periodic", "periodic", "reflecting"), ("periodic", "periodic", "absorbing")))])
def test_pec_current_matches_oracle(orc, dev, case, pbc):
    n, ng = (20, 12, 33), (5, 5, 5)
    cfg = BND[case]
    geom = abi.make_geom(n, (0, 0, 0), (1, 1, 1), periodic=cfg["periodic"])
    bnd = abi.make_boundaries(cfg["field_lo"], cfg["field_hi"], *(pbc or (None, None)))
    rng = np.random.default_rng(32)
    blo, bhi = box(n)
    J = [orc.HostFab(blo, bhi, ng, abi.YEE_STAG[6 + c]) for c in range(3)]
    for f in J:
        f.a[...] = rng.standard_normal(f.a.shape)
    arr, tens = dev.fabs(J)
    dev.ok(dev.L.pic_apply_pec_current(arr, C.byref(geom), C.byref(bnd), dev.stream))
    dev.sync()
    orc.lib().orc_apply_pec_current(orc.fab_array(J), C.byref(geom), C.byref(bnd))
    for f, t in zip(J, tens):
        assert np.array_equal(t.cpu().numpy(), f.a)


@pytest.mark.parametrize("shift", [1, 2, -1])
@pytest.mark.parametrize("comp", [0, 2, 4, 6, 8])
def test_shift_fab_matches_oracle(orc, dev, shift, comp):
    n, ng = (20, 12, 33), (4, 4, 5) if comp >= 6 else (4, 4, 4)
    periodic = (1, 1, 0)
    geom = abi.make_geom(n, (0, 0, 0), (1, 1, 1), periodic=periodic)
    rng = np.random.default_rng(33)
    blo, bhi = box(n)
    f = orc.HostFab(blo, bhi, ng, abi.YEE_STAG[comp])
    f.a[...] = rng.standard_normal(f.a.shape)
    _sync_periodic_duplicates(f, periodic)
    arr, tens = dev.fabs([f])
    tmp = dev.t.empty(f.a.size, dtype=dev.t.float64, device="cuda")
    dev.ok(dev.L.pic_shift_fab(C.byref(arr[0]), tmp.data_ptr(), C.byref(geom), shift, 2, 0.25, dev.stream))
    dev.sync()
    orc.lib().orc_shift_fab(C.byref(f.desc), C.byref(geom), shift, 2, 0.25)
    assert np.array_equal(tens[0].cpu().numpy(), f.a)


def test_fill_boundary_leaves_corners_beyond_a_wall_alone(orc, dev):
    """FillBoundary on a box with a non-periodic z: the x / y guards are refreshed only on valid z
    indices (AMReX fills a guard point only where its periodic image is a valid point)."""
    n, ng = (12, 10, 14), (4, 4, 4)
    geom = abi.make_geom(n, (0, 0, 0), (1, 1, 1), periodic=(1, 1, 0))
    rng = np.random.default_rng(34)
    blo, bhi = box(n)
    f = orc.HostFab(blo, bhi, ng, abi.YEE_STAG[0])
    f.a[...] = rng.standard_normal(f.a.shape)
    _sync_periodic_duplicates(f, (1, 1, 0))
    orig = f.a.copy()
    arr, tens = dev.fabs([f])
    for dim in (0, 1):
        dev.ok(dev.L.pic_fill_boundary_local(C.byref(arr[0]), dim, 2, C.byref(geom), dev.stream))
    dev.sync()
    orc.lib().orc_fill_boundary(orc.fab_array([f]), 1, abi.int3((2, 2, 2)), C.byref(geom))
    # the x / y guards beyond ng = 2 are outside FillBoundary(2): the axis sweeps drag stale values
    # through them, the oracle leaves them alone; every z plane (guards beyond the walls included) counts
    sl = (slice(None), slice(2, -2), slice(2, -2))
    got = tens[0].cpu().numpy()
    assert np.array_equal(got[sl], f.a[sl])
    kz = ng[2]
    assert np.array_equal(got[:kz], orig[:kz]) and np.array_equal(got[-kz:], orig[-kz:])   # beyond the walls: untouched


def _laser():
    la = workloads.laser_acceleration_3d()["lasers"][0]
    return abi.make_laser(la["position"], la["direction"], la["polarization"], la["wavelength"], la["e_max"],
                          la["waist"], la["duration"], la["t_peak"], la["focal_distance"])


def test_laser_antenna_push_matches_oracle(orc, dev):
    wl = workloads.laser_acceleration_3d()
    dx = [(wl["prob_hi"][d] - wl["prob_lo"][d]) / wl["n_cell"][d] for d in range(3)]
    las, dxa = _laser(), abi.dbl3(dx)
    n = dev.L.pic_laser_antenna_particles(C.byref(las), dxa, abi.dbl3(wl["prob_lo"]), abi.dbl3(wl["prob_hi"]),
                                          None, None, None, None, 0)
    assert n == 2048
    A = [np.empty(n) for _ in range(4)]
    assert dev.L.pic_laser_antenna_particles(C.byref(las), dxa, abi.dbl3(wl["prob_lo"]), abi.dbl3(wl["prob_hi"]),
                                             *[a.ctypes.data for a in A], n) == n
    z = np.zeros(n)
    dt = 8.687655225973464e-16
    for t in (0.0, 17 * dt, 30.e-15, 61 * dt):
        P = orc.HostParticles(x=A[0], y=A[1], z=A[2], w=A[3], ux=z, uy=z, uz=z)
        soa, buf = dev.soa(P)
        dev.ok(dev.L.pic_laser_antenna_push(C.byref(las), dxa, C.byref(soa), t, dt, dev.stream))
        dev.sync()
        orc.lib().orc_antenna_push(C.byref(las), dxa, C.byref(P.soa), t, dt)
        got = buf.cpu().numpy()
        umax = max(np.max(np.abs(P.ux)), np.max(np.abs(P.uy)), np.max(np.abs(P.uz)))
        assert umax > 0
        for k, name in enumerate(("x", "y", "z", "w", "ux", "uy", "uz")):
            ref = getattr(P, name)
            tol = 1e-13 * umax if name.startswith("u") else (1e-13 * umax * dt if name in "xyz" else 0.0)
            assert np.max(np.abs(got[k] - ref)) <= tol, (t, name)


@pytest.mark.parametrize("ppc", [(1, 1, 1), (2, 2, 2), (1, 2, 3)])
@pytest.mark.parametrize("slab", ["domain", "top_slab", "cut"])
def test_add_plasma_matches_oracle(orc, dev, ppc, slab):
    n_cell, prob_lo, prob_hi = (24, 18, 40), (-30.e-6, -20.e-6, -56.e-6 + 3.1e-7), (30.e-6, 25.e-6, 12.e-6 + 3.1e-7)
    geom = abi.make_geom(n_cell, prob_lo, prob_hi, periodic=(1, 1, 0))
    dz = (prob_hi[2] - prob_lo[2]) / n_cell[2]
    blo, bhi = (-20.e-6, -20.e-6, 0.0), (20.e-6, 11.e-6, math.inf)
    if slab == "cut":
        blo, bhi = (-7.3e-6, -3.1e-6, -31.7e-6), (9.9e-6, 8.4e-6, -2.2e-6)
    inj = abi.make_injector(ppc, blo, bhi, 2.e23, True)
    plo, phi = list(prob_lo), list(prob_hi)
    if slab == "top_slab":
        plo[2] = prob_hi[2] - dz
    cap = n_cell[0] * n_cell[1] * n_cell[2] * ppc[0] * ppc[1] * ppc[2] + 7
    t = dev.t
    buf = t.full((7, cap), float("nan"), dtype=t.float64, device="cuda")
    ids = t.zeros(cap, dtype=t.int64, device="cuda")
    soa = abi.pic_soa()
    for k, name in enumerate(("x", "y", "z", "w", "ux", "uy", "uz")):
        setattr(soa, name, buf[k].data_ptr())
    soa.idcpu, soa.np = ids.data_ptr(), 7
    n = dev.L.pic_add_plasma(C.byref(inj), C.byref(geom), None, None, None, abi.dbl3(plo), abi.dbl3(phi), C.byref(soa), cap, 500, 0.0, dev.stream)
    assert n > 0, dev.L.pic_last_error().decode()
    dev.sync()
    B = [np.empty(cap) for _ in range(4)]
    dp = lambda a: a.ctypes.data_as(abi.c_double_p)   # noqa: E731
    m = orc.lib().orc_add_plasma(C.byref(inj), C.byref(geom), abi.dbl3(plo), abi.dbl3(phi), *[dp(b) for b in B], cap, 0.0, None)
    assert m == n
    got = buf.cpu().numpy()
    for k, b in enumerate(B):
        assert np.array_equal(got[k, 7:7 + n], b[:n]), k
    assert np.all(got[4:7, 7:7 + n] == 0.0)
    assert np.all(np.isnan(got[:, :7])) and np.all(np.isnan(got[:, 7 + n:]))
    assert np.array_equal(ids.cpu().numpy()[7:7 + n], 500 + np.arange(n))


@pytest.mark.parametrize("pbc_z", [("absorbing", "absorbing"), ("reflecting", "absorbing")])
def test_particle_boundaries_match_oracle(orc, dev, pbc_z):
    rng = np.random.default_rng(41)
    n = 200000
    geom = abi.make_geom((8, 8, 8), (-1.0, -1.0, -2.0), (1.0, 1.0, 2.0), periodic=(1, 1, 0))
    bnd = abi.make_boundaries(("periodic", "periodic", "pec"), ("periodic", "periodic", "pec"),
                              ("periodic", "periodic", pbc_z[0]), ("periodic", "periodic", pbc_z[1]))
    arr = {k: rng.uniform(-1.2, 1.2, n) for k in ("x", "y")}
    arr["z"] = rng.uniform(-2.3, 2.3, n)
    arr["z"][-4000:] = 2.2
    arr["z"][-8000:-6000] = 0.0
    for k in ("w", "ux", "uy", "uz"):
        arr[k] = rng.standard_normal(n)
    P = orc.HostParticles(**arr)
    keep = C.create_string_buffer(n)
    orc.lib().orc_apply_particle_boundaries(C.byref(P.soa), C.byref(geom), C.byref(bnd), keep)
    keep = np.frombuffer(keep.raw, dtype=np.int8)[:n].astype(bool)
    Q = orc.HostParticles(**arr)
    soa, buf = dev.soa(Q)
    t = dev.t
    ids = t.arange(n, dtype=t.int64, device="cuda")
    soa.idcpu = ids.data_ptr()
    cap = 1 << 16
    work = t.zeros(dev.L.pic_particles_boundary_workspace_ints(cap), dtype=t.int32, device="cuda")
    dev.ok(dev.L.pic_particles_boundary_mark(C.byref(soa), C.byref(geom), C.byref(bnd), work.data_ptr(), cap, dev.stream))
    n_lost = int(work[0].item())
    assert n_lost == int(np.sum(~keep)) and 0 < n_lost < cap
    dev.ok(dev.L.pic_particles_boundary_compact(C.byref(soa), work.data_ptr(), cap, n_lost, dev.stream))
    dev.sync()
    m = n - n_lost
    got_ids = ids.cpu().numpy()[:m]
    assert np.array_equal(np.sort(got_ids), np.flatnonzero(keep))
    order = np.argsort(got_ids)
    got = buf.cpu().numpy()
    for k, name in enumerate(("x", "y", "z", "w", "ux", "uy", "uz")):
        assert np.array_equal(got[k, :m][order], getattr(P, name)[keep]), name


def test_particle_energy_matches_oracle(orc, dev):
    """pic_particle_energy against ParticleEnergy restated (ReducedDiags/ParticleEnergy.cpp:86-170)."""
    rng = np.random.default_rng(51)
    n = 300001
    arr = {k: rng.standard_normal(n) for k in ("x", "y", "z")}
    arr["w"] = rng.uniform(0.5, 2.0, n)
    for k in ("ux", "uy", "uz"):
        arr[k] = rng.standard_normal(n) * 2.0 * workloads.C
    P = orc.HostParticles(**arr)
    soa, buf = dev.soa(P)
    out = dev.t.zeros(2, dtype=dev.t.float64, device="cuda")
    dev.ok(dev.L.pic_particle_energy(C.byref(soa), workloads.M_E, out.data_ptr(), dev.stream))
    ref = (C.c_double * 2)()
    orc.lib().orc_particle_energy(C.byref(P.soa), workloads.M_E, ref)
    got = out.cpu().numpy()
    assert got[0] == pytest.approx(ref[0], rel=1e-12) and got[1] == pytest.approx(ref[1], rel=1e-12)


# ---------------------------------------------------------------------------------------------
def make_lwfa_sim(wl, capacity):
    from warpx_b200.engine import Simulation, max_dt, nci_godfrey_stencils
    from warpx_b200.lib import lib as piclib
    from test_oracle import _nci_lines
    nci = None
    if wl.get("use_fdtd_nci_corr"):
        dx = [(wl["prob_hi"][d] - wl["prob_lo"][d]) / wl["n_cell"][d] for d in range(3)]
        nci = nci_godfrey_stencils(piclib(), _nci_lines(), workloads.C * wl["cfl"] * max_dt(wl["solver"], dx) / dx[2])
    sim = Simulation(wl["n_cell"], wl["prob_lo"], wl["prob_hi"], nox=wl["nox"], cfl=wl["cfl"],
                     solver=wl["solver"], pusher=wl["pusher"], use_filter=wl["use_filter"], sort_interval=4,
                     boundaries=abi.make_boundaries(wl["field_lo"], wl["field_hi"]),
                     moving_window=(wl["moving_window_dir"], wl["moving_window_v"]),
                     gamma_boost=wl.get("gamma_boost", 1.0), nci_stencils=nci)
    for s in wl["species"]:
        sim.add_plasma_species(s["name"], s["q"], s["m"],
                               abi.make_injector(s["ppc"], s["bound_lo"], s["bound_hi"], s["density"],
                                                 s["do_continuous_injection"]), capacity)
    for la in wl["lasers"]:
        sim.add_laser(abi.make_laser(la["position"], la["direction"], la["polarization"], la["wavelength"],
                                     la["e_max"], la["waist"], la["duration"], la["t_peak"], la["focal_distance"]))
    return sim


@pytest.mark.parametrize("solver,pusher", [(abi.SOLVER_YEE, abi.PUSHER_BORIS), (abi.SOLVER_CKC, abi.PUSHER_VAY)])
def test_laser_acceleration_loop_matches_oracle(orc, cuda, solver, pusher):
    """30 steps of the laser-acceleration deck (order 3, filter, PEC z, moving window, antenna,
    continuous injection) through the C++ driver against the oracle: fields, every electron by id,
    the antenna particles, the moving domain and the time.  Yee / Boris is the deck itself; CKC / Vay
    is what BASELINE.json's config 4 names (dt = dz / c: the window moves one cell every step)."""
    wl = workloads.laser_acceleration_3d(max_step=30, solver=solver, pusher=pusher)
    sim = make_lwfa_sim(wl, capacity=22 * 22 * 256)
    osim = make_lwfa_oracle(orc, wl)
    assert sim.ng_EB == osim.guards()["ng_EB"] and sim.ng_J == osim.guards()["ng_J"]
    assert sim.species[0].np == osim.L.orc_sim_np(osim.h, 0) == 22 * 22 * 45
    for chunk, sync in ((17, False), (13, True)):       # two Evolve calls: the window state carries over
        sim.Evolve(chunk, synchronize_last=sync)
        osim.evolve(chunk, synchronize_last=sync)
    cuda.cuda.synchronize()
    assert sim.time == pytest.approx(osim.time(), rel=1e-15)
    plo, phi = osim.prob_domain()
    assert sim.prob_lo == pytest.approx(plo, rel=0, abs=1e-20) and sim.prob_hi == pytest.approx(phi, rel=0, abs=1e-20)
    for c in range(9):
        d, a = sim.field_numpy(c)
        _, oa = osim.fab(c)
        scale = np.max(np.abs(oa[d.valid_slices()]))
        assert scale > 0, abi.COMP_NAMES[c]
        assert rel_linf(a[d.valid_slices()], oa[d.valid_slices()]) <= 1e-9, abi.COMP_NAMES[c]
    A = sim.species_numpy(0, sort_by_id=True)
    B = osim.particles(0)
    assert len(A["x"]) == len(B["x"]) and np.array_equal(A["id"], np.arange(len(B["x"])))
    assert np.array_equal(A["w"], B["w"])
    for k in ("x", "y", "z"):
        assert np.max(np.abs(A[k] - B[k])) / sim.dx[2] <= 1e-10, k
    for k in ("ux", "uy", "uz"):
        assert np.max(np.abs(A[k] - B[k])) / workloads.C <= 1e-10, k
    ke, ws = sim.particle_energy()[0]
    oke, ows = osim.particle_energy(0)
    assert ke == pytest.approx(oke, rel=1e-9) and ws == pytest.approx(ows, rel=1e-13)
    e, b = sim.field_energy()
    eo, bo = osim.field_energy()
    assert e == pytest.approx(eo, rel=1e-9) and b == pytest.approx(bo, rel=1e-9)
    LA, LB = sim.laser_numpy(0), osim.laser_particles(0)
    assert len(LA["x"]) == len(LB["x"]) == 2048
    for k in ("x", "y", "z"):
        assert np.max(np.abs(LA[k] - LB[k])) / sim.dx[2] <= 1e-10, k
    for k in ("ux", "uy", "uz"):
        assert np.max(np.abs(LA[k] - LB[k])) / workloads.C <= 1e-10, k


def test_boosted_laser_acceleration_loop_matches_oracle(orc, cuda):
    """BASELINE.json config 4 in the small (workloads.laser_acceleration_boosted_3d: gamma_boost = 10, CKC, Vay,
    order 3, filter, NCI corrector, PEC z, moving window, boosted antenna, electrons + ions injected continuously from the
    lab-frame plasma bounds): 40 steps through the C++ driver against the oracle -- fields, every particle of
    both species by id, the drifting antenna, the moving domain."""
    wl = workloads.laser_acceleration_boosted_3d(use_fdtd_nci_corr=True)
    sim = make_lwfa_sim(wl, capacity=16 * 16 * 200)
    osim = make_lwfa_oracle(orc, wl)
    assert sim.gamma_boost == 10.0 and sim.species[0].np == osim.L.orc_sim_np(osim.h, 0) == 0
    for chunk, sync in ((23, False), (17, True)):
        sim.Evolve(chunk, synchronize_last=sync)
        osim.evolve(chunk, synchronize_last=sync)
    cuda.cuda.synchronize()
    assert sim.time == pytest.approx(osim.time(), rel=1e-15)
    plo, phi = osim.prob_domain()
    assert sim.prob_lo == pytest.approx(plo, rel=0, abs=1e-18) and sim.prob_hi == pytest.approx(phi, rel=0, abs=1e-18)
    for group in ((0, 1, 2), (3, 4, 5), (6, 7, 8)):        # E, B, J: each component against the scale of its vector
        got, want = [], []                                 # (the small components are cancellation noise of the antenna's
        for c in group:                                    #  +-w drift currents next to the wall)
            d, a = sim.field_numpy(c)
            _, oa = osim.fab(c)
            got.append(a[d.valid_slices()])
            want.append(oa[d.valid_slices()])
        scale = max(np.max(np.abs(w)) for w in want)
        assert scale > 0
        for c, g, w in zip(group, got, want):
            assert np.max(np.abs(g - w)) <= 1e-8 * scale, abi.COMP_NAMES[c]
    for isp in (0, 1):
        A = sim.species_numpy(isp, sort_by_id=True)
        B = osim.particles(isp)
        assert len(A["x"]) == len(B["x"]) > 0 and np.array_equal(A["id"], np.arange(len(B["x"])))
        assert np.array_equal(A["w"], B["w"])
        for k in ("x", "y", "z"):
            assert np.max(np.abs(A[k] - B[k])) / sim.dx[2] <= 1e-9, k
        for k in ("ux", "uy", "uz"):
            assert np.max(np.abs(A[k] - B[k])) / (10.0 * workloads.C) <= 1e-9, k
    LA, LB = sim.laser_numpy(0), osim.laser_particles(0)
    assert len(LA["x"]) == len(LB["x"]) > 0
    for k in ("x", "y", "z"):
        assert np.max(np.abs(LA[k] - LB[k])) / sim.dx[2] <= 1e-10, k
    for k in ("ux", "uy", "uz"):
        assert np.max(np.abs(LA[k] - LB[k])) / (10.0 * workloads.C) <= 1e-10, k


@pytest.mark.parametrize("comp", range(6))
def test_nci_filter_matches_oracle(orc, dev, comp):
    """pic_apply_nci_filter on the device against the oracle (the reference's summation order; the final
    multiply-add may contract to an FMA on the device: 1e-15 of the data scale)."""
    from test_oracle import oracle_nci_stencils
    n, ng, nox = (40, 24, 33), (4, 4, 8), 3
    rng = np.random.default_rng(200 + comp)
    stz = (C.c_double * 5)(*oracle_nci_stencils(orc, 0.98)[0 if comp in (0, 1, 5) else 1])
    src = orc.HostFab((0, 0, 0), (n[0] - 1, n[1] - 1, n[2] - 1), ng, abi.YEE_STAG[comp])
    src.a[:] = rng.standard_normal(src.a.shape)
    want = orc.HostFab((0, 0, 0), (n[0] - 1, n[1] - 1, n[2] - 1), ng, abi.YEE_STAG[comp])
    want.a[:] = -7.0
    tlo = [-nox] * 3
    thi = [n[d] - 1 + nox + abi.YEE_STAG[comp][d] for d in range(3)]
    orc.lib().orc_apply_nci_filter(C.byref(src.desc), C.byref(want.desc), stz, abi.int3(tlo), abi.int3(thi))
    out = orc.HostFab((0, 0, 0), (n[0] - 1, n[1] - 1, n[2] - 1), ng, abi.YEE_STAG[comp])
    out.a[:] = -7.0
    arr, tens = dev.fabs([src, out])
    dev.ok(dev.L.pic_apply_nci_filter(C.byref(arr[0]), C.byref(arr[1]), stz, abi.int3((0, 0, 0)), abi.int3((n[0] - 1, n[1] - 1, n[2] - 1)), nox, dev.stream))
    dev.sync()
    got = tens[1].cpu().numpy()
    assert np.max(np.abs(got - want.a)) <= 1e-15 * np.max(np.abs(src.a)) * 8
    assert np.array_equal(got == -7.0, want.a == -7.0)


def test_nci_corrected_loop_matches_oracle(orc, cuda):
    """particles.use_fdtd_nci_corr through the C++ driver: the streaming-plasma deck of the oracle's known-answer
    test (periodic box, CKC at c dt = dz, Vay, order 3, filter), 60 steps, fields and particles against the oracle;
    the guard cells grow by 4 along z."""
    from test_oracle import nci_streaming_plasma, _nci_lines, oracle_nci_stencils
    from warpx_b200.engine import Simulation, nci_godfrey_stencils
    from warpx_b200.lib import lib as piclib
    wl = nci_streaming_plasma()
    stencils = nci_godfrey_stencils(piclib(), _nci_lines(), 1.0)
    assert list(stencils[0]) == oracle_nci_stencils(orc, 1.0)[0]
    sim = Simulation(wl["n_cell"], wl["prob_lo"], wl["prob_hi"], nox=3, cfl=1.0, solver=wl["solver"], pusher=wl["pusher"],
                     use_filter=True, sort_interval=4, nci_stencils=stencils)
    osim = orc.OracleSim(wl["n_cell"], wl["prob_lo"], wl["prob_hi"], nox=3, cfl=1.0, use_filter=1, solver=wl["solver"],
                         pusher=wl["pusher"])
    osim.set_nci_corrector(*oracle_nci_stencils(orc, 1.0))
    assert sim.ng_EB == osim.guards()["ng_EB"] == [4, 4, 8] and sim.ng_FG == osim.guards()["ng_FG"] == [2, 2, 6]
    for s in wl["species"]:
        sim.add_species(s["name"], s["q"], s["m"], s["x"], s["y"], s["z"], s["w"], s["ux"], s["uy"], s["uz"])
        osim.add_species(s["q"], s["m"], s["x"], s["y"], s["z"], s["w"], s["ux"], s["uy"], s["uz"])
    sim.Evolve(60)
    osim.evolve(60)
    cuda.cuda.synchronize()
    for c in range(9):
        d, a = sim.field_numpy(c)
        _, oa = osim.fab(c)
        assert rel_linf(a[d.valid_slices()], oa[d.valid_slices()]) <= 1e-7, abi.COMP_NAMES[c]
    for isp in (0, 1):
        A = sim.species_numpy(isp, sort_by_id=True)
        B = osim.particles(isp)
        for k in ("x", "y", "z"):
            assert np.max(np.abs(A[k] - B[k])) / sim.dx[2] <= 1e-9, k
        for k in ("ux", "uy", "uz"):
            assert np.max(np.abs(A[k] - B[k])) / (30.0 * workloads.C) <= 1e-10, k
    e, b = sim.field_energy()
    eo, bo = osim.field_energy()
    assert e + b == pytest.approx(eo + bo, rel=1e-6)


def test_laser_acceleration_golden_checksums(orc, cuda, golden):
    """The full deck (100 steps) on the GPU against WarpX's regression checksums at WarpX's rtol 1e-9."""
    wl = workloads.laser_acceleration_3d()
    sim = make_lwfa_sim(wl, capacity=22 * 22 * 256)
    sim.Evolve(wl["max_step"])
    cuda.cuda.synchronize()
    g = golden["test_3d_laser_acceleration"]
    L = orc.lib()
    for c, name in enumerate(abi.COMP_NAMES):
        d, a = sim.field_numpy(c)
        hf = orc.HostFab(sim.box_lo, sim.box_hi, d.ng, abi.YEE_STAG[c], data=a)
        cs = L.orc_checksum_cell_centered(C.byref(hf.desc), abi.int3(sim.box_lo), abi.int3(sim.box_hi))
        assert abs(cs - g["lev=0"][name]) <= 1e-9 * abs(g["lev=0"][name]), name
    P = sim.species_numpy(0)
    vals = {"particle_position_x": P["x"], "particle_position_y": P["y"], "particle_position_z": P["z"],
            "particle_momentum_x": P["ux"] * workloads.M_E, "particle_momentum_y": P["uy"] * workloads.M_E,
            "particle_momentum_z": P["uz"] * workloads.M_E, "particle_weight": P["w"]}
    for key, arr in vals.items():
        assert abs(float(np.sum(np.abs(arr))) - g["electrons"][key]) <= 1e-9 * abs(g["electrons"][key]), key
    assert len(P["x"]) == 22 * 22 * (45 + 98)


def test_rho_diagnostic_golden_checksum(orc, cuda, golden):
    """The `rho` key of test_3d_laser_acceleration.json on the GPU: charge deposition of the electrons and
    of the antenna, PEC image charge, bilinear filter, SumBoundary (Simulation.rho_numpy)."""
    wl = workloads.laser_acceleration_3d()
    sim = make_lwfa_sim(wl, capacity=22 * 22 * 256)
    sim.Evolve(wl["max_step"])
    d, a = sim.rho_numpy()
    cuda.cuda.synchronize()
    hf = orc.HostFab(sim.box_lo, sim.box_hi, d.ng, (1, 1, 1), data=a)
    cs = orc.lib().orc_checksum_cell_centered(C.byref(hf.desc), abi.int3(sim.box_lo), abi.int3(sim.box_hi))
    g = golden["test_3d_laser_acceleration"]["lev=0"]["rho"]
    assert abs(cs - g) <= 1e-9 * g


def test_absorbing_walls_remove_particles_like_the_oracle(orc, cuda):
    """A drifting plasma slab in a PEC box without a moving window: particles cross the absorbing z
    walls and are removed; fields and the surviving particles follow the oracle."""
    from warpx_b200.engine import Simulation
    n_cell, lo, hi = (16, 16, 32), (-2.e-6, -2.e-6, -4.e-6), (2.e-6, 2.e-6, 4.e-6)
    bnd = abi.make_boundaries(("periodic", "periodic", "pec"), ("periodic", "periodic", "pec"))
    wl = workloads.uniform_plasma_3d(n_cell=n_cell, ppc=(2, 2, 2), u_th=0.05, lx=(4.e-6, 4.e-6, 8.e-6), density=1.e25)
    s = wl["species"][0]
    s["uz"] = s["uz"] + 0.6 * workloads.C * np.sign(s["z"])      # both halves fly towards their wall
    sim = Simulation(n_cell, lo, hi, nox=3, sort_interval=4, boundaries=bnd)
    osim = orc.OracleSim(n_cell, lo, hi, nox=3)
    osim.set_boundaries(bnd)
    sim.add_species(s["name"], s["q"], s["m"], s["x"], s["y"], s["z"], s["w"], s["ux"], s["uy"], s["uz"])
    osim.add_species(s["q"], s["m"], s["x"], s["y"], s["z"], s["w"], s["ux"], s["uy"], s["uz"])
    n0 = len(s["x"])
    sim.Evolve(12)
    osim.evolve(12)
    cuda.cuda.synchronize()
    B = osim.particles(0)
    A = sim.species_numpy(0, sort_by_id=True)
    assert 0 < len(B["x"]) < n0 and len(A["x"]) == len(B["x"])
    for c in range(9):
        d, a = sim.field_numpy(c)
        _, oa = osim.fab(c)
        assert rel_linf(a[d.valid_slices()], oa[d.valid_slices()]) <= 1e-9, abi.COMP_NAMES[c]
    # the oracle keeps creation order among the survivors; ids on the GPU are creation indices
    for k in ("x", "y", "z"):
        assert np.max(np.abs(A[k] - B[k])) / sim.dx[2] <= 1e-9, k
    for k in ("ux", "uy", "uz"):
        assert np.max(np.abs(A[k] - B[k])) / workloads.C <= 1e-9, k


# ---------------------------------------------------------------------------------------------
# algo.particle_shape = 4 (the order-agnostic kernels; the supercell / run kernels cover 1..3)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("galerkin", [1, 0])
def test_order4_gather_push_matches_oracle(orc, dev, galerkin):
    L = orc.lib()
    n, lx = (20, 16, 12), 1e-5
    box_lo, box_hi = box(n)
    wl, sp = _particles(orc, n, (2, 1, 2), 0.3, lx, shuffle=True)
    prob_lo = wl["prob_lo"]
    dx = [(wl["prob_hi"][d] - prob_lo[d]) / n[d] for d in range(3)]
    dinv = [1.0 / v for v in dx]
    ngEB = (4, 4, 4)
    xyzmin, lo = lower_corner(prob_lo, dx, box_lo, ngEB)
    F = random_fields(orc, box_lo, box_hi, ngEB, 5, comps=range(6), scale=[1e10] * 3 + [30.0] * 3)
    P = orc.HostParticles(**{k: sp[k] for k in orc.HostParticles.NAMES})
    arr, _ = dev.fabs(F)
    E, B = (abi.pic_fab * 3)(*arr[0:3]), (abi.pic_fab * 3)(*arr[3:6])
    dt = 0.9 * dx[0] / workloads.C
    Pd, buf = dev.soa(P)
    for push_position in (1, 0):
        dev.ok(dev.L.pic_gather_push(C.byref(Pd), 0, P.np, E, B, abi.dbl3(dinv), abi.dbl3(xyzmin), abi.int3(lo),
                                     sp["q"], sp["m"], dt, 4, galerkin, abi.PUSHER_BORIS, push_position, None, None,
                                     dev.stream))
        assert L.orc_gather_push(C.byref(P.soa), 0, P.np, orc.fab_array(F[0:3]), orc.fab_array(F[3:6]), abi.dbl3(dinv),
                                 abi.dbl3(xyzmin), abi.int3(lo), sp["q"], sp["m"], dt, 4, galerkin, abi.PUSHER_BORIS,
                                 push_position) == 0
    dev.sync()
    got = buf.cpu().numpy()
    for i, k in enumerate(orc.HostParticles.NAMES):
        assert rel_linf(got[i], getattr(P, k)) <= 3e-13, k       # 5^3-point stencils: a little above the order-3 bound


def test_order4_deposit_matches_oracle(orc, dev):
    L = orc.lib()
    n, lx = (20, 16, 12), 1e-5
    box_lo, box_hi = box(n)
    wl, sp = _particles(orc, n, (2, 2, 2), 0.5, lx, shuffle=True)
    prob_lo = wl["prob_lo"]
    dx = [(wl["prob_hi"][d] - prob_lo[d]) / n[d] for d in range(3)]
    dinv = [1.0 / v for v in dx]
    dt = 0.95 / (np.sqrt(sum(1.0 / v ** 2 for v in dx)) * workloads.C)
    ngJ = (5, 5, 5)
    xyzmin, lo = lower_corner(prob_lo, dx, box_lo, ngJ)
    J = [orc.HostFab(box_lo, box_hi, ngJ, abi.YEE_STAG[c]) for c in (6, 7, 8)]
    P = orc.HostParticles(**{k: sp[k] for k in orc.HostParticles.NAMES})
    arr, tens = dev.fabs(J)
    Pd, buf = dev.soa(P)
    dev.ok(dev.L.pic_deposit_esirkepov(C.byref(Pd), 0, P.np, (abi.pic_fab * 3)(*arr), abi.dbl3(dinv), abi.dbl3(xyzmin),
                                       abi.int3(lo), sp["q"], dt, -0.5 * dt, 4, None, dev.stream))
    dev.sync()
    assert L.orc_deposit_esirkepov(C.byref(P.soa), 0, P.np, orc.fab_array(J), abi.dbl3(dinv), abi.dbl3(xyzmin),
                                   abi.int3(lo), sp["q"], dt, -0.5 * dt, 4) == 0
    for c in range(3):
        assert rel_linf(tens[c].cpu().numpy(), J[c].a) <= 1e-12, "j" + "xyz"[c]


def test_order4_loop_matches_oracle(orc, cuda):
    """Whole loop at order 4 through the C++ driver (bins present: the driver falls back to the
    order-agnostic kernels for this order)."""
    wl = workloads.uniform_plasma_3d(n=24, ppc=(2, 2, 2), u_th=0.01, lx=3.75e-6, perturbation=0.01)
    sim, osim = _run_both(orc, cuda, wl, 4, 8, sort_interval=4)
    assert sim.ng_EB == [4, 4, 4] and sim.ng_J == osim.guards()["ng_J"] == [5, 5, 5]
    for c in range(9):
        d, a = sim.field_numpy(c)
        _, oa = osim.fab(c)
        tol = 1e-9 if c not in (3, 4, 5) else 1e-7
        assert rel_linf(a[d.valid_slices()], oa[d.valid_slices()]) <= tol, abi.COMP_NAMES[c]
    A, B = _match_particles(sim, osim, 0)
    for k in ("x", "y", "z"):
        assert np.max(np.abs(A[k] - B[k])) / sim.dx[0] <= 1e-10
    for k in ("ux", "uy", "uz"):
        assert np.max(np.abs(A[k] - B[k])) / workloads.C <= 1e-10


# ---------------------------------------------------------------------------------------------
# The reference's PEC decks (Examples/Tests/pec) on the GPU against WarpX's golden checksums
# ---------------------------------------------------------------------------------------------
def test_pec_field_golden_checksums(orc, cuda, golden):
    from warpx_b200.engine import Simulation
    wl = workloads.pec_field_3d()
    sim = Simulation(wl["n_cell"], wl["prob_lo"], wl["prob_hi"], nox=wl["nox"], cfl=wl["cfl"], use_filter=wl["use_filter"],
                     boundaries=abi.make_boundaries(wl["field_lo"], wl["field_hi"]))
    for c, fn in wl["init_fields"].items():
        d, a = sim.field_numpy(c)
        sim.set_field(c, fn(*workloads.staggered_coordinates(d, wl["prob_lo"], sim.dx)) + 0.0 * a)
    sim.Evolve(wl["max_step"])
    cuda.cuda.synchronize()
    g = golden["test_3d_pec_field"]["lev=0"]
    L = orc.lib()
    for name, c in (("Ey", 1), ("Bx", 3)):
        d, a = sim.field_numpy(c)
        hf = orc.HostFab(sim.box_lo, sim.box_hi, d.ng, abi.YEE_STAG[c], data=a)
        cs = L.orc_checksum_cell_centered(C.byref(hf.desc), abi.int3(sim.box_lo), abi.int3(sim.box_hi))
        assert abs(cs - g[name]) <= 1e-9 * abs(g[name]), name


def test_laser_injection_golden_checksums(orc, cuda, golden):
    """Examples/Tests/laser_injection (order 1, no filter, antenna in vacuum, moving window) on the GPU."""
    wl = workloads.laser_injection_3d()
    sim = make_lwfa_sim(wl, capacity=1)
    sim.Evolve(wl["max_step"])
    cuda.cuda.synchronize()
    g = golden["test_3d_laser_injection"]["lev=0"]
    L = orc.lib()
    for c, name in enumerate(abi.COMP_NAMES):
        if name not in g:
            continue
        d, a = sim.field_numpy(c)
        hf = orc.HostFab(sim.box_lo, sim.box_hi, d.ng, abi.YEE_STAG[c], data=a)
        cs = L.orc_checksum_cell_centered(C.byref(hf.desc), abi.int3(sim.box_lo), abi.int3(sim.box_hi))
        assert abs(cs - g[name]) <= 1e-9 * abs(g[name]) + 1e-40, name


def test_pec_particle_golden_checksums(orc, cuda, golden):
    """PEC walls in x with two particles 2 nm from the wall (Vay, order 3, filter).  jx is a unit-in-the-last-place
    artefact of the deposition coordinate that depends on the box decomposition: on one box it is exactly twice
    the value stored for the reference's two-box run (tests/test_oracle.py::test_pec_particle_golden_checksums)."""
    from warpx_b200.engine import Simulation
    wl = workloads.pec_particle_3d()
    sim = Simulation(wl["n_cell"], wl["prob_lo"], wl["prob_hi"], nox=wl["nox"], cfl=wl["cfl"], use_filter=wl["use_filter"],
                     pusher=abi.PUSHER_VAY, boundaries=abi.make_boundaries(wl["field_lo"], wl["field_hi"]))
    for s in wl["species"]:
        sim.add_species(s["name"], s["q"], s["m"], s["x"], s["y"], s["z"], s["w"], s["ux"], s["uy"], s["uz"])
    sim.Evolve(wl["max_step"])
    cuda.cuda.synchronize()
    L = orc.lib()

    def checksum(c):
        d, a = sim.field_numpy(c)
        hf = orc.HostFab(sim.box_lo, sim.box_hi, d.ng, abi.YEE_STAG[c], data=a)
        return L.orc_checksum_cell_centered(C.byref(hf.desc), abi.int3(sim.box_lo), abi.int3(sim.box_hi))

    ratio = check_pec_particle(golden, checksum, lambda isp: sim.species_numpy(isp), wl["mass"])
    # What is physical: the current normal to the wall vanishes against the tangential one (the stored file: 1e-11).
    jx, jy = checksum(6), checksum(7)
    assert abs(jx) <= 1e-9 * abs(jy)
    # jx itself is units in the last place of Esirkepov's grid coordinates (133 on one box: one unit = 2^-45, twice the
    # unit of the reference's two-box run).  The CPU oracle, compiled without FMA contraction, gives exactly 2.0; nvcc
    # contracts the shape-factor polynomials, which adds last-place units of the fractional coordinate (2^-53, i.e.
    # 2^-8 of the first unit: the device run of round 1 gave 2.0039).  The deposition coordinates themselves are
    # evaluated with individually rounded operations (pic_common.cuh deposit_coords), as the reference's CPU build does.
    assert ratio == pytest.approx(2.0, rel=2e-2)


def test_particle_boundaries_golden_checksums(orc, cuda, golden):
    """Examples/Tests/boundaries (reflecting x, absorbing y, periodic z, neutral particles) on the GPU."""
    from warpx_b200.engine import Simulation
    wl = workloads.particle_boundaries_3d()
    sim = Simulation(wl["n_cell"], wl["prob_lo"], wl["prob_hi"], nox=wl["nox"], cfl=wl["cfl"], use_filter=wl["use_filter"],
                     boundaries=abi.make_boundaries(wl["field_lo"], wl["field_hi"], wl["particle_lo"], wl["particle_hi"]))
    for s in wl["species"]:
        sim.add_species(s["name"], s["q"], s["m"], s["x"], s["y"], s["z"], s["w"], s["ux"], s["uy"], s["uz"])
    sim.Evolve(wl["max_step"])
    cuda.cuda.synchronize()
    check_particle_boundaries(golden, lambda isp: sim.species_numpy(isp))
    assert sim.species[1].np == 1


@pytest.mark.parametrize("nox", [1, 3, 4])
def test_charge_deposition_matches_oracle(orc, dev, nox):
    rng = np.random.default_rng(90 + nox)
    n, ng = (20, 16, 24), (5, 5, 5)
    prob_lo, prob_hi = (-1.0, -2.0, 0.5), (1.5, 2.0, 3.5)
    dx = [(prob_hi[d] - prob_lo[d]) / n[d] for d in range(3)]
    dinv = [1.0 / v for v in dx]
    npart = 200000
    arr = {k: rng.uniform(prob_lo[d], prob_hi[d], npart) for d, k in enumerate(("x", "y", "z"))}
    arr["w"] = rng.uniform(0.5, 2.0, npart)
    for k in ("ux", "uy", "uz"):
        arr[k] = np.zeros(npart)
    P = orc.HostParticles(**arr)
    lo = [-ng[d] for d in range(3)]
    xyzmin = [prob_lo[d] + dx[d] * lo[d] for d in range(3)]
    geom = abi.make_geom(n, prob_lo, prob_hi, periodic=(1, 1, 0))
    bnd = abi.make_boundaries(("periodic", "periodic", "pec"), ("periodic", "periodic", "pec"))
    A = orc.HostFab(*box(n), ng, (1, 1, 1))
    arr_d, tens = dev.fabs([A])
    soa, buf = dev.soa(P)
    dev.ok(dev.L.pic_deposit_charge(C.byref(soa), 0, npart, C.byref(arr_d[0]), abi.dbl3(dinv), abi.dbl3(xyzmin),
                                    abi.int3(lo), -1.6e-19, nox, dev.stream))
    dev.ok(dev.L.pic_apply_pec_rho(C.byref(arr_d[0]), C.byref(geom), C.byref(bnd), dev.stream))
    dev.sync()
    assert orc.lib().orc_deposit_charge(C.byref(P.soa), C.byref(A.desc), abi.dbl3(dinv), abi.dbl3(xyzmin), abi.int3(lo),
                                        -1.6e-19, nox) == 0
    orc.lib().orc_apply_pec_rho(C.byref(A.desc), C.byref(geom), C.byref(bnd))
    assert rel_linf(tens[0].cpu().numpy(), A.a) <= 1e-12


# ---------------------------------------------------------------------------------------------
# Experimental variants of the register-run deposition (pic_set_deposit_mode 2, 3, 4):
# same parity bar as the default kernel.  (Checked under the SIMT emulator on the host: test_simt_host.py.)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
@pytest.mark.parametrize("nox,kind", [(3, "sorted"), (3, "drifted"), (1, "sorted"), (2, "sorted"), (3, "relativistic")])
def test_deposit_variants_match_oracle(orc, dev, mode, nox, kind):
    if mode in (5, 6) and nox != 3:
        pytest.skip("four lines per lane: order 3 only")
    L = orc.lib()
    n, lx = (20, 16, 12), 1e-5
    box_lo, box_hi = box(n)
    u_th = {"relativistic": 3.0, "sorted": 0.02}.get(kind, 0.5)
    wl, sp = _particles(orc, n, (2, 2, 2), u_th, lx)
    prob_lo = wl["prob_lo"]
    dx = [(wl["prob_hi"][d] - prob_lo[d]) / n[d] for d in range(3)]
    dinv = [1.0 / v for v in dx]
    dt = 0.95 / (np.sqrt(sum(1.0 / v ** 2 for v in dx)) * workloads.C)
    ngJ = tuple(nox + 1 + (1 if kind == "drifted" else 0) for _ in range(3))
    xyzmin, lo = lower_corner(prob_lo, dx, box_lo, ngJ)
    J = [orc.HostFab(box_lo, box_hi, ngJ, abi.YEE_STAG[c]) for c in (6, 7, 8)]
    P = orc.HostParticles(**{k: sp[k] for k in orc.HostParticles.NAMES})
    arr, tens = dev.fabs(J)
    Pd, buf, bins_s, _, _ = _sorted_device_species(dev, P, n, prob_lo, wl["prob_hi"])
    if kind == "drifted":
        buf[0:3] += dev.t.tensor([[0.8 * dx[0]], [-0.9 * dx[1]], [0.6 * dx[2]]], device="cuda")
    host = buf.cpu().numpy()
    P = orc.HostParticles(**{k: host[i] for i, k in enumerate(orc.HostParticles.NAMES)})
    dev.L.pic_set_deposit_mode(mode)
    try:
        dev.ok(dev.L.pic_deposit_esirkepov(C.byref(Pd), 0, P.np, (abi.pic_fab * 3)(*arr), abi.dbl3(dinv), abi.dbl3(xyzmin),
                                           abi.int3(lo), sp["q"], dt, -0.5 * dt, nox, C.byref(bins_s), dev.stream))
        dev.sync()
    finally:
        dev.L.pic_set_deposit_mode(0)
    L.orc_deposit_esirkepov(C.byref(P.soa), 0, P.np, orc.fab_array(J), abi.dbl3(dinv), abi.dbl3(xyzmin),
                            abi.int3(lo), sp["q"], dt, -0.5 * dt, nox)
    for c in range(3):
        assert rel_linf(tens[c].cpu().numpy(), J[c].a) <= 1e-12, "j" + "xyz"[c]


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("nox,kind", [(3, "sorted"), (3, "drifted"), (1, "drifted"), (2, "sorted")])
def test_gather_variants_match_oracle(orc, dev, mode, nox, kind):
    """pic_set_gather_mode(PIC_GATHER_PAIRS / _WIDE): two particles of a cell per lane, against the oracle's gather +
    push, cell-sorted and after a drift of most of a cell (stray lists, lone survivors); order 2 keeps the default kernel."""
    L = orc.lib()
    n, lx = (24, 16, 16), 1.2e-5
    box_lo, box_hi = box(n)
    wl, sp = _particles(orc, n, (2, 2, 2), 0.3, lx)
    prob_lo = wl["prob_lo"]
    dx = [(wl["prob_hi"][d] - prob_lo[d]) / n[d] for d in range(3)]
    dinv = [1.0 / v for v in dx]
    ngEB = (4, 4, 4)
    xyzmin, lo = lower_corner(prob_lo, dx, box_lo, ngEB)
    F = random_fields(orc, box_lo, box_hi, ngEB, 5, comps=range(6), scale=[1e10] * 3 + [30.0] * 3)
    P = orc.HostParticles(**{k: sp[k] for k in orc.HostParticles.NAMES})
    arr, _ = dev.fabs(F)
    E, B = (abi.pic_fab * 3)(*arr[0:3]), (abi.pic_fab * 3)(*arr[3:6])
    dt = 0.9 * dx[0] / workloads.C
    Pd, buf, bins_s, _, _ = _sorted_device_species(dev, P, n, prob_lo, wl["prob_hi"])
    if kind == "drifted":
        buf[0:3] += dev.t.tensor([[0.9 * dx[0]], [-0.7 * dx[1]], [0.8 * dx[2]]], device="cuda")
    host = buf.cpu().numpy()
    P = orc.HostParticles(**{k: host[i] for i, k in enumerate(orc.HostParticles.NAMES)})
    dev.L.pic_set_gather_mode(mode)
    try:
        for push_position in (1, 0):
            dev.ok(dev.L.pic_gather_push(C.byref(Pd), 0, P.np, E, B, abi.dbl3(dinv), abi.dbl3(xyzmin), abi.int3(lo),
                                         sp["q"], sp["m"], dt, nox, 1, abi.PUSHER_BORIS, push_position, C.byref(bins_s),
                                         None, dev.stream))
            L.orc_gather_push(C.byref(P.soa), 0, P.np, orc.fab_array(F[0:3]), orc.fab_array(F[3:6]), abi.dbl3(dinv),
                              abi.dbl3(xyzmin), abi.int3(lo), sp["q"], sp["m"], dt, nox, 1, abi.PUSHER_BORIS, push_position)
        dev.sync()
    finally:
        dev.L.pic_set_gather_mode(0)
    got = buf.cpu().numpy()
    for i, k in enumerate(("x", "y", "z")):
        assert np.max(np.abs(got[i] - getattr(P, k))) <= 1e-13 * lx, k
    for i, k in ((4, "ux"), (5, "uy"), (6, "uz")):
        assert rel_linf(got[i], getattr(P, k)) <= 1e-13, k


def test_loop_with_pair_gather_matches_oracle(orc, cuda):
    """Ten steps of the uniform-plasma deck (order 3, sort every 4 steps) with the pair gather against the oracle."""
    from warpx_b200.lib import lib as piclib
    wl = workloads.uniform_plasma_3d(n=16, ppc=(2, 2, 2), u_th=0.05, perturbation=0.01)
    piclib().pic_set_gather_mode(1)
    try:
        sim, osim = _run_both(orc, cuda, wl, 3, 10, sort_interval=4)
    finally:
        piclib().pic_set_gather_mode(0)
    for c in range(9):
        d, a = sim.field_numpy(c)
        _, oa = osim.fab(c)
        assert rel_linf(a[d.valid_slices()], oa[d.valid_slices()]) <= 1e-9, abi.COMP_NAMES[c]
    A, B = _match_particles(sim, osim, 0)
    for k in ("x", "y", "z"):
        assert np.max(np.abs(A[k] - B[k])) / sim.dx[0] <= 1e-10, k
    for k in ("ux", "uy", "uz"):
        assert np.max(np.abs(A[k] - B[k])) / workloads.C <= 1e-11, k
