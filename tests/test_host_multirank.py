"""CPU tests (-m "not gpu"): the multi-rank paths of the C++ step driver on the host.  The "ranks" are threads of
this process, each with its own engine instance of the host library (tests/host_harness: every csrc file compiled
by g++ against the SIMT emulator); the eight NCCL entry points the driver binds at run time resolve to
tests/host_harness/fake_nccl.cpp (PIC_NCCL_LIBRARY).  Checked against the single-box oracle: guard-cell sweeps and
particle migration between bricks, and -- never exercised on a device so far -- slabs along a non-periodic axis
with PEC walls, the moving window pulling planes from the next slab, continuous injection on the top slab with
global ids, and the antenna replicated on every rank.  TEST INFRASTRUCTURE; no product code runs this way."""
import numpy as np
import pytest

from warpx_b200 import abi, parallel, workloads


@pytest.fixture(scope="module")
def hh():
    from host_harness import harness
    return harness


def _gather_fields(res, osim, world):
    """max |rank value - oracle value| per component over the valid points of every brick, and the oracle's scale."""
    err, scale = [], []
    for c in range(9):
        od, oa = osim.fab(c)
        worst = 0.0
        for r in range(world):
            (d, a), (blo, bhi) = res[r]["fields"][c], res[r]["box"]
            sl = tuple(slice(od.ng[2 - ax] + blo[2 - ax], od.ng[2 - ax] + bhi[2 - ax] + 1 + abi.YEE_STAG[c][2 - ax])
                       for ax in range(3))
            worst = max(worst, float(np.max(np.abs(a[d.valid_slices()] - oa[sl]))))
        err.append(worst)
        scale.append(float(np.max(np.abs(oa))))
    return err, scale


def test_two_bricks_periodic_loop_matches_oracle(orc, hh):
    """Config 1 in the small (16^3 Langmuir wave) on two bricks: NCCL halo sweeps (FillBoundary / SumBoundary) and
    particle migration of the C++ driver, 8 steps, against the single-box oracle; particle count conserved."""
    HS = hh.host_simulation_class()
    n, world = 16, 2
    full = workloads.langmuir_3d(n=n)

    def rank_fn(rank, dist):
        dec = parallel.Decomposition((n, n, n), parallel.brick_grid(world), rank)
        sim = HS(full["n_cell"], full["prob_lo"], full["prob_hi"], nox=1, dist=dist, sort_interval=4)
        for s in full["species"]:
            cell = [np.floor((s[k] - full["prob_lo"][d]) / sim.dx[d]).astype(int) for d, k in enumerate("xyz")]
            m = np.ones(len(s["x"]), dtype=bool)
            for d in range(3):
                m &= (cell[d] >= dec.box_lo[d]) & (cell[d] <= dec.box_hi[d])
            sim.add_species(s["name"], s["q"], s["m"], *[s[k][m] for k in ("x", "y", "z", "w", "ux", "uy", "uz")])
        n0 = sim.total_particles()
        sim.Evolve(8)
        assert sim.total_particles() == n0 == 2 * n ** 3
        return dict(fields={c: sim.field_numpy(c) for c in range(9)}, box=(sim.box_lo, sim.box_hi))

    res = hh.run_ranks(world, rank_fn)
    osim = orc.OracleSim(full["n_cell"], full["prob_lo"], full["prob_hi"], nox=1)
    for s in full["species"]:
        osim.add_species(s["q"], s["m"], s["x"], s["y"], s["z"], s["w"], s["ux"], s["uy"], s["uz"])
    osim.evolve(8)
    err, scale = _gather_fields(res, osim, world)
    for c in (0, 1, 2, 6, 7, 8):             # B is round-off in this electrostatic mode
        assert err[c] <= 1e-10 * scale[c], abi.COMP_NAMES[c]


@pytest.mark.parametrize("world", [4])
def test_z_slabs_with_moving_window_match_oracle(orc, hh, world):
    """The laser-acceleration deck in the small (12 x 12 x 64, order 3, filter, PEC z, moving window at c, antenna,
    continuous injection) on slabs along z: every field and every electron (by id) against the single-box oracle."""
    from test_oracle import make_lwfa_oracle
    HS = hh.host_simulation_class()
    nsteps = 10
    wl = workloads.laser_acceleration_3d(n_cell=(12, 12, 64), max_step=nsteps)

    def rank_fn(rank, dist):
        sim = HS(wl["n_cell"], wl["prob_lo"], wl["prob_hi"], nox=wl["nox"], cfl=wl["cfl"], dist=dist, use_filter=True,
                 sort_interval=4, nb=(1, 1, world), boundaries=abi.make_boundaries(wl["field_lo"], wl["field_hi"]),
                 moving_window=(wl["moving_window_dir"], wl["moving_window_v"]))
        sp = wl["species"][0]
        sim.add_plasma_species(sp["name"], sp["q"], sp["m"],
                               abi.make_injector(sp["ppc"], sp["bound_lo"], sp["bound_hi"], sp["density"], True),
                               capacity=12 * 12 * 200)
        la = wl["lasers"][0]
        sim.add_laser(abi.make_laser(la["position"], la["direction"], la["polarization"], la["wavelength"], la["e_max"],
                                     la["waist"], la["duration"], la["t_peak"], la["focal_distance"]))
        sim.Evolve(nsteps)
        return dict(fields={c: sim.field_numpy(c) for c in range(9)}, box=(sim.box_lo, sim.box_hi),
                    parts=sim.species_numpy(0, sort_by_id=False), dz=sim.dx[2], prob=(sim.prob_lo, sim.prob_hi))

    res = hh.run_ranks(world, rank_fn)
    osim = make_lwfa_oracle(orc, wl)
    osim.evolve(nsteps)
    plo, phi = osim.prob_domain()
    assert res[0]["prob"][0] == pytest.approx(plo, rel=0, abs=1e-20) and res[-1]["prob"][1] == pytest.approx(phi, rel=0, abs=1e-20)
    err, scale = _gather_fields(res, osim, world)
    for group in ((0, 1, 2), (3, 4, 5), (6, 7, 8)):
        s = max(scale[c] for c in group)
        assert s > 0
        for c in group:
            assert err[c] <= 1e-11 * s, abi.COMP_NAMES[c]
    B = osim.particles(0)
    ids = np.concatenate([res[r]["parts"]["id"] for r in range(world)])
    order = np.argsort(ids)
    assert np.array_equal(ids[order], np.arange(len(B["x"])))          # global ids: every electron exactly once
    assert sum(len(res[r]["parts"]["x"]) > 0 for r in range(world)) >= 2    # the plasma straddles a slab boundary: migration
    for k in ("x", "y", "z"):
        a = np.concatenate([res[r]["parts"][k] for r in range(world)])[order]
        assert np.max(np.abs(a - B[k])) / res[0]["dz"] <= 1e-10, k
    for k in ("ux", "uy", "uz"):
        a = np.concatenate([res[r]["parts"][k] for r in range(world)])[order]
        assert np.max(np.abs(a - B[k])) / workloads.C <= 1e-10, k
