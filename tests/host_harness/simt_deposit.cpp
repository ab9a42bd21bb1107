// TEST INFRASTRUCTURE -- runs warpx_b200/csrc/deposit_runs.cu (kernel source unmodified) under the SIMT
// emulator of simt_host.h and exposes it with the argument set of pic_deposit_esirkepov (host pointers).
#include "../../warpx_b200/csrc/deposit_runs.cu"

extern "C" int simt_deposit_runs(const pic_soa* p, long offset, long np, const pic_fab J[3], const double dinv[3],
                                 const double xyzmin[3], const int lo[3], double q, double dt, double relative_time,
                                 int nox, int variant) {
    pic::g_runs_variant = variant;
    pic::DepositGeom dg;
    for (int d = 0; d < 3; ++d) { dg.dinv[d] = dinv[d]; dg.xyzmin[d] = xyzmin[d]; dg.lo[d] = lo[d]; }
    dg.q = q; dg.dt = dt; dg.tshift = relative_time + 0.5 * dt;
    dg.invdtd[0] = (1.0 / dt) * dinv[1] * dinv[2];
    dg.invdtd[1] = (1.0 / dt) * dinv[0] * dinv[2];
    dg.invdtd[2] = (1.0 / dt) * dinv[0] * dinv[1];
    return pic::deposit_runs_launch(p, offset, np, J, dg, nox, nullptr);
}
