// TEST INFRASTRUCTURE -- the order-agnostic kernels (one thread per particle) of warpx_b200/csrc/deposit.cu and
// gather_push.cu under the SIMT emulator of simt_host.h: the code path of particle shape order 4 and of callers
// without cell bins.  Kernel source unmodified; argument sets of pic_deposit_esirkepov / pic_gather_push.
#include "../../warpx_b200/csrc/deposit.cu"
#include "../../warpx_b200/csrc/gather_push.cu"

extern "C" int simt_deposit_global(const pic_soa* p, const pic_fab J[3], const double dinv[3], const double xyzmin[3],
                                   const int lo[3], double q, double dt, double relative_time, int nox) {
    using namespace pic;
    DepositGeom dg;
    for (int d = 0; d < 3; ++d) { dg.dinv[d] = dinv[d]; dg.xyzmin[d] = xyzmin[d]; dg.lo[d] = lo[d]; }
    dg.q = q; dg.dt = dt; dg.tshift = relative_time + 0.5 * dt;
    dg.invdtd[0] = (1.0 / dt) * dinv[1] * dinv[2];
    dg.invdtd[1] = (1.0 / dt) * dinv[0] * dinv[2];
    dg.invdtd[2] = (1.0 / dt) * dinv[0] * dinv[1];
    SoaView P = make_soa(*p, 0);
    const long np = p->np;
    const int tpb = 128;
    const unsigned nblk = (unsigned)((np + tpb - 1) / tpb);
    FabView jx = make_view(J[0]), jy = make_view(J[1]), jz = make_view(J[2]);
#define GO(N) ::simt::launch(dim3(nblk), dim3(tpb), 64, [&] { deposit_global<N>(P, np, jx, jy, jz, dg); })
    if (nox == 1) GO(1); else if (nox == 2) GO(2); else if (nox == 3) GO(3); else if (nox == 4) GO(4); else return 1;
#undef GO
    return 0;
}

extern "C" int simt_gather_push_global(const pic_soa* p, const pic_fab E[3], const pic_fab B[3], const double dinv[3],
                                       const double xyzmin[3], const int lo[3], double q, double m, double dt, int nox,
                                       int galerkin, int pusher, int push_position) {
    using namespace pic;
    GatherGeom gg;
    for (int d = 0; d < 3; ++d) { gg.dinv[d] = dinv[d]; gg.xyzmin[d] = xyzmin[d]; gg.lo[d] = lo[d]; }
    for (int c = 0; c < 3; ++c)
        for (int d = 0; d < 3; ++d) { gg.stag[c][d] = E[c].stag[d]; gg.stag[3 + c][d] = B[c].stag[d]; }
    GlobalFields fld;
    for (int c = 0; c < 3; ++c) { fld.v[c] = make_view(E[c]); fld.v[3 + c] = make_view(B[c]); }
    SoaView P = make_soa(*p, 0);
    const long np = p->np;
    const double qdt2m = 0.5 * q * dt / m;
    const EscapeView esc = make_escape(nullptr, push_position);
    const bool yee = is_yee(E, B);
    const int tpb = 128;
    const unsigned nblk = (unsigned)((np + tpb - 1) / tpb);
#define GO(N, G) do { if (yee) ::simt::launch(dim3(nblk), dim3(tpb), 64, [&] { gather_push_global<N, G, true>(P, np, fld, gg, qdt2m, dt, pusher, push_position, esc, 0); }); \
                      else ::simt::launch(dim3(nblk), dim3(tpb), 64, [&] { gather_push_global<N, G, false>(P, np, fld, gg, qdt2m, dt, pusher, push_position, esc, 0); }); } while (0)
    if (nox == 1 && galerkin) GO(1, 1); else if (nox == 1) GO(1, 0);
    else if (nox == 2 && galerkin) GO(2, 1); else if (nox == 2) GO(2, 0);
    else if (nox == 3 && galerkin) GO(3, 1); else if (nox == 3) GO(3, 0);
    else if (nox == 4 && galerkin) GO(4, 1); else if (nox == 4) GO(4, 0);
    else return 1;
#undef GO
    return 0;
}
