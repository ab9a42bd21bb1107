// TEST INFRASTRUCTURE -- runs warpx_b200/csrc/gather_push_tile.cu (the supercell gather + push kernels, source
// unmodified) under the SIMT emulator of simt_host.h, with the argument set of pic_gather_push (host pointers).
#include "../../warpx_b200/csrc/gather_push_tile.cu"

extern "C" int simt_gather_push(const pic_soa* p, const pic_fab E[3], const pic_fab B[3], const double dinv[3],
                                const double xyzmin[3], const int lo[3], double q, double m, double dt, int nox,
                                int galerkin, int pusher, int push_position, const pic_bins* bins, int mode) {
    pic::g_gather_mode = mode;
    pic::GatherGeom gg;
    for (int d = 0; d < 3; ++d) { gg.dinv[d] = dinv[d]; gg.xyzmin[d] = xyzmin[d]; gg.lo[d] = lo[d]; }
    for (int c = 0; c < 3; ++c)
        for (int d = 0; d < 3; ++d) { gg.stag[c][d] = E[c].stag[d]; gg.stag[3 + c][d] = B[c].stag[d]; }
    const pic::EscapeView esc = pic::make_escape(nullptr, push_position);
    return pic::gather_push_tile_launch(p, 0, p->np, E, B, gg, 0.5 * q * dt / m, dt, nox, galerkin, pusher,
                                        push_position, bins, esc, nullptr);
}
