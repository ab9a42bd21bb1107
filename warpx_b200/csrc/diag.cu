// Reduced diagnostics used as parity metrics (FieldEnergy, ParticleEnergy).
// Source/Diagnostics/ReducedDiags/FieldEnergy.cpp:120-144: MultiFab::norm2(0, periodicity)^2,
// i.e. the sum of squares with every periodic / shared nodal duplicate counted once.
#include "pic_common.cuh"

namespace pic {

__global__ void sumsq_kernel(FabView F, int s0, int s1, int s2, int n0, int n1, int n2, double* out) {
    __shared__ double red[256];
    const long total = (long)n0 * n1 * n2;
    double acc = 0.0;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int a = (int)(t % n0), b = (int)((t / n0) % n1), c = (int)(t / ((long)n0 * n1));
        const double v = F(s0 + a, s1 + b, s2 + c);
        acc += v * v;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

// ParticleEnergy (Source/Diagnostics/ReducedDiags/ParticleEnergy.cpp:86-170): sum of
// w * KineticEnergy(u, m) (Source/Particles/Algorithms/KineticEnergy.H:31-46: m u^2 / (1 + gamma)) and of w.
__global__ void particle_energy_kernel(const double* __restrict__ w, const double* __restrict__ ux,
                                       const double* __restrict__ uy, const double* __restrict__ uz, long np,
                                       double mass, double* out) {
    __shared__ double red_e[256], red_w[256];
    double e = 0.0, ws = 0.0;
    for (long ip = (long)blockIdx.x * blockDim.x + threadIdx.x; ip < np; ip += (long)gridDim.x * blockDim.x) {
        const double u2 = ux[ip] * ux[ip] + uy[ip] * uy[ip] + uz[ip] * uz[ip];
        const double gamma = sqrt(1.0 + u2 * INV_C2);
        e += w[ip] * (1.0 / (1.0 + gamma) * mass * u2);
        ws += w[ip];
    }
    red_e[threadIdx.x] = e; red_w[threadIdx.x] = ws;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { red_e[threadIdx.x] += red_e[threadIdx.x + s]; red_w[threadIdx.x] += red_w[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(out, red_e[0]); atomicAdd(out + 1, red_w[0]); }
}

}  // namespace pic

using namespace pic;

extern "C" int pic_particle_energy(const pic_soa* p, double mass, double* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(out, 0, 2 * sizeof(double), st);
    if (p->np == 0) return 0;
    particle_energy_kernel<<<pic::NUM_SMS * 4, 256, 0, st>>>(p->w, p->ux, p->uy, p->uz, p->np, mass, out);
    count_launch();
    return check_launch("pic_particle_energy") ? 0 : 1;
}

extern "C" int pic_sum_squares_unique(const pic_fab* f, const pic_geom* g, double* out, void* stream) {
    int n[3], s[3];
    for (int d = 0; d < 3; ++d) {
        s[d] = vlo(*f, d);
        int top = vhi(*f, d);
        // the upper nodal layer is a duplicate (periodic image or the neighbour's first node)
        // unless it is the physical upper boundary of a non-periodic domain
        if (f->stag[d] && !(!g->periodic[d] && top == g->n_cell[d])) top -= 1;
        n[d] = top - s[d] + 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(out, 0, sizeof(double), st);
    sumsq_kernel<<<pic::NUM_SMS * 4, 256, 0, st>>>(make_view(*f), s[0], s[1], s[2], n[0], n[1], n[2], out);
    count_launch();
    return check_launch("pic_sum_squares_unique") ? 0 : 1;
}
