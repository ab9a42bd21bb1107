// Neighbour migration of particles between ranks, entirely on the device.
//
// Replaces the pack / unpack phases of AMReX ParticleContainer::Redistribute as used by
// WarpX::HandleParticlesAtBoundaries (Source/Evolve/WarpXEvolve.cpp:550-559, RedistributeLocal(1)).
// Particles move less than one cell per step, so only a thin layer next to a brick face leaves.
// Design: only those particles are touched.
//   1. pic_particles_classify (particles_misc.cu) lists the indices leaving to the low / high
//      neighbour (device counters).
//   2. pic_migrate_pack gathers them into a FIXED-SIZE message (count in the header), so the NCCL
//      send/recv sizes are known to the host without a round trip.
//   3. pic_migrate_unpack drops the arrivals into the holes the departures left, appends the rest,
//      or -- when more left than arrived -- moves tail particles into the remaining holes, and
//      writes the new particle count to device memory (the host reads it once per sweep).
// Particle order -- and therefore the cell bins -- stays valid for everything that did not move.
#include "pic_common.cuh"

namespace pic {

constexpr int MSG_HEADER = 8;   // doubles; [0] = particle count
constexpr int MSG_ROWS = 8;     // x y z w ux uy uz id(bit pattern)

struct SoaPtrs { double* a[7]; uint64_t* id; };
static SoaPtrs soa_ptrs(const pic_soa& p) {
    SoaPtrs s;
    s.a[0] = p.x; s.a[1] = p.y; s.a[2] = p.z; s.a[3] = p.w; s.a[4] = p.ux; s.a[5] = p.uy; s.a[6] = p.uz;
    s.id = p.idcpu;
    return s;
}

__global__ void pack_kernel(SoaPtrs P, const int* __restrict__ idx, const int* __restrict__ count, int cap,
                            double* __restrict__ msg) {
    const int n = min(*count, cap);
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) msg[0] = (double)(*count);      // the true count: the receiver detects overflow
    if (j >= n) return;
    const int ip = idx[j];
    double* body = msg + MSG_HEADER;
#pragma unroll
    for (int r = 0; r < 7; ++r) body[(long)r * cap + j] = P.a[r][ip];
    body[(long)7 * cap + j] = P.id ? __longlong_as_double((long long)P.id[ip]) : 0.0;
}

// work layout (ints): [0] np_new  [1] status (sticky)  [2] #survivors  [3] #low holes  [4] np_old  [5] np_new
//                     [8, 8+2cap) survivors   [.., +2cap) low holes   [.., +2cap) tail marks
struct UnpackArgs {
    SoaPtrs P;
    const int* counts;        // leaving: [0] low, [1] high
    const int* idx_lo; const int* idx_hi;
    const double* msg_lo; const double* msg_hi;   // arrivals from the low / high neighbour
    int cap; long np_host; const int* np_dev; long capacity;
    int* work;
    __device__ __forceinline__ long np_old() const { return np_dev ? (long)*np_dev : np_host; }
};

__device__ __forceinline__ int hole_index(const UnpackArgs& a, int j, int n0) {
    return j < n0 ? a.idx_lo[j] : a.idx_hi[j - n0];
}

__device__ __forceinline__ void write_particle(const SoaPtrs& P, long dst, const double* body, int cap, int j) {
#pragma unroll
    for (int r = 0; r < 7; ++r) P.a[r][dst] = body[(long)r * cap + j];
    if (P.id) P.id[dst] = (uint64_t)__double_as_longlong(body[(long)7 * cap + j]);
}

// arrivals -> holes, then appended at the end
__global__ void unpack_fill_kernel(UnpackArgs a) {
    const int n0 = min(a.counts[0], a.cap), n1 = min(a.counts[1], a.cap);
    const int r0 = (int)a.msg_lo[0], r1 = (int)a.msg_hi[0];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_holes = n0 + n1, n_arr = min(r0, a.cap) + min(r1, a.cap);
    const long np_old = a.np_old();
    if (j == 0) {
        int status = 0;
        if (a.counts[0] > a.cap || a.counts[1] > a.cap || r0 > a.cap || r1 > a.cap) status |= 1;   // list overflow
        const long np_new = np_old + (long)n_arr - (long)n_holes;
        if (np_new > a.capacity) status |= 2;                                                 // capacity
        a.work[4] = (int)np_old;     // kept for the follow-up kernels (work[0] may alias np_dev)
        a.work[5] = (int)np_new;
        a.work[1] |= status;
        a.work[2] = 0; a.work[3] = 0;
    }
    if (j >= n_arr) return;
    const int rr0 = min(r0, a.cap);
    const double* body = (j < rr0 ? a.msg_lo : a.msg_hi) + MSG_HEADER;
    const int jj = j < rr0 ? j : j - rr0;
    long dst;
    if (j < n_holes) dst = hole_index(a, j, n0);
    else dst = np_old + (j - n_holes);
    if (dst < a.capacity) write_particle(a.P, dst, body, a.cap, jj);
}

// more departures than arrivals: mark the open holes that lie in the tail [np_new, np_old)
__global__ void mark_tail_kernel(UnpackArgs a) {
    const int n0 = min(a.counts[0], a.cap), n1 = min(a.counts[1], a.cap);
    const int n_holes = n0 + n1, n_arr = min((int)a.msg_lo[0], a.cap) + min((int)a.msg_hi[0], a.cap);
    const int m = n_holes - n_arr;
    if (m <= 0) return;
    const long np_new = (long)a.work[4] - m;
    int* survivors = a.work + 8; int* lows = survivors + 2 * a.cap; int* marks = lows + 2 * a.cap;
    (void)survivors;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int h = hole_index(a, n_arr + t, n0);         // an open hole
    if (h >= np_new) marks[h - np_new] = 1;
    else lows[atomicAdd(&a.work[3], 1)] = h;
}
__global__ void list_survivors_kernel(UnpackArgs a) {
    const int n0 = min(a.counts[0], a.cap), n1 = min(a.counts[1], a.cap);
    const int m = n0 + n1 - (min((int)a.msg_lo[0], a.cap) + min((int)a.msg_hi[0], a.cap));
    if (m <= 0) return;
    const long np_new = (long)a.work[4] - m;
    int* survivors = a.work + 8; int* marks = survivors + 4 * a.cap;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    if (!marks[t]) survivors[atomicAdd(&a.work[2], 1)] = (int)(np_new + t);
}
__global__ void move_survivors_kernel(UnpackArgs a) {
    const int n = a.work[2];                              // == work[3]
    int* survivors = a.work + 8; int* lows = survivors + 2 * a.cap;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) a.work[0] = a.work[5];                    // publish the new count last
    if (s >= n) return;
    const int src = survivors[s], dst = lows[s];
#pragma unroll
    for (int r = 0; r < 7; ++r) a.P.a[r][dst] = a.P.a[r][src];
    if (a.P.id) a.P.id[dst] = a.P.id[src];
}

}  // namespace pic

using namespace pic;

extern "C" long pic_migrate_message_doubles(int cap) { return (long)MSG_HEADER + (long)MSG_ROWS * cap; }
extern "C" long pic_migrate_workspace_bytes(int cap) { return (long)sizeof(int) * (8 + 6L * cap); }

extern "C" int pic_migrate_pack(const pic_soa* p, const int* idx, const int* count, int cap, double* msg,
                                void* stream) {
    pack_kernel<<<(cap + 255) / 256, 256, 0, (cudaStream_t)stream>>>(soa_ptrs(*p), idx, count, cap, msg);
    count_launch();
    return check_launch("pic_migrate_pack") ? 0 : 1;
}

extern "C" int pic_migrate_unpack(const pic_soa* p, const int* counts, const int* idx_lo, const int* idx_hi,
                                  const double* msg_lo, const double* msg_hi, int cap, long capacity,
                                  void* work, const int* np_dev, void* stream) {
    UnpackArgs a;
    a.P = soa_ptrs(*p); a.counts = counts; a.idx_lo = idx_lo; a.idx_hi = idx_hi;
    a.msg_lo = msg_lo; a.msg_hi = msg_hi; a.cap = cap; a.np_host = p->np; a.np_dev = np_dev; a.capacity = capacity;
    a.work = (int*)work;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g2 = (unsigned)((2L * cap + 255) / 256);
    cudaMemsetAsync((int*)work + 8 + 4L * cap, 0, sizeof(int) * (size_t)(2L * cap), s);     // tail marks
    unpack_fill_kernel<<<g2, 256, 0, s>>>(a);
    mark_tail_kernel<<<g2, 256, 0, s>>>(a);
    list_survivors_kernel<<<g2, 256, 0, s>>>(a);
    move_survivors_kernel<<<g2, 256, 0, s>>>(a);
    count_launch(4);
    return check_launch("pic_migrate_unpack") ? 0 : 1;
}
