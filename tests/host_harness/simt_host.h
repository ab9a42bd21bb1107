// TEST INFRASTRUCTURE -- a small SIMT emulator for the warp-level kernels (g++ only, never part of the
// product).  The authoring container has no GPU; this header lets the CPU test-suite run the UNMODIFIED
// kernel source of e.g. warpx_b200/csrc/deposit_runs.cu: every CUDA thread becomes a cooperative fiber
// (ucontext), the 32 lanes of a warp run in lock step between warp collectives (__shfl_*_sync,
// __ballot_sync, __syncwarp), __syncthreads is a barrier over the fibers of the block, shared memory
// is a per-block buffer, atomics are plain updates (one OS thread).  What it checks: indexing, lane
// roles, run / segment logic, arithmetic.  What it cannot check: memory-model races, launch bounds,
// performance.  Force-included (-include) before the kernel source, with -DPIC_SIMT_HOST.
#ifndef PIC_SIMT_HOST_H_
#define PIC_SIMT_HOST_H_
#ifndef PIC_SIMT_HOST
#error "simt_host.h is for -DPIC_SIMT_HOST builds"
#endif

#include <cuda_runtime.h>      // vector types, host API declarations (dim3, double2, cudaStream_t ...)
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <string>
#include <vector>

namespace simt {

struct Barrier {
    int expected = 0, arrived = 0;
    unsigned long gen = 0;
};

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;           // from the per-thread pool below (reused by every block, never zero-filled)
    bool done = false;
    int warp = 0, lane = 0, tid = 0;
};
constexpr size_t STACK_BYTES = 1 << 17;
inline char* pooled_stack(int tid) {
    static thread_local std::vector<char*> pool;
    while ((int)pool.size() <= tid) pool.push_back(static_cast<char*>(std::malloc(STACK_BYTES)));
    return pool[(size_t)tid];
}

struct Block {
    std::vector<Fiber> fibers;
    std::vector<Barrier> warp_bar;       // one per warp
    Barrier block_bar;
    std::vector<char> smem;
    ucontext_t sched;
    int current = -1;
    dim3 block_idx, block_dim, grid_dim;
    // collective scratch, one slot per lane of a warp
    std::vector<std::vector<unsigned long long>> slot;   // [warp][32]
    std::function<void()> body;
    std::map<std::string, std::vector<char>> statics;    // __shared__ variables declared inside the kernel
};

inline Block*& cur_block() { static thread_local Block* b = nullptr; return b; }
inline Fiber& me() { Block* b = cur_block(); return b->fibers[(size_t)b->current]; }

inline void yield() {                        // back to the scheduler
    Block* b = cur_block();
    swapcontext(&b->fibers[(size_t)b->current].ctx, &b->sched);
}

// wait until every live participant has arrived (cooperative: keep yielding until the generation turns)
inline void barrier_wait(Barrier& bar) {
    const unsigned long g = bar.gen;
    if (++bar.arrived == bar.expected) { bar.arrived = 0; ++bar.gen; return; }
    while (bar.gen == g) yield();
}

inline void trampoline() {
    Block* b = cur_block();
    b->body();
    Fiber& f = b->fibers[(size_t)b->current];
    f.done = true;
    // a finished lane no longer takes part in barriers (the kernels under test exit warp-uniformly)
    b->warp_bar[(size_t)f.warp].expected -= 1;
    b->block_bar.expected -= 1;
    swapcontext(&f.ctx, &b->sched);
}

inline void* dynamic_smem() { return cur_block()->smem.data(); }
// a statically declared __shared__ array: one zero-initialised buffer per block, keyed by its name
inline void* block_static(const char* name, size_t bytes) {
    std::vector<char>& v = cur_block()->statics[name];
    if (v.size() < bytes) v.assign(bytes, 0);
    return v.data();
}

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& kernel_call) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const int nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                Block blk;
                blk.block_idx = dim3(bx, by, bz); blk.block_dim = block; blk.grid_dim = grid;
                blk.smem.assign(smem_bytes + 64, 0);
                blk.fibers.resize((size_t)nthreads);
                blk.warp_bar.resize((size_t)nwarps);
                blk.slot.assign((size_t)nwarps, std::vector<unsigned long long>(32, 0ull));
                blk.block_bar.expected = nthreads;
                blk.body = kernel_call;
                cur_block() = &blk;
                for (int t = 0; t < nthreads; ++t) {
                    Fiber& f = blk.fibers[(size_t)t];
                    f.tid = t; f.warp = t / 32; f.lane = t % 32;
                    blk.warp_bar[(size_t)f.warp].expected += 1;
                    f.stack = pooled_stack(t);
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = STACK_BYTES;
                    f.ctx.uc_link = &blk.sched;
                    makecontext(&f.ctx, (void (*)())trampoline, 0);
                }
                int live = nthreads;
                while (live > 0) {
                    live = 0;
                    for (int t = 0; t < nthreads; ++t) {
                        if (blk.fibers[(size_t)t].done) continue;
                        blk.current = t;
                        swapcontext(&blk.sched, &blk.fibers[(size_t)t].ctx);
                        if (!blk.fibers[(size_t)t].done) ++live;
                    }
                }
                cur_block() = nullptr;
            }
}

// ---- warp collectives -----------------------------------------------------------------------
template <class T>
inline T exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "collective payload");
    Block* b = cur_block();
    Fiber& f = me();
    unsigned long long raw = 0;
    __builtin_memcpy(&raw, &v, sizeof(T));
    b->slot[(size_t)f.warp][(size_t)f.lane] = raw;
    barrier_wait(b->warp_bar[(size_t)f.warp]);          // everybody wrote
    T out = v;
    if (src_lane >= 0 && src_lane < 32) {
        const unsigned long long r = b->slot[(size_t)f.warp][(size_t)src_lane];
        __builtin_memcpy(&out, &r, sizeof(T));
    }
    barrier_wait(b->warp_bar[(size_t)f.warp]);          // everybody read
    return out;
}

}  // namespace simt

// ---- the CUDA surface the kernels use ---------------------------------------------------------
struct SimtIdx { unsigned x, y, z; };
#define threadIdx (SimtIdx{(unsigned)::simt::me().tid % ::simt::cur_block()->block_dim.x, \
                           ((unsigned)::simt::me().tid / ::simt::cur_block()->block_dim.x) % ::simt::cur_block()->block_dim.y, \
                           (unsigned)::simt::me().tid / (::simt::cur_block()->block_dim.x * ::simt::cur_block()->block_dim.y)})
#define blockIdx (::simt::cur_block()->block_idx)
#define blockDim (::simt::cur_block()->block_dim)
#define gridDim (::simt::cur_block()->grid_dim)

template <class T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return ::simt::exchange(v, src); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    const int l = ::simt::me().lane;
    return ::simt::exchange(v, l - (int)d >= 0 ? l - (int)d : l);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    const int l = ::simt::me().lane;
    return ::simt::exchange(v, l + (int)d < 32 ? l + (int)d : l);
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return ::simt::exchange(v, ::simt::me().lane ^ m); }
inline unsigned __ballot_sync(unsigned, int pred) {
    simt::Block* b = simt::cur_block();
    simt::Fiber& f = simt::me();
    b->slot[(size_t)f.warp][(size_t)f.lane] = pred ? 1ull : 0ull;
    simt::barrier_wait(b->warp_bar[(size_t)f.warp]);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) if (b->slot[(size_t)f.warp][(size_t)l]) m |= (1u << l);
    // lanes beyond the block size / finished lanes contribute 0
    simt::barrier_wait(b->warp_bar[(size_t)f.warp]);
    return m;
}
inline unsigned __match_any_sync(unsigned, int value) {
    simt::Block* b = simt::cur_block();
    simt::Fiber& f = simt::me();
    b->slot[(size_t)f.warp][(size_t)f.lane] = (unsigned long long)(unsigned)value | (1ull << 40);   // bit 40: the lane took part
    simt::barrier_wait(b->warp_bar[(size_t)f.warp]);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) if (b->slot[(size_t)f.warp][(size_t)l] == ((unsigned long long)(unsigned)value | (1ull << 40))) m |= (1u << l);
    simt::barrier_wait(b->warp_bar[(size_t)f.warp]);
    // the slots of lanes that do not exist / already left must not match on the next collective
    b->slot[(size_t)f.warp][(size_t)f.lane] = 0;
    return m;
}
inline unsigned __activemask() { return 0xffffffffu; }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::barrier_wait(simt::cur_block()->warp_bar[(size_t)simt::me().warp]); }
inline void __syncthreads() { simt::barrier_wait(simt::cur_block()->block_bar); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline double atomicAdd(double* p, double v) { const double o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline int atomicMax(int* p, int v) { const int o = *p; if (v > o) *p = v; return o; }
inline int atomicMin(int* p, int v) { const int o = *p; if (v < o) *p = v; return o; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline double __longlong_as_double(long long v) { double d; __builtin_memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; __builtin_memcpy(&v, &d, 8); return v; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline double min(double a, double b) { return a < b ? a : b; }
inline double max(double a, double b) { return a > b ? a : b; }

#undef __launch_bounds__
#define __launch_bounds__(...)

// ---- what tests/host_harness/cuda2host.py turns `kernel<<<grid, block, smem, stream>>>(args)` into ----
#define SIMT_LAUNCH(grid, block, smem, stream, ...) \
    ::simt::launch(dim3(grid), dim3(block), (size_t)(smem) + 64, [&] { __VA_ARGS__; })

// ---- CUDA runtime calls of the host code: "device" memory is host memory, streams do not exist ----
namespace simt {
inline cudaError_t Malloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t Malloc(T** p, size_t n) { return Malloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t Free(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t Memset(void* p, int v, size_t n, cudaStream_t = nullptr) { __builtin_memset(p, v, n); return cudaSuccess; }
inline cudaError_t Memcpy(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { __builtin_memmove(d, s, n); return cudaSuccess; }
inline cudaError_t Sync(cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t LastError() { return cudaSuccess; }
}  // namespace simt
#define cudaMalloc ::simt::Malloc
#define cudaMallocHost ::simt::Malloc
#define cudaFree ::simt::Free
#define cudaFreeHost ::simt::Free
#define cudaMemsetAsync ::simt::Memset
#define cudaMemcpyAsync ::simt::Memcpy
#define cudaStreamSynchronize ::simt::Sync
#define cudaDeviceSynchronize ::simt::Sync
#define cudaGetLastError ::simt::LastError
#define cudaFuncSetAttribute(...) cudaSuccess
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

// ---- the one CUB call of the sort (two-phase API) ----
namespace cub {
struct DeviceScan {
    template <class In, class Out>
    static cudaError_t ExclusiveSum(void* temp, size_t& temp_bytes, In in, Out out, int n, cudaStream_t = nullptr) {
        if (temp == nullptr) { temp_bytes = 256; return cudaSuccess; }
        long run = 0;
        for (int i = 0; i < n; ++i) { const auto v = in[i]; out[i] = (decltype(v))run; run += v; }
        return cudaSuccess;
    }
};
}  // namespace cub
#define PIC_SIMT_NO_CUB_INCLUDE 1

#endif
