// TEST INFRASTRUCTURE -- an in-process stand-in for the eight NCCL entry points the step driver binds at run time
// (warpx_b200/csrc/comm.cu), for the CPU test-suite: the "ranks" are THREADS of one process, each driving its own
// engine instance of the host library (tests/host_harness/harness.py), and exchange through mailboxes.
// Semantics kept: messages between a pair of ranks arrive in the order they were sent; sends are buffered, receives
// posted inside a group complete at ncclGroupEnd (outside a group, immediately); ncclAllReduce(max / sum) over all
// ranks of the communicator.  Select with PIC_NCCL_LIBRARY=<this library>.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <vector>

namespace {
struct Message { std::vector<char> data; };
struct World {
    std::mutex mu;
    std::condition_variable cv;
    std::map<std::pair<int, int>, std::deque<Message>> box;     // (src, dst) -> FIFO
    // all-reduce rendezvous
    int arrived = 0, departed = 0;
    unsigned long generation = 0;
    std::vector<char> acc;
};
std::mutex g_mu;
std::map<std::vector<char>, World*> g_worlds;                   // by unique id
int g_next_id = 1;

struct Comm { World* w; int nranks, rank; };
struct PendingRecv { void* p; size_t bytes; int peer; Comm* c; };
thread_local int t_group = 0;
thread_local std::vector<PendingRecv> t_pending;

size_t dtype_bytes(int dt) { return dt == 2 ? 4 : dt == 8 ? 8 : dt == 7 ? 4 : 1; }   // int32, float64, float32

// a peer that died (a failed assertion in its thread) must not hang the test-suite: give up after two minutes
constexpr std::chrono::seconds PATIENCE(120);

int complete(const PendingRecv& r) {
    World* w = r.c->w;
    std::unique_lock<std::mutex> lk(w->mu);
    auto& q = w->box[{r.peer, r.c->rank}];
    if (!w->cv.wait_for(lk, PATIENCE, [&] { return !q.empty(); })) return 1;
    Message m = std::move(q.front());
    q.pop_front();
    std::memcpy(r.p, m.data.data(), std::min(r.bytes, m.data.size()));
    return 0;
}
}  // namespace

extern "C" {
struct FakeId { char internal[128]; };

int ncclGetUniqueId(FakeId* id) {
    std::lock_guard<std::mutex> lk(g_mu);
    std::memset(id->internal, 0, 128);
    std::memcpy(id->internal, &g_next_id, sizeof(int));
    ++g_next_id;
    return 0;
}
int ncclCommInitRank(void** comm, int nranks, FakeId id, int rank) {
    std::lock_guard<std::mutex> lk(g_mu);
    std::vector<char> key(id.internal, id.internal + 128);
    World*& w = g_worlds[key];
    if (!w) w = new World();
    *comm = new Comm{w, nranks, rank};
    return 0;
}
int ncclCommDestroy(void* comm) { delete static_cast<Comm*>(comm); return 0; }
int ncclGroupStart() { ++t_group; return 0; }
int ncclGroupEnd() {
    if (--t_group == 0) {
        std::vector<PendingRecv> todo;
        todo.swap(t_pending);
        int rc = 0;
        for (const auto& r : todo) rc |= complete(r);
        return rc;
    }
    return 0;
}
int ncclSend(const void* p, size_t count, int dtype, int peer, void* comm, void*) {
    Comm* c = static_cast<Comm*>(comm);
    Message m;
    m.data.assign(static_cast<const char*>(p), static_cast<const char*>(p) + count * dtype_bytes(dtype));
    {
        std::lock_guard<std::mutex> lk(c->w->mu);
        c->w->box[{c->rank, peer}].push_back(std::move(m));
    }
    c->w->cv.notify_all();
    return 0;
}
int ncclRecv(void* p, size_t count, int dtype, int peer, void* comm, void*) {
    PendingRecv r{p, count * dtype_bytes(dtype), peer, static_cast<Comm*>(comm)};
    if (t_group > 0) { t_pending.push_back(r); return 0; }
    return complete(r);
}
int ncclAllReduce(const void* in, void* out, size_t count, int dtype, int op, void* comm, void*) {
    Comm* c = static_cast<Comm*>(comm);
    World* w = c->w;
    const size_t bytes = count * dtype_bytes(dtype);
    std::unique_lock<std::mutex> lk(w->mu);
    if (!w->cv.wait_for(lk, PATIENCE, [&] { return w->departed == 0 || w->arrived > 0; })) return 1;   // previous round fully drained
    const unsigned long gen = w->generation;
    if (w->arrived == 0) w->acc.assign(static_cast<const char*>(in), static_cast<const char*>(in) + bytes);
    else
        for (size_t i = 0; i < count; ++i) {
            if (dtype == 2) { int* a = reinterpret_cast<int*>(w->acc.data()); const int v = static_cast<const int*>(in)[i];
                              a[i] = op == 2 ? (v > a[i] ? v : a[i]) : a[i] + v; }
            else { double* a = reinterpret_cast<double*>(w->acc.data()); const double v = static_cast<const double*>(in)[i];
                   a[i] = op == 2 ? (v > a[i] ? v : a[i]) : a[i] + v; }
        }
    if (++w->arrived == c->nranks) { w->departed = c->nranks; w->arrived = 0; ++w->generation; w->cv.notify_all(); }
    else if (!w->cv.wait_for(lk, PATIENCE, [&] { return w->generation != gen; })) return 1;
    std::memcpy(out, w->acc.data(), bytes);
    if (--w->departed == 0) w->cv.notify_all();
    return 0;
}
const char* ncclGetErrorString(int) { return "stand-in NCCL: a peer did not answer"; }
}
