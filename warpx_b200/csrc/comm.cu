// NCCL transport of the C++ step driver: one communicator per process (one process per GPU).
//
// Replaces the MPI layer under ablastr::utils::communication::FillBoundary / SumBoundary
// (Source/ablastr/utils/Communication.cpp:71-175) and AMReX ParticleContainer::Redistribute
// (Source/Evolve/WarpXEvolve.cpp:550-559) for a node of NVSwitch-connected GPUs.  NCCL is bound at
// run time (dlopen of the libnccl.so.2 already loaded by the host application, e.g. PyTorch's), so
// the library itself has no link-time dependency and single-GPU use needs no NCCL at all.
// Bootstrap: rank 0 calls pic_comm_unique_id, the 128-byte id reaches the other ranks by whatever
// out-of-band channel the host has (MPI_Bcast in WarpX, torch.distributed here), then every rank
// calls pic_comm_create.
#include "pic_common.cuh"
#include "comm.cuh"
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>

namespace pic {

NcclApi g_nccl;

static void* nccl_handle() {
    static void* h = nullptr;
    if (h) return h;
    // PIC_NCCL_LIBRARY: bind this library instead of the libnccl.so.2 already loaded / found by the loader
    if (const char* path = getenv("PIC_NCCL_LIBRARY")) { h = dlopen(path, RTLD_NOW | RTLD_GLOBAL); return h; }
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (h) return h; }   // already loaded?
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) return h; }
    return nullptr;
}

bool nccl_load() {
    if (g_nccl.loaded) return true;
    void* h = nccl_handle();
    if (!h) return false;
#define PIC_SYM(field, name) do { *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return false; } while (0)
    PIC_SYM(GetUniqueId, "ncclGetUniqueId");
    PIC_SYM(CommInitRank, "ncclCommInitRank");
    PIC_SYM(CommDestroy, "ncclCommDestroy");
    PIC_SYM(Send, "ncclSend");
    PIC_SYM(Recv, "ncclRecv");
    PIC_SYM(GroupStart, "ncclGroupStart");
    PIC_SYM(GroupEnd, "ncclGroupEnd");
    PIC_SYM(AllReduce, "ncclAllReduce");
    PIC_SYM(GetErrorString, "ncclGetErrorString");
#undef PIC_SYM
    g_nccl.loaded = true;
    return true;
}

int nccl_fail(const char* what, int rc) {
    return fail("%s: NCCL error %d (%s)", what, rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
}

}  // namespace pic

using namespace pic;

extern "C" int pic_comm_unique_id(unsigned char out[128]) {
    if (!nccl_load()) return fail("pic_comm_unique_id: libnccl.so.2 not found");
    PicNcclUniqueId id;
    if (int rc = g_nccl.GetUniqueId(&id)) return nccl_fail("pic_comm_unique_id", rc);
    std::memcpy(out, id.internal, 128);
    return 0;
}

extern "C" void* pic_comm_create(const unsigned char id_bytes[128], int nranks, int rank) {
    if (!nccl_load()) { fail("pic_comm_create: libnccl.so.2 not found"); return nullptr; }
    PicNcclUniqueId id;
    std::memcpy(id.internal, id_bytes, 128);
    Comm* c = new Comm();
    c->nranks = nranks; c->rank = rank; c->comm = nullptr;
    if (int rc = g_nccl.CommInitRank(&c->comm, nranks, id, rank)) {
        nccl_fail("pic_comm_create", rc);
        delete c;
        return nullptr;
    }
    return c;
}

extern "C" void pic_comm_destroy(void* h) {
    Comm* c = static_cast<Comm*>(h);
    if (!c) return;
    if (c->comm && g_nccl.loaded) g_nccl.CommDestroy(c->comm);
    delete c;
}
