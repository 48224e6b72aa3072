// Multi-GPU plumbing of libpdeb200.so (include/pdeb200.h "multi-GPU"; SURVEY.md 8b `comm_init`, 8e).
//
// One process per GPU.  pdeb200_comm_init() joins the ranks with NCCL (ncclCommInitRank on the caller-distributed
// unique id -- bootstrap, and the transport of last resort), then sets up the peer-memory exchange of comm.cuh:
// every rank allocates an exchange buffer, the cudaIpc handles are all-gathered through NCCL, and each rank maps
// its peers' buffers (NVLink 5 / NVSwitch peer access).  From then on pdeb200_sample / pdeb200_ddpg_update /
// pdeb200_train_updates are collectives: the statistics and gradient sums are exchanged INSIDE their kernels.
//
// NCCL is resolved with dlopen at comm_init time (the copy already loaded in the process -- torch's bundled
// libnccl.so.2, or NCCL_jll's under Julia -- else $PDEB200_NCCL_LIB, else the system libnccl.so.2), so the library has
// no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "comm.cuh"
#include "ctx.hpp"

namespace pdeb200 {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string err;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy the host process already uses
    if (!h) {
        const char* e = getenv("PDEB200_NCCL_LIB");
        if (e && *e) h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
    }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return &api; }
    api.handle = h;
#define PDEB_SYM(field, name)                                                       \
    do {                                                                            \
        *(void**)(&api.field) = dlsym(h, name);                                     \
        if (!api.field) { api.err = std::string("libnccl.so.2 lacks ") + name; api.handle = nullptr; return &api; } \
    } while (0)
    PDEB_SYM(GetUniqueId, "ncclGetUniqueId");
    PDEB_SYM(CommInitRank, "ncclCommInitRank");
    PDEB_SYM(CommDestroy, "ncclCommDestroy");
    PDEB_SYM(AllReduce, "ncclAllReduce");
    PDEB_SYM(AllGather, "ncclAllGather");
    PDEB_SYM(GetErrorString, "ncclGetErrorString");
    PDEB_SYM(GetVersion, "ncclGetVersion");
#undef PDEB_SYM
    return &api;
}

struct Comm {
    ncclComm_t nccl = nullptr;
    int rank = 0, nranks = 1;
    int transport = PDEB200_COMM_NONE;
    char* buf = nullptr; size_t buf_bytes = 0;
    char* peer[kMaxRanks] = {nullptr};
    unsigned int* epoch = nullptr;
    int* err = nullptr;
    int cap = 0;
    unsigned long long timeout_ns = 0;
    double* scratch = nullptr;           // 64 doubles for the host-facing scalar allreduce
};

Comm* cm(const pdeb200_ctx* c) { return static_cast<Comm*>(c->comm); }

#define PDEB_NCCL(ctx, expr)                                                                              \
    do {                                                                                                  \
        ncclResult_t r__ = (expr);                                                                        \
        if (r__ != ncclSuccess)                                                                           \
            return fail(ctx, PDEB200_ECOMM, std::string(#expr) + ": " + nccl_api()->GetErrorString(r__)); \
    } while (0)

}  // namespace

CommDev comm_dev(const pdeb200_ctx* c) {
    CommDev d;
    const Comm* m = cm(c);
    if (!m || m->transport != PDEB200_COMM_PEER) return d;
    d.rank = m->rank; d.nranks = m->nranks; d.cap = m->cap; d.epoch = m->epoch; d.err = m->err; d.timeout_ns = m->timeout_ns;
    for (int r = 0; r < m->nranks; ++r) d.peer[r] = m->peer[r];
    return d;
}

int comm_nranks(const pdeb200_ctx* c) { return cm(c) ? cm(c)->nranks : 1; }
int comm_transport(const pdeb200_ctx* c) { return cm(c) ? cm(c)->transport : PDEB200_COMM_NONE; }
int comm_cap(const pdeb200_ctx* c) { return cm(c) ? cm(c)->cap : 0; }

int32_t comm_allreduce_f32(pdeb200_ctx* c, float* dev, size_t n) {
    Comm* m = cm(c);
    if (!m || m->nranks <= 1) return PDEB200_OK;
    PDEB_NCCL(c, nccl_api()->AllReduce(dev, dev, n, ncclFloat32, ncclSum, m->nccl, c->stream));
    return PDEB200_OK;
}

int32_t comm_allreduce_f64(pdeb200_ctx* c, double* dev, size_t n) {
    Comm* m = cm(c);
    if (!m || m->nranks <= 1) return PDEB200_OK;
    PDEB_NCCL(c, nccl_api()->AllReduce(dev, dev, n, ncclFloat64, ncclSum, m->nccl, c->stream));
    return PDEB200_OK;
}

int32_t comm_check(pdeb200_ctx* c) {
    Comm* m = cm(c);
    if (!m || !m->err) return PDEB200_OK;
    int e = 0;
    PDEB_CUDA(c, cudaMemcpyAsync(&e, m->err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (e) return fail(c, PDEB200_ECOMM, "peer exchange timed out (a rank died or never issued the matching collective call)");
    return PDEB200_OK;
}

void comm_free(pdeb200_ctx* c) {
    Comm* m = cm(c);
    if (!m) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int r = 0; r < m->nranks; ++r)
        if (r != m->rank && m->peer[r]) cudaIpcCloseMemHandle(m->peer[r]);
    for (void* p : {(void*)m->buf, (void*)m->epoch, (void*)m->err, (void*)m->scratch})
        if (p) cudaFree(p);
    if (m->nccl) nccl_api()->CommDestroy(m->nccl);
    delete m;
    c->comm = nullptr;
}

}  // namespace pdeb200

using namespace pdeb200;

extern "C" {

int32_t pdeb200_comm_unique_id(uint8_t* id128) {
    if (!id128) return fail(nullptr, PDEB200_EINVAL, "comm_unique_id: null argument");
    NcclApi* api = nccl_api();
    if (!api->handle) return fail(nullptr, PDEB200_ECOMM, api->err);
    static_assert(sizeof(ncclUniqueId) == PDEB200_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, PDEB200_ECOMM, std::string("ncclGetUniqueId: ") + api->GetErrorString(r));
    std::memcpy(id128, &id, sizeof(id));
    return PDEB200_OK;
}

int32_t pdeb200_comm_init(pdeb200_ctx* c, const uint8_t* id128, int32_t rank, int32_t nranks) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(c, PDEB200_EINVAL, "comm_init: bad argument");
    if (c->comm) return fail(c, PDEB200_ESTATE, "comm_init: communicator already initialised (pdeb200_comm_destroy first)");
    cudaSetDevice(c->device);
    NcclApi* api = nccl_api();
    if (!api->handle) return fail(c, PDEB200_ECOMM, api->err);
    Comm* m = new Comm();
    c->comm = m;
    m->rank = rank; m->nranks = nranks;
    auto bail = [&](int32_t rc) { std::string e = c->err; comm_free(c); c->err = e; g_last_error = e; return rc; };
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    {
        ncclResult_t r = api->CommInitRank(&m->nccl, nranks, id, rank);
        if (r != ncclSuccess) return bail(fail(c, PDEB200_ECOMM, std::string("ncclCommInitRank: ") + api->GetErrorString(r)));
    }
    m->transport = PDEB200_COMM_NCCL;
    const char* tenv = getenv("PDEB200_COMM_TRANSPORT");
    const bool want_peer = nranks > 1 && nranks <= kMaxRanks && !(tenv && std::strcmp(tenv, "nccl") == 0);
    const char* cenv = getenv("PDEB200_COMM_CAP_FLOATS");
    m->cap = cenv && atoi(cenv) > 0 ? atoi(cenv) : 65536;
    const char* toenv = getenv("PDEB200_COMM_TIMEOUT_MS");
    m->timeout_ns = (unsigned long long)(toenv && atoi(toenv) > 0 ? atoi(toenv) : 30000) * 1000000ull;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e == cudaSuccess) return true;
        fail(c, PDEB200_ECUDA, std::string("comm_init: ") + what + ": " + cudaGetErrorString(e));
        return false;
    };
    if (!cu(cudaMalloc(&m->scratch, 64 * sizeof(double)), "cudaMalloc")) return bail(PDEB200_ECUDA);
    if (nranks == 1) { m->transport = PDEB200_COMM_NONE; return PDEB200_OK; }
    // ---- peer-memory exchange buffers -----------------------------------------------------------------------
    int ok = want_peer ? 1 : 0;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok) {
        m->buf_bytes = kCommHeaderBytes + (size_t)2 * kMaxRanks * m->cap * sizeof(float);
        ok = cudaMalloc(&m->buf, m->buf_bytes) == cudaSuccess && cudaMemsetAsync(m->buf, 0, m->buf_bytes, c->stream) == cudaSuccess &&
             cudaMalloc(&m->epoch, 4 * sizeof(unsigned int)) == cudaSuccess &&
             cudaMemsetAsync(m->epoch, 0, 4 * sizeof(unsigned int), c->stream) == cudaSuccess &&
             cudaMalloc(&m->err, sizeof(int)) == cudaSuccess && cudaMemsetAsync(m->err, 0, sizeof(int), c->stream) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine, m->buf) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    // all-gather {ok, handle} through NCCL: 128 bytes per rank
    constexpr size_t REC = 128;
    static_assert(sizeof(cudaIpcMemHandle_t) <= REC - 8, "ipc handle size");
    char* d_rec = nullptr;
    if (!cu(cudaMalloc(&d_rec, REC * (size_t)(nranks + 1)), "cudaMalloc")) return bail(PDEB200_ECUDA);
    std::vector<char> rec(REC * (size_t)(nranks + 1), 0);
    std::memcpy(rec.data(), &ok, sizeof(int));
    std::memcpy(rec.data() + 8, &mine, sizeof(mine));
    cudaMemcpyAsync(d_rec, rec.data(), REC, cudaMemcpyHostToDevice, c->stream);
    {
        ncclResult_t r = api->AllGather(d_rec, d_rec + REC, REC, ncclChar, m->nccl, c->stream);
        if (r != ncclSuccess) { cudaFree(d_rec); return bail(fail(c, PDEB200_ECOMM, std::string("ncclAllGather: ") + api->GetErrorString(r))); }
    }
    cudaMemcpyAsync(rec.data(), d_rec, rec.size(), cudaMemcpyDeviceToHost, c->stream);
    if (!cu(cudaStreamSynchronize(c->stream), "all-gather of the ipc handles")) { cudaFree(d_rec); return bail(PDEB200_ECUDA); }
    int all_ok = 1;
    for (int r = 0; r < nranks; ++r) { int o; std::memcpy(&o, rec.data() + REC * (r + 1), sizeof(int)); all_ok &= o; }
    int opened = all_ok;
    if (all_ok) {
        for (int r = 0; r < nranks && opened; ++r) {
            if (r == rank) { m->peer[r] = m->buf; continue; }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, rec.data() + REC * (r + 1) + 8, sizeof(h));
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; break; }
            m->peer[r] = (char*)p;
        }
    }
    // second round: did EVERY rank map every peer?  (also the barrier after which peers may write into our buffer)
    std::memset(rec.data(), 0, REC);
    std::memcpy(rec.data(), &opened, sizeof(int));
    cudaMemcpyAsync(d_rec, rec.data(), REC, cudaMemcpyHostToDevice, c->stream);
    {
        ncclResult_t r = api->AllGather(d_rec, d_rec + REC, REC, ncclChar, m->nccl, c->stream);
        if (r != ncclSuccess) { cudaFree(d_rec); return bail(fail(c, PDEB200_ECOMM, std::string("ncclAllGather: ") + api->GetErrorString(r))); }
    }
    cudaMemcpyAsync(rec.data(), d_rec, rec.size(), cudaMemcpyDeviceToHost, c->stream);
    if (!cu(cudaStreamSynchronize(c->stream), "all-gather of the mapping status")) { cudaFree(d_rec); return bail(PDEB200_ECUDA); }
    cudaFree(d_rec);
    int all_opened = 1;
    for (int r = 0; r < nranks; ++r) { int o; std::memcpy(&o, rec.data() + REC * (r + 1), sizeof(int)); all_opened &= o; }
    if (all_opened) m->transport = PDEB200_COMM_PEER;
    else {
        // plain NCCL on the context's stream (allreduce between the gradient kernels and the optimiser kernels)
        for (int r = 0; r < nranks; ++r)
            if (r != rank && m->peer[r]) { cudaIpcCloseMemHandle(m->peer[r]); }
        for (int r = 0; r < kMaxRanks; ++r) m->peer[r] = nullptr;
        if (tenv && std::strcmp(tenv, "peer") == 0)
            return bail(fail(c, PDEB200_ECOMM, "comm_init: PDEB200_COMM_TRANSPORT=peer but peer memory could not be mapped on every rank"));
    }
    return PDEB200_OK;
}

int32_t pdeb200_comm_destroy(pdeb200_ctx* c) {
    if (!c) return PDEB200_EINVAL;
    cudaSetDevice(c->device);
    comm_free(c);
    return PDEB200_OK;
}

int32_t pdeb200_comm_info(const pdeb200_ctx* c, int32_t* rank, int32_t* nranks, int32_t* transport) {
    if (!c) return PDEB200_EINVAL;
    const Comm* m = cm(c);
    if (rank) *rank = m ? m->rank : 0;
    if (nranks) *nranks = m ? m->nranks : 1;
    if (transport) *transport = m ? m->transport : PDEB200_COMM_NONE;
    return PDEB200_OK;
}

int32_t pdeb200_comm_allreduce_f64(pdeb200_ctx* c, double* host_inout, int32_t n) {
    if (!c || !host_inout || n < 1 || n > 64) return fail(c, PDEB200_EINVAL, "comm_allreduce_f64: bad argument (1 <= n <= 64)");
    Comm* m = cm(c);
    if (!m || m->nranks <= 1) return PDEB200_OK;
    cudaSetDevice(c->device);
    PDEB_CUDA(c, cudaMemcpyAsync(m->scratch, host_inout, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    int32_t rc = comm_allreduce_f64(c, m->scratch, n);
    if (rc) return rc;
    PDEB_CUDA(c, cudaMemcpyAsync(host_inout, m->scratch, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PDEB_CUDA(c, cudaStreamSynchronize(c->stream));
    return PDEB200_OK;
}

}  // extern "C"
