// K4: the path's only exchange step.  Batch shards are independent; per
// training step one fp32 buffer (head-gradient sums + loss sums + denominators)
// and, per epoch or step, one int64 buffer (confusion counts) are sum-all-reduced
// over NCCL (NVLink 5 / NVSwitch) inside a single group call.
//
// NCCL is bound at run time with dlopen so the library has no link-time
// dependency: inside a PyTorch process "libnccl.so.2" resolves to the copy
// torch already loaded (nvidia/nccl/lib, 2.28.x).
#include "nkbk_common.cuh"

#include <dlfcn.h>
#include <string.h>

namespace nkbk {

// Minimal NCCL surface (matches nccl.h 2.x; the ABI of these entry points is stable).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[NKBK_UNIQUE_ID_BYTES]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclInt64 = 4, ncclFloat32 = 7 };  // ncclDataType_t
enum { ncclSum = 0 };                     // ncclRedOp_t

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_world = 0, g_rank = -1;

static int load_nccl() {
    if (g_nccl.handle) return NKBK_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("nkbk_comm: cannot dlopen libnccl.so.2 (%s)", dlerror());
        return NKBK_E_NCCL;
    }
#define NKBK_SYM(field, name)                                              \
    *(void**)(&g_nccl.field) = dlsym(h, name);                             \
    if (!g_nccl.field) {                                                   \
        set_error("nkbk_comm: symbol %s missing from libnccl", name);      \
        return NKBK_E_NCCL;                                                \
    }
    NKBK_SYM(GetUniqueId, "ncclGetUniqueId");
    NKBK_SYM(CommInitRank, "ncclCommInitRank");
    NKBK_SYM(CommDestroy, "ncclCommDestroy");
    NKBK_SYM(AllReduce, "ncclAllReduce");
    NKBK_SYM(GroupStart, "ncclGroupStart");
    NKBK_SYM(GroupEnd, "ncclGroupEnd");
    NKBK_SYM(GetErrorString, "ncclGetErrorString");
#undef NKBK_SYM
    g_nccl.handle = h;
    return NKBK_OK;
}

#define NKBK_CHECK_NCCL(expr)                                                            \
    do {                                                                                 \
        ncclResult_t _r = (expr);                                                        \
        if (_r != ncclSuccess) {                                                         \
            set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r));                \
            return NKBK_E_NCCL;                                                          \
        }                                                                                \
    } while (0)

}  // namespace nkbk

using namespace nkbk;

extern "C" int nkbk_comm_unique_id(void* out_id_host) {
    NKBK_CHECK_ARG(out_id_host != nullptr, "nkbk_comm_unique_id: NULL output");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    NKBK_CHECK_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out_id_host, &id, NKBK_UNIQUE_ID_BYTES);
    return NKBK_OK;
}

extern "C" int nkbk_comm_init(int rank, int world, const void* id_host, int device) {
    NKBK_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && id_host, "nkbk_comm_init: rank=%d world=%d", rank, world);
    if (g_comm) {
        set_error("nkbk_comm_init: communicator already initialised (world %d)", g_world);
        return NKBK_E_NCCL;
    }
    int rc = load_nccl();
    if (rc) return rc;
    NKBK_CHECK_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, id_host, NKBK_UNIQUE_ID_BYTES);
    NKBK_CHECK_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
    g_world = world;
    g_rank = rank;
    return NKBK_OK;
}

extern "C" int nkbk_comm_world(void) { return g_comm ? g_world : 0; }

extern "C" int nkbk_allreduce_heads(float* reduce_buf, int64_t n_f32, int64_t* cm, int64_t n_i64, void* stream) {
    if (!g_comm) {
        set_error("nkbk_allreduce_heads: communicator not initialised (call nkbk_comm_init)");
        return NKBK_E_NCCL;
    }
    NKBK_CHECK_ARG(n_f32 >= 0 && n_i64 >= 0, "nkbk_allreduce_heads: negative count");
    NKBK_CHECK_ARG((n_f32 == 0 || reduce_buf) && (n_i64 == 0 || cm), "nkbk_allreduce_heads: NULL buffer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    NKBK_CHECK_NCCL(g_nccl.GroupStart());
    if (n_f32 > 0) NKBK_CHECK_NCCL(g_nccl.AllReduce(reduce_buf, reduce_buf, (size_t)n_f32, ncclFloat32, ncclSum, g_comm, st));
    if (n_i64 > 0) NKBK_CHECK_NCCL(g_nccl.AllReduce(cm, cm, (size_t)n_i64, ncclInt64, ncclSum, g_comm, st));
    NKBK_CHECK_NCCL(g_nccl.GroupEnd());
    count_launch((n_f32 > 0) + (n_i64 > 0));
    return NKBK_OK;
}

extern "C" int nkbk_comm_shutdown(void) {
    if (g_comm) {
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
        g_world = 0;
        g_rank = -1;
    }
    return NKBK_OK;
}
