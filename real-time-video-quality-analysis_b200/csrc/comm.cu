// Multi-GPU close of a clip (SURVEY.md 8 e): every rank holds, per clip, the partial weighted sums of its
// frame range (vqa_ewm_partial) and integer side totals; ONE ncclAllReduce(sum) over NVLink / NVSwitch on the
// context's stream turns them into the clip result on every rank.  The reference's counterpart is the
// order-preserving gather of its process pool (complexity_metrics.py:128-148) followed by np.mean at
// :301-310; here the mean(ewm(x)) of a10 is a fixed weighted sum, so a sum-reduce of 8 doubles per clip is
// the whole exchange.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy PyTorch already loaded when the caller is a
// torch process, the system library otherwise), so libvqa_b200.so has no link-time dependency on it and a
// single-GPU user never loads it.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>

#include "vqa_common.cuh"

namespace vqa {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    char err[256] = {0};
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        snprintf(api.err, sizeof(api.err), "libnccl.so.2 not found: %s", dlerror());
        return &api;
    }
#define NCCL_SYM(field, name)                                                          \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));           \
    if (!api.field) { snprintf(api.err, sizeof(api.err), "symbol %s missing in libnccl", name); api.lib = nullptr; return &api; }
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllReduce, "ncclAllReduce")
    NCCL_SYM(Send, "ncclSend")
    NCCL_SYM(Recv, "ncclRecv")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(CommCount, "ncclCommCount")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    return &api;
}

#define VQA_NCCL(c, api, call)                                                                          \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess)                                                                         \
            return set_err((c), VQA_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, (api)->GetErrorString(r__)); \
    } while (0)

void comm_release(vqa_ctx *c)
{
    if (c && c->comm) {
        NcclApi *api = nccl_api();
        if (api->lib) api->CommDestroy((ncclComm_t)c->comm);
        c->comm = nullptr;
        c->comm_world = 0;
    }
}

}  // namespace vqa

using namespace vqa;

static_assert(sizeof(ncclUniqueId) == VQA_COMM_ID_BYTES, "ncclUniqueId size");

extern "C" {

int vqa_comm_unique_id(uint8_t *id_out)
{
    if (!id_out) return VQA_E_INVALID;
    NcclApi *api = nccl_api();
    if (!api->lib) return set_err(nullptr, VQA_E_UNSUPPORTED, "%s", api->err);
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return set_err(nullptr, VQA_E_CUDA, "ncclGetUniqueId -> %s", api->GetErrorString(r));
    memcpy(id_out, &id, sizeof(id));
    return VQA_OK;
}

int vqa_comm_init(vqa_ctx *c, const uint8_t *id_bytes, int rank, int world)
{
    if (!c) return VQA_E_INVALID;
    if (!id_bytes || world < 1 || rank < 0 || rank >= world) return set_err(c, VQA_E_INVALID, "vqa_comm_init: bad argument");
    NcclApi *api = nccl_api();
    if (!api->lib) return set_err(c, VQA_E_UNSUPPORTED, "%s", api->err);
    VQA_CUDA(c, cudaSetDevice(c->device));
    comm_release(c);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    VQA_NCCL(c, api, api->CommInitRank(&comm, world, id, rank));
    c->comm = comm;
    c->comm_rank = rank;
    c->comm_world = world;
    return VQA_OK;
}

int vqa_comm_destroy(vqa_ctx *c)
{
    if (!c) return VQA_E_INVALID;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    comm_release(c);
    return VQA_OK;
}

// One fused buffer [n_f64 doubles | n_i64 integers carried as doubles], ONE ncclAllReduce(sum, f64).  Integer
// totals (edge / keypoint / frame counts) are exact in a double up to 2^53; the call refuses anything larger
// instead of rounding, so integer outputs stay identical for every rank count.
int vqa_clip_reduce(vqa_ctx *c, void *nccl_comm, double *partials, int n_f64, int64_t *ints, int n_i64)
{
    if (!c) return VQA_E_INVALID;
    if (n_f64 < 0 || n_i64 < 0 || (n_f64 > 0 && !partials) || (n_i64 > 0 && !ints))
        return set_err(c, VQA_E_INVALID, "vqa_clip_reduce: bad argument");
    const int total = n_f64 + n_i64;
    if (total == 0) return VQA_OK;
    ncclComm_t comm = nccl_comm ? (ncclComm_t)nccl_comm : (ncclComm_t)c->comm;
    if (!comm) return set_err(c, VQA_E_INVALID, "vqa_clip_reduce: no communicator (pass one or call vqa_comm_init)");
    NcclApi *api = nccl_api();
    if (!api->lib) return set_err(c, VQA_E_UNSUPPORTED, "%s", api->err);
    VQA_CUDA(c, cudaSetDevice(c->device));
    int world = 1;
    VQA_NCCL(c, api, api->CommCount(comm, &world));
    const double lim = 9007199254740992.0 / (double)world;                 // 2^53 / world: the SUM stays exact
    double *hb = (double *)pinned_buf(c, "comm.host", sizeof(double) * (size_t)total);
    if (!hb) return VQA_E_NOMEM;
    for (int i = 0; i < n_f64; i++) hb[i] = partials[i];
    for (int i = 0; i < n_i64; i++) {
        if (fabs((double)ints[i]) >= lim) return set_err(c, VQA_E_UNSUPPORTED, "vqa_clip_reduce: integer total %lld too large for an exact reduce", (long long)ints[i]);
        hb[n_f64 + i] = (double)ints[i];
    }
    VQA_BUF(c, db, double, "comm.buf", total);
    VQA_CUDA(c, cudaMemcpyAsync(db, hb, sizeof(double) * (size_t)total, cudaMemcpyHostToDevice, c->stream));
    VQA_NCCL(c, api, api->AllReduce(db, db, (size_t)total, ncclDouble, ncclSum, comm, c->stream));
    VQA_CUDA(c, cudaMemcpyAsync(hb, db, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, c->stream));
    VQA_CUDA(c, wait_stream(c));
    for (int i = 0; i < n_f64; i++) partials[i] = hb[i];
    for (int i = 0; i < n_i64; i++) ints[i] = (int64_t)llround(hb[n_f64 + i]);
    return VQA_OK;
}

// Frame-range sharding (SURVEY.md 8 e): rank r receives the LAST frame of rank r-1's range as its one-frame
// halo.  `send` (may be NULL on the last rank) goes to rank+1, `recv` (may be NULL on rank 0) comes from
// rank-1; device pointers, one grouped ncclSend/ncclRecv pair on the context's stream over NVLink.
int vqa_comm_halo_exchange(vqa_ctx *c, void *nccl_comm, const uint8_t *send, uint8_t *recv, size_t bytes)
{
    if (!c) return VQA_E_INVALID;
    ncclComm_t comm = nccl_comm ? (ncclComm_t)nccl_comm : (ncclComm_t)c->comm;
    if (!comm) return set_err(c, VQA_E_INVALID, "vqa_comm_halo_exchange: no communicator");
    if (nccl_comm) return set_err(c, VQA_E_UNSUPPORTED, "vqa_comm_halo_exchange needs the context's own communicator (rank / world known)");
    NcclApi *api = nccl_api();
    if (!api->lib) return set_err(c, VQA_E_UNSUPPORTED, "%s", api->err);
    VQA_CUDA(c, cudaSetDevice(c->device));
    const int rank = c->comm_rank, world = c->comm_world;
    VQA_NCCL(c, api, api->GroupStart());
    if (send && rank + 1 < world) VQA_NCCL(c, api, api->Send(send, bytes, ncclUint8, rank + 1, comm, c->stream));
    if (recv && rank > 0) VQA_NCCL(c, api, api->Recv(recv, bytes, ncclUint8, rank - 1, comm, c->stream));
    VQA_NCCL(c, api, api->GroupEnd());
    return VQA_OK;
}

}  // extern "C"
