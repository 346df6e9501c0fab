// comm.cu — the path's multi-GPU exchange behind the C ABI (new work, the reference is single-GPU).
//
// One process per GPU, one NCCL communicator per solb_ctx.  Two exchanges exist on this path:
//   * frames split (sample split): every rank renders frames f = rank (mod N) into a per-rank SUM target; ONE float32
//     sum-reduce of the W x H x float4 buffers onto the root over NVLink, then the resolve (sum / count, gamma, rgba8) on the
//     root in the same call, on the ctx stream (solb_reduce_accum);
//   * tile split (single-sample interactive frames): rank r traces the 8-row bands r, r + N, ... of a frame; its rows are
//     packed, all-gathered and scattered back into every rank's full-size target (solb_allgather_rows).
// NCCL is bound at run time (dlopen): a host that already carries a copy (PyTorch bundles its own libnccl.so.2) must not get a
// second one mapped next to it, so the copy already in the process is used when there is one, else SOLB_NCCL_LIB, else the
// system library.  Nothing here runs without a communicator: no host-staged fallback.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; every call goes through the table below

#include "solb_handles.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char *names[3] = { nullptr, getenv("SOLB_NCCL_LIB"), "libnccl.so.2" };
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host application already loaded
    for (int i = 1; i < 3 && !h; i++)
        if (names[i] && *names[i]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        api.error = std::string("NCCL not found (libnccl.so.2; set SOLB_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
        return &api;
    }
    api.handle = h;
#define SOLB_SYM(field, name)                                             \
    api.field = (decltype(api.field))dlsym(h, name);                      \
    if (!api.field) { api.error = std::string("NCCL symbol missing: ") + name; api.handle = nullptr; return &api; }
    SOLB_SYM(GetUniqueId, "ncclGetUniqueId")
    SOLB_SYM(CommInitRank, "ncclCommInitRank")
    SOLB_SYM(CommDestroy, "ncclCommDestroy")
    SOLB_SYM(Reduce, "ncclReduce")
    SOLB_SYM(AllGather, "ncclAllGather")
    SOLB_SYM(GetErrorString, "ncclGetErrorString")
    SOLB_SYM(GetVersion, "ncclGetVersion")
#undef SOLB_SYM
    return &api;
}

int fail_nccl(solb_ctx *ctx, ncclResult_t r, const char *what) {
    NcclApi *n = nccl_api();
    return fail(ctx, SOLB_ERR_CUDA, std::string(what) + ": " + (n->GetErrorString ? n->GetErrorString(r) : "NCCL error"));
}
#define NC(ctx, call)                                              \
    do {                                                           \
        ncclResult_t r__ = (call);                                 \
        if (r__ != ncclSuccess) return fail_nccl(ctx, r__, #call); \
    } while (0)

// Rows of the interleaved bands owned by `owner` <-> a contiguous chunk.  One thread per 16-byte (or smaller) element.
//   band b (rows [b * band_rows, (b + 1) * band_rows), clipped to height) belongs to rank b % world; the chunk holds the
//   owner's bands in ascending order, band_rows rows each (the last band of the image may be short: its tail is padding).
template <bool PACK>
__global__ void __launch_bounds__(256) k_band_rows(uint4 *image, uint4 *chunks, uint32_t row_elems, uint32_t height, uint32_t band_rows,
                                                   uint32_t world, uint32_t bands_per_rank, uint32_t only_owner, uint32_t skip_owner) {
    const uint32_t rows_per_chunk = bands_per_rank * band_rows;
    const uint64_t total = (uint64_t)(PACK ? 1u : world) * rows_per_chunk * row_elems;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % row_elems);
        const uint32_t cr = (uint32_t)(i / row_elems);
        const uint32_t owner = PACK ? only_owner : cr / rows_per_chunk;
        if (!PACK && owner == skip_owner) continue;  // this rank's own rows are already in place
        const uint32_t r = cr - (PACK ? 0u : owner * rows_per_chunk);
        const uint32_t band = (r / band_rows) * world + owner, y = band * band_rows + r % band_rows;
        if (y >= height) continue;
        const size_t ci = ((size_t)owner * rows_per_chunk + r) * row_elems + x, ii = (size_t)y * row_elems + x;
        if (PACK) chunks[ci] = image[ii];
        else image[ii] = chunks[ci];
    }
}

}  // namespace

extern "C" {

SOLB_API int solb_comm_unique_id(uint8_t *id_out) {
    if (!id_out) return fail(nullptr, SOLB_ERR_INVALID, "solb_comm_unique_id: null argument");
    NcclApi *n = nccl_api();
    if (!n->handle) return fail(nullptr, SOLB_ERR_UNSUPPORTED, n->error);
    static_assert(sizeof(ncclUniqueId) == SOLB_COMM_ID_BYTES, "SOLB_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
    ncclUniqueId id;
    NC(nullptr, n->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return SOLB_OK;
}

SOLB_API int solb_comm_init(solb_ctx *ctx, const uint8_t *id_bytes, int rank, int world) {
    if (!ctx || !id_bytes) return fail(ctx, SOLB_ERR_INVALID, "solb_comm_init: null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, SOLB_ERR_INVALID, "solb_comm_init: bad rank / world");
    if (ctx->nccl_comm) return fail(ctx, SOLB_ERR_INVALID, "solb_comm_init: this ctx already has a communicator");
    NcclApi *n = nccl_api();
    if (!n->handle) return fail(ctx, SOLB_ERR_UNSUPPORTED, n->error);
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    NC(ctx, n->CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    // NCCL connects the ranks lazily, at the first collective of each kind (measured: 180 - 410 ms inside the first
    // ncclReduce of a fresh communicator).  Pay for it here, where set-up belongs, with one tiny reduce and all-gather.
    {
        float *scratch = nullptr;
        CU(ctx, cudaMalloc((void **)&scratch, sizeof(float) * 4 * (size_t)world));
        CU(ctx, cudaMemsetAsync(scratch, 0, sizeof(float) * 4 * (size_t)world, ctx->stream));
        ncclResult_t r = n->Reduce(scratch, scratch, 4, ncclFloat32, ncclSum, 0, comm, ctx->stream);
        if (r == ncclSuccess) r = n->AllGather(scratch + 4 * rank, scratch, 4, ncclFloat32, comm, ctx->stream);
        const cudaError_t e = cudaStreamSynchronize(ctx->stream);
        cudaFree(scratch);
        if (r != ncclSuccess) return fail_nccl(ctx, r, "communicator warm-up");
        if (e != cudaSuccess) return fail_cuda(ctx, e, "communicator warm-up");
    }
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

SOLB_API int solb_comm_info(solb_ctx *ctx, int *rank, int *world, int *nccl_version) {
    if (!ctx) return fail(nullptr, SOLB_ERR_INVALID, "null ctx");
    if (rank) *rank = ctx->comm_rank;
    if (world) *world = ctx->nccl_comm ? ctx->comm_world : 1;
    if (nccl_version) {
        *nccl_version = 0;
        NcclApi *n = nccl_api();
        if (n->handle) n->GetVersion(nccl_version);
    }
    return SOLB_OK;
}

SOLB_API int solb_comm_destroy(solb_ctx *ctx) {
    if (!ctx) return SOLB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->nccl_comm) {
        nccl_api()->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    cudaFree(ctx->comm_stage);
    ctx->comm_stage = nullptr;
    ctx->comm_stage_bytes = 0;
    ctx->comm_rank = 0;
    ctx->comm_world = 1;
    return SOLB_OK;
}

// Frames split: sum the per-rank SOLB_ACCUM_SUM targets onto `root` (in place) and resolve there, all on the ctx stream.
SOLB_API int solb_reduce_accum(solb_ctx *ctx, solb_target *sum, int root, solb_target *accum_out, solb_target *render) {
    if (!ctx || !sum) return fail(ctx, SOLB_ERR_INVALID, "solb_reduce_accum: null argument");
    if (sum->ctx != ctx || sum->format != SOLB_FORMAT_RGBA32F) return fail(ctx, SOLB_ERR_INVALID, "solb_reduce_accum: sum must be an rgba32f target of this ctx");
    const int world = ctx->nccl_comm ? ctx->comm_world : 1;
    if (root < 0 || root >= world) return fail(ctx, SOLB_ERR_INVALID, "solb_reduce_accum: bad root");
    if (world > 1 && !ctx->nccl_comm) return fail(ctx, SOLB_ERR_INVALID, "solb_reduce_accum before solb_comm_init");
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    if (world > 1) {
        NC(ctx, nccl_api()->Reduce(sum->dev, sum->dev, (size_t)sum->width * sum->height * 4, ncclFloat32, ncclSum, root,
                                   (ncclComm_t)ctx->nccl_comm, ctx->stream));
        ctx->launches += 1;
    }
    if (ctx->comm_rank == root && (accum_out || render)) return solb_resolve_sum(ctx, sum, accum_out, render);
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

// Tile split: every rank has traced its interleaved bands (SolbTraceParams tile_row_begin = rank * band_rows, tile_row_count =
// band_rows, tile_row_stride = world * band_rows) into `target`; afterwards every rank holds the whole image.
SOLB_API int solb_allgather_rows(solb_ctx *ctx, solb_target *target, uint32_t band_rows) {
    if (!ctx || !target) return fail(ctx, SOLB_ERR_INVALID, "solb_allgather_rows: null argument");
    if (target->ctx != ctx) return fail(ctx, SOLB_ERR_INVALID, "solb_allgather_rows: target belongs to another ctx");
    if (band_rows == 0) return fail(ctx, SOLB_ERR_INVALID, "solb_allgather_rows: band_rows must be > 0");
    const int world = ctx->nccl_comm ? ctx->comm_world : 1;
    if (world == 1) return SOLB_OK;
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t bpp = target->format == SOLB_FORMAT_RGBA32F ? 16 : (target->format == SOLB_FORMAT_RGBA8 ? 4 : 8);
    const size_t row_bytes = (size_t)target->width * bpp;
    if (row_bytes % 16) return fail(ctx, SOLB_ERR_UNSUPPORTED, "solb_allgather_rows: row size must be a multiple of 16 bytes");
    const uint32_t n_bands = (target->height + band_rows - 1) / band_rows;
    const uint32_t bands_per_rank = (n_bands + world - 1) / world;
    const size_t chunk_bytes = (size_t)bands_per_rank * band_rows * row_bytes;
    if (ctx->comm_stage_bytes < chunk_bytes * world) {
        cudaFree(ctx->comm_stage);
        ctx->comm_stage = nullptr;
        ctx->comm_stage_bytes = 0;
        CU(ctx, cudaMalloc(&ctx->comm_stage, chunk_bytes * world));
        CU(ctx, cudaMemsetAsync(ctx->comm_stage, 0, chunk_bytes * world, ctx->stream));  // padding rows are sent too
        ctx->comm_stage_bytes = chunk_bytes * world;
    }
    const uint32_t row_elems = (uint32_t)(row_bytes / 16);
    const int grid = ctx->sm_count * 4;
    k_band_rows<true><<<grid, 256, 0, ctx->stream>>>((uint4 *)target->dev, (uint4 *)ctx->comm_stage, row_elems, target->height, band_rows,
                                                    (uint32_t)world, bands_per_rank, (uint32_t)ctx->comm_rank, 0u);
    CU(ctx, cudaGetLastError());
    NC(ctx, nccl_api()->AllGather((const char *)ctx->comm_stage + chunk_bytes * ctx->comm_rank, ctx->comm_stage, chunk_bytes, ncclChar,
                                  (ncclComm_t)ctx->nccl_comm, ctx->stream));
    k_band_rows<false><<<grid, 256, 0, ctx->stream>>>((uint4 *)target->dev, (uint4 *)ctx->comm_stage, row_elems, target->height, band_rows,
                                                     (uint32_t)world, bands_per_rank, 0u, (uint32_t)ctx->comm_rank);
    CU(ctx, cudaGetLastError());
    ctx->launches += 3;
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

}  // extern "C"
