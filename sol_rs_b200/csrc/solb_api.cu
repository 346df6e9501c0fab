// solb_api.cu — the extern "C" boundary declared in include/solb.h.
// Handles are heap objects; every entry point validates, sets the device, catches everything and
// returns a SolbStatus.  There is no CPU path: without a usable CUDA device solb_ctx_create fails.
#include "solb_handles.h"

#include "host_math.h"

thread_local std::string g_last_error;

static size_t format_bytes(uint32_t f) { return f == SOLB_FORMAT_RGBA32F ? 16 : (f == SOLB_FORMAT_RGBA8 ? 4 : 8); }

extern "C" {

SOLB_API uint32_t solb_version(void) { return SOLB_VERSION; }

SOLB_API const char *solb_last_error(solb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

SOLB_API int solb_ctx_create(int device, void *stream, solb_ctx **out) {
    if (!out) return fail(nullptr, SOLB_ERR_INVALID, "solb_ctx_create: out is NULL");
    *out = nullptr;
    SOLB_TRY
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, SOLB_ERR_CUDA, std::string("no CUDA device: libsolb has no CPU fallback (") +
                                                (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") + ")");
    if (device < 0 || device >= count) return fail(nullptr, SOLB_ERR_INVALID, "solb_ctx_create: bad device index");
    CU(nullptr, cudaSetDevice(device));
    solb_ctx *c = new solb_ctx();
    c->device = device;
    if (stream) c->stream = (cudaStream_t)stream;
    else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail_cuda(nullptr, e, "cudaStreamCreate"); }
        c->own_stream = true;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    {
        // Build scratch comes from a stream-ordered pool.  With the default release threshold (0) every synchronisation hands
        // the freed scratch back to the driver and the next build pays for mapping it again (measured: 485 ms to release and
        // 210 ms to re-map the 6 GB a 20 M-triangle build uses); keep it cached and let the caller return it with
        // solb_ctx_trim.  The pool is PRIVATE to the ctx: the device's default pool is process-global state of the embedding
        // application (torch, another library) and is left alone.
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&c->build_pool, &props) == cudaSuccess) {
            uint64_t threshold = ~0ull;
            cudaMemPoolSetAttribute(c->build_pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        } else {
            c->build_pool = nullptr;  // fall back to the default pool with ITS settings untouched
        }
        cudaGetLastError();
    }
    {
        auto env_int = [](const char *name, int dflt, int lo, int hi) {
            const char *v = getenv(name);
            if (!v || !*v) return dflt;
            const int x = atoi(v);
            return x < lo ? lo : (x > hi ? hi : x);
        };
        c->tune.fetch_idle = env_int("SOLB_FETCH_IDLE", c->tune.fetch_idle, 1, 32);
        c->tune.tri_weight = env_int("SOLB_TRI_WEIGHT", c->tune.tri_weight, 1, 64);
        c->tune.node_weight = env_int("SOLB_NODE_WEIGHT", c->tune.node_weight, 1, 64);
        c->tune.ctas_per_sm = env_int("SOLB_CTAS_PER_SM", c->tune.ctas_per_sm, 1, 16);
        c->tune.check_every = env_int("SOLB_CHECK_EVERY", c->tune.check_every, 1, 1024);
        c->tune.overlap = env_int("SOLB_OVERLAP", c->tune.overlap, 1, WF_MAX_PARTS);
        c->tune.sort_shade = env_int("SOLB_SORT_SHADE", c->tune.sort_shade, 0, 1);
        c->tune.pool = env_int("SOLB_POOL", c->tune.pool, 0, 1);
        c->tune.pool_ctas_per_sm = env_int("SOLB_POOL_CTAS_PER_SM", c->tune.pool_ctas_per_sm, 1, 6);
        c->tune.pool_refill = env_int("SOLB_POOL_REFILL", c->tune.pool_refill, 1, 64);
        c->tune.ctas_per_sm_overlap = env_int("SOLB_CTAS_PER_SM_OVERLAP", c->tune.ctas_per_sm_overlap, 1, 16);
        c->tune.min_batch = env_int("SOLB_WF_MIN_BATCH", c->tune.min_batch, 32, 128) & ~31;
        c->tune.min_pixels_per_part = env_int("SOLB_MIN_PIXELS_PER_PART", c->tune.min_pixels_per_part, 1, 1 << 30);
        c->tune.coop_tri = env_int("SOLB_COOP_TRI", c->tune.coop_tri, 0, 1);
        c->tune.coop_block = env_int("SOLB_COOP_BLOCK", c->tune.coop_block, 1, 32);
        c->tune.shade_ctas_per_sm_overlap = env_int("SOLB_SHADE_CTAS", c->tune.shade_ctas_per_sm_overlap, 1, 16);
        c->tune.mega_persistent = env_int("SOLB_MEGA_PERSISTENT", c->tune.mega_persistent, 0, 1);
        c->tune.async_poll = env_int("SOLB_ASYNC_POLL", c->tune.async_poll, 0, 1);
        c->tune.max_ahead = env_int("SOLB_MAX_AHEAD", c->tune.max_ahead, 1, 4096);
        c->tune.shade_priority = env_int("SOLB_SHADE_PRIORITY", c->tune.shade_priority, 0, 1);
        c->tune.part_priority = env_int("SOLB_PART_PRIORITY", c->tune.part_priority, 0, 1);
        c->tune.mega_vote = env_int("SOLB_MEGA_VOTE", c->tune.mega_vote, -1, 1);
        c->tune.mega_ctas_per_sm = env_int("SOLB_MEGA_CTAS_PER_SM", c->tune.mega_ctas_per_sm, 1, 16);
        c->tune.mega_fetch_idle = env_int("SOLB_MEGA_FETCH_IDLE", c->tune.mega_fetch_idle, 1, 32);
        c->auto_wide_schedule = env_int("SOLB_AUTO_SCHEDULE", (int)c->auto_wide_schedule, 0, 3) == 0 ? SOLB_SCHEDULE_WAVEFRONT : SOLB_SCHEDULE_WARPFRONT;
        c->tune.wl_ctas_per_sm = env_int("SOLB_WL_CTAS_PER_SM", c->tune.wl_ctas_per_sm, 1, 16);
        c->tune.wl_fetch_idle = env_int("SOLB_WL_FETCH_IDLE", c->tune.wl_fetch_idle, 1, 32);
        c->tune.wl_fetch_idle_large = env_int("SOLB_WL_FETCH_IDLE_LARGE", c->tune.wl_fetch_idle_large, 1, 32);
        c->tune.wl_starve_idle = env_int("SOLB_WL_STARVE_IDLE", c->tune.wl_starve_idle, 1, 32);
        c->tune.wl_gen_min = env_int("SOLB_WL_GEN_MIN", c->tune.wl_gen_min, 1, WL_POOL);
        c->tune.wl_batch = env_int("SOLB_WL_BATCH", c->tune.wl_batch, 32, 1024) & ~31;
        c->tune.wl_region_slots_per_warp = env_int("SOLB_WL_REGION_SLOTS", c->tune.wl_region_slots_per_warp, 32, 4096);
        c->tune.wl_warps_per_sm = env_int("SOLB_WL_WARPS_PER_SM", c->tune.wl_warps_per_sm, 0, 64);
        c->use_hi_stream = env_int("SOLB_HI_STREAM", c->use_hi_stream, 0, 1);
        c->tune.wl_frames_in_flight = env_int("SOLB_WL_FRAMES_IN_FLIGHT", c->tune.wl_frames_in_flight, 1, WL_MAX_FRAMES);
    }
    e = cudaMalloc((void **)&c->d_stats, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&c->pinned_count, 64);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e != cudaSuccess) { delete c; return fail_cuda(nullptr, e, "ctx setup"); }
    *out = c;
    return SOLB_OK;
    SOLB_CATCH(nullptr)
}

static void free_wavefront(solb_ctx *c) {
    WavefrontState &w = c->ws;
    cudaFree(w.ray_o); cudaFree(w.ray_d); cudaFree(w.thr); cudaFree(w.pix); cudaFree(w.hit);
    cudaFree(w.queue[0]); cudaFree(w.queue[1]); cudaFree(w.counters);
    for (int k = 1; k < WF_MAX_PARTS; k++) {
        cudaFree(c->part_queue[k][0]); cudaFree(c->part_queue[k][1]); cudaFree(c->part_counters[k]);
        c->part_queue[k][0] = c->part_queue[k][1] = nullptr; c->part_counters[k] = nullptr;
    }
    for (int k = 0; k < WF_MAX_PARTS; k++) { cudaFree(c->part_spill[k]); c->part_spill[k] = nullptr; }
    w = WavefrontState{};
}

static void free_warpfront(solb_ctx *c) {
    for (int k = 0; k < WL_MAX_FRAMES; k++) {
        if (c->frame_stream[k]) cudaStreamSynchronize(c->frame_stream[k]);
        WarpfrontState &w = c->wl[k];
        cudaFree(w.ray_o); cudaFree(w.ray_d); cudaFree(w.ray_i); cudaFree(w.frame0); cudaFree(w.frame1);
        cudaFree(w.thr); cudaFree(w.pix); cudaFree(w.hit); cudaFree(w.cursor);
        w = WarpfrontState{};
        cudaFree(c->frame_sum[k]);
        c->frame_sum[k] = nullptr;
        c->frame_sum_pixels[k] = 0;
        if (c->frame_stream[k]) cudaStreamDestroy(c->frame_stream[k]);
        c->frame_stream[k] = nullptr;
        cudaEvent_t *evs[] = { &c->ev_trace_done[k], &c->ev_resolve_done[k], &c->ev_k0[k], &c->ev_k1[k] };
        for (cudaEvent_t *e : evs) { if (*e) cudaEventDestroy(*e); *e = nullptr; }
        c->have_resolve_done[k] = false;
    }
    if (c->hi_stream) { cudaStreamSynchronize(c->hi_stream); cudaStreamDestroy(c->hi_stream); c->hi_stream = nullptr; }
    if (c->ev_hi_in) cudaEventDestroy(c->ev_hi_in);
    if (c->ev_hi_out) cudaEventDestroy(c->ev_hi_out);
    c->ev_hi_in = c->ev_hi_out = nullptr;
    if (c->ev_side_barrier) cudaEventDestroy(c->ev_side_barrier);
    if (c->ev_serial) cudaEventDestroy(c->ev_serial);
    c->ev_side_barrier = c->ev_serial = nullptr;
    c->have_side_barrier = false;
}

// Work just enqueued on the ctx stream that later frames (launched on the side streams) must see: scene uploads, builds,
// TLAS regenerates, stats resets.
static void mark_side_barrier(solb_ctx *c) {
    if (!c->ev_side_barrier && cudaEventCreateWithFlags(&c->ev_side_barrier, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaEventRecord(c->ev_side_barrier, c->stream) == cudaSuccess) c->have_side_barrier = true;
}

static void ctx_release(solb_ctx *ctx) {
    if (--ctx->refs > 0) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_wavefront(ctx);
    free_warpfront(ctx);
    solb_comm_destroy(ctx);
    if (ctx->build_pool) cudaMemPoolDestroy(ctx->build_pool);
    cudaFree(ctx->d_stats);
    cudaFree(ctx->d_blue);
    cudaFreeHost(ctx->pinned_count);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int k = 1; k < WF_MAX_PARTS; k++) {
        if (ctx->part_stream[k]) cudaStreamDestroy(ctx->part_stream[k]);
        if (ctx->ev_join[k]) cudaEventDestroy(ctx->ev_join[k]);
        if (ctx->ev_poll[k]) cudaEventDestroy(ctx->ev_poll[k]);
    }
    if (ctx->ev_poll[0]) cudaEventDestroy(ctx->ev_poll[0]);
    for (int k = 0; k < WF_MAX_PARTS; k++) {
        if (ctx->shade_stream[k]) cudaStreamDestroy(ctx->shade_stream[k]);
        if (ctx->ev_ts[k]) cudaEventDestroy(ctx->ev_ts[k]);
        if (ctx->ev_st[k]) cudaEventDestroy(ctx->ev_st[k]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// Scenes and targets created from the ctx keep it alive (the reference's wrappers hold Arc<Context>:
// src/buffer.rs:291-303), so destroy order does not matter.
SOLB_API int solb_ctx_destroy(solb_ctx *ctx) {
    if (!ctx) return SOLB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx_release(ctx);
    return SOLB_OK;
}

SOLB_API int solb_synchronize(solb_ctx *ctx) {
    if (!ctx) return fail(nullptr, SOLB_ERR_INVALID, "null ctx");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->p2p_err_host && *ctx->p2p_err_host) return fail(ctx, SOLB_ERR_CUDA, "a peer did not deliver its bands within 10 min (solb_allgather_rows)");
    return SOLB_OK;
}

SOLB_API int solb_fence_create(solb_ctx *ctx, solb_fence **out) {
    if (!ctx || !out) return fail(ctx, SOLB_ERR_INVALID, "null argument");
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    cudaEvent_t ev = nullptr;
    CU(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync));
    solb_fence *f = new solb_fence();
    f->ctx = ctx;
    f->ev = ev;
    *out = f;
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

SOLB_API int solb_fence_signal(solb_fence *f) {
    if (!f) return fail(nullptr, SOLB_ERR_INVALID, "null fence");
    CU(f->ctx, cudaSetDevice(f->ctx->device));
    CU(f->ctx, cudaEventRecord(f->ev, f->ctx->stream));  // every frame's resolve is on this stream, after its kernel
    f->armed = true;
    return SOLB_OK;
}

SOLB_API int solb_fence_wait(solb_fence *f) {
    if (!f) return fail(nullptr, SOLB_ERR_INVALID, "null fence");
    if (!f->armed) return SOLB_OK;
    CU(f->ctx, cudaSetDevice(f->ctx->device));
    CU(f->ctx, cudaEventSynchronize(f->ev));
    f->armed = false;
    return SOLB_OK;
}

SOLB_API int solb_fence_destroy(solb_fence *f) {
    if (!f) return SOLB_OK;
    cudaSetDevice(f->ctx->device);
    if (f->ev) cudaEventDestroy(f->ev);
    delete f;
    return SOLB_OK;
}

SOLB_API int solb_host_alloc(solb_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out || !bytes) return fail(ctx, SOLB_ERR_INVALID, "null argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return SOLB_OK;
}

SOLB_API int solb_host_free(solb_ctx *ctx, void *ptr) {
    if (!ctx) return fail(nullptr, SOLB_ERR_INVALID, "null ctx");
    if (!ptr) return SOLB_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaFreeHost(ptr));
    return SOLB_OK;
}

SOLB_API int solb_ctx_trim(solb_ctx *ctx) {
    if (!ctx) return fail(nullptr, SOLB_ERR_INVALID, "null ctx");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->build_pool) CU(ctx, cudaMemPoolTrimTo(ctx->build_pool, 0));
    return SOLB_OK;
}

// Pipeline::new compiles the GLSL stages when the pipeline is created (src/ray/pipeline.rs:61-115); the counterpart here is
// loading the AOT-compiled kernels.  CUDA loads a kernel lazily at its first launch, so without this the first
// solb_accel_build of a process also paid for ~25 module loads and the creation of the scratch pool (the driver's bench recorded
// 549 ms for a 3 ms build).  The preload pushes a two-triangle scene through every launch path once: both build modes, the
// three path-tracing schedules, ao, debug and trace_rays.
SOLB_API int solb_ctx_preload(solb_ctx *ctx) {
    if (!ctx) return fail(nullptr, SOLB_ERR_INVALID, "null ctx");
    if (ctx->preloaded) return SOLB_OK;
    SolbModelVertex v[4];
    memset(v, 0, sizeof(v));
    const float q[4][3] = { { -1, 0, -1 }, { 1, 0, -1 }, { 1, 0, 1 }, { -1, 0, 1 } };
    for (int i = 0; i < 4; i++) {
        for (int k = 0; k < 3; k++) v[i].pos[k] = q[i][k];
        v[i].pos[3] = 1.0f;
        v[i].color[0] = v[i].color[1] = v[i].color[2] = v[i].color[3] = 1.0f;
        v[i].normal[1] = 1.0f;
    }
    const uint32_t idx[6] = { 0, 1, 2, 0, 2, 3 };
    SolbSection sec = { 0, 4, 0, 6, 0 };
    SolbMeshDesc mesh;
    memset(&mesh, 0, sizeof(mesh));
    mesh.vertices = v; mesh.n_vertices = 4; mesh.indices = idx; mesh.n_indices = 6; mesh.sections = &sec; mesh.n_sections = 1;
    mesh.transform[0] = mesh.transform[5] = mesh.transform[10] = mesh.transform[15] = 1.0f;
    SolbMaterialInfo mat;
    memset(&mat, 0, sizeof(mat));
    mat.base_color[0] = mat.base_color[1] = mat.base_color[2] = mat.base_color[3] = 0.5f;
    mat.roughness = 1.0f;
    SolbSceneUniforms u;
    memset(&u, 0, sizeof(u));
    for (int i = 0; i < 16; i += 5) u.view_inverse[i] = u.projection_inverse[i] = 1.0f;
    u.view_inverse[13] = 2.0f;  // camera above the quad, looking down -z of an identity view: rays mostly miss, which is fine
    u.frame[0] = u.frame[1] = 8;
    solb_scene *sc = nullptr;
    solb_target *acc = nullptr, *ren = nullptr, *ids = nullptr;
    int rc = solb_scene_create(ctx, &mesh, 1, &mat, 1, &sc);
    const uint64_t launches = ctx->launches;
    const float build_ms = ctx->last_build_ms;
    if (!rc) rc = solb_target_create(ctx, 8, 8, SOLB_FORMAT_RGBA32F, &acc);
    if (!rc) rc = solb_target_create(ctx, 8, 8, SOLB_FORMAT_RGBA8, &ren);
    if (!rc) rc = solb_target_create(ctx, 8, 8, SOLB_FORMAT_RG32UI, &ids);
    for (uint32_t mode = 0; mode < 2 && !rc; mode++) {
        rc = solb_scene_set_accel_mode(sc, mode ? SOLB_ACCEL_FLAT : SOLB_ACCEL_TWO_LEVEL);
        if (!rc) rc = solb_accel_build(sc);
        SolbTraceParams p;
        solb_trace_params_default(&p, 0);
        p.max_bounces = 2;
        for (uint32_t sch = 0; sch < 4 && !rc; sch++) {
            if (sch == SOLB_SCHEDULE_AUTO) continue;
            p.schedule = sch;
            rc = solb_trace_pathtrace(sc, &u, &p, acc, ren);
        }
        if (!rc) rc = solb_trace_debug(sc, &u, ren, ids, nullptr);
        float ray[8] = { 0, 1, 0, 0.001f, 0, -1, 0, 100.0f };
        uint32_t hit[4];
        if (!rc) rc = solb_trace_rays(sc, ray, 1, hit, nullptr);
        if (!rc && ctx->d_blue) {
            solb_trace_params_default(&p, 1);
            rc = solb_trace_ao(sc, &u, &p, acc);
        }
    }
    if (!rc) rc = solb_resolve_sum(ctx, acc, acc, ren);
    if (!rc) rc = solb_synchronize(ctx);
    solb_target_destroy(ids); solb_target_destroy(ren); solb_target_destroy(acc);
    solb_scene_destroy(sc);
    solb_stats_reset(ctx);
    ctx->launches = launches;  // the preload is not part of anybody's frame
    ctx->last_build_ms = build_ms;
    if (!rc) ctx->preloaded = true;
    return rc;
}

SOLB_API int solb_set_timing(solb_ctx *ctx, int enabled) {
    if (!ctx) return fail(nullptr, SOLB_ERR_INVALID, "null ctx");
    ctx->timing = enabled != 0;
    return SOLB_OK;
}

SOLB_API int solb_stats_get(solb_ctx *ctx, SolbStats *out) {
    if (!ctx || !out) return fail(ctx, SOLB_ERR_INVALID, "solb_stats_get: null argument");
    CU(ctx, cudaSetDevice(ctx->device));
    unsigned long long h[8];
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaMemcpy(h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    out->rays = h[0]; out->hits = h[1]; out->paths = h[2]; out->nodes_visited = h[3]; out->tris_tested = h[4];
    out->kernel_launches = ctx->launches;
    out->last_build_ms = ctx->last_build_ms;
    out->last_trace_ms = ctx->last_trace_ms;
    out->trace_kernel_ms_total = ctx->trace_kernel_ms_total;
    out->trace_kernel_launches = ctx->trace_kernel_launches;
    return SOLB_OK;
}

SOLB_API int solb_stats_reset(solb_ctx *ctx) {
    if (!ctx) return fail(nullptr, SOLB_ERR_INVALID, "null ctx");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemsetAsync(ctx->d_stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
    mark_side_barrier(ctx);  // frames launched on the side streams from now on count into the cleared counters
    ctx->launches = 0;
    ctx->trace_kernel_ms_total = 0.0f;
    ctx->trace_kernel_launches = 0;
    return SOLB_OK;
}

// ---- scene -------------------------------------------------------------------------------------------

SOLB_API int solb_scene_create(solb_ctx *ctx, const SolbMeshDesc *meshes, uint32_t n_meshes, const SolbMaterialInfo *materials,
                               uint32_t n_materials, solb_scene **out) {
    if (!ctx || !out || (n_meshes && !meshes)) return fail(ctx, SOLB_ERR_INVALID, "solb_scene_create: null argument");
    *out = nullptr;
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    std::vector<float> verts;        // 16 floats per vertex
    std::vector<uint32_t> indices;
    solb_scene *s = new solb_scene();
    s->ctx = ctx;
    ctx->refs++;
    struct Guard { solb_scene *s; ~Guard() { if (s) solb_scene_destroy(s); } } guard{ s };
    s->h_first_tri.push_back(0);
    for (uint32_t m = 0; m < n_meshes; m++) {
        const SolbMeshDesc &md = meshes[m];
        if ((md.n_vertices && !md.vertices) || (md.n_sections && !md.sections))
            return fail(ctx, SOLB_ERR_INVALID, "solb_scene_create: mesh with null vertex/section array");
        const uint32_t vbase = (uint32_t)(verts.size() / 16), ibase = (uint32_t)indices.size();
        verts.insert(verts.end(), (const float *)md.vertices, (const float *)md.vertices + 16 * (size_t)md.n_vertices);
        if (md.n_indices) {
            if (!md.indices) return fail(ctx, SOLB_ERR_INVALID, "solb_scene_create: null index array");
            indices.insert(indices.end(), md.indices, md.indices + md.n_indices);
        }
        for (uint32_t k = 0; k < md.n_sections; k++) {
            const SolbSection &sec = md.sections[k];
            // src/ray/mod.rs:103-108: sections without indices never get an index descriptor (App.A item 4)
            if (sec.n_indices == 0 || md.n_indices == 0)
                return fail(ctx, SOLB_ERR_UNSUPPORTED, "non-indexed primitive section: unsupported on the ray-tracing path");
            if (sec.n_indices % 3) return fail(ctx, SOLB_ERR_INVALID, "section index count is not a multiple of 3");
            if ((uint64_t)sec.first_index + sec.n_indices > md.n_indices || (uint64_t)sec.first_vertex + sec.n_vertices > md.n_vertices)
                return fail(ctx, SOLB_ERR_INVALID, "section range outside its mesh");
            if (sec.material_index >= n_materials || !materials)
                return fail(ctx, SOLB_ERR_INVALID, "section material index out of range (the reference unwrap()s here)");
            for (uint32_t i = 0; i < sec.n_indices; i++)
                if (md.indices[sec.first_index + i] >= sec.n_vertices)
                    return fail(ctx, SOLB_ERR_INVALID, "index outside its section's vertex range");
            SolbSceneInstance si;
            memset(&si, 0, sizeof(si));
            si.id = (uint32_t)s->instances.size();  // running count: src/ray/mod.rs:113
            memcpy(si.transform, md.transform, sizeof(si.transform));
            float inv[16];
            mat4_inverse(md.transform, inv);
            mat4_transpose(inv, si.transform_it);  // src/ray/mod.rs:116
            s->instances.push_back(si);
            DeviceBlas db;
            db.first_vertex = vbase + sec.first_vertex;
            db.first_index = ibase + sec.first_index;
            db.n_indices = sec.n_indices;
            db.first_tri = s->n_geom_tris;
            DeviceInstance di;
            memset(&di, 0, sizeof(di));
            di.first_vertex = db.first_vertex;
            di.first_index = db.first_index;
            di.n_indices = db.n_indices;
            di.material = sec.material_index;
            di.blas = (uint32_t)s->h_blas.size();  // one BLAS and one instance per section (src/ray/mod.rs:122)
            di.shade_first_tri = db.first_tri;
            s->h_blas.push_back(db);
            s->n_geom_tris += sec.n_indices / 3;
            memcpy(di.transform, si.transform, sizeof(di.transform));
            memcpy(di.transform_it, si.transform_it, sizeof(di.transform_it));
            memcpy(di.mat, &materials[sec.material_index], sizeof(di.mat));  // materials[gl_InstanceID]
            di.mat[10] = 0.0f;  // MaterialInfo::padding1: libsolb's texture slot (solb_scene_set_textures)
            s->h_inst.push_back(di);
            s->n_tris += sec.n_indices / 3;
            s->h_first_tri.push_back(s->n_tris);
        }
    }
    s->n_vertices = (uint32_t)(verts.size() / 16);
    s->n_indices = (uint32_t)indices.size();
    if (materials && n_materials) s->materials.assign(materials, materials + n_materials);
    const size_t ni = s->h_inst.size();
    s->inst_capacity = std::max<size_t>(ni, 1);
    CU(ctx, cudaMalloc((void **)&s->d_vertices, std::max<size_t>(verts.size(), 16) * sizeof(float)));
    CU(ctx, cudaMalloc((void **)&s->d_indices, std::max<size_t>(indices.size(), 1) * sizeof(uint32_t)));
    CU(ctx, cudaMalloc((void **)&s->d_first_tri, (s->inst_capacity + 1) * sizeof(uint32_t)));
    CU(ctx, cudaMalloc((void **)&s->d_inst, s->inst_capacity * sizeof(DeviceInstance)));
    CU(ctx, cudaMalloc((void **)&s->d_blas, std::max<size_t>(s->h_blas.size(), 1) * sizeof(DeviceBlas)));
    CU(ctx, cudaMalloc((void **)&s->d_shade, std::max<size_t>(s->n_geom_tris, 1) * sizeof(ShadeRecord)));
    cudaStream_t st = ctx->stream;
    if (!verts.empty()) CU(ctx, cudaMemcpyAsync(s->d_vertices, verts.data(), verts.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    if (!indices.empty()) CU(ctx, cudaMemcpyAsync(s->d_indices, indices.data(), indices.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(s->d_first_tri, s->h_first_tri.data(), (ni + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (ni) CU(ctx, cudaMemcpyAsync(s->d_inst, s->h_inst.data(), ni * sizeof(DeviceInstance), cudaMemcpyHostToDevice, st));
    if (ni) CU(ctx, cudaMemcpyAsync(s->d_blas, s->h_blas.data(), s->h_blas.size() * sizeof(DeviceBlas), cudaMemcpyHostToDevice, st));
    CU(ctx, launch_build_shade_records(st, s->view(), s->d_shade));
    ctx->launches += s->n_geom_tris ? 1 : 0;
    CU(ctx, cudaStreamSynchronize(st));  // host vectors go out of scope
    guard.s = nullptr;
    *out = s;
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

static void free_textures(solb_scene *s) {
    for (float4 *t : s->d_texels) cudaFree(t);
    s->d_texels.clear();
    cudaFree(s->d_tex);
    s->d_tex = nullptr;
    s->n_tex = 0;
}

SOLB_API int solb_scene_destroy(solb_scene *s) {
    if (!s) return SOLB_OK;
    if (s->ctx) { cudaSetDevice(s->ctx->device); cudaStreamSynchronize(s->ctx->stream); }
    free_textures(s);
    cudaFree(s->d_vertices); cudaFree(s->d_indices); cudaFree(s->d_first_tri); cudaFree(s->d_inst); cudaFree(s->d_blas);
    cudaFree(s->d_shade);
    s->accel.release();
    solb_ctx *ctx = s->ctx;
    delete s;
    if (ctx) ctx_release(ctx);
    return SOLB_OK;
}

// instance records + per-instance triangle prefix -> device (grows the arrays after solb_scene_add_instance)
static int upload_instances(solb_scene *s) {
    solb_ctx *ctx = s->ctx;
    const size_t ni = s->h_inst.size();
    if (ni > s->inst_capacity) {
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(s->d_inst); cudaFree(s->d_first_tri);
        s->d_inst = nullptr; s->d_first_tri = nullptr;
        s->inst_capacity = ni + ni / 2;
        CU(ctx, cudaMalloc((void **)&s->d_first_tri, (s->inst_capacity + 1) * sizeof(uint32_t)));
        CU(ctx, cudaMalloc((void **)&s->d_inst, s->inst_capacity * sizeof(DeviceInstance)));
    }
    CU(ctx, cudaMemcpyAsync(s->d_first_tri, s->h_first_tri.data(), (ni + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (ni) CU(ctx, cudaMemcpyAsync(s->d_inst, s->h_inst.data(), ni * sizeof(DeviceInstance), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SOLB_OK;
}

// sRGB electro-optical transfer function (IEC 61966-2-1), the decode glTF prescribes for base-colour textures
static float srgb_to_linear(uint8_t c) {
    const float x = (float)c / 255.0f;
    return x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f);
}

SOLB_API int solb_scene_set_textures(solb_scene *s, const SolbTextureDesc *textures, uint32_t n_textures, const uint32_t *material_texture,
                                     uint32_t n_materials) {
    if (!s) return fail(nullptr, SOLB_ERR_INVALID, "null scene");
    solb_ctx *ctx = s->ctx;
    if (n_textures && !textures) return fail(ctx, SOLB_ERR_INVALID, "solb_scene_set_textures: null texture array");
    if (n_materials != s->materials.size()) return fail(ctx, SOLB_ERR_INVALID, "solb_scene_set_textures: one texture index per material of the scene");
    if (n_materials && !material_texture) return fail(ctx, SOLB_ERR_INVALID, "solb_scene_set_textures: null material_texture");
    for (uint32_t t = 0; t < n_textures; t++) {
        const SolbTextureDesc &d = textures[t];
        if (!d.rgba8 || d.width == 0 || d.height == 0 || d.width > 32768u || d.height > 32768u) return fail(ctx, SOLB_ERR_INVALID, "solb_scene_set_textures: bad texture");
        for (uint32_t wmode : { d.wrap_s, d.wrap_t })
            if (wmode != 0u && wmode != 10497u && wmode != 33071u && wmode != 33648u) return fail(ctx, SOLB_ERR_INVALID, "solb_scene_set_textures: unknown wrap mode");
    }
    for (uint32_t m = 0; m < n_materials; m++)
        if (material_texture[m] != SOLB_NO_TEXTURE && material_texture[m] >= n_textures) return fail(ctx, SOLB_ERR_INVALID, "solb_scene_set_textures: texture index out of range");
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));  // frames in flight may still sample the textures being replaced
    free_textures(s);
    float lut[256];
    for (int i = 0; i < 256; i++) lut[i] = srgb_to_linear((uint8_t)i);
    std::vector<TexDesc> descs(n_textures);
    std::vector<float> texels;
    for (uint32_t t = 0; t < n_textures; t++) {
        const SolbTextureDesc &d = textures[t];
        const size_t n = (size_t)d.width * d.height;
        texels.resize(n * 4);
        for (size_t i = 0; i < n; i++) {
            for (int c = 0; c < 3; c++) texels[4 * i + c] = d.srgb ? lut[d.rgba8[4 * i + c]] : (float)d.rgba8[4 * i + c] / 255.0f;
            texels[4 * i + 3] = (float)d.rgba8[4 * i + 3] / 255.0f;
        }
        float4 *dev = nullptr;
        CU(ctx, cudaMalloc((void **)&dev, n * sizeof(float4)));
        s->d_texels.push_back(dev);
        CU(ctx, cudaMemcpy(dev, texels.data(), n * sizeof(float4), cudaMemcpyHostToDevice));
        descs[t].texels = dev;
        descs[t].width = d.width;
        descs[t].height = d.height;
        descs[t].wrap_s = d.wrap_s ? d.wrap_s : 10497u;
        descs[t].wrap_t = d.wrap_t ? d.wrap_t : 10497u;
    }
    if (n_textures) {
        CU(ctx, cudaMalloc((void **)&s->d_tex, n_textures * sizeof(TexDesc)));
        CU(ctx, cudaMemcpy(s->d_tex, descs.data(), n_textures * sizeof(TexDesc), cudaMemcpyHostToDevice));
    }
    s->n_tex = n_textures;
    s->material_texture.assign(material_texture, material_texture + n_materials);
    // SceneInstance::texture_offset (src/ray/mod.rs:20) finally means something: the texture of the instance's material
    for (size_t i = 0; i < s->h_inst.size(); i++) {
        const uint32_t m = s->h_inst[i].material;
        const uint32_t tex = (n_textures && m < n_materials) ? material_texture[m] : SOLB_NO_TEXTURE;
        s->instances[i].texture_offset = n_textures ? tex : 0u;  // no textures bound: the reference's default 0
        const uint32_t t1 = tex == SOLB_NO_TEXTURE ? 0u : tex + 1u;
        memcpy(&s->h_inst[i].mat[10], &t1, sizeof(t1));
    }
    return upload_instances(s);
    SOLB_CATCH(ctx)
}

static int do_build(solb_scene *s) {
    solb_ctx *ctx = s->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    BuildOptions opt;
    // Treelet passes cost ~12 ns/triangle each on big scenes (244 ms per pass at 20 M) for a few % of trace speed there:
    // one pass above 4 M triangles, two below (profiles/r01_configs.jsonl).
    if (s->n_tris > (4u << 20)) opt.treelet_passes = 1;
    if (const char *v = getenv("SOLB_TREELET_PASSES")) opt.treelet_passes = std::max(0, std::min(8, atoi(v)));
    if (const char *v = getenv("SOLB_TREELET_COOP")) opt.coop_treelet = atoi(v) != 0;
    if (const char *v = getenv("SOLB_DP_COLLAPSE")) opt.dp_collapse = atoi(v) != 0;
    if (const char *v = getenv("SOLB_TREELET_GAMMA")) opt.treelet_gamma = std::max(3, std::min(1 << 20, atoi(v)));
    // PLOC instead of LBVH + treelets above 4 M triangles: 20 M triangles build in 24.5 ms instead of 60.9 (warm pool) for 5.7 %
    // more node visits per ray; below that the treelet build is a few milliseconds anyway and traces faster (tunnel.gltf:
    // 7.30 vs 7.80 nodes per ray).  SOLB_PLOC=0 / 1 forces either.
    opt.ploc = s->n_tris > (4u << 20) && s->accel_mode != SOLB_ACCEL_TWO_LEVEL;
    if (const char *v = getenv("SOLB_PLOC")) opt.ploc = atoi(v) != 0;
    // a wide node keeps its triangle (or instance-leaf) base in 28 bits (bvh.cuh: SOLB_TRI_BASE_MASK)
    if (s->n_tris > SOLB_TRI_BASE_MASK || s->h_inst.size() > SOLB_TRI_BASE_MASK)
        return fail(ctx, SOLB_ERR_OVERFLOW, "more than 2^28 - 1 triangles or instances");
    int rc = upload_instances(s);
    if (rc != SOLB_OK) return rc;
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    set_build_pool(ctx->build_pool);
    if (ctx->build_pool && s->n_tris > (1u << 20)) {
        // Grow the scratch pool in ONE step before a large build: left to the individual allocations the first 20 M-triangle
        // build of a process spent ~1 s mapping its ~9 GB of scratch piecemeal (bvh_build_ms_first in bench.py's extras).
        uint64_t reserved = 0;
        const uint64_t want = (uint64_t)s->n_tris * 480ull;
        if (cudaMemPoolGetAttribute(ctx->build_pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess && reserved < want) {
            void *grow = nullptr;
            if (cudaMallocFromPoolAsync(&grow, want, ctx->build_pool, ctx->stream) == cudaSuccess) cudaFreeAsync(grow, ctx->stream);
            cudaGetLastError();
        }
    }
    cudaError_t e = s->accel_mode == SOLB_ACCEL_TWO_LEVEL ? build_accel_two_level(ctx->stream, s->view(), s->accel, opt, &ctx->launches)
                                                          : build_accel(ctx->stream, s->view(), s->accel, opt, &ctx->launches);
    if (e == cudaErrorLaunchOutOfResources) return fail(ctx, SOLB_ERR_OVERFLOW, "acceleration structure deeper than the traversal stack");
    if (e != cudaSuccess) return fail_cuda(ctx, e, "build_accel");
    CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CU(ctx, cudaEventSynchronize(ctx->ev1));
    CU(ctx, cudaEventElapsedTime(&ctx->last_build_ms, ctx->ev0, ctx->ev1));
    s->built = true;
    s->dirty = false;
    s->needs_build = false;
    return SOLB_OK;
}

SOLB_API int solb_accel_build(solb_scene *s) {
    if (!s) return fail(nullptr, SOLB_ERR_INVALID, "null scene");
    SOLB_TRY
    return do_build(s);
    SOLB_CATCH(s->ctx)
}

SOLB_API int solb_instance_set_transform(solb_scene *s, uint32_t index, const float transform[16]) {
    if (!s || !transform) return fail(s ? s->ctx : nullptr, SOLB_ERR_INVALID, "solb_instance_set_transform: null argument");
    if (index >= s->instances.size()) return fail(s->ctx, SOLB_ERR_INVALID, "instance index out of range");
    SolbSceneInstance &si = s->instances[index];
    memcpy(si.transform, transform, sizeof(si.transform));
    float inv[16];
    mat4_inverse(transform, inv);
    mat4_transpose(inv, si.transform_it);  // SceneInstance::update_transform, src/ray/mod.rs:27-30
    memcpy(s->h_inst[index].transform, si.transform, sizeof(si.transform));
    memcpy(s->h_inst[index].transform_it, si.transform_it, sizeof(si.transform_it));
    s->dirty = true;
    return SOLB_OK;
}

SOLB_API int solb_scene_update(solb_scene *s) {
    if (!s) return fail(nullptr, SOLB_ERR_INVALID, "null scene");
    solb_ctx *ctx = s->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    return upload_instances(s);
}

SOLB_API int solb_tlas_regenerate(solb_scene *s) {
    if (!s) return fail(nullptr, SOLB_ERR_INVALID, "null scene");
    if (!s->built) return fail(s->ctx, SOLB_ERR_NOT_BUILT, "solb_tlas_regenerate before solb_accel_build");
    if (!s->dirty && !s->needs_build) return SOLB_OK;
    SOLB_TRY
    solb_ctx *ctx = s->ctx;
    if (s->needs_build || s->accel_mode != SOLB_ACCEL_TWO_LEVEL || !s->accel.two_level) return do_build(s);  // flattened: the rebuild bakes the transforms
    // two-level: the BLASes are untouched, only the TLAS over the moved instances is rebuilt
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = upload_instances(s);
    if (rc != SOLB_OK) return rc;
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    set_build_pool(ctx->build_pool);
    cudaError_t e = rebuild_tlas(ctx->stream, s->view(), s->accel, &ctx->launches);
    if (e == cudaErrorLaunchOutOfResources) return fail(ctx, SOLB_ERR_OVERFLOW, "acceleration structure deeper than the traversal stack");
    if (e != cudaSuccess) return fail_cuda(ctx, e, "rebuild_tlas");
    CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CU(ctx, cudaEventSynchronize(ctx->ev1));
    CU(ctx, cudaEventElapsedTime(&ctx->last_build_ms, ctx->ev0, ctx->ev1));
    s->dirty = false;
    return SOLB_OK;
    SOLB_CATCH(s->ctx)
}

SOLB_API int solb_scene_add_instance(solb_scene *s, uint32_t source_instance, const float transform[16], uint32_t material_index,
                                     uint32_t *out_index) {
    if (!s || !transform) return fail(s ? s->ctx : nullptr, SOLB_ERR_INVALID, "solb_scene_add_instance: null argument");
    if (source_instance >= s->h_inst.size()) return fail(s->ctx, SOLB_ERR_INVALID, "source instance out of range");
    if (material_index >= s->materials.size()) return fail(s->ctx, SOLB_ERR_INVALID, "material index out of range");
    SOLB_TRY
    const DeviceInstance src = s->h_inst[source_instance];
    if ((uint64_t)s->n_tris + src.n_indices / 3 > 0x7fffffffull) return fail(s->ctx, SOLB_ERR_INVALID, "too many triangles");
    SolbSceneInstance si;
    memset(&si, 0, sizeof(si));
    si.id = (uint32_t)s->instances.size();
    memcpy(si.transform, transform, sizeof(si.transform));
    float inv[16];
    mat4_inverse(transform, inv);
    mat4_transpose(inv, si.transform_it);
    DeviceInstance di = src;  // same BLAS: geometry range, blas, shade_first_tri
    di.material = material_index;
    memcpy(di.transform, si.transform, sizeof(di.transform));
    memcpy(di.transform_it, si.transform_it, sizeof(di.transform_it));
    memcpy(di.mat, &s->materials[material_index], sizeof(di.mat));
    di.mat[10] = 0.0f;
    if (material_index < s->material_texture.size() && s->material_texture[material_index] < s->n_tex) {
        si.texture_offset = s->material_texture[material_index];
        const uint32_t t1 = si.texture_offset + 1u;
        memcpy(&di.mat[10], &t1, sizeof(t1));
    } else if (s->n_tex) si.texture_offset = SOLB_NO_TEXTURE;
    s->instances.push_back(si);
    s->h_inst.push_back(di);
    s->n_tris += di.n_indices / 3;
    s->h_first_tri.push_back(s->n_tris);
    s->needs_build = true;
    if (out_index) *out_index = si.id;
    return SOLB_OK;
    SOLB_CATCH(s->ctx)
}

SOLB_API int solb_scene_set_accel_mode(solb_scene *s, uint32_t mode) {
    if (!s) return fail(nullptr, SOLB_ERR_INVALID, "null scene");
    if (mode > SOLB_ACCEL_TWO_LEVEL) return fail(s->ctx, SOLB_ERR_INVALID, "unknown acceleration-structure mode");
    if (mode != s->accel_mode) { s->accel_mode = mode; s->needs_build = true; }
    return SOLB_OK;
}

SOLB_API int solb_scene_instance_count(solb_scene *s, uint32_t *out) {
    if (!s || !out) return fail(s ? s->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    *out = (uint32_t)s->instances.size();
    return SOLB_OK;
}

SOLB_API int solb_scene_get_instances(solb_scene *s, SolbSceneInstance *out, uint32_t capacity) {
    if (!s || (!out && capacity)) return fail(s ? s->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    if (capacity < s->instances.size()) return fail(s->ctx, SOLB_ERR_INVALID, "instance buffer too small");
    if (!s->instances.empty()) memcpy(out, s->instances.data(), s->instances.size() * sizeof(SolbSceneInstance));
    return SOLB_OK;
}

SOLB_API int solb_scene_instance_triangles(solb_scene *s, uint32_t *out, uint32_t capacity) {
    if (!s || (!out && capacity)) return fail(s ? s->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    if (capacity < s->h_inst.size()) return fail(s->ctx, SOLB_ERR_INVALID, "buffer too small");
    for (size_t i = 0; i < s->h_inst.size(); i++) out[i] = s->h_inst[i].n_indices / 3;
    return SOLB_OK;
}

SOLB_API int solb_accel_info(solb_scene *s, SolbAccelInfo *out) {
    if (!s || !out) return fail(s ? s->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    if (!s->built) return fail(s->ctx, SOLB_ERR_NOT_BUILT, "acceleration structure not built");
    memset(out, 0, sizeof(*out));
    out->n_instances = (uint32_t)s->instances.size();
    out->n_triangles = s->accel.n_tris;
    out->mode = s->accel.two_level ? SOLB_ACCEL_TWO_LEVEL : SOLB_ACCEL_FLAT;
    out->n_blas = (uint32_t)s->h_blas.size();
    out->n_tlas_nodes = s->accel.n_tlas_wide;
    out->tlas_depth = s->accel.tlas_depth;
    out->n_wide_nodes = s->accel.n_wide;
    out->wide_depth = s->accel.depth;
    out->n_binary_nodes = s->accel.n_binary;
    out->sah_cost_binary = s->accel.sah_final;
    out->sah_cost_lbvh = s->accel.sah_lbvh;
    for (int k = 0; k < 3; k++) { out->scene_lo[k] = s->accel.lo[k]; out->scene_hi[k] = s->accel.hi[k]; }
    return SOLB_OK;
}

SOLB_API int solb_accel_read_nodes(solb_scene *s, void *host, size_t bytes) {
    if (!s || !host) return fail(s ? s->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    if (!s->built) return fail(s->ctx, SOLB_ERR_NOT_BUILT, "acceleration structure not built");
    if (bytes != (size_t)s->accel.n_wide * sizeof(Node8)) return fail(s->ctx, SOLB_ERR_INVALID, "size must be n_wide_nodes * 80");
    CU(s->ctx, cudaSetDevice(s->ctx->device));
    CU(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    CU(s->ctx, cudaMemcpy(host, s->accel.nodes, bytes, cudaMemcpyDeviceToHost));
    return SOLB_OK;
}

SOLB_API int solb_accel_read_triangles(solb_scene *s, void *host, size_t bytes) {
    if (!s || (!host && bytes)) return fail(s ? s->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    if (!s->built) return fail(s->ctx, SOLB_ERR_NOT_BUILT, "acceleration structure not built");
    if (bytes != (size_t)s->accel.n_tris * sizeof(Tri48)) return fail(s->ctx, SOLB_ERR_INVALID, "size must be n_triangles * 48");
    CU(s->ctx, cudaSetDevice(s->ctx->device));
    CU(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    if (bytes) CU(s->ctx, cudaMemcpy(host, s->accel.tris, bytes, cudaMemcpyDeviceToHost));
    return SOLB_OK;
}

// ---- targets -----------------------------------------------------------------------------------------

SOLB_API int solb_target_create(solb_ctx *ctx, uint32_t width, uint32_t height, uint32_t format, solb_target **out) {
    if (!ctx || !out) return fail(ctx, SOLB_ERR_INVALID, "solb_target_create: null argument");
    *out = nullptr;
    if (format > SOLB_FORMAT_RG32UI) return fail(ctx, SOLB_ERR_INVALID, "unknown target format");
    if (width == 0 || height == 0) return fail(ctx, SOLB_ERR_INVALID, "zero-sized target (src/texture.rs:45 asserts extent > 0)");
    if ((uint64_t)width * height > 0x7fffffffull) return fail(ctx, SOLB_ERR_INVALID, "target too large");
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    solb_target *t = new solb_target();
    t->ctx = ctx; t->width = width; t->height = height; t->format = format;
    t->bytes = (size_t)width * height * format_bytes(format);
    cudaError_t e = cudaMalloc(&t->dev, t->bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(t->dev, 0, t->bytes, ctx->stream);
    if (e != cudaSuccess) { cudaFree(t->dev); delete t; return fail_cuda(ctx, e, "target alloc"); }
    ctx->refs++;
    *out = t;
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

SOLB_API int solb_target_destroy(solb_target *t) {
    if (!t) return SOLB_OK;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->dev);
    solb_ctx *ctx = t->ctx;
    delete t;
    ctx_release(ctx);
    return SOLB_OK;
}

SOLB_API int solb_target_clear(solb_target *t) {
    if (!t) return fail(nullptr, SOLB_ERR_INVALID, "null target");
    CU(t->ctx, cudaSetDevice(t->ctx->device));
    CU(t->ctx, cudaMemsetAsync(t->dev, 0, t->bytes, t->ctx->stream));
    return SOLB_OK;
}

SOLB_API int solb_target_readback(solb_target *t, void *host, size_t bytes) {
    if (!t || !host) return fail(t ? t->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    if (bytes != t->bytes) return fail(t->ctx, SOLB_ERR_INVALID, "readback size must equal width*height*texel size");
    CU(t->ctx, cudaSetDevice(t->ctx->device));
    CU(t->ctx, cudaMemcpyAsync(host, t->dev, bytes, cudaMemcpyDeviceToHost, t->ctx->stream));
    CU(t->ctx, cudaStreamSynchronize(t->ctx->stream));
    return SOLB_OK;
}

SOLB_API int solb_target_readback_async(solb_target *t, void *host, size_t bytes) {
    if (!t || !host) return fail(t ? t->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    if (bytes != t->bytes) return fail(t->ctx, SOLB_ERR_INVALID, "readback size must equal width*height*texel size");
    CU(t->ctx, cudaSetDevice(t->ctx->device));
    CU(t->ctx, cudaMemcpyAsync(host, t->dev, bytes, cudaMemcpyDeviceToHost, t->ctx->stream));
    return SOLB_OK;
}

SOLB_API int solb_target_upload(solb_target *t, const void *host, size_t bytes) {
    if (!t || !host) return fail(t ? t->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    if (bytes != t->bytes) return fail(t->ctx, SOLB_ERR_INVALID, "upload size must equal width*height*texel size");
    CU(t->ctx, cudaSetDevice(t->ctx->device));
    CU(t->ctx, cudaMemcpyAsync(t->dev, host, bytes, cudaMemcpyHostToDevice, t->ctx->stream));
    CU(t->ctx, cudaStreamSynchronize(t->ctx->stream));
    return SOLB_OK;
}

SOLB_API int solb_target_device_ptr(solb_target *t, void **out) {
    if (!t || !out) return fail(t ? t->ctx : nullptr, SOLB_ERR_INVALID, "null argument");
    *out = t->dev;
    return SOLB_OK;
}

SOLB_API int solb_target_info(solb_target *t, uint32_t *width, uint32_t *height, uint32_t *format) {
    if (!t) return fail(nullptr, SOLB_ERR_INVALID, "null target");
    if (width) *width = t->width;
    if (height) *height = t->height;
    if (format) *format = t->format;
    return SOLB_OK;
}

// ---- trace -------------------------------------------------------------------------------------------

SOLB_API void solb_trace_params_default(SolbTraceParams *p, int pipeline) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->accum_start_frame = 0;
    p->enable_sky = 0;
    p->samples_per_frame = pipeline == 1 ? 4u : 8u;  // ao.rgen:43 / pathtrace.rgen:44
    p->max_bounces = pipeline == 1 ? 4u : 32u;       // ao.rgen:42 / pathtrace.rgen:43
    p->schedule = SOLB_SCHEDULE_AUTO;
    p->accum_mode = SOLB_ACCUM_MIX;
}

static void fill_frame_consts(FrameConsts &fc, const SolbSceneUniforms *u, uint32_t w, uint32_t h) {
    memset(&fc, 0, sizeof(fc));
    memcpy(fc.view_inv, u->view_inverse, sizeof(fc.view_inv));
    memcpy(fc.proj_inv, u->projection_inverse, sizeof(fc.proj_inv));
    // origin = view_inverse * vec4(0,0,0,1) (pathtrace.rgen:55)
    fc.origin = make_float3(u->view_inverse[12], u->view_inverse[13], u->view_inverse[14]);
    const float len = sqrtf(fc.origin.x * fc.origin.x + fc.origin.y * fc.origin.y + fc.origin.z * fc.origin.z);
    fc.tmin = fmaxf(1.0f, len) * 1e-3f;  // preparePayload, pathtrace.rgen:35
    fc.tmax = 10000.0f;
    // gl_LaunchSizeEXT = the extent passed to cmd_trace_rays = the target size (src/ray/sbt.rs:175-177)
    fc.width = w;
    fc.height = h;
    fc.row_begin = 0;
    fc.band_rows = h;
    fc.band_stride = 0;
    fc.n_bands = 1;
    fc.frame = u->frame[2];
}

// SolbTraceParams::tile_row_begin / tile_row_count -> FrameConsts rows (0 rows = the whole image)
static int apply_tile(solb_ctx *ctx, FrameConsts &fc, const SolbTraceParams *p) {
    if (p->tile_row_count == 0 && p->tile_row_begin == 0 && p->tile_row_stride == 0) return SOLB_OK;
    if (p->tile_row_count == 0 || p->tile_row_begin >= fc.height ||
        (p->tile_row_stride == 0 && (uint64_t)p->tile_row_begin + p->tile_row_count > fc.height))
        return fail(ctx, SOLB_ERR_INVALID, "tile rows outside the target");
    if (p->tile_row_stride != 0 && p->tile_row_stride < p->tile_row_count)
        return fail(ctx, SOLB_ERR_INVALID, "tile_row_stride smaller than tile_row_count: bands would overlap");
    fc.row_begin = p->tile_row_begin;
    fc.band_rows = p->tile_row_count;
    fc.band_stride = p->tile_row_stride;
    fc.n_bands = p->tile_row_stride ? (fc.height - p->tile_row_begin + p->tile_row_stride - 1) / p->tile_row_stride : 1u;
    return SOLB_OK;
}

static int check_trace_args(solb_scene *s, const SolbSceneUniforms *u) {
    if (!s || !u) return fail(s ? s->ctx : nullptr, SOLB_ERR_INVALID, "trace: null scene/uniforms");
    if (!s->built) return fail(s->ctx, SOLB_ERR_NOT_BUILT, "trace before solb_accel_build");
    if (s->needs_build) return fail(s->ctx, SOLB_ERR_NOT_BUILT, "instances were added or the mode changed: call solb_accel_build / solb_tlas_regenerate first");
    return SOLB_OK;
}

static int check_target(solb_ctx *ctx, solb_target *t, uint32_t format, const solb_target *like, const char *name) {
    if (!t) return fail(ctx, SOLB_ERR_INVALID, std::string(name) + " target is NULL");
    if (t->ctx != ctx) return fail(ctx, SOLB_ERR_INVALID, std::string(name) + " target belongs to another ctx");
    if (t->format != format) return fail(ctx, SOLB_ERR_INVALID, std::string(name) + " target has the wrong format");
    if (like && (like->width != t->width || like->height != t->height))
        return fail(ctx, SOLB_ERR_INVALID, std::string(name) + " target size differs from the other targets");
    return SOLB_OK;
}

static int ensure_wavefront(solb_ctx *ctx, uint32_t n_pixels) {
    WavefrontState &w = ctx->ws;
    if (w.capacity >= n_pixels && w.ray_o) return SOLB_OK;
    free_wavefront(ctx);
    const size_t n = n_pixels;
    CU(ctx, cudaMalloc((void **)&w.ray_o, n * sizeof(float4)));
    CU(ctx, cudaMalloc((void **)&w.ray_d, n * sizeof(float4)));
    CU(ctx, cudaMalloc((void **)&w.thr, n * sizeof(float4)));
    CU(ctx, cudaMalloc((void **)&w.pix, n * sizeof(float4)));
    CU(ctx, cudaMalloc((void **)&w.hit, n * sizeof(uint4)));
    CU(ctx, cudaMalloc((void **)&w.queue[0], n * sizeof(uint32_t)));
    CU(ctx, cudaMalloc((void **)&w.queue[1], n * sizeof(uint32_t)));
    CU(ctx, cudaMalloc((void **)&w.counters, 4 * sizeof(uint32_t)));
    for (int k = 1; k < WF_MAX_PARTS; k++) {
        CU(ctx, cudaMalloc((void **)&ctx->part_queue[k][0], n * sizeof(uint32_t)));
        CU(ctx, cudaMalloc((void **)&ctx->part_queue[k][1], n * sizeof(uint32_t)));
        CU(ctx, cudaMalloc((void **)&ctx->part_counters[k], 4 * sizeof(uint32_t)));
        if (!ctx->part_stream[k]) {
            int prio_lo = 0, prio_hi = 0;  // numerically lower = higher priority
            CU(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            const int prio = ctx->tune.part_priority ? std::max(prio_hi, prio_lo - k) : prio_lo;
            CU(ctx, cudaStreamCreateWithPriority(&ctx->part_stream[k], cudaStreamNonBlocking, prio));
            CU(ctx, cudaEventCreateWithFlags(&ctx->ev_join[k], cudaEventDisableTiming));
        }
    }
    for (int k = 0; k < WF_MAX_PARTS; k++)
        if (!ctx->ev_poll[k]) CU(ctx, cudaEventCreateWithFlags(&ctx->ev_poll[k], cudaEventDisableTiming));
    for (int k = 0; k < WF_MAX_PARTS && ctx->tune.shade_priority; k++)
        if (!ctx->shade_stream[k]) {
            int prio_lo = 0, prio_hi = 0;
            CU(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CU(ctx, cudaStreamCreateWithPriority(&ctx->shade_stream[k], cudaStreamNonBlocking, prio_hi));
            CU(ctx, cudaEventCreateWithFlags(&ctx->ev_ts[k], cudaEventDisableTiming));
            CU(ctx, cudaEventCreateWithFlags(&ctx->ev_st[k], cudaEventDisableTiming));
        }
    if (!ctx->ev_fork) CU(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    if (ctx->tune.pool)
        for (int k = 0; k < WF_MAX_PARTS; k++) CU(ctx, cudaMalloc((void **)&ctx->part_spill[k], pool_spill_bytes(ctx->sm_count, ctx->tune)));
    w.capacity = n_pixels;
    return SOLB_OK;
}

static int ensure_warpfront_slot(solb_ctx *ctx, int k, size_t n_pixels);

// Every frame slot in use is sized by the first frame that needs it (not slot k by the k-th frame): allocation and the
// synchronisation it needs then fall into the first frame of a size - a warm-up frame - instead of the first three.
static int ensure_warpfront(solb_ctx *ctx, size_t n_pixels) {
    const int n_slots = std::max(1, std::min(ctx->tune.wl_frames_in_flight, (int)WL_MAX_FRAMES));
    for (int k = 0; k < n_slots; k++) {
        const int rc = ensure_warpfront_slot(ctx, k, n_pixels);
        if (rc != SOLB_OK) return rc;
    }
    return SOLB_OK;
}

static int ensure_warpfront_slot(solb_ctx *ctx, int k, size_t n_pixels) {
    WarpfrontState &w = ctx->wl[k];
    const uint32_t n_warps = warpfront_grid_warps(ctx->sm_count, ctx->tune);
    if (!ctx->frame_stream[k]) {
        CU(ctx, cudaStreamCreateWithFlags(&ctx->frame_stream[k], cudaStreamNonBlocking));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_trace_done[k], cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&ctx->ev_resolve_done[k], cudaEventDisableTiming));
        CU(ctx, cudaEventCreate(&ctx->ev_k0[k]));
        CU(ctx, cudaEventCreate(&ctx->ev_k1[k]));
    }
    if (!ctx->ev_serial) CU(ctx, cudaEventCreateWithFlags(&ctx->ev_serial, cudaEventDisableTiming));
    if (!(w.n_warps >= n_warps && w.ray_o)) {
        CU(ctx, cudaStreamSynchronize(ctx->frame_stream[k]));
        cudaFree(w.ray_o); cudaFree(w.ray_d); cudaFree(w.ray_i); cudaFree(w.frame0); cudaFree(w.frame1);
        cudaFree(w.thr); cudaFree(w.pix); cudaFree(w.hit); cudaFree(w.cursor);
        w = WarpfrontState{};
        const size_t n = (size_t)n_warps * WL_POOL;
        CU(ctx, cudaMalloc((void **)&w.ray_o, n * sizeof(float4)));
        CU(ctx, cudaMalloc((void **)&w.ray_d, n * sizeof(float4)));
        CU(ctx, cudaMalloc((void **)&w.ray_i, n * sizeof(float4)));
        CU(ctx, cudaMalloc((void **)&w.frame0, n * sizeof(float4)));
        CU(ctx, cudaMalloc((void **)&w.frame1, n * sizeof(float4)));
        CU(ctx, cudaMalloc((void **)&w.thr, n * sizeof(float4)));
        CU(ctx, cudaMalloc((void **)&w.pix, n * sizeof(float4)));
        CU(ctx, cudaMalloc((void **)&w.hit, n * sizeof(uint4)));
        CU(ctx, cudaMalloc((void **)&w.cursor, sizeof(uint32_t)));
        w.n_warps = n_warps;
    }
    if (ctx->frame_sum_pixels[k] < n_pixels) {
        // (the previous users of this buffer, a kernel on the side stream and a resolve on the ctx stream, must be done)
        CU(ctx, cudaStreamSynchronize(ctx->frame_stream[k]));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->frame_sum[k]);
        ctx->frame_sum[k] = nullptr;
        ctx->frame_sum_pixels[k] = 0;
        CU(ctx, cudaMalloc((void **)&ctx->frame_sum[k], n_pixels * sizeof(float4)));
        ctx->frame_sum_pixels[k] = n_pixels;
    }
    return SOLB_OK;
}

struct TraceTimer {
    solb_ctx *c;
    explicit TraceTimer(solb_ctx *ctx) : c(ctx) { if (c->timing) cudaEventRecord(c->ev0, c->stream); }
    void stop() {
        if (!c->timing) return;
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        cudaEventElapsedTime(&c->last_trace_ms, c->ev0, c->ev1);
    }
};

SOLB_API int solb_trace_pathtrace(solb_scene *s, const SolbSceneUniforms *u, const SolbTraceParams *params, solb_target *accum,
                                  solb_target *render) {
    int rc = check_trace_args(s, u);
    if (rc) return rc;
    solb_ctx *ctx = s->ctx;
    if (!params) return fail(ctx, SOLB_ERR_INVALID, "trace: null params");
    if ((rc = check_target(ctx, accum, SOLB_FORMAT_RGBA32F, nullptr, "accum"))) return rc;
    if (render && (rc = check_target(ctx, render, SOLB_FORMAT_RGBA8, accum, "render"))) return rc;
    if (params->samples_per_frame == 0 || params->samples_per_frame > 0xffffu || params->max_bounces > 0xfff0u)
        return fail(ctx, SOLB_ERR_INVALID, "samples_per_frame must be 1..65535 and max_bounces <= 65520");
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    FrameConsts fc;
    fill_frame_consts(fc, u, accum->width, accum->height);
    fc.texb = s->texb();
    fc.accum_start = params->accum_start_frame;
    fc.enable_sky = params->enable_sky;
    fc.spp = params->samples_per_frame;
    fc.max_bounces = params->max_bounces;
    fc.accum_mode = params->accum_mode;
    if ((rc = apply_tile(ctx, fc, params))) return rc;
    TraceTimer timer(ctx);
    uint32_t n_ev = 0;
    uint32_t schedule = params->schedule;
    if (schedule == SOLB_SCHEDULE_AUTO) schedule = s->accel.n_wide <= 8 ? SOLB_SCHEDULE_MEGAKERNEL : ctx->auto_wide_schedule;
    if (schedule > SOLB_SCHEDULE_WARPFRONT) return fail(ctx, SOLB_ERR_INVALID, "trace: unknown schedule");
    if (schedule == SOLB_SCHEDULE_WARPFRONT) {
        // The kernel runs on a side stream and writes only its own path state and the frame's per-pixel sums; the resolve into
        // the targets follows on the ctx stream.  The two frame slots alternate, so the kernel of this frame may start while
        // the previous frame's kernel drains (its last pixels are chains of ~50 rays each: ~9 % of a 1080p frame with the SMs
        // emptying), and everything later on the ctx stream is ordered after this frame's resolve as before.
        const int k = (int)(ctx->frame_slot = (ctx->frame_slot + 1u) % (uint32_t)std::max(ctx->tune.wl_frames_in_flight, 1));
        if ((rc = ensure_warpfront(ctx, (size_t)fc.width * fc.height))) return rc;
        cudaStream_t side = ctx->frame_stream[k];
        const bool serial = ctx->timing || ctx->tune.wl_frames_in_flight < 2;
        if (serial) {  // after everything already on the ctx stream (incl. the previous frame's resolve)
            CU(ctx, cudaEventRecord(ctx->ev_serial, ctx->stream));
            CU(ctx, cudaStreamWaitEvent(side, ctx->ev_serial, 0));
        } else {
            if (ctx->have_side_barrier) CU(ctx, cudaStreamWaitEvent(side, ctx->ev_side_barrier, 0));  // scene / stats changes
            if (ctx->have_resolve_done[k]) CU(ctx, cudaStreamWaitEvent(side, ctx->ev_resolve_done[k], 0));  // frame_sum[k] consumed
        }
        if (ctx->timing) CU(ctx, cudaEventRecord(ctx->ev_k0[k], side));
        CU(ctx, launch_pathtrace_warpfront(side, fc, s->accel, s->d_inst, s->d_shade, ctx->wl[k], ctx->frame_sum[k], ctx->d_stats,
                                           params->collect_stats != 0, ctx->sm_count, ctx->tune));
        if (ctx->timing) CU(ctx, cudaEventRecord(ctx->ev_k1[k], side));
        CU(ctx, cudaEventRecord(ctx->ev_trace_done[k], side));
        cudaStream_t rs;
        CU(ctx, hi_begin(ctx, &rs));
        CU(ctx, cudaStreamWaitEvent(rs, ctx->ev_trace_done[k], 0));
        CU(ctx, launch_warpfront_resolve(rs, fc, ctx->frame_sum[k], (float4 *)accum->dev, render ? (uint32_t *)render->dev : nullptr));
        CU(ctx, cudaEventRecord(ctx->ev_resolve_done[k], rs));
        CU(ctx, hi_end(ctx));
        ctx->have_resolve_done[k] = true;
        ctx->launches += 2;
        timer.stop();
        if (ctx->timing) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ctx->ev_k0[k], ctx->ev_k1[k]) == cudaSuccess) ctx->trace_kernel_ms_total += ms;
            ctx->trace_kernel_launches += 1;
        }
        return SOLB_OK;
    }
    if (schedule == SOLB_SCHEDULE_MEGAKERNEL) {
        CU(ctx, launch_pathtrace_mega(ctx->stream, fc, s->accel, s->d_inst, s->d_shade, (float4 *)accum->dev,
                                      render ? (uint32_t *)render->dev : nullptr, ctx->d_stats, params->collect_stats != 0,
                                      (uint32_t *)(ctx->d_stats + 7), ctx->sm_count, ctx->tune));  // stats slot 7: pixel counter
        ctx->launches += 1;
        timer.stop();
        if (ctx->timing) { ctx->trace_kernel_ms_total += ctx->last_trace_ms; ctx->trace_kernel_launches += 1; }
        return SOLB_OK;
    } else {
        if ((rc = ensure_wavefront(ctx, fc.width * fc.height))) return rc;
        WavefrontLaunch L;
        L.stream[0] = ctx->stream;
        L.fork = ctx->ev_fork;
        for (int k = 0; k < WF_MAX_PARTS; k++) L.spill[k] = ctx->part_spill[k];
        L.ws[0] = ctx->ws;
        for (int k = 1; k < WF_MAX_PARTS; k++) {
            L.stream[k] = ctx->part_stream[k];
            L.join[k] = ctx->ev_join[k];
            L.ws[k] = ctx->ws;
            L.ws[k].queue[0] = ctx->part_queue[k][0];
            L.ws[k].queue[1] = ctx->part_queue[k][1];
            L.ws[k].counters = ctx->part_counters[k];
        }
        for (int k = 0; k < WF_MAX_PARTS; k++) {
            L.poll[k] = ctx->ev_poll[k];
            L.shade_stream[k] = ctx->shade_stream[k];
            L.ev_ts[k] = ctx->ev_ts[k];
            L.ev_st[k] = ctx->ev_st[k];
        }
        L.host_counts = ctx->pinned_count;
        L.sm_count = ctx->sm_count;
        CU(ctx, launch_pathtrace_wavefront(L, fc, s->accel, s->d_inst, s->d_shade, (float4 *)accum->dev,
                                           render ? (uint32_t *)render->dev : nullptr, ctx->d_stats, params->collect_stats != 0,
                                           &ctx->launches, ctx->timing ? &ctx->ev_pool : nullptr, &n_ev, ctx->tune));
    }
    timer.stop();  // synchronises in timing mode
    for (uint32_t i = 0; ctx->timing && i + 1 < n_ev; i += 2) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]) == cudaSuccess) ctx->trace_kernel_ms_total += ms;
        ctx->trace_kernel_launches += 1;
    }
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

SOLB_API int solb_set_blue_noise(solb_ctx *ctx, const uint8_t *rgba8, uint32_t width, uint32_t height) {
    if (!ctx || !rgba8 || !width || !height) return fail(ctx, SOLB_ERR_INVALID, "solb_set_blue_noise: bad argument");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_blue);
    ctx->d_blue = nullptr;
    CU(ctx, cudaMalloc((void **)&ctx->d_blue, (size_t)width * height * 4));
    CU(ctx, cudaMemcpy(ctx->d_blue, rgba8, (size_t)width * height * 4, cudaMemcpyHostToDevice));
    ctx->blue_w = width;
    ctx->blue_h = height;
    return SOLB_OK;
}

SOLB_API int solb_trace_ao(solb_scene *s, const SolbSceneUniforms *u, const SolbTraceParams *params, solb_target *image) {
    int rc = check_trace_args(s, u);
    if (rc) return rc;
    solb_ctx *ctx = s->ctx;
    if (!params) return fail(ctx, SOLB_ERR_INVALID, "trace: null params");
    if ((rc = check_target(ctx, image, SOLB_FORMAT_RGBA32F, nullptr, "image"))) return rc;
    if (!ctx->d_blue) return fail(ctx, SOLB_ERR_INVALID, "solb_trace_ao before solb_set_blue_noise (binding 2, examples/4-ray-ao.rs:282)");
    if (params->samples_per_frame == 0 || params->max_bounces == 0) return fail(ctx, SOLB_ERR_INVALID, "ao: zero samples");
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    FrameConsts fc;
    fill_frame_consts(fc, u, image->width, image->height);
    fc.accum_start = params->accum_start_frame;
    fc.spp = params->samples_per_frame;
    fc.max_bounces = params->max_bounces;
    if ((rc = apply_tile(ctx, fc, params))) return rc;
    TraceTimer timer(ctx);
    CU(ctx, launch_ao(ctx->stream, fc, s->accel, s->d_inst, s->d_shade, ctx->d_blue, ctx->blue_w, ctx->blue_h, (float4 *)image->dev,
                      ctx->d_stats, ctx->tune));
    ctx->launches += 1;
    timer.stop();
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

SOLB_API int solb_trace_debug(solb_scene *s, const SolbSceneUniforms *u, solb_target *render, solb_target *ids, solb_target *attribs) {
    int rc = check_trace_args(s, u);
    if (rc) return rc;
    solb_ctx *ctx = s->ctx;
    solb_target *first = render ? render : (ids ? ids : attribs);
    if (!first) return fail(ctx, SOLB_ERR_INVALID, "solb_trace_debug: no output target");
    if (render && (rc = check_target(ctx, render, SOLB_FORMAT_RGBA8, first, "render"))) return rc;
    if (ids && (rc = check_target(ctx, ids, SOLB_FORMAT_RG32UI, first, "ids"))) return rc;
    if (attribs && (rc = check_target(ctx, attribs, SOLB_FORMAT_RGBA32F, first, "attribs"))) return rc;
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    FrameConsts fc;
    fill_frame_consts(fc, u, first->width, first->height);
    TraceTimer timer(ctx);
    CU(ctx, launch_debug(ctx->stream, fc, s->accel, render ? (uint32_t *)render->dev : nullptr, ids ? (uint2 *)ids->dev : nullptr,
                         attribs ? (float4 *)attribs->dev : nullptr, ctx->d_stats, ctx->tune));
    ctx->launches += 1;
    timer.stop();
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

SOLB_API int solb_trace_rays(solb_scene *s, const float *rays, uint32_t n, uint32_t *hits, float *t_out) {
    if (!s) return fail(nullptr, SOLB_ERR_INVALID, "null scene");
    solb_ctx *ctx = s->ctx;
    if (!s->built || s->needs_build) return fail(ctx, SOLB_ERR_NOT_BUILT, "trace before solb_accel_build");
    if (n == 0) return SOLB_OK;
    if (!rays || !hits) return fail(ctx, SOLB_ERR_INVALID, "solb_trace_rays: null argument");
    SOLB_TRY
    CU(ctx, cudaSetDevice(ctx->device));
    float4 *d_rays = nullptr;
    uint4 *d_hits = nullptr;
    float *d_t = nullptr;
    struct Free { void *a, *b, *c; ~Free() { cudaFree(a); cudaFree(b); cudaFree(c); } };
    CU(ctx, cudaMalloc((void **)&d_rays, (size_t)n * 32));
    Free fr{ d_rays, nullptr, nullptr };
    CU(ctx, cudaMalloc((void **)&d_hits, (size_t)n * 16));
    fr.b = d_hits;
    CU(ctx, cudaMalloc((void **)&d_t, (size_t)n * 4));
    fr.c = d_t;
    CU(ctx, cudaMemcpyAsync(d_rays, rays, (size_t)n * 32, cudaMemcpyHostToDevice, ctx->stream));
    TraceTimer timer(ctx);
    CU(ctx, launch_trace_rays(ctx->stream, s->accel, d_rays, n, d_hits, d_t, ctx->d_stats, ctx->tune));
    ctx->launches += 1;
    timer.stop();
    CU(ctx, cudaMemcpyAsync(hits, d_hits, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (t_out) CU(ctx, cudaMemcpyAsync(t_out, d_t, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SOLB_OK;
    SOLB_CATCH(ctx)
}

SOLB_API int solb_test_sort_pairs(solb_ctx *ctx, uint64_t *keys, uint32_t *values, uint32_t n, int key_bits) {
    if (!ctx || ((!keys || !values) && n)) return fail(ctx, SOLB_ERR_INVALID, "solb_test_sort_pairs: null argument");
    if (n == 0) return SOLB_OK;
    if (key_bits < 1 || key_bits > 64) return fail(ctx, SOLB_ERR_INVALID, "key_bits must be 1..64");
    CU(ctx, cudaSetDevice(ctx->device));
    uint64_t *dk = nullptr;
    uint32_t *dv = nullptr;
    struct Free { void *a, *b; ~Free() { cudaFree(a); cudaFree(b); } } fr{ nullptr, nullptr };
    CU(ctx, cudaMalloc((void **)&dk, (size_t)n * 8));
    fr.a = dk;
    CU(ctx, cudaMalloc((void **)&dv, (size_t)n * 4));
    fr.b = dv;
    CU(ctx, cudaMemcpyAsync(dk, keys, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(dv, values, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    set_build_pool(ctx->build_pool);
    CU(ctx, sort_pairs_device(ctx->stream, dk, dv, n, key_bits));
    CU(ctx, cudaMemcpyAsync(keys, dk, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(values, dv, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return SOLB_OK;
}

SOLB_API int solb_resolve_sum(solb_ctx *ctx, solb_target *sum, solb_target *accum_out, solb_target *render) {
    if (!ctx) return fail(nullptr, SOLB_ERR_INVALID, "null ctx");
    int rc;
    if ((rc = check_target(ctx, sum, SOLB_FORMAT_RGBA32F, nullptr, "sum"))) return rc;
    if (accum_out && (rc = check_target(ctx, accum_out, SOLB_FORMAT_RGBA32F, sum, "accum_out"))) return rc;
    if (render && (rc = check_target(ctx, render, SOLB_FORMAT_RGBA8, sum, "render"))) return rc;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, launch_resolve_sum(ctx->stream, (const float4 *)sum->dev, accum_out ? (float4 *)accum_out->dev : nullptr,
                               render ? (uint32_t *)render->dev : nullptr, sum->width * sum->height));
    ctx->launches += 1;
    return SOLB_OK;
}

}  // extern "C"
