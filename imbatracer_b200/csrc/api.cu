/* api.cu -- C-ABI layer, part 1: error channel, scene upload, traversal on device and host buffers.
 * Signatures: include/imba_b200.h.  There is no CPU fallback anywhere in this library: without a CUDA
 * device every entry point fails (L2: negative status + imba_last_error(); L1: message + abort()). */
#include "api_common.h"
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <sched.h>

std::atomic<uint64_t> g_imba_launches{ 0 };
static thread_local std::string t_error;

void imba_set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    t_error = buf;
}

extern "C" const char* imba_last_error(void) { return t_error.c_str(); }
extern "C" uint64_t imba_launch_count(void) { return g_imba_launches.load(); }

template <typename T>
static int upload(T** dst, const void* src, size_t bytes) {
    *dst = nullptr;
    if (bytes == 0) bytes = 16;
    IMBA_CUDA(cudaMalloc((void**)dst, bytes), IMBA_ERR_CUDA);
    if (src) IMBA_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice), IMBA_ERR_CUDA);
    else     IMBA_CUDA(cudaMemset(*dst, 0, bytes), IMBA_ERR_CUDA);
    return IMBA_OK;
}

int imba_scene_upload_accel(imba_scene* s, const Bvh8& b) {
    if (b.max_depth + 11 > IMBA_TRAVERSAL_STACK) {   // + sentinel, remainders and the five entries of the saved world-space ray
        imba_set_error("acceleration structure too deep for the traversal stack (depth %d, stack %d)", b.max_depth, IMBA_TRAVERSAL_STACK);
        return IMBA_ERR_OVERFLOW;
    }
    int rc;
    if ((rc = upload(&s->d_nodes, b.nodes.data(), b.nodes.size() * sizeof(Node8)))) return rc;
    if ((rc = upload(&s->d_tris, b.tris.data(), b.tris.size() * sizeof(Tri48)))) return rc;
    if ((rc = upload(&s->d_insts, b.insts.data(), b.insts.size() * sizeof(Inst64)))) return rc;
    if ((rc = upload(&s->d_counters, nullptr, imba_scene::kCounters * sizeof(unsigned)))) return rc;
    s->top_root = b.flat ? b.flat_root : b.top_root; s->flat_inst = b.flat && !getenv("IMBA_NO_FLAT") ? b.flat_inst : -1;
    if (s->flat_inst < 0) s->top_root = b.top_root;
    s->stats[0] = (int64_t)b.nodes.size(); s->stats[1] = (int64_t)b.tris.size() - 1; s->stats[2] = (int64_t)b.insts.size();
    s->stats[3] = (int64_t)(b.nodes.size() * sizeof(Node8) + b.tris.size() * sizeof(Tri48) + b.insts.size() * sizeof(Inst64));
    s->stats[4] = b.max_depth;
    return IMBA_OK;
}

static int init_device(imba_scene* s) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { imba_set_error("no CUDA device (this library has no CPU path)"); return IMBA_ERR_NO_DEVICE; }
    IMBA_CUDA(cudaGetDevice(&s->device), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device), IMBA_ERR_CUDA);
    return IMBA_OK;
}

int imba_shading_upload(imba_scene* s, const imba_scene_desc* d); // shade_data.cu
void imba_shading_free(imba_scene* s);

extern "C" imba_scene* imba_scene_create(const imba_scene_desc* d) {
    if (!d) { imba_set_error("null scene description"); return nullptr; }
    imba_scene* s = new imba_scene;
    if (init_device(s) != IMBA_OK) { delete s; return nullptr; }
    if (d->accel_layout == IMBA_ACCEL_DEVICE_BUILD) { // no reference-built structure: LBVH + 8-wide collapse on the device
        if (imba_scene_build_accel_device(s, d, s->build_info) != IMBA_OK) { imba_scene_destroy(s); return nullptr; }
    } else {
        Bvh8 b; std::string err;
        if (!imba_build_bvh8(d->accel_layout, d->root, d->nodes, d->tris, d->inst_nodes, b, err)) {
            imba_set_error("BVH conversion failed: %s", err.c_str()); delete s; return nullptr;
        }
        if (imba_scene_upload_accel(s, b) != IMBA_OK) { imba_scene_destroy(s); return nullptr; }
    }
    // opacity masks: active only if some descriptor is not the always-opaque (1,1,0) (core/mask.h:14-26)
    for (int i = 0; i < d->mask_count; ++i)
        if (!(d->masks[i].width == 1 && d->masks[i].height == 1 && d->masks[i].offset == 0)) s->has_masks = true;
    if (s->has_masks) {
        if (!d->mask_indices || !d->mask_texcoords || !d->mask_buffer) { imba_set_error("mask arrays missing"); imba_scene_destroy(s); return nullptr; }
        const size_t ntri = d->mesh_count > 0 && d->mesh_tri_base ? (size_t)d->mesh_tri_base[d->mesh_count] : 0;
        if (upload(&s->d_mask_indices, d->mask_indices, ntri * 16) || upload(&s->d_mask_texcoords, d->mask_texcoords, (size_t)d->mask_texcoord_count * 8) ||
            upload(&s->d_masks, d->masks, (size_t)d->mask_count * 12) || upload(&s->d_mask_buffer, d->mask_buffer, (size_t)d->mask_buffer_size)) {
            imba_scene_destroy(s); return nullptr;
        }
    }
    if (d->mesh_count > 0 && d->vertices && imba_shading_upload(s, d) != IMBA_OK) { imba_scene_destroy(s); return nullptr; }
    return s;
}

extern "C" void imba_scene_destroy(imba_scene* s) {
    if (!s) return;
    cudaFree(s->d_nodes); cudaFree(s->d_tris); cudaFree(s->d_insts); cudaFree(s->d_counters);
    if (s->owns_masks) { cudaFree(s->d_mask_indices); cudaFree(s->d_mask_texcoords); cudaFree(s->d_masks); cudaFree(s->d_mask_buffer); }
    for (auto& set : s->stage) for (int i = 0; i < imba_scene::kStageSlots; ++i) {
        cudaFree(set.d_rays[i]); cudaFree(set.d_hits[i]);
        if (set.stream[i]) cudaStreamDestroy(set.stream[i]);
        if (set.done[i]) cudaEventDestroy(set.done[i]);
    }
    imba_shading_free(s);
    delete s;
}

extern "C" int imba_scene_build_info(const imba_scene* s, float out[2]) {
    if (!s || !out) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    out[0] = s->build_info[0]; out[1] = s->build_info[1];
    return IMBA_OK;
}

extern "C" int imba_scene_stats(const imba_scene* s, int64_t out[5]) {
    if (!s || !out) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    memcpy(out, s->stats, sizeof(s->stats));
    return IMBA_OK;
}

int imba_scene_trace_ex(imba_scene* s, const TraceRequest& rq, cudaStream_t stream) {
    if (!s || rq.capacity < 0 || (rq.capacity > 0 && !rq.d_rays)) { imba_set_error("invalid traversal arguments"); return IMBA_ERR_INVALID; }
    if (rq.capacity == 0) return IMBA_OK;
    const bool masks = s->has_masks, stats = rq.d_counters != nullptr;
    const int variant = rq.mode * 4 + (masks ? 2 : 0) + (stats ? 1 : 0);
    if (s->grid_cache[variant] == 0) {
        const int per_sm = imba_trace_blocks_per_sm(rq.mode, masks, stats, s->flat_inst >= 0);
        if (per_sm <= 0) { imba_set_error("traversal kernel cannot be resident (occupancy 0): %s", cudaGetErrorString(cudaGetLastError())); return IMBA_ERR_CUDA; }
        s->grid_cache[variant] = per_sm * s->sm_count;
    }
    const int warps_per_block = imba_trace_block_threads() / 32;
    const int needed = (rq.capacity + 32 * warps_per_block - 1) / (32 * warps_per_block);
    const int grid = needed < s->grid_cache[variant] ? needed : s->grid_cache[variant];
    unsigned* ctr = rq.d_work_counter ? rq.d_work_counter : s->d_counters + (s->next_counter.fetch_add(1) % imba_scene::kCounters);
    IMBA_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned), stream), IMBA_ERR_CUDA);
    TraceArgs a;
    a.nodes = s->d_nodes; a.tris = s->d_tris; a.insts = s->d_insts; a.top_root = s->top_root; a.flat_inst = s->flat_inst;
    a.rays = (const float4*)rq.d_rays; a.hits = (int4*)rq.d_hits; a.count = rq.capacity; a.count_ptr = rq.d_count; a.count_mul = rq.count_mul; a.counter = ctr; a.stats = rq.d_counters;
    a.indices = s->d_mask_indices; a.texcoords = s->d_mask_texcoords; a.masks = s->d_masks; a.mask_buffer = s->d_mask_buffer;
    a.shadow_states = rq.d_shadow_states; a.framebuffer = rq.d_framebuffer;
    IMBA_CUDA(imba_launch_trace(a, rq.mode, masks, stats, grid, stream), IMBA_ERR_CUDA);
    g_imba_launches.fetch_add(1);
    return IMBA_OK;
}

/* Scene part of the traversal arguments, for kernels that trace rays inline (traverse_core.cuh: trace_single). */
void imba_scene_trace_args(const imba_scene* s, TraceArgs* a) {
    memset(a, 0, sizeof(*a));
    a->nodes = s->d_nodes; a->tris = s->d_tris; a->insts = s->d_insts; a->top_root = s->top_root; a->flat_inst = s->flat_inst;
    a->indices = s->d_mask_indices; a->texcoords = s->d_mask_texcoords; a->masks = s->d_masks; a->mask_buffer = s->d_mask_buffer;
}

int imba_scene_trace(imba_scene* s, const imba_ray* d_rays, imba_hit* d_hits, int count, int any_hit, uint32_t* d_counters,
                     cudaStream_t stream) {
    if (count > 0 && !d_hits) { imba_set_error("invalid traversal arguments"); return IMBA_ERR_INVALID; }
    TraceRequest rq; rq.d_rays = d_rays; rq.d_hits = d_hits; rq.capacity = count; rq.mode = any_hit ? 1 : 0; rq.d_counters = d_counters;
    return imba_scene_trace_ex(s, rq, stream);
}

extern "C" int imba_trace(imba_scene* s, const imba_ray* d_rays, imba_hit* d_hits, int count, int any_hit,
                          uint32_t* d_counters, void* stream) {
    return imba_scene_trace(s, d_rays, d_hits, count, any_hit, d_counters, (cudaStream_t)stream);
}

/* RayQueue::traverse_gpu / traverse_cpu (ray_queue.h:247-343): rays and hits live in host memory.  A call takes one of the
 * scene's staging sets (concurrent callers get different ones and overlap; a fifth caller waits) and streams its chunks
 * through it. */
extern "C" int imba_trace_host(imba_scene* s, const imba_ray* h_rays, imba_hit* h_hits, int count, int any_hit) {
    if (!s || count < 0 || (count > 0 && (!h_rays || !h_hits))) { imba_set_error("invalid traversal arguments"); return IMBA_ERR_INVALID; }
    if (count == 0) return IMBA_OK;
    IMBA_CUDA(cudaSetDevice(s->device), IMBA_ERR_CUDA);   // callers may be worker threads that never selected a device
    imba_scene::StageSet* set = nullptr;
    {
        std::unique_lock<std::mutex> lock(s->stage_mutex);
        s->stage_free.wait(lock, [&] { for (auto& c : s->stage) if (!c.busy) return true; return false; });
        for (auto& c : s->stage) if (!c.busy) { set = &c; break; }
        set->busy = true;
    }
    struct Release { imba_scene* s; imba_scene::StageSet* set; ~Release() { { std::lock_guard<std::mutex> l(s->stage_mutex); set->busy = false; } s->stage_free.notify_one(); } } release{ s, set };
    const int chunk = imba_scene::kChunkRays;
    for (int i = 0; i < imba_scene::kStageSlots; ++i) {
        if (set->d_rays[i]) continue;
        IMBA_CUDA(cudaMalloc((void**)&set->d_rays[i], (size_t)chunk * sizeof(imba_ray)), IMBA_ERR_CUDA);
        IMBA_CUDA(cudaMalloc((void**)&set->d_hits[i], (size_t)chunk * sizeof(imba_hit)), IMBA_ERR_CUDA);
        IMBA_CUDA(cudaStreamCreateWithFlags(&set->stream[i], cudaStreamNonBlocking), IMBA_ERR_CUDA);
        IMBA_CUDA(cudaEventCreateWithFlags(&set->done[i], cudaEventDisableTiming | cudaEventBlockingSync), IMBA_ERR_CUDA);
    }
    int k = 0;
    for (int first = 0; first < count; first += chunk, k = (k + 1) % imba_scene::kStageSlots) {
        const int n = count - first < chunk ? count - first : chunk;
        cudaStream_t st = set->stream[k];
        IMBA_CUDA(cudaMemcpyAsync(set->d_rays[k], h_rays + first, (size_t)n * sizeof(imba_ray), cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
        const int rc = imba_scene_trace(s, (const imba_ray*)set->d_rays[k], (imba_hit*)set->d_hits[k], n, any_hit, nullptr, st);
        if (rc != IMBA_OK) return rc;
        IMBA_CUDA(cudaMemcpyAsync(h_hits + first, set->d_hits[k], (size_t)n * sizeof(imba_hit), cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    }
    // Wait for the last chunks: poll-and-yield for a short while (as quick as spinning when cores are free: small queues finish
    // within it), then SLEEP on the event (cudaEventBlockingSync).  Eight ranks with four calling threads each on a 32-core host
    // otherwise keep every core spinning in sched_yield() while the copy engines do the work (round 1: 3.2 instead of 8 x 1.4
    // Grays/s at 8 GPUs), and the threads that have to enqueue the next chunks wait for a time slice.
    for (int i = 0; i < imba_scene::kStageSlots; ++i) IMBA_CUDA(cudaEventRecord(set->done[i], set->stream[i]), IMBA_ERR_CUDA);
    static const int spin_polls = getenv("IMBA_HOST_SPIN_POLLS") ? atoi(getenv("IMBA_HOST_SPIN_POLLS")) : 256;
    for (int i = 0; i < imba_scene::kStageSlots; ++i) {
        cudaError_t e = cudaErrorNotReady;
        for (int k = 0; k < spin_polls && (e = cudaEventQuery(set->done[i])) == cudaErrorNotReady; ++k) sched_yield();
        if (e == cudaErrorNotReady) e = cudaEventSynchronize(set->done[i]);
        IMBA_CUDA(e, IMBA_ERR_CUDA);
    }
    return IMBA_OK;
}

/* Page-locks a host buffer that is handed to imba_trace_host repeatedly (a RayQueue's ray and hit arrays,
 * ray_queue.h:78-79) so that the chunk copies are truly asynchronous. */
extern "C" int imba_host_register(void* ptr, size_t bytes) {
    IMBA_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault), IMBA_ERR_CUDA);
    return IMBA_OK;
}
extern "C" int imba_host_unregister(void* ptr) {
    IMBA_CUDA(cudaHostUnregister(ptr), IMBA_ERR_CUDA);
    return IMBA_OK;
}

/* Host-only conversion (no CUDA device needed): lets a host program cache or inspect the converted structure. */
struct imba_bvh8 { Bvh8 b; };

extern "C" imba_bvh8* imba_bvh8_build(int accel_layout, int root, const void* nodes, const imba_vec4* tris,
                                      const imba_instance_node* inst_nodes) {
    imba_bvh8* h = new imba_bvh8; std::string err;
    if (!imba_build_bvh8(accel_layout, root, nodes, tris, inst_nodes, h->b, err)) {
        imba_set_error("BVH conversion failed: %s", err.c_str()); delete h; return nullptr;
    }
    return h;
}
extern "C" void imba_bvh8_free(imba_bvh8* h) { delete h; }
extern "C" int imba_bvh8_arrays(const imba_bvh8* h, const void** nodes, int64_t* node_count, const void** tris,
                                int64_t* tri_count, const void** insts, int64_t* inst_count, uint32_t* top_root,
                                int32_t* max_depth) {
    if (!h) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    if (nodes) *nodes = h->b.nodes.data();
    if (node_count) *node_count = (int64_t)h->b.nodes.size();
    if (tris) *tris = h->b.tris.data();
    if (tri_count) *tri_count = (int64_t)h->b.tris.size();
    if (insts) *insts = h->b.insts.data();
    if (inst_count) *inst_count = (int64_t)h->b.insts.size();
    if (top_root) *top_root = h->b.top_root;
    if (max_depth) *max_depth = h->b.max_depth;
    return IMBA_OK;
}
