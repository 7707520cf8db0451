/* integrator.cu -- host side of the device-resident integrators behind Integrator::{render,reset,preprocess,
 * pixel_size} (render/integrators/integrator.h:19-39) and RayScheduler::run_iteration (ray_scheduler.h:29-32).
 *
 * The reference runs one primary + one shadow queue per CPU thread over 256x256 tiles (tile_scheduler.h:100-144).  Here
 * one iteration is ONE wavefront over all width*height*spp samples, every stage a kernel on one stream, queue sizes in
 * device memory: nothing synchronises with the host until the caller downloads the framebuffer.
 *   per wavefront:  closest-hit trace -> sort hits by material -> shade (emit / NEE / bounce) ->
 *                   any-hit trace fused with the shadow resolve -> stable compaction into the ping-pong queue
 */
#include "integrator.cuh"
#include <cstring>

static DevCamera make_camera(const imba_render_settings& s) {
    // CameraControl::setup (frontend/main.cpp:64-71) followed by PerspectiveCamera::move (camera.h:27-50), in double
    auto norm = [](double* v) { const double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); v[0] /= l; v[1] /= l; v[2] /= l; };
    auto cross3 = [](const double* a, const double* b, double* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
    double dir[3] = { s.cam_dir[0], s.cam_dir[1], s.cam_dir[2] }, up[3] = { s.cam_up[0], s.cam_up[1], s.cam_up[2] }, right[3], down[3];
    norm(dir); norm(up);
    cross3(dir, up, right); norm(right);
    cross3(right, dir, up); norm(up);       // orthogonalised up (main.cpp:69)
    cross3(dir, up, right); norm(right);    // camera.h:29
    cross3(dir, right, down);               // camera.h:30: "up" of the camera matrix points down the image
    DevCamera c;
    c.px = s.cam_pos[0]; c.py = s.cam_pos[1]; c.pz = s.cam_pos[2];
    c.dx = (float)dir[0]; c.dy = (float)dir[1]; c.dz = (float)dir[2];
    c.rx = (float)right[0]; c.ry = (float)right[1]; c.rz = (float)right[2];
    c.ux = (float)down[0]; c.uy = (float)down[1]; c.uz = (float)down[2];
    c.tan_half = (float)tan((double)s.fov * 3.14159265358979323846 / 360.0);
    c.width = s.width; c.height = s.height;
    c.img_plane_dist = (float)s.width / (2.0f * c.tan_half);
    return c;
}

uint32_t imba_iteration_seed(uint64_t seed, uint64_t global_iteration, uint32_t stream_id) {
    // splitmix64 of (seed, iteration, stream): replaces std::random_device (ray_gen.h:49-50) so runs are reproducible
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (global_iteration * 4 + stream_id + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (uint32_t)z;
}

extern "C" void imba_integrator_destroy(imba_integrator* it) {
    if (!it) return;
    imba_vcm_free(it);   // also points the queue members back at the buffers allocated at creation
    for (int i = 0; i < 2; ++i) { cudaFree(it->rays[i]); cudaFree(it->states[i]); cudaFree(it->count[i]); }
    cudaFree(it->hits); cudaFree(it->shadow_rays); cudaFree(it->shadow_states);
    imba_partition_free(it->part);
    if (it->owns_fb) cudaFree(it->fb);
    cudaFree(it->d_stats); cudaFree(it->d_probe);
    if (it->shadow_stream) cudaStreamDestroy(it->shadow_stream);
    if (it->ev_shaded) cudaEventDestroy(it->ev_shaded);
    if (it->ev_shadow_done) cudaEventDestroy(it->ev_shadow_done);
    if (it->stream) cudaStreamDestroy(it->stream);
    delete it;
}

extern "C" imba_integrator* imba_integrator_create(imba_scene* scene, const imba_render_settings* s) {
    if (!scene || !s || !imba_shading_view(scene)) { imba_set_error("integrator needs a scene created with geometry, materials and lights"); return nullptr; }
    if (s->width <= 0 || s->height <= 0 || s->spp <= 0 || s->max_path_len <= 0) { imba_set_error("invalid render settings"); return nullptr; }
    if (s->algorithm < IMBA_ALGO_PT || s->algorithm > IMBA_ALGO_LT) { imba_set_error("unknown algorithm %d", s->algorithm); return nullptr; }
    const bool vcm = s->algorithm != IMBA_ALGO_PT;
    if (vcm && (s->light_path_count <= 0 || s->num_connections < 1 || s->num_knn < 1 || s->num_knn > 32)) { imba_set_error("invalid VCM settings (light paths > 0, connections >= 1, 1 <= knn <= 32)"); return nullptr; }
    if (imba_shading_view(scene)->light_count <= 0) { imba_set_error("scene has no lights"); return nullptr; }
    if (imba_shading_view(scene)->material_count > 1023) { imba_set_error("more than 1023 materials"); return nullptr; }
    imba_integrator* it = new imba_integrator;
    it->scene = scene; it->cfg = *s; it->cam = make_camera(*s);
    if (it->cfg.rank_count <= 0) { it->cfg.rank_count = 1; it->cfg.rank = 0; }
    it->cfg.iterations_per_pass = (vcm && s->algorithm != IMBA_ALGO_LT) ? 1 : std::min(std::max(it->cfg.iterations_per_pass, 1), 16);
    it->pixel_capacity = s->width * s->height * s->spp;
    it->state_bytes = vcm ? (int)sizeof(VCMState) : (int)sizeof(PTState);
    it->shadow_slots = vcm ? s->num_connections + 1 : 1;
    const long long cap = std::max((long long)s->width * s->height * s->spp * (vcm ? 1 : it->cfg.iterations_per_pass),
                                   vcm ? (long long)s->light_path_count * it->cfg.iterations_per_pass : 0ll);
    if (cap * it->shadow_slots > (1ll << 30)) { imba_set_error("queue capacity too large"); delete it; return nullptr; }
    it->capacity = (int)cap;
    bool ok = cudaStreamCreateWithFlags(&it->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&it->shadow_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&it->ev_shaded, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&it->ev_shadow_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventRecord(it->ev_shadow_done, it->shadow_stream) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
        ok = ok && cudaMalloc((void**)&it->rays[i], (size_t)cap * 32) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&it->states[i], (size_t)cap * it->state_bytes) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&it->count[i], 16) == cudaSuccess;
    }
    ok = ok && cudaMalloc((void**)&it->hits, (size_t)cap * 16) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&it->shadow_rays, (size_t)cap * it->shadow_slots * 32) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&it->shadow_states, (size_t)cap * it->shadow_slots * 16) == cudaSuccess;
    ok = ok && imba_partition_alloc(it->part, it->capacity, imba_shading_view(scene)->material_count + 1) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&it->fb, (size_t)s->width * s->height * 3 * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&it->d_stats, 16) == cudaSuccess && cudaMalloc((void**)&it->d_probe, 16) == cudaSuccess;
    ok = ok && cudaMemset(it->fb, 0, (size_t)s->width * s->height * 3 * sizeof(float)) == cudaSuccess && cudaMemset(it->d_stats, 0, 16) == cudaSuccess;
    if (!ok) { imba_set_error("integrator allocation failed: %s", cudaGetErrorString(cudaGetLastError())); imba_integrator_destroy(it); return nullptr; }
    return it;
}

extern "C" int imba_integrator_set_framebuffer(imba_integrator* it, float* d_rgb) {
    if (!it || !d_rgb) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    if (d_rgb == it->fb) return IMBA_OK;
    // the tail graph of the VCM light pass captured launches that write to the current framebuffer: park it under that
    // pointer and pick up the one captured for the new framebuffer, if any (callers alternate between two buffers to overlap
    // the multi-GPU sum with the next iteration)
    if (it->tail_graph) {
        int slot = -1;
        for (int i = 0; i < imba_integrator::kParkedGraphs && slot < 0; ++i) if (!it->parked_graph[i]) slot = i;
        if (slot < 0) { cudaGraphExecDestroy(it->parked_graph[0]); slot = 0; }
        it->parked_fb[slot] = it->fb; it->parked_graph[slot] = it->tail_graph; it->tail_graph = nullptr;
    }
    for (int i = 0; i < imba_integrator::kParkedGraphs; ++i)
        if (it->parked_graph[i] && it->parked_fb[i] == d_rgb) { it->tail_graph = it->parked_graph[i]; it->parked_graph[i] = nullptr; it->parked_fb[i] = nullptr; }
    if (it->owns_fb) cudaFree(it->fb);
    it->fb = d_rgb; it->owns_fb = false;
    return IMBA_OK;
}

extern "C" void* imba_integrator_stream(imba_integrator* it) { return it ? (void*)it->stream : nullptr; }

/* The one collective of the multi-GPU path, for host programs that are not Python: sums the integrator's W x H x 3
 * float framebuffer over all ranks of `nccl_comm` in place (ncclAllReduce, ncclFloat32, ncclSum).  The library does not link
 * NCCL: the host program created the communicator, so its NCCL is in the process already and is looked up at the first call
 * (dlopen RTLD_NOLOAD, then by name).  stream = null: the integrator's own stream, i.e. ordered after the rendering. */
#include <dlfcn.h>
typedef int (*NcclAllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
static NcclAllReduceFn nccl_allreduce_fn() {
    static NcclAllReduceFn fn = nullptr;
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    if (fn) return fn;
    fn = (NcclAllReduceFn)dlsym(RTLD_DEFAULT, "ncclAllReduce");           // linked into the host program or loaded RTLD_GLOBAL
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (int pass = 0; pass < 2 && !fn; ++pass)
        for (const char* n : names) {
            void* h = dlopen(n, RTLD_NOW | (pass == 0 ? RTLD_NOLOAD : RTLD_GLOBAL));
            if (h && (fn = (NcclAllReduceFn)dlsym(h, "ncclAllReduce"))) break;
        }
    return fn;
}

extern "C" int imba_integrator_allreduce(imba_integrator* it, void* nccl_comm, void* stream) {
    if (!it || !nccl_comm) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    NcclAllReduceFn fn = nccl_allreduce_fn();
    if (!fn) { imba_set_error("ncclAllReduce not found: no NCCL library in this process (%s)", dlerror() ? dlerror() : "libnccl.so.2"); return IMBA_ERR_INVALID; }
    const size_t count = (size_t)it->cfg.width * it->cfg.height * 3;
    const int rc = fn(it->fb, it->fb, count, 7 /* ncclFloat32 */, 0 /* ncclSum */, nccl_comm, stream ? (cudaStream_t)stream : it->stream);
    if (rc != 0) { imba_set_error("ncclAllReduce failed with ncclResult_t %d", rc); return IMBA_ERR_CUDA; }
    return IMBA_OK;
}

extern "C" int imba_integrator_preprocess(imba_integrator* it) {
    if (!it) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    const int quads = ((it->cam.width + 7) / 8) * ((it->cam.height + 7) / 8);
    const int n = std::min(quads * 4, it->capacity / 4 * 4); // capacity >= width * height >= 4 * quads
    if (n <= 0) {   // fewer than four samples in the whole image: nothing to probe with (integrator.cpp:68-70 leaves the size at 1)
        it->pixel_size = 1.0f;
        return it->cfg.algorithm != IMBA_ALGO_PT ? imba_vcm_preprocess(it) : IMBA_OK;
    }
    IMBA_CUDA(cudaMemsetAsync(it->d_probe, 0, 16, it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(imba_probe_raygen(it->cam, it->rays[0], n, it->stream), IMBA_ERR_CUDA);
    int rc = imba_scene_trace(it->scene, (const imba_ray*)it->rays[0], (imba_hit*)it->hits, n, 0, nullptr, it->stream);
    if (rc != IMBA_OK) return rc;
    IMBA_CUDA(imba_probe_reduce(it->rays[0], it->hits, n / 4, it->d_probe, (int*)(it->d_probe + 1), it->stream), IMBA_ERR_CUDA);
    g_imba_launches.fetch_add(2);
    float h[2]; IMBA_CUDA(cudaMemcpyAsync(h, it->d_probe, 8, cudaMemcpyDeviceToHost, it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamSynchronize(it->stream), IMBA_ERR_CUDA);
    int cnt; memcpy(&cnt, &h[1], 4);
    it->pixel_size = cnt == 0 ? 1.0f : h[0] / (float)cnt;
    if (it->cfg.algorithm != IMBA_ALGO_PT) return imba_vcm_preprocess(it);
    return IMBA_OK;
}

extern "C" float imba_integrator_pixel_size(const imba_integrator* it) { return it ? it->pixel_size : 0.0f; }
extern "C" int imba_integrator_vertex_cache_size(const imba_integrator* it) { return it ? it->cache_capacity : 0; }
extern "C" int imba_integrator_reset(imba_integrator* it) { if (!it) return IMBA_ERR_INVALID; it->iteration = 0; return IMBA_OK; }

/* Camera moved (CameraControl -> PerspectiveCamera::move, camera.h:27-50; the reference then calls Integrator::reset,
 * render_window.cpp:97-101): new position / direction / up, restart of the progressive schedule.  Waits for pending work;
 * captured launch graphs hold the old camera by value and are dropped. */
extern "C" int imba_integrator_set_camera(imba_integrator* it, const float pos[3], const float dir[3], const float up[3]) {
    if (!it || !pos || !dir || !up) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    int rc = imba_integrator_sync(it);
    if (rc != IMBA_OK) return rc;
    for (int k = 0; k < 3; ++k) { it->cfg.cam_pos[k] = pos[k]; it->cfg.cam_dir[k] = dir[k]; it->cfg.cam_up[k] = up[k]; }
    it->cam = make_camera(it->cfg);
    if (it->tail_graph) { cudaGraphExecDestroy(it->tail_graph); it->tail_graph = nullptr; }
    for (int i = 0; i < imba_integrator::kParkedGraphs; ++i)
        if (it->parked_graph[i]) { cudaGraphExecDestroy(it->parked_graph[i]); it->parked_graph[i] = nullptr; it->parked_fb[i] = nullptr; }
    it->iteration = 0;
    return IMBA_OK;
}

static int pt_wavefront(imba_integrator* it, int cur, int capacity) {
    const DeviceShading& S = *imba_shading_view(it->scene);
    TraceRequest rq; rq.d_rays = (const imba_ray*)it->rays[cur]; rq.d_hits = (imba_hit*)it->hits; rq.capacity = capacity; rq.d_count = it->count[cur]; rq.mode = 0;
    int rc = imba_scene_trace_ex(it->scene, rq, it->stream);
    if (rc != IMBA_OK) return rc;
    // IMBA_PT_SORT=1 restores the reference's order of work (compact_hits + sort_by_material before shading); the default shades
    // in queue order (see pt_shade_kernel): same per-path results, 11 % less time per iteration on stilllife-proxy
    const bool sort = getenv("IMBA_PT_SORT") != nullptr;
    const int* sorted = sort ? it->part.indices : nullptr;
    // (the sort pass resets the pixel ids of the shadow slots, which the previous wave's shadow + resolve traversal on the other
    //  stream still reads: it has to wait for that traversal here already, not only in front of the shading kernel)
    if (sort) IMBA_CUDA(cudaStreamWaitEvent(it->stream, it->ev_shadow_done, 0), IMBA_ERR_CUDA);
    if (sort) IMBA_CUDA(imba_sort_hits_by_material(it->part, it->hits, S.indices, S.material_count, (int*)it->shadow_states, 1, it->count[cur], 0, capacity, it->stream, false), IMBA_ERR_CUDA);
    IMBA_CUDA(imba_pt_miss(S, it->rays[cur], it->hits, (const PTState*)it->states[cur], sorted, it->part.bin_start + S.material_count, it->count[cur], capacity, it->fb, it->stream), IMBA_ERR_CUDA);
    // the previous wave's shadow traversal (other stream) reads the shadow buffers and that wave's queue count, which is this
    // wave's output count: both are rewritten below
    IMBA_CUDA(cudaStreamWaitEvent(it->stream, it->ev_shadow_done, 0), IMBA_ERR_CUDA);
    // survivors are pushed straight into the other queue by the shading kernel (fused compaction, replaces compact_rays)
    IMBA_CUDA(cudaMemsetAsync(it->count[cur ^ 1], 0, sizeof(int), it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(imba_pt_shade(S, it->rays[cur], it->hits, (PTState*)it->states[cur], sorted, it->part.bin_start + S.material_count, it->count[cur], capacity,
                            it->shadow_rays, it->shadow_states, it->fb, it->cfg.max_path_len, it->d_stats, it->rays[cur ^ 1], (PTState*)it->states[cur ^ 1],
                            it->count[cur ^ 1], it->stream), IMBA_ERR_CUDA);
    // shadow rays + resolve on the second stream: the next wave's closest-hit traversal does not depend on them, and the two
    // persistent kernels fill each other's tails (small late wavefronts leave most SMs idle otherwise)
    IMBA_CUDA(cudaEventRecord(it->ev_shaded, it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamWaitEvent(it->shadow_stream, it->ev_shaded, 0), IMBA_ERR_CUDA);
    TraceRequest sq; sq.d_rays = (const imba_ray*)it->shadow_rays; sq.capacity = capacity; sq.d_count = it->count[cur]; sq.mode = 2;
    sq.d_shadow_states = it->shadow_states; sq.d_framebuffer = it->fb;
    rc = imba_scene_trace_ex(it->scene, sq, it->shadow_stream);
    if (rc != IMBA_OK) return rc;
    IMBA_CUDA(cudaEventRecord(it->ev_shadow_done, it->shadow_stream), IMBA_ERR_CUDA);
    g_imba_launches.fetch_add((sort ? 3 : 0) + 1 + 2 + (S.env_tex >= 0 ? 1 : 0));
    return IMBA_OK;
}

extern "C" int imba_integrator_render(imba_integrator* it, int iterations) {
    if (!it || iterations < 0) { imba_set_error("invalid argument"); return IMBA_ERR_INVALID; }
    for (int k = 0; k < iterations;) {
        if (it->cfg.algorithm != IMBA_ALGO_PT) return imba_vcm_render(it, iterations);
        // a group of consecutive iterations as one pass (iterations_per_pass); each keeps its own seed base
        const int group = std::min(it->cfg.iterations_per_pass, iterations - k);
        PtSeeds seeds;
        for (int g = 0; g < group; ++g) {
            const uint64_t global_iter = (it->iteration + g) * (uint64_t)it->cfg.rank_count + (uint64_t)it->cfg.rank;
            seeds.base[g] = imba_iteration_seed(it->cfg.seed, global_iter, 0);
        }
        const int rays_in_pass = it->pixel_capacity * group;
        IMBA_CUDA(imba_pt_raygen(it->cam, it->cfg.spp, seeds, it->pixel_capacity, 0, rays_in_pass, it->rays[0], (PTState*)it->states[0], it->count[0], it->stream), IMBA_ERR_CUDA);
        g_imba_launches.fetch_add(1);
        int cur = 0;
        for (int wave = 0; wave < it->cfg.max_path_len; ++wave, cur ^= 1) {
            const int rc = pt_wavefront(it, cur, rays_in_pass);
            if (rc != IMBA_OK) return rc;
        }
        k += group; it->iteration += group;
        // whoever consumes the framebuffer next does so on it->stream: order the last shadow traversal before it
        IMBA_CUDA(cudaStreamWaitEvent(it->stream, it->ev_shadow_done, 0), IMBA_ERR_CUDA);
    }
    return IMBA_OK;
}

extern "C" float* imba_integrator_framebuffer(imba_integrator* it) { return it ? it->fb : nullptr; }

extern "C" int imba_integrator_sync(imba_integrator* it) {
    if (!it) return IMBA_ERR_INVALID;
    IMBA_CUDA(cudaStreamSynchronize(it->stream), IMBA_ERR_CUDA);
    return IMBA_OK;
}

extern "C" int imba_integrator_download(imba_integrator* it, float* h_rgb) {
    if (!it || !h_rgb) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    IMBA_CUDA(cudaMemcpyAsync(h_rgb, it->fb, (size_t)it->cfg.width * it->cfg.height * 3 * sizeof(float), cudaMemcpyDeviceToHost, it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamSynchronize(it->stream), IMBA_ERR_CUDA);
    return IMBA_OK;
}

extern "C" int imba_integrator_clear(imba_integrator* it) {
    if (!it) return IMBA_ERR_INVALID;
    IMBA_CUDA(cudaMemsetAsync(it->fb, 0, (size_t)it->cfg.width * it->cfg.height * 3 * sizeof(float), it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemsetAsync(it->d_stats, 0, 16, it->stream), IMBA_ERR_CUDA);
    return IMBA_OK;
}

extern "C" int imba_integrator_ray_counts(imba_integrator* it, uint64_t out[2]) {
    if (!it || !out) return IMBA_ERR_INVALID;
    unsigned long long h[2];
    IMBA_CUDA(cudaMemcpyAsync(h, it->d_stats, 16, cudaMemcpyDeviceToHost, it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamSynchronize(it->stream), IMBA_ERR_CUDA);
    out[0] = h[0]; out[1] = h[1];
    return IMBA_OK;
}

/* Camera rays of the integrator's camera for given raster positions (parity harness for camera.h:57-65). */
extern "C" int imba_integrator_camera_rays(imba_integrator* it, int count, const float* raster_xy, imba_ray* h_rays) {
    if (!it || !raster_xy || !h_rays) return IMBA_ERR_INVALID;
    for (int i = 0; i < count; ++i) {
        const V3 d = it->cam.ray_dir(raster_xy[2 * i], raster_xy[2 * i + 1]);
        h_rays[i].org = imba_vec4{ it->cam.px, it->cam.py, it->cam.pz, 0.0f };
        h_rays[i].dir = imba_vec4{ d.x, d.y, d.z, FLT_MAX };
    }
    return IMBA_OK;
}

/* RNG states of primary rays first .. first+count-1 for a given seed_base (parity harness for ray_gen.h:57-72). */
extern "C" int imba_seed_states(uint32_t seed_base, int first, int count, int spp, uint64_t* out) {
    if (!out || spp <= 0) return IMBA_ERR_INVALID;
    for (int i = first; i < first + count; ++i) out[i - first] = seed_primary(seed_base, i, i / spp);
    return IMBA_OK;
}

/* PT shading stage on an injected batch: sort by material, shade, compact -- the same kernels render() enqueues.
 * Outputs: continuation queue (stable order), shadow queue (slot order, empty slots removed); contributions are added
 * to the integrator's accumulation buffer. */
extern "C" int imba_pt_shade_batch(imba_integrator* it, int count, const imba_ray* h_rays, const imba_hit* h_hits, const void* h_states,
                                   imba_ray* h_rays_out, void* h_states_out, int* out_count, imba_ray* h_shadow_rays, void* h_shadow_states,
                                   int* shadow_count) {
    if (!it || count < 0 || count > it->capacity || it->cfg.algorithm != IMBA_ALGO_PT) { imba_set_error("needs a PT integrator whose queue capacity covers the batch"); return IMBA_ERR_INVALID; }
    const DeviceShading& S = *imba_shading_view(it->scene);
    cudaStream_t st = it->stream;
    IMBA_CUDA(cudaMemcpyAsync(it->rays[0], h_rays, (size_t)count * 32, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(it->hits, h_hits, (size_t)count * 16, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(it->states[0], h_states, (size_t)count * sizeof(PTState), cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(it->count[0], &count, 4, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(imba_sort_hits_by_material(it->part, it->hits, S.indices, S.material_count, (int*)it->shadow_states, 1, it->count[0], 0, count, st), IMBA_ERR_CUDA);
    IMBA_CUDA(imba_pt_miss(S, it->rays[0], it->hits, (const PTState*)it->states[0], it->part.indices, it->part.bin_start + S.material_count, it->count[0], count, it->fb, st), IMBA_ERR_CUDA);
    IMBA_CUDA(imba_pt_shade(S, it->rays[0], it->hits, (PTState*)it->states[0], it->part.indices, it->part.bin_start + S.material_count, it->count[0], count,
                            it->shadow_rays, it->shadow_states, it->fb, it->cfg.max_path_len, it->d_stats, nullptr, nullptr, nullptr, st), IMBA_ERR_CUDA);
    // shadow queue: compact the valid slots (pixel_id >= 0) into rays[1] / states[1]
    IMBA_CUDA(imba_compact_alive(it->part, nullptr, it->shadow_rays, it->shadow_states, sizeof(ShadowState), it->count[0], 0, count,
                                 it->rays[1], it->states[1], it->count[1], st), IMBA_ERR_CUDA);
    int nshadow = 0;
    IMBA_CUDA(cudaMemcpyAsync(&nshadow, it->count[1], 4, cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
    if (count == 0) nshadow = 0;
    IMBA_CUDA(cudaMemcpyAsync(h_shadow_rays, it->rays[1], (size_t)nshadow * 32, cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(h_shadow_states, it->states[1], (size_t)nshadow * sizeof(ShadowState), cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
    // continuation queue
    IMBA_CUDA(imba_compact_alive(it->part, it->hits, it->rays[0], it->states[0], sizeof(PTState), it->count[0], 0, count,
                                 it->rays[1], it->states[1], it->count[1], st), IMBA_ERR_CUDA);
    int ncont = 0;
    IMBA_CUDA(cudaMemcpyAsync(&ncont, it->count[1], 4, cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
    if (count == 0) ncont = 0;
    IMBA_CUDA(cudaMemcpyAsync(h_rays_out, it->rays[1], (size_t)ncont * 32, cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(h_states_out, it->states[1], (size_t)ncont * sizeof(PTState), cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
    *out_count = ncont; *shadow_count = nshadow;
    g_imba_launches.fetch_add(3 + 1 + 4 + 4);
    return IMBA_OK;
}

/* ---- standalone queue primitives (ray_queue.h:161-244) on caller-owned device arrays -------------------------------- */
extern "C" int imba_queue_sort_by_material(imba_scene* scene, const imba_hit* d_hits, int count, int* d_sorted_indices, int* d_hit_count, void* stream) {
    if (!scene || !imba_shading_view(scene) || count < 0) { imba_set_error("invalid argument"); return IMBA_ERR_INVALID; }
    const DeviceShading& S = *imba_shading_view(scene);
    cudaStream_t st = (cudaStream_t)stream;
    if (count == 0) { IMBA_CUDA(cudaMemsetAsync(d_hit_count, 0, 4, st), IMBA_ERR_CUDA); IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA); return IMBA_OK; }
    PartitionBuffers b;
    IMBA_CUDA(imba_partition_alloc(b, count, S.material_count + 1), IMBA_ERR_CUDA);
    cudaError_t e = imba_sort_hits_by_material(b, (const int4*)d_hits, S.indices, S.material_count, nullptr, 1, nullptr, count, count, st);
    if (e == cudaSuccess && count > 0) e = cudaMemcpyAsync(d_sorted_indices, b.indices, sizeof(int) * (size_t)count, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_hit_count, b.bin_start + S.material_count, sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    imba_partition_free(b);
    g_imba_launches.fetch_add(3);
    IMBA_CUDA(e, IMBA_ERR_CUDA);
    return IMBA_OK;
}

extern "C" int imba_queue_compact_rays(const imba_ray* d_rays, const void* d_states, int state_bytes, imba_ray* d_rays_out, void* d_states_out,
                                       int count, int* d_out_count, void* stream) {
    if (count < 0 || state_bytes < 8 || state_bytes % 8) { imba_set_error("invalid argument (state size must be a multiple of 8)"); return IMBA_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    if (count == 0) { IMBA_CUDA(cudaMemsetAsync(d_out_count, 0, 4, st), IMBA_ERR_CUDA); IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA); return IMBA_OK; }
    PartitionBuffers b;
    IMBA_CUDA(imba_partition_alloc(b, count, 2), IMBA_ERR_CUDA);
    cudaError_t e = imba_compact_alive(b, nullptr, (const float4*)d_rays, d_states, state_bytes, nullptr, count, count, (float4*)d_rays_out, d_states_out, d_out_count, st);
    if (e == cudaSuccess && count == 0) e = cudaMemsetAsync(d_out_count, 0, 4, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    imba_partition_free(b);
    g_imba_launches.fetch_add(4);
    IMBA_CUDA(e, IMBA_ERR_CUDA);
    return IMBA_OK;
}

/* compact_hits (ray_queue.h:161-185): stable partition, hits first (material-sorted order is given by
 * imba_queue_sort_by_material; this entry point physically permutes rays / hits / states like the reference does). */
extern "C" int imba_queue_compact_hits(imba_scene* scene, const imba_ray* d_rays, const imba_hit* d_hits, const void* d_states, int state_bytes,
                                       imba_ray* d_rays_out, imba_hit* d_hits_out, void* d_states_out, int count, int* d_hit_count, void* stream) {
    if (!scene || !imba_shading_view(scene) || count < 0 || state_bytes < 8 || state_bytes % 8) { imba_set_error("invalid argument"); return IMBA_ERR_INVALID; }
    const DeviceShading& S = *imba_shading_view(scene);
    cudaStream_t st = (cudaStream_t)stream;
    if (count == 0) { IMBA_CUDA(cudaMemsetAsync(d_hit_count, 0, 4, st), IMBA_ERR_CUDA); IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA); return IMBA_OK; }
    PartitionBuffers b;
    IMBA_CUDA(imba_partition_alloc(b, count, 2), IMBA_ERR_CUDA);
    // two bins: hit / miss  (material count 1 -> every hit falls into bin 0)
    cudaError_t e = imba_sort_hits_by_material(b, (const int4*)d_hits, S.indices, 1, nullptr, 1, nullptr, count, count, st);
    if (e == cudaSuccess) e = imba_gather_queue(b.indices, nullptr, count, count, (const float4*)d_rays, (const int4*)d_hits, d_states, state_bytes,
                                                (float4*)d_rays_out, (int4*)d_hits_out, d_states_out, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_hit_count, b.bin_start + 1, sizeof(int), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    imba_partition_free(b);
    g_imba_launches.fetch_add(4);
    IMBA_CUDA(e, IMBA_ERR_CUDA);
    return IMBA_OK;
}
