/* vcm_host.cu -- host driver of the VCM family on the device: VCMIntegrator::{preprocess, render, trace_light_paths,
 * trace_camera_paths} (vcm.h:74-81, vcm.cpp:26-130) and LightVertices::{compute_cache_size, build, clear}
 * (light_vertices.cpp:15-129, light_vertices.h:70-116).
 *
 * Light pass: ONE wavefront over all light paths (the reference tiles them per light, tile_gen.h:102-177); light paths
 * have no length cap (only Russian roulette ends them, vcm.cpp:133-139), so the host reads the queue size back every 2-16
 * wavefronts and hands the last <= 128 Ki paths to the per-path tail kernels (vcm.cu).  The camera pass is capped by
 * max_path_len and is enqueued without host sync.  imba_vcm_render pipelines the iterations of one call: camera pass of
 * iteration k beside the light-pass tail of iteration k+1 (BPT / VCM / PPM), light paths of several iterations as one pass (LT).
 */
#include <algorithm>
#include <vector>
#include "vcm.cuh"
#include <cstdlib>

static VcmParams iteration_params(const imba_integrator* it, uint64_t iteration_1based) {
    // vcm.cpp:28-46
    const float radius_alpha = 0.75f;
    VcmParams P;
    P.base_radius = it->base_radius;
    P.pm_radius = it->base_radius / powf((float)iteration_1based, 0.5f * (1.0f - radius_alpha));
    P.pm_radius = std::max(P.pm_radius, 1e-7f);
    const float eta_vcm = IMBA_PI * sqr(P.pm_radius) * it->cfg.light_path_count;
    P.mis_eta_vc = 1.0f / eta_vcm;
    P.mis_eta_vm = it->cfg.algorithm == IMBA_ALGO_BPT ? 0.0f : eta_vcm;
    P.light_path_count = it->cfg.light_path_count; P.num_connections = it->cfg.num_connections; P.num_knn = it->cfg.num_knn;
    P.max_path_len = it->cfg.max_path_len;
    return P;
}

static VcmLaunch base_launch(imba_integrator* it, int cur, int capacity, bool sorted = true) {
    const DeviceShading* S = imba_shading_view(it->scene);
    VcmLaunch a;
    a.S = S; a.P = it->d_params; a.cam = it->cam;
    a.rays = it->rays[cur]; a.hits = it->hits; a.states = (VCMState*)it->states[cur]; a.sorted = sorted ? it->part.indices : nullptr;
    a.hit_count_ptr = it->part.bin_start + S->material_count; a.queue_count_ptr = it->count[cur]; a.capacity = capacity;
    a.shadow_rays = it->shadow_rays; a.shadow_states = it->shadow_states;
    a.cache = it->cache; a.cache_capacity = it->cache_capacity; a.cache_last = it->cache_last;
    a.grid = it->grid; a.cells = it->cells; a.table = it->table; a.photons = it->photons;
    a.photon_pos = it->photon_pos; a.merge_found = it->merge_found; a.merge_stride = it->capacity; a.knn = it->cfg.num_knn;
    a.fb = it->fb; a.ray_stats = it->d_stats;
    a.out_rays = it->rays[cur ^ 1]; a.out_states = (VCMState*)it->states[cur ^ 1]; a.out_count = it->count[cur ^ 1];
    return a;
}

/* One wavefront of the light pass (light != 0) or of the camera pass. */
static int vcm_wavefront(imba_integrator* it, int algo, bool light, int cur, int capacity, bool overlap_shadows = false) {
    const DeviceShading& S = *imba_shading_view(it->scene);
    const int slots = light ? 1 : it->shadow_slots;
    TraceRequest rq; rq.d_rays = (const imba_ray*)it->rays[cur]; rq.d_hits = (imba_hit*)it->hits; rq.capacity = capacity; rq.d_count = it->count[cur]; rq.mode = 0;
    if (it->tail_counter_next >= 0) rq.d_work_counter = it->tail_counters + it->tail_counter_next++;
    int rc = imba_scene_trace_ex(it->scene, rq, it->stream);
    if (rc != IMBA_OK) return rc;
    // IMBA_VCM_SORT=1: the reference's order of work (compact_hits + sort_by_material, vcm.cpp:197-210, 327-340); default: queue order
    const bool sort = getenv("IMBA_VCM_SORT") != nullptr;
    // the previous wave's shadow traversal on the second stream still reads the shadow buffers and its queue count; the sort pass
    // already rewrites the pixel ids of the shadow slots, so the wait comes in front of it
    if (overlap_shadows) IMBA_CUDA(cudaStreamWaitEvent(it->stream, it->ev_shadow_done, 0), IMBA_ERR_CUDA);
    if (sort) IMBA_CUDA(imba_sort_hits_by_material(it->part, it->hits, S.indices, S.material_count, (int*)it->shadow_states, slots, it->count[cur], 0, capacity, it->stream, false), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemsetAsync(it->count[cur ^ 1], 0, sizeof(int), it->stream), IMBA_ERR_CUDA);
    const VcmLaunch a = base_launch(it, cur, capacity, sort);
    IMBA_CUDA(light ? imba_vcm_light_shade(algo, a, it->stream) : imba_vcm_camera_shade(algo, a, it->stream), IMBA_ERR_CUDA);
    const bool shadows = light ? algo != IMBA_ALGO_PPM : algo != IMBA_ALGO_PPM; // PPM emits no shadow rays in either pass
    if (shadows) {
        TraceRequest sq; sq.d_rays = (const imba_ray*)it->shadow_rays; sq.capacity = capacity * slots; sq.d_count = it->count[cur]; sq.count_mul = slots; sq.mode = 2;
        sq.d_shadow_states = it->shadow_states; sq.d_framebuffer = it->fb;
        if (it->tail_counter_next >= 0) sq.d_work_counter = it->tail_counters + it->tail_counter_next++;
        if (overlap_shadows) { // as in the PT integrator: beside the next wave's closest-hit traversal
            IMBA_CUDA(cudaEventRecord(it->ev_shaded, it->stream), IMBA_ERR_CUDA);
            IMBA_CUDA(cudaStreamWaitEvent(it->shadow_stream, it->ev_shaded, 0), IMBA_ERR_CUDA);
        }
        rc = imba_scene_trace_ex(it->scene, sq, overlap_shadows ? it->shadow_stream : it->stream);
        if (rc != IMBA_OK) return rc;
        if (overlap_shadows) IMBA_CUDA(cudaEventRecord(it->ev_shadow_done, it->shadow_stream), IMBA_ERR_CUDA);
    }
    // survivors were pushed into the other queue by the shading kernel (fused, unordered compaction)
    g_imba_launches.fetch_add(3 + 1 + 1 + (shadows ? 1 : 0));
    return IMBA_OK;
}

static void use_set(imba_integrator* it, int b) {
    const imba_integrator::VcmSet& v = it->vset[b];
    it->cache = v.cache; it->cache_last = v.cache_last; it->grid = v.grid; it->cells = v.cells; it->photons = v.photons; it->photon_pos = v.photon_pos; it->d_params = v.d_params;
}
static void use_queues(imba_integrator* it, const imba_integrator::QueueSet& q) {
    for (int i = 0; i < 2; ++i) { it->rays[i] = q.rays[i]; it->states[i] = q.states[i]; it->count[i] = q.count[i]; }
    it->hits = q.hits; it->shadow_rays = q.shadow_rays; it->shadow_states = q.shadow_states;
}

/* Parameters of the iteration into the current set's device block.  The host copy outlives the call (pageable memory: the
 * driver has staged it when cudaMemcpyAsync returns), so no stream synchronisation is needed. */
static int set_params(imba_integrator* it, const VcmParams& P, int b = 0) {
    it->h_params[b] = P;
    IMBA_CUDA(cudaMemcpyAsync(it->d_params, &it->h_params[b], sizeof(P), cudaMemcpyHostToDevice, it->stream), IMBA_ERR_CUDA);
    return IMBA_OK;
}

constexpr int TAIL_RAYS = 65536, TAIL_WAVES = 32;
constexpr int TAIL_MEGA_PATHS = 131072;  // below this many live light paths the per-path tail kernel takes over

/* Light paths have no length cap (vcm.cpp:133-139): in scenes with glass a few paths bounce hundreds of times.  Once the
 * queue is small the wavefront is launch-bound, so TAIL_WAVES wavefronts (sized for TAIL_RAYS rays, counts in device
 * memory) are captured once into a CUDA graph and replayed until the queue is empty. */
static int build_tail_graph(imba_integrator* it, int algo) {
    cudaGraph_t graph = nullptr;
    if (!it->tail_counters) IMBA_CUDA(cudaMalloc((void**)&it->tail_counters, sizeof(unsigned) * 2 * TAIL_WAVES), IMBA_ERR_CUDA);
    it->tail_counter_next = 0;
    IMBA_CUDA(cudaStreamBeginCapture(it->stream, cudaStreamCaptureModeThreadLocal), IMBA_ERR_CUDA);
    int rc = IMBA_OK, cur = 0;
    for (int k = 0; k < TAIL_WAVES && rc == IMBA_OK; ++k, cur ^= 1) rc = vcm_wavefront(it, algo, true, cur, std::min(TAIL_RAYS, it->capacity));
    const cudaError_t e = cudaStreamEndCapture(it->stream, &graph);
    it->tail_counter_next = -1;
    if (rc != IMBA_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    IMBA_CUDA(e, IMBA_ERR_CUDA);
    const cudaError_t e2 = cudaGraphInstantiate(&it->tail_graph, graph, 0);
    cudaGraphDestroy(graph);
    IMBA_CUDA(e2, IMBA_ERR_CUDA);
    return IMBA_OK;
}

/* Traces all light paths of one iteration; returns the number of wavefronts through *waves. */
static int light_pass(imba_integrator* it, int algo, const PtSeeds& seeds, int per_iteration, int path_count, int* waves, bool use_graph, cudaStream_t tail_stream = nullptr) {
    const DeviceShading& S = *imba_shading_view(it->scene);
    IMBA_CUDA(cudaMemsetAsync(it->cache_last, 0, sizeof(int), it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(imba_vcm_light_raygen(S, it->d_params, seeds, per_iteration, path_count, it->rays[0], (VCMState*)it->states[0], it->count[0], it->stream), IMBA_ERR_CUDA);
    g_imba_launches.fetch_add(1);
    // The queue size is read back only every `batch` wavefronts (the grid stays sized for the last known count; kernels on
    // an already empty queue return at once), so long light paths do not cost one host round trip per bounce.
    int cur = 0, n = path_count, w = 0, batch = 2;
    const int mega = getenv("IMBA_TAIL_PATHS") ? atoi(getenv("IMBA_TAIL_PATHS")) : TAIL_MEGA_PATHS;
    while (n > 0) {
        if (use_graph && n <= mega) { // few paths left: one thread per path follows it to its end (vcm.cu: vcm_light_tail_kernel)
            IMBA_CUDA(cudaStreamWaitEvent(it->stream, it->ev_shadow_done, 0), IMBA_ERR_CUDA); // last wavefront's shadow traversal reads the queue counts the tail resets
            TraceArgs T; imba_scene_trace_args(it->scene, &T);
            const VcmLaunch a = base_launch(it, cur, n);
            if (tail_stream) { // pipelined: the tail runs beside the previous iteration's camera pass (imba_vcm_render)
                IMBA_CUDA(cudaEventRecord(it->ev_light_done, it->stream), IMBA_ERR_CUDA);
                IMBA_CUDA(cudaStreamWaitEvent(tail_stream, it->ev_light_done, 0), IMBA_ERR_CUDA);
            }
            IMBA_CUDA(imba_vcm_light_tail(algo, it->scene->has_masks, a, T, n, tail_stream ? tail_stream : it->stream), IMBA_ERR_CUDA);
            if (tail_stream) IMBA_CUDA(cudaEventRecord(it->ev_tail_done, tail_stream), IMBA_ERR_CUDA);
            g_imba_launches.fetch_add(2); ++w;
            break;
        }
        const int tail_rays = std::min(TAIL_RAYS, it->capacity);
        // the captured wavefront graph holds the cache / parameter pointers of ONE buffer set: it is only used when the iterations are
        // not pipelined over two sets (IMBA_VCM_NO_PIPELINE, or integrators without a second set)
        const bool two_sets = it->vset[1].cache != nullptr && !getenv("IMBA_VCM_NO_PIPELINE") && !getenv("IMBA_VCM_SORT");
        if (use_graph && !two_sets && n <= tail_rays && cur == 0) {
            IMBA_CUDA(cudaStreamWaitEvent(it->stream, it->ev_shadow_done, 0), IMBA_ERR_CUDA);
            if (!it->tail_graph) { const int rc = build_tail_graph(it, algo); if (rc != IMBA_OK) return rc; }
            IMBA_CUDA(cudaGraphLaunch(it->tail_graph, it->stream), IMBA_ERR_CUDA);
            g_imba_launches.fetch_add(TAIL_WAVES * 6);
            w += TAIL_WAVES; // an even number of wavefronts: the live queue is queue 0 again
        } else {
            const int k_end = (use_graph && !two_sets && n <= tail_rays) ? 1 : batch; // make the parity even before switching to the graph
            for (int k = 0; k < k_end; ++k, cur ^= 1, ++w) {
                const int rc = vcm_wavefront(it, algo, true, cur, n, true);   // camera-connection shadow rays beside the next wave's traversal
                if (rc != IMBA_OK) return rc;
            }
            if (batch < 16) batch *= 2;
        }
        IMBA_CUDA(cudaMemcpyAsync(&n, it->count[cur], sizeof(int), cudaMemcpyDeviceToHost, it->stream), IMBA_ERR_CUDA);
        IMBA_CUDA(cudaStreamSynchronize(it->stream), IMBA_ERR_CUDA);
        if (w > 65536) { imba_set_error("light pass did not terminate"); return IMBA_ERR_OVERFLOW; }
    }
    IMBA_CUDA(cudaStreamWaitEvent(it->stream, it->ev_shadow_done, 0), IMBA_ERR_CUDA);
    if (waves) *waves = w;
    if (getenv("IMBA_VERBOSE")) fprintf(stderr, "[imba] light pass: %d paths, %d wavefronts\n", path_count, w);
    return IMBA_OK;
}

void imba_vcm_free(imba_integrator* it) {
    if (it->tail_graph) { cudaGraphExecDestroy(it->tail_graph); it->tail_graph = nullptr; }
    for (int i = 0; i < imba_integrator::kParkedGraphs; ++i) if (it->parked_graph[i]) { cudaGraphExecDestroy(it->parked_graph[i]); it->parked_graph[i] = nullptr; }
    cudaFree(it->tail_counters); it->tail_counters = nullptr;
    for (auto& v : it->vset) {
        cudaFree(v.cache); cudaFree(v.cache_last); cudaFree(v.grid); cudaFree(v.cells); cudaFree(v.photons); cudaFree(v.photon_pos); cudaFree(v.d_params);
        v = imba_integrator::VcmSet{};
    }
    cudaFree(it->grid_bbox); cudaFree(it->block_sums); cudaFree(it->merge_found); cudaFree(it->d_table_state);
    it->grid_bbox = nullptr; it->block_sums = nullptr; it->merge_found = nullptr; it->d_table_state = nullptr;
    it->cache = nullptr; it->cache_last = nullptr; it->grid = nullptr; it->cells = nullptr; it->photons = nullptr; it->photon_pos = nullptr; it->d_params = nullptr;
    delete[] it->h_params; it->h_params = nullptr;
    // camera queues of the pipelined loop; the members go back to the buffers the integrator was created with
    if (it->light_q.rays[0]) use_queues(it, it->light_q);
    for (int i = 0; i < 2; ++i) { cudaFree(it->cam_q.rays[i]); cudaFree(it->cam_q.states[i]); cudaFree(it->cam_q.count[i]); }
    cudaFree(it->cam_q.hits); cudaFree(it->cam_q.shadow_rays); cudaFree(it->cam_q.shadow_states);
    it->cam_q = imba_integrator::QueueSet{};
    if (it->tail_stream) { cudaStreamDestroy(it->tail_stream); it->tail_stream = nullptr; }
    if (it->ev_light_done) { cudaEventDestroy(it->ev_light_done); it->ev_light_done = nullptr; }
    if (it->ev_tail_done) { cudaEventDestroy(it->ev_tail_done); it->ev_tail_done = nullptr; }
    it->camera_pending = false;
}

static bool pipelined_algo(int algo) { return algo == IMBA_ALGO_BPT || algo == IMBA_ALGO_VCM || algo == IMBA_ALGO_PPM; }

int imba_vcm_allocate(imba_integrator* it) {
    imba_vcm_free(it);
    const int cap = std::max(it->cache_capacity, 1);
    it->table = (unsigned)cap * 2u; // inv_load_factor = 2 (rangesearch.h:26)
    // the queue buffers of the integrator are the light pass's; remember them before any swap
    for (int i = 0; i < 2; ++i) { it->light_q.rays[i] = it->rays[i]; it->light_q.states[i] = it->states[i]; it->light_q.count[i] = it->count[i]; }
    it->light_q.hits = it->hits; it->light_q.shadow_rays = it->shadow_rays; it->light_q.shadow_states = it->shadow_states;
    const bool two = pipelined_algo(it->cfg.algorithm) && it->cache_capacity > 0;   // second set + camera queues only where iterations are pipelined
    bool ok = true;
    for (int b = 0; b < (two ? 2 : 1) && ok; ++b) {
        imba_integrator::VcmSet& v = it->vset[b];
        ok = ok && cudaMalloc((void**)&v.cache, (size_t)cap * sizeof(DevLightVertex)) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&v.d_params, sizeof(VcmParams)) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&v.cache_last, 16) == cudaSuccess && cudaMemset(v.cache_last, 0, 16) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&v.grid, sizeof(GridParams)) == cudaSuccess && cudaMemset(v.grid, 0, sizeof(GridParams)) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&v.cells, sizeof(int) * ((size_t)it->table + 1)) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&v.photons, (size_t)cap * sizeof(DevPhoton)) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&v.photon_pos, (size_t)cap * sizeof(float4)) == cudaSuccess;
    }
    ok = ok && cudaMalloc((void**)&it->grid_bbox, 32) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&it->d_table_state, 16) == cudaSuccess && cudaMemset(it->d_table_state, 0, 16) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&it->block_sums, sizeof(int) * ((size_t)it->table / 2048 + 2)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&it->merge_found, (size_t)(IMBA_MAX_KNN + 1) * std::max(it->capacity, 1) * sizeof(int)) == cudaSuccess;
    it->h_params = new VcmParams[2];
    if (two) {
        const size_t n = (size_t)std::max(it->pixel_capacity, 1);
        for (int i = 0; i < 2 && ok; ++i) {
            ok = ok && cudaMalloc((void**)&it->cam_q.rays[i], n * 32) == cudaSuccess;
            ok = ok && cudaMalloc((void**)&it->cam_q.states[i], n * it->state_bytes) == cudaSuccess;
            ok = ok && cudaMalloc((void**)&it->cam_q.count[i], 16) == cudaSuccess;
        }
        ok = ok && cudaMalloc((void**)&it->cam_q.hits, n * 16) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&it->cam_q.shadow_rays, n * it->shadow_slots * 32) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&it->cam_q.shadow_states, n * it->shadow_slots * 16) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&it->tail_stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&it->ev_light_done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&it->ev_tail_done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventRecord(it->ev_tail_done, it->tail_stream) == cudaSuccess;
    }
    if (!ok) { imba_set_error("VCM allocation failed: %s", cudaGetErrorString(cudaGetLastError())); return IMBA_ERR_CUDA; }
    use_set(it, 0);
    return IMBA_OK;
}

/* VCMIntegrator::preprocess (vcm.h:74-81): base radius from the pixel size; LightVertices::compute_cache_size: trace
 * LIGHT_PATH_LEN_PROBES light paths, count the vertices that would be stored, size the cache to
 * margin * ceil(average) * path_count with margin 1.1 (10 for tiny path counts), light_vertices.cpp:121-128.
 * The probe is the reference's own simplified light tracer (vcm.cu: vcm_probe_*), not the light pass: its average path length is
 * what the slot count is a step function of (tests/test_vcm_gpu.py compares the counts with the reference's for six scenes). */
int imba_vcm_preprocess(imba_integrator* it) {
    it->base_radius = it->pixel_size * it->cfg.radius_factor;
    const int algo = it->cfg.algorithm;
    if (algo == IMBA_ALGO_LT || algo == IMBA_ALGO_VCM_PT) { it->cache_capacity = 0; return imba_vcm_allocate(it); }
    const int probes = std::min(10000, it->capacity);   // LIGHT_PATH_LEN_PROBES (light_vertices.h:61)
    it->cache_capacity = 0;
    int rc = imba_vcm_allocate(it);
    if (rc != IMBA_OK) return rc;
    // the reference's own probe (light_vertices.cpp:15-129; vcm.cu: vcm_probe_*): wavefronts until 256 or fewer paths are left
    const DeviceShading& S = *imba_shading_view(it->scene);
    int* d_vertex_count = it->cache_last;   // scratch: nothing is cached yet
    IMBA_CUDA(cudaMemsetAsync(d_vertex_count, 0, sizeof(int), it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(imba_vcm_probe_raygen(S, imba_iteration_seed(it->cfg.seed, ~0ull, 3), probes, it->rays[0], (PTState*)it->states[0], it->count[0], it->stream), IMBA_ERR_CUDA);
    g_imba_launches.fetch_add(1);
    int cur = 0, n = probes;
    for (int wave = 0; n > 256; ++wave, cur ^= 1) {
        TraceRequest rq; rq.d_rays = (const imba_ray*)it->rays[cur]; rq.d_hits = (imba_hit*)it->hits; rq.capacity = n; rq.d_count = it->count[cur]; rq.mode = 0;
        if ((rc = imba_scene_trace_ex(it->scene, rq, it->stream)) != IMBA_OK) return rc;
        IMBA_CUDA(cudaMemsetAsync(it->count[cur ^ 1], 0, sizeof(int), it->stream), IMBA_ERR_CUDA);
        IMBA_CUDA(imba_vcm_probe_shade(S, it->rays[cur], it->hits, (const PTState*)it->states[cur], it->count[cur], n, it->rays[cur ^ 1], (PTState*)it->states[cur ^ 1],
                                       it->count[cur ^ 1], d_vertex_count, it->stream), IMBA_ERR_CUDA);
        g_imba_launches.fetch_add(1);
        IMBA_CUDA(cudaMemcpyAsync(&n, it->count[cur ^ 1], sizeof(int), cudaMemcpyDeviceToHost, it->stream), IMBA_ERR_CUDA);
        IMBA_CUDA(cudaStreamSynchronize(it->stream), IMBA_ERR_CUDA);
        if (wave > 4096) { imba_set_error("cache-size probe did not terminate"); return IMBA_ERR_OVERFLOW; }
    }
    int vertex_count = 0;
    IMBA_CUDA(cudaMemcpy(&vertex_count, d_vertex_count, sizeof(int), cudaMemcpyDeviceToHost), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemset(d_vertex_count, 0, sizeof(int)), IMBA_ERR_CUDA);
    const float avg_len = (float)vertex_count / (float)probes;
    const float margin = (it->cfg.light_path_count < 10000 / 10) ? 10.0f : 1.1f;
    const double size = (double)margin * std::ceil(avg_len) * (double)it->cfg.light_path_count;
    if (size > 1.5e9) { imba_set_error("light vertex cache too large"); return IMBA_ERR_OVERFLOW; }
    it->cache_capacity = (int)size;
    return imba_vcm_allocate(it);
}

/* Camera pass of one iteration on the current parameter / cache set and the current queues. */
static int camera_pass(imba_integrator* it, int algo, uint64_t global_iter) {
    IMBA_CUDA(imba_vcm_camera_raygen(it->cam, it->d_params, it->cfg.spp, imba_iteration_seed(it->cfg.seed, global_iter, 2), it->pixel_capacity, it->rays[0],
                                     (VCMState*)it->states[0], it->count[0], it->stream), IMBA_ERR_CUDA);
    g_imba_launches.fetch_add(1);
    int cur = 0;
    for (int wave = 0; wave < it->cfg.max_path_len; ++wave, cur ^= 1) {
        const int rc = vcm_wavefront(it, algo, false, cur, it->pixel_capacity, true);
        if (rc != IMBA_OK) return rc;
    }
    IMBA_CUDA(cudaStreamWaitEvent(it->stream, it->ev_shadow_done, 0), IMBA_ERR_CUDA); // framebuffer consumers follow on it->stream
    return IMBA_OK;
}

static void grid_stats(imba_integrator* it) { // occupancy of the photon hash grid (diagnosis only)
    GridParams g; cudaMemcpy(&g, it->grid, sizeof(g), cudaMemcpyDeviceToHost);
    std::vector<int> ends(g.table);
    cudaMemcpy(ends.data(), it->cells, sizeof(int) * g.table, cudaMemcpyDeviceToHost);
    long long used = 0, sq = 0; int mx = 0, prev = 0; std::vector<int> sizes;
    for (unsigned c = 0; c < g.table; ++c) { const int n = ends[c] - prev; prev = ends[c]; if (n > 0) { ++used; sq += (long long)n * n; mx = std::max(mx, n); sizes.push_back(n); } }
    std::sort(sizes.begin(), sizes.end());
    fprintf(stderr, "[imba] photon grid: %d photons, table %u, %lld cells used, largest %d, median %d, p99 %d, mean photons scanned per visited photon's cell %.1f, radius %g\n",
            g.count, g.table, used, mx, sizes.empty() ? 0 : sizes[sizes.size() / 2], sizes.empty() ? 0 : sizes[sizes.size() * 99 / 100], g.count ? (double)sq / g.count : 0.0, sqrt(g.radius_sqr));
}

/* `iterations` progressive iterations (VCMIntegrator::render, vcm.cpp:26-55, once per iteration).  Within one call consecutive
 * iterations are pipelined (BPT, VCM, PPM): the camera pass of iteration k is enqueued only after the light wavefronts of
 * iteration k+1, so that it runs on the main stream BESIDE the tail of that light pass (the few paths trapped in glass, one
 * warp each, latency-bound for ~10 ms at 1080p: vcm.cu) on a side stream.  What a camera pass reads is double-buffered; when
 * the call returns no camera pass is pending.  IMBA_VCM_NO_PIPELINE=1 (and the sorted order, which shares its scratch
 * buffers) runs the passes strictly one after the other as the reference does. */
static int vcm_render_iterations(imba_integrator* it, int iterations);
int imba_vcm_render(imba_integrator* it, int iterations) {
    const int rc = vcm_render_iterations(it, iterations);
    if (rc != IMBA_OK) {   // leave no half-finished pipeline behind: no pending camera pass, light-pass queues and set 0 current
        it->camera_pending = false;
        if (it->cache_last) { use_set(it, 0); use_queues(it, it->light_q); }
    }
    return rc;
}
static int vcm_render_iterations(imba_integrator* it, int iterations) {
    if (!it->cache_last) { imba_set_error("imba_integrator_preprocess must be called before rendering with a VCM-family integrator"); return IMBA_ERR_INVALID; }
    const int algo = it->cfg.algorithm;
    const bool pipelined = it->vset[1].cache != nullptr && !getenv("IMBA_VCM_NO_PIPELINE") && !getenv("IMBA_VCM_SORT");
    int rc = IMBA_OK;
    for (int k = 0; k < iterations; ++k, ++it->iteration) {
        const uint64_t global_iter = it->iteration * (uint64_t)it->cfg.rank_count + (uint64_t)it->cfg.rank;
        const int b = pipelined ? (int)(it->iteration & 1) : 0;
        use_set(it, b); use_queues(it, it->light_q);
        // the radius schedule follows the GLOBAL iteration index when iterations are sharded over ranks
        const VcmParams P = iteration_params(it, global_iter + 1);
        if ((rc = set_params(it, P, b)) != IMBA_OK) return rc;
        if (algo != IMBA_ALGO_VCM_PT) {
            const bool beside = pipelined && it->camera_pending;
            // the light tracer has no camera pass: the light paths of up to iterations_per_pass iterations form ONE pass
            const int group = algo == IMBA_ALGO_LT ? std::min(std::max(it->cfg.iterations_per_pass, 1), iterations - k) : 1;
            PtSeeds seeds;
            for (int g = 0; g < group; ++g)
                seeds.base[g] = imba_iteration_seed(it->cfg.seed, (it->iteration + g) * (uint64_t)it->cfg.rank_count + (uint64_t)it->cfg.rank, 1);
            rc = light_pass(it, algo, seeds, it->cfg.light_path_count, it->cfg.light_path_count * group, nullptr, !getenv("IMBA_NO_GRAPH"), beside ? it->tail_stream : nullptr);
            k += group - 1; it->iteration += group - 1;
            if (rc != IMBA_OK) return rc;
            if (beside) { // the previous iteration's camera pass, while the tail of this light pass crawls along on the side stream
                use_set(it, it->pending_set); use_queues(it, it->cam_q);
                rc = camera_pass(it, algo, it->pending_global_iter);
                it->camera_pending = false;
                use_set(it, b); use_queues(it, it->light_q);
                if (rc != IMBA_OK) return rc;
                IMBA_CUDA(cudaStreamWaitEvent(it->stream, it->ev_tail_done, 0), IMBA_ERR_CUDA);
            }
            if (algo != IMBA_ALGO_LT) { // light_vertices_.build(pm_radius_, algo != ALGO_BPT), vcm.cpp:98-99
                IMBA_CUDA(imba_grid_build(it->cache, it->cache_last, it->cache_capacity, P.pm_radius, it->grid, it->grid_bbox, it->cells, it->table, it->block_sums,
                                          it->photons, it->photon_pos, algo != IMBA_ALGO_BPT, it->d_table_state, it->stream), IMBA_ERR_CUDA);
                g_imba_launches.fetch_add(algo != IMBA_ALGO_BPT ? 7 : 1);
                if (getenv("IMBA_GRID_STATS") && algo != IMBA_ALGO_BPT) grid_stats(it);
            }
        }
        if (algo != IMBA_ALGO_LT) {
            if (pipelined && k + 1 < iterations) { it->camera_pending = true; it->pending_set = b; it->pending_global_iter = global_iter; }
            else {
                if (pipelined) use_queues(it, it->cam_q);
                rc = camera_pass(it, algo, global_iter);
                use_queues(it, it->light_q);
                if (rc != IMBA_OK) return rc;
            }
        }
    }
    return IMBA_OK;
}

/* ---- parity harness: one VCM stage on an injected batch -----------------------------------------------------------------
 * stage 0 = process_light_rays (vcm.cpp:193-258), stage 1 = process_camera_rays (vcm.cpp:323-439) for the integrator's
 * algorithm at progressive iteration `iteration` (1-based, as cur_iteration_ after the increment in render()).
 * Host pointers, synchronous.  Light stage: the vertex cache is cleared first and returned (112-byte DevLightVertex
 * records, append order).  Camera stage: `h_cache_in` is installed as the vertex cache (and the photon grid is built on
 * it for PPM / VCM) before shading.  Outputs like imba_pt_shade_batch: continuation queue (stable order), shadow queue
 * (slot order, empty slots removed); contributions go to the accumulation buffer. */
extern "C" int imba_vcm_stage_batch(imba_integrator* it, int stage, int iteration, float base_radius, int count, const imba_ray* h_rays,
                                    const imba_hit* h_hits, const void* h_states, const void* h_cache_in, int cache_in_count,
                                    imba_ray* h_rays_out, void* h_states_out, int* out_count, imba_ray* h_shadow_rays, void* h_shadow_states,
                                    int* shadow_count, void* h_cache_out, int* cache_out_count) {
    if (!it || it->cfg.algorithm == IMBA_ALGO_PT || !it->cache_last || count < 0 || count > it->capacity || iteration < 1) {
        imba_set_error("needs a preprocessed VCM-family integrator whose queue capacity covers the batch"); return IMBA_ERR_INVALID;
    }
    const int algo = it->cfg.algorithm;
    const DeviceShading& S = *imba_shading_view(it->scene);
    cudaStream_t st = it->stream;
    it->base_radius = base_radius;
    const VcmParams P = iteration_params(it, (uint64_t)iteration);
    int rc = set_params(it, P);
    if (rc != IMBA_OK) return rc;
    const int slots = stage == 0 ? 1 : it->shadow_slots;
    IMBA_CUDA(cudaMemcpyAsync(it->rays[0], h_rays, (size_t)count * 32, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(it->hits, h_hits, (size_t)count * 16, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(it->states[0], h_states, (size_t)count * sizeof(VCMState), cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(it->count[0], &count, 4, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    if (stage == 0) {
        IMBA_CUDA(cudaMemsetAsync(it->cache_last, 0, sizeof(int), st), IMBA_ERR_CUDA);
    } else {
        if (cache_in_count > it->cache_capacity) { imba_set_error("vertex cache too small for the injected vertices"); return IMBA_ERR_OVERFLOW; }
        IMBA_CUDA(cudaMemcpyAsync(it->cache, h_cache_in, (size_t)cache_in_count * sizeof(DevLightVertex), cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
        IMBA_CUDA(cudaMemcpyAsync(it->cache_last, &cache_in_count, 4, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
        IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
        IMBA_CUDA(imba_grid_build(it->cache, it->cache_last, it->cache_capacity, P.pm_radius, it->grid, it->grid_bbox, it->cells, it->table, it->block_sums,
                                  it->photons, it->photon_pos, algo != IMBA_ALGO_BPT && algo != IMBA_ALGO_VCM_PT, it->d_table_state, st), IMBA_ERR_CUDA);
    }
    IMBA_CUDA(imba_sort_hits_by_material(it->part, it->hits, S.indices, S.material_count, (int*)it->shadow_states, slots, it->count[0], 0, count, st), IMBA_ERR_CUDA);
    VcmLaunch a = base_launch(it, 0, count);
    a.out_rays = nullptr; a.out_states = nullptr; a.out_count = nullptr; // continuation in place, then the stable compaction
    IMBA_CUDA(stage == 0 ? imba_vcm_light_shade(algo, a, st) : imba_vcm_camera_shade(algo, a, st), IMBA_ERR_CUDA);
    // shadow queue: valid slots compacted into rays[1] / states[1]
    int n = 0;
    const int nslots = count * slots;
    if (nslots > it->capacity * it->shadow_slots) { imba_set_error("shadow queue too small"); return IMBA_ERR_OVERFLOW; }
    {   // the partition buffers cover `capacity` entries: compact the shadow slots in pieces of at most `capacity`
        int done = 0, total = 0;
        while (done < nslots) {
            const int piece = std::min(it->capacity, nslots - done);
            IMBA_CUDA(imba_compact_alive(it->part, nullptr, it->shadow_rays + 2 * (size_t)done, it->shadow_states + done, sizeof(ShadowState), nullptr, piece, piece,
                                         it->rays[1], it->states[1], it->count[1], st), IMBA_ERR_CUDA);
            IMBA_CUDA(cudaMemcpyAsync(&n, it->count[1], 4, cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
            IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
            if (piece == 0) n = 0;
            IMBA_CUDA(cudaMemcpyAsync(h_shadow_rays + total, it->rays[1], (size_t)n * 32, cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
            IMBA_CUDA(cudaMemcpyAsync((char*)h_shadow_states + (size_t)total * sizeof(ShadowState), it->states[1], (size_t)n * sizeof(ShadowState), cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
            IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
            total += n; done += piece;
            if (piece == 0) break;
        }
        *shadow_count = total;
    }
    IMBA_CUDA(imba_compact_alive(it->part, it->hits, it->rays[0], it->states[0], sizeof(VCMState), it->count[0], 0, count,
                                 it->rays[1], it->states[1], it->count[1], st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(&n, it->count[1], 4, cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
    if (count == 0) n = 0;
    IMBA_CUDA(cudaMemcpyAsync(h_rays_out, it->rays[1], (size_t)n * 32, cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(h_states_out, it->states[1], (size_t)n * sizeof(VCMState), cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    *out_count = n;
    int nv = 0;
    if (stage == 0 && h_cache_out) {
        IMBA_CUDA(cudaMemcpyAsync(&nv, it->cache_last, 4, cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
        IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
        nv = std::min(nv, it->cache_capacity);
        IMBA_CUDA(cudaMemcpyAsync(h_cache_out, it->cache, (size_t)nv * sizeof(DevLightVertex), cudaMemcpyDeviceToHost, st), IMBA_ERR_CUDA);
    }
    if (cache_out_count) *cache_out_count = nv;
    IMBA_CUDA(cudaStreamSynchronize(st), IMBA_ERR_CUDA);
    g_imba_launches.fetch_add(12);
    return IMBA_OK;
}

/* Parity harness: the rays and 48-byte states the light (light = 1) or camera (light = 0) ray generation produces for
 * `count` paths / samples with the given seed base and iteration constants (vcm.cpp:64-96, 109-129). */
extern "C" int imba_vcm_raygen_batch(imba_integrator* it, int light, int iteration, float base_radius, uint32_t seed_base, int count, imba_ray* h_rays, void* h_states) {
    if (!it || it->cfg.algorithm == IMBA_ALGO_PT || !it->d_params || count < 0 || count > it->capacity || iteration < 1) {
        imba_set_error("needs a preprocessed VCM-family integrator whose queue capacity covers the batch"); return IMBA_ERR_INVALID;
    }
    it->base_radius = base_radius;
    const VcmParams P = iteration_params(it, (uint64_t)iteration);
    int rc = set_params(it, P);
    if (rc != IMBA_OK) return rc;
    const DeviceShading& S = *imba_shading_view(it->scene);
    PtSeeds one_seed; one_seed.base[0] = seed_base;
    if (light) IMBA_CUDA(imba_vcm_light_raygen(S, it->d_params, one_seed, count, count, it->rays[0], (VCMState*)it->states[0], it->count[0], it->stream), IMBA_ERR_CUDA);
    else       IMBA_CUDA(imba_vcm_camera_raygen(it->cam, it->d_params, it->cfg.spp, seed_base, count, it->rays[0], (VCMState*)it->states[0], it->count[0], it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(h_rays, it->rays[0], (size_t)count * 32, cudaMemcpyDeviceToHost, it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(h_states, it->states[0], (size_t)count * sizeof(VCMState), cudaMemcpyDeviceToHost, it->stream), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaStreamSynchronize(it->stream), IMBA_ERR_CUDA);
    g_imba_launches.fetch_add(1);
    return IMBA_OK;
}
