/* integrator.cuh -- shared declarations of the device integrators (pt.cu, vcm.cu, integrator.cu). */
#ifndef IMBA_INTEGRATOR_CUH
#define IMBA_INTEGRATOR_CUH
#include <algorithm>
#include "api_common.h"
#include "queue.cuh"
#include "shade.cuh"

/* PerspectiveCamera (render/ray_gen/camera.h:12-112) in closed form.  The reference builds 4x4 raster<->world
 * matrices and inverts them numerically; the quirks are kept: aspect fixed at 1 (the image is stretched when
 * width != height), raster y runs along cross(dir, right) (downwards), image-plane distance = width / (2 tan(fov/2)). */
struct DevCamera {
    float px, py, pz;       // position
    float dx, dy, dz;       // forward (unit)
    float rx, ry, rz;       // right = normalize(cross(dir, up))
    float ux, uy, uz;       // cross(dir, right): raster +y
    float tan_half;
    int   width, height;
    float img_plane_dist;
    __host__ __device__ __forceinline__ V3 ray_dir(float sx, float sy) const {
        const float a = (2.0f * sx / (float)width - 1.0f) * tan_half, b = (2.0f * sy / (float)height - 1.0f) * tan_half;
        return normalize(v3(dx + a * rx + b * ux, dy + a * ry + b * uy, dz + a * rz + b * uz));
    }
    /* world_to_raster (camera.h:67-70); depth <= 0 means behind the camera */
    __host__ __device__ __forceinline__ float2 to_raster(V3 p, float& depth) const {
        const V3 d = v3(p.x - px, p.y - py, p.z - pz);
        depth = d.x * dx + d.y * dy + d.z * dz;
        const float a = (d.x * rx + d.y * ry + d.z * rz) / (depth * tan_half), b = (d.x * ux + d.y * uy + d.z * uz) / (depth * tan_half);
        return make_float2((a + 1.0f) * 0.5f * (float)width, (b + 1.0f) * 0.5f * (float)height);
    }
};

/* Seeding of a primary / light ray (ray_gen.h:57-72, 139-147): Bernstein hash of (seed_base, i) in 32-bit int
 * arithmetic, RNG(seed) with the int sign-extended, then discard((seed % 5) + 16 + k % 5).  Returns the RNG state. */
__host__ __device__ __forceinline__ uint64_t seed_primary(uint32_t seed_base, int i, int k) {
    uint32_t seed = seed_base;
    seed = (33u * seed) ^ (uint32_t)i; seed = (33u * seed) ^ (uint32_t)i; seed = (33u * seed) ^ (uint32_t)i; seed = (33u * seed) ^ (uint32_t)i;
    const int s = (int)seed;
    uint64_t state = (uint64_t)(int64_t)s;
    const int n = (s % 5) + 16 + k % 5;
    for (int j = 0; j < n; ++j) {
        const uint32_t c = (uint32_t)(state >> 32), x = (uint32_t)(state & 0xffffffffu);
        state = (uint64_t)x * 4294883355ull + c;
    }
    return state;
}

const DeviceShading* imba_shading_view(const imba_scene* s);

struct imba_integrator {
    imba_scene* scene = nullptr;
    imba_render_settings cfg;
    DevCamera cam;
    cudaStream_t stream = nullptr;
    // PT: the shadow + resolve traversal of wave k runs on a second stream beside the closest-hit traversal of wave k + 1
    cudaStream_t shadow_stream = nullptr; cudaEvent_t ev_shaded = nullptr, ev_shadow_done = nullptr;
    int capacity = 0;
    int state_bytes = sizeof(PTState);
    float4* rays[2] = { nullptr, nullptr };
    char*   states[2] = { nullptr, nullptr };
    int*    count[2] = { nullptr, nullptr };
    int4*   hits = nullptr;
    float4* shadow_rays = nullptr; ShadowState* shadow_states = nullptr;
    PartitionBuffers part;
    float*  fb = nullptr; bool owns_fb = true;
    unsigned long long* d_stats = nullptr;   // [0] primary rays, [1] shadow rays
    float*  d_probe = nullptr;               // [0] total (float), [1] count (int)
    uint64_t iteration = 0;
    float   pixel_size = 1.0f;
    // VCM family (vcm_host.cu)
    int     shadow_slots = 1;                // num_connections + 1 shadow rays per camera vertex (frontend/main.cpp:132)
    int     pixel_capacity = 0;              // width * height * spp
    struct DevLightVertex* cache = nullptr; int cache_capacity = 0; int* cache_last = nullptr;
    struct GridParams* grid = nullptr; int* grid_bbox = nullptr; int* cells = nullptr; unsigned table = 0; int* block_sums = nullptr;
    struct DevPhoton* photons = nullptr;
    float4* photon_pos = nullptr; int* merge_found = nullptr; // SoA copy of the photon positions; k-NN lists of the camera wavefront
    float   base_radius = 0.0f;
    struct VcmParams* d_params = nullptr;    // per-iteration constants in device memory
    // Pipelining of consecutive iterations inside one render(n) call (vcm_host.cu: imba_vcm_render): the camera pass of iteration
    // k-1 runs on the main stream beside the latency-bound tail of the light pass of iteration k (side stream).  The data a
    // camera pass reads (vertex cache, photon grid, parameters) is therefore double-buffered, and the camera pass has queues of
    // its own; the members above (cache, grid, d_params, rays, ...) always point at the set the next launches are to use.
    struct VcmSet { struct DevLightVertex* cache; int* cache_last; struct GridParams* grid; int* cells; struct DevPhoton* photons; float4* photon_pos; struct VcmParams* d_params; };
    struct QueueSet { float4* rays[2]; char* states[2]; int* count[2]; int4* hits; float4* shadow_rays; ShadowState* shadow_states; };
    VcmSet vset[2] = {};
    QueueSet light_q = {}, cam_q = {};
    struct VcmParams* h_params = nullptr;    // host copies of the two parameter blocks
    unsigned* d_table_state = nullptr;       // current size of the photon hash table (only grows, rangesearch.h:33-43), shared by both grids
    cudaStream_t tail_stream = nullptr; cudaEvent_t ev_light_done = nullptr, ev_tail_done = nullptr;
    bool camera_pending = false; uint64_t pending_global_iter = 0; int pending_set = 0;
    unsigned* tail_counters = nullptr; int tail_counter_next = -1; // dedicated traversal work counters of the captured graph
    cudaGraphExec_t tail_graph = nullptr;    // captured light-pass wavefronts for the long tail of small queues
    // the captured launches hold the framebuffer pointer: graphs of other framebuffers the caller alternates between are parked here
    static constexpr int kParkedGraphs = 4;
    float* parked_fb[kParkedGraphs] = {}; cudaGraphExec_t parked_graph[kParkedGraphs] = {};
};
int imba_vcm_allocate(imba_integrator* it);
void imba_vcm_free(imba_integrator* it);
int imba_vcm_preprocess(imba_integrator* it);
int imba_vcm_render(imba_integrator* it, int iterations);
uint32_t imba_iteration_seed(uint64_t seed, uint64_t global_iteration, uint32_t stream_id);




struct PtSeeds { uint32_t base[16]; };   // seed base of every progressive iteration of one pass
cudaError_t imba_pt_raygen(const DevCamera& cam, int spp, const PtSeeds& seeds, int per_iteration, int first, int count, float4* rays, PTState* states, int* count_out, cudaStream_t st);
cudaError_t imba_pt_shade(const DeviceShading& S, float4* rays, const int4* hits, PTState* states, const int* sorted, const int* hit_count_ptr, const int* queue_count_ptr, int capacity,
                          float4* shadow_rays, ShadowState* shadow_states, float* fb, int max_path_len, unsigned long long* ray_stats,
                          float4* out_rays, PTState* out_states, int* out_count, cudaStream_t st);

/* Warp-aggregated queue push (RayQueue::push, ray_queue.h:123-130): one atomicAdd per converged group of lanes. */
__device__ __forceinline__ int queue_push_slot(int* count) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}
/* Statistics counter (TileScheduler::total_shadow_rays_, tile_scheduler.h:117-127): one atomic per converged group of lanes instead
 * of one per ray on a single global address. */
__device__ __forceinline__ void count_rays(unsigned long long* counter) {
    const unsigned m = __activemask();
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(counter, (unsigned long long)__popc(m));
}
cudaError_t imba_pt_miss(const DeviceShading& S, const float4* rays, const int4* hits, const PTState* states, const int* sorted, const int* hit_count_ptr, const int* queue_count_ptr,
                         int capacity, float* fb, cudaStream_t st);
cudaError_t imba_probe_raygen(const DevCamera& cam, float4* rays, int count, cudaStream_t st);
cudaError_t imba_probe_reduce(const float4* rays, const int4* hits, int quads, float* total, int* count, cudaStream_t st);
cudaError_t imba_clear_shadow(ShadowState* s, int n, cudaStream_t st);

#endif
