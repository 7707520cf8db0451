/* api_common.h -- shared state of the C-ABI layer (error channel, launch counter, scene object). */
#ifndef IMBA_API_COMMON_H
#define IMBA_API_COMMON_H
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <mutex>
#include <cstdio>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/imba_b200.h"
#include "bvh8_build.h"
#include "kernels.h"

void imba_set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_imba_launches;

#define IMBA_CUDA(expr, ret)                                                                         \
    do { cudaError_t e__ = (expr);                                                                   \
         if (e__ != cudaSuccess) { imba_set_error("%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); return ret; } \
    } while (0)

struct DeviceShading; // shade.cu

struct imba_scene {
    int       device = 0;
    int       sm_count = 0;
    // acceleration structure
    uint4*    d_nodes = nullptr; float4* d_tris = nullptr; float4* d_insts = nullptr;
    uint32_t  top_root = 0; int flat_inst = -1;
    int64_t   stats[5] = { 0, 0, 0, 0, 0 };
    float     build_info[2] = { 0.0f, 0.0f }; // device build: seconds of device time, levels of the deepest mesh tree
    // opacity masks
    bool      has_masks = false; bool owns_masks = true;
    int*      d_mask_indices = nullptr; float2* d_mask_texcoords = nullptr; int* d_masks = nullptr; uint8_t* d_mask_buffer = nullptr;
    // work counters (one per in-flight launch)
    unsigned* d_counters = nullptr; std::atomic<unsigned> next_counter{ 0 };
    static constexpr unsigned kCounters = 4096;
    // staging for host-buffer traversal: kStageSets independent pipelines (one per concurrent caller: the reference's
    // TileScheduler threads call traverse concurrently on disjoint queues, tile_scheduler.h:69-78), each rotating its chunks over
    // kStageSlots streams so that H2D, traversal and D2H of different chunks overlap
#ifndef IMBA_CHUNK_LOG2
#define IMBA_CHUNK_LOG2 19
#endif
    static constexpr int kChunkRays = 1 << IMBA_CHUNK_LOG2;   // 2^19: 16 MiB of rays up, 8 MiB of hits down per chunk
    static constexpr int kStageSlots = 4;
    static constexpr int kStageSets = 4;
    struct StageSet { float4* d_rays[kStageSlots] = {}; int4* d_hits[kStageSlots] = {}; cudaStream_t stream[kStageSlots] = {};
                      cudaEvent_t done[kStageSlots] = {}; bool busy = false; };
    StageSet stage[kStageSets];
    std::mutex stage_mutex; std::condition_variable stage_free;
    // shading data (geometry, materials, lights, textures), null until uploaded
    DeviceShading* shading = nullptr;
    int grid_cache[12] = { 0 };
};

/* Uploads a converted structure; fills scene->d_nodes etc. */
int imba_scene_upload_accel(imba_scene* s, const Bvh8& b);
/* Builds the structure on the device from meshes + instances (bvh8_gpu.cu). */
int imba_scene_build_accel_device(imba_scene* s, const imba_scene_desc* d, float* seconds);
/* Launches traversal on device buffers. */
struct TraceRequest {
    const imba_ray* d_rays = nullptr; imba_hit* d_hits = nullptr;
    int capacity = 0;               // upper bound of the ray count (sizes the grid)
    const int* d_count = nullptr;   // optional device-resident ray count
    int count_mul = 1;              // ray count = *d_count * count_mul
    int mode = 0;                   // 0 closest, 1 any hit, 2 any hit + shadow resolve
    uint32_t* d_counters = nullptr;
    unsigned* d_work_counter = nullptr; // optional dedicated work counter (launches captured into a CUDA graph)
    const void* d_shadow_states = nullptr; float* d_framebuffer = nullptr;
};
int imba_scene_trace_ex(imba_scene* s, const TraceRequest& rq, cudaStream_t stream);
void imba_scene_trace_args(const imba_scene* s, TraceArgs* a);
int imba_scene_trace(imba_scene* s, const imba_ray* d_rays, imba_hit* d_hits, int count, int any_hit, uint32_t* d_counters,
                     cudaStream_t stream);

/* Two-level structure in the reference's BVH2 layout (accel_build.cpp, lbvh.cu). */
struct imba_accel {
    std::vector<imba_node2> nodes; std::vector<imba_vec4> tris; std::vector<imba_instance_node> inst_nodes;
    int root = 0;
};
int imba_accel_finish_top(imba_accel* out, const imba_scene_desc* d, const int* mesh_root, const float* mesh_box6);

#endif
