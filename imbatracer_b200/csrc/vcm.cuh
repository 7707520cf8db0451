/* vcm.cuh -- parameters shared by the VCM kernels (vcm.cu) and their host driver (integrator.cu). */
#ifndef IMBA_VCM_CUH
#define IMBA_VCM_CUH
#include "integrator.cuh"

#define IMBA_MAX_KNN 32

struct VcmParams {            // per-iteration constants of VCMIntegrator::render (vcm.cpp:26-46)
    float pm_radius, base_radius, mis_eta_vc, mis_eta_vm;
    int   light_path_count, num_connections, num_knn, max_path_len;
};
struct GridParams {           // HashGrid members (rangesearch.h:148-156) + the vertex count of the iteration
    float lo[3], hi[3];
    float radius_sqr, inv_cell_size;
    int   count;
    unsigned table;           // cell_ends_.size(): 2 * photon count, only ever grows (rangesearch.h:26,34-35)
};
struct VcmLaunch {
    const DeviceShading* S; const VcmParams* P; DevCamera cam;   // P lives in device memory: launches (and captured graphs) stay valid across iterations
    float4* rays; const int4* hits; VCMState* states; const int* sorted; const int* hit_count_ptr; const int* queue_count_ptr; int capacity;
    float4* shadow_rays; ShadowState* shadow_states;
    DevLightVertex* cache; int cache_capacity; int* cache_last;
    const GridParams* grid; const int* cells; unsigned table; const DevPhoton* photons;
    const float4* photon_pos;   // positions only, in cell order: what the k-NN search streams through
    int* merge_found;           // k-NN results per queue slot: row 0 = count, rows 1..k = photon indices; row stride = merge_stride
    int merge_stride; int knn;
    float* fb; unsigned long long* ray_stats;
    float4* out_rays; VCMState* out_states; int* out_count; // next wavefront's queue (fused compaction)
};

/* LightVertices::compute_cache_size (light_vertices.cpp:15-129): probe paths and one wavefront of the probe's own shading loop */
cudaError_t imba_vcm_probe_raygen(const DeviceShading& S, uint32_t seed_base, int count, float4* rays, PTState* states, int* count_out, cudaStream_t st);
cudaError_t imba_vcm_probe_shade(const DeviceShading& S, const float4* rays, const int4* hits, const PTState* states, const int* count_ptr, int capacity,
                                 float4* out_rays, PTState* out_states, int* out_count, int* vertex_count, cudaStream_t st);
cudaError_t imba_vcm_light_raygen(const DeviceShading& S, const VcmParams* P, const PtSeeds& seeds, int per_iteration, int total, float4* rays, VCMState* states, int* count_out, cudaStream_t st);
cudaError_t imba_vcm_camera_raygen(const DevCamera& cam, const VcmParams* P, int spp, uint32_t seed_base, int count, float4* rays, VCMState* states, int* count_out, cudaStream_t st);
cudaError_t imba_vcm_light_shade(int algo, const VcmLaunch& a, cudaStream_t st);
cudaError_t imba_vcm_camera_shade(int algo, const VcmLaunch& a, cudaStream_t st);
cudaError_t imba_vcm_light_tail(int algo, bool masks, const VcmLaunch& a, const TraceArgs& T, int capacity, cudaStream_t st);
cudaError_t imba_grid_build(const DevLightVertex* cache, const int* cache_last, int cache_capacity, float radius, GridParams* grid, int* bbox, int* cells, unsigned table,
                            int* block_sums, DevPhoton* photons, float4* photon_pos, bool build_cells, unsigned* table_state, cudaStream_t st);
#endif
