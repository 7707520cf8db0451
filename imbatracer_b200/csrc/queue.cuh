/* queue.cuh -- device ray-queue primitives (queue.cu). */
#ifndef IMBA_QUEUE_CUH
#define IMBA_QUEUE_CUH
#include <algorithm>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int PART_THREADS = 256;
constexpr int PART_CHUNK = 8192;   // elements per block

struct PartitionBuffers {
    int* keys = nullptr;        // [capacity]
    int* indices = nullptr;     // [capacity]  sorted element indices (the reference's sorted_indices_)
    int* block_hist = nullptr;  // [max_bins][max_blocks]
    int* bin_start = nullptr;   // [max_bins + 1]
    int  max_blocks = 0, max_bins = 0;
};

cudaError_t imba_partition_alloc(PartitionBuffers& b, int capacity, int max_bins);
void        imba_partition_free(PartitionBuffers& b);

/* compact_hits + sort_by_material: b.indices[0 .. b.bin_start[nmat]) = hit rays grouped by material (stable);
 * b.bin_start[nmat] = hit count.  shadow_pixel (optional): ShadowState array (shadow_slots per ray) whose pixel_id is reset to -1. */
cudaError_t imba_sort_hits_by_material(const PartitionBuffers& b, const int4* hits, const int4* indices, int nmat, int* shadow_pixel, int shadow_slots,
                                       const int* count_ptr, int count_imm, int capacity, cudaStream_t st, bool stable = true);
/* compact_rays: stable copy of the live rays + states (hits == nullptr: pixel_id >= 0; else also hit.tri_id >= 0). */
cudaError_t imba_compact_alive(const PartitionBuffers& b, const int4* hits, const float4* rays, const void* states, int state_bytes,
                               const int* count_ptr, int count_imm, int capacity, float4* rays_out, void* states_out, int* count_out, cudaStream_t st);
cudaError_t imba_gather_queue(const int* idx, const int* n_ptr, int n_imm, int capacity, const float4* rays, const int4* hits, const void* states,
                              int state_bytes, float4* rays_out, int4* hits_out, void* states_out, cudaStream_t st);
#endif
