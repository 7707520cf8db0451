/* kernels.h -- internal launch interface between the C-ABI layer (api.cu) and the CUDA kernels. */
#ifndef IMBA_KERNELS_H
#define IMBA_KERNELS_H
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef IMBA_TRAVERSAL_STACK
#define IMBA_TRAVERSAL_STACK 40 /* 8-byte entries per ray; scenes needing more are rejected at upload */
#endif

struct TraceArgs {
    const uint4*   nodes;      // Node8[] as 5 x uint4
    const float4*  tris;       // Tri48[] as 3 x float4
    const float4*  insts;      // Inst64[] as 4 x float4
    uint32_t       top_root;
    int            flat_inst;  // >= 0: single identity instance, top_root is the mesh root and rays never change space
    const float4*  rays;       // imba_ray[] as 2 x float4
    int4*          hits;       // imba_hit[]
    int            count;
    const int*     count_ptr;  // if non-null the ray count is read from device memory (wavefront loops without host sync)
    int            count_mul;  // ray count = *count_ptr * count_mul (several shadow slots per queue entry)
    unsigned*      counter;    // work counter, zero at launch
    uint32_t*      stats;      // 4 x u32 per ray or null
    // opacity masks (core/mask.h), only read by the MASKS instantiations
    const int*     indices;    // 4 ints per global triangle
    const float2*  texcoords;
    const int*     masks;      // 3 ints per material
    const uint8_t* mask_buffer;
    // mode 2 (any hit + resolve): ShadowState[] (16 bytes: pixel_id, r, g, b) and the W*H*3 accumulation buffer
    const void*    shadow_states;
    float*         framebuffer;
};

int         imba_trace_block_threads();
/* mode: 0 closest hit, 1 any hit, 2 any hit + shadow resolve */
int         imba_trace_blocks_per_sm(int mode, bool masks, bool stats, bool flat);
cudaError_t imba_launch_trace(const TraceArgs& a, int mode, bool masks, bool stats, int grid, cudaStream_t s);

#endif
