/* queue.cu -- ray-queue primitives on device arrays: stable partition / counting sort by a small key.
 *
 * Replaces RayQueue::compact_hits (ray_queue.h:161-185, serial), sort_by_material (ray_queue.h:211-244, two TBB passes
 * + serial scan, racy scatter) and compact_rays (ray_queue.h:188-207, serial) by one three-kernel counting sort:
 *   1. part_hist<KeyFn>: every block takes a contiguous chunk, computes the key of each element (material id of the
 *      hit, "miss", "terminated"), stores it and builds a shared-memory histogram -> block_hist[bin][block];
 *   2. part_scan: exclusive scan over (bin-major, block-minor) -> start offset of every (bin, block) and of every bin;
 *   3. part_scatter: each block re-reads its keys and writes element indices to their final position; the rank inside a
 *      block comes from warp match/ballot/popc, so the result is STABLE (deterministic), unlike the reference's atomic
 *      scatter.
 * Nothing is moved: the shading kernels gather through the index array (the reference's sorted_indices_, ray_queue.h:
 * 114-116); compact_rays gathers rays + states into the other half of a ping-pong pair.
 * Ray counts live in device memory (count_ptr), so a whole wavefront loop is enqueued without host synchronisation.
 */
#include "queue.cuh"
#include "api_common.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;

struct HitMaterialKey {   // compact_hits + sort_by_material: key = material id, misses go to the last bin
    const int4* hits; const int4* indices; int nmat; int* shadow_pixel; int shadow_slots; // shadow_pixel: optional, stride 4 ints
    __device__ __forceinline__ int operator()(int i) const {
        const int tri = __ldg(&hits[i].x);
        if (shadow_pixel) for (int k = 0; k < shadow_slots; ++k) shadow_pixel[4 * ((size_t)i * shadow_slots + k)] = -1; // no shadow rays yet
        if (tri < 0) return nmat;
        const int m = __ldg(&indices[tri].w);
        return m < 0 ? 0 : (m >= nmat ? nmat - 1 : m);
    }
};

struct AliveKey {         // compact_rays: keep rays that hit something and were not terminated (pixel_id >= 0)
    const int4* hits; const char* states; int state_bytes;
    __device__ __forceinline__ int operator()(int i) const {
        const bool alive = __ldg(&hits[i].x) >= 0 && *reinterpret_cast<const int*>(states + (size_t)i * state_bytes) >= 0;
        return alive ? 0 : 1;
    }
};

struct PixelAliveKey {    // standalone compact_rays (no hit array): keep states with pixel_id >= 0
    const char* states; int state_bytes;
    __device__ __forceinline__ int operator()(int i) const {
        return *reinterpret_cast<const int*>(states + (size_t)i * state_bytes) >= 0 ? 0 : 1;
    }
};

struct ArrayKey {         // sort_by_material with the keys computed by the caller (the reference's GetMatIDFn runs on the host)
    const int* keys_in; int nbins;
    __device__ __forceinline__ int operator()(int i) const { const int k = __ldg(keys_in + i); return k < 0 ? 0 : (k >= nbins ? nbins - 1 : k); }
};

struct HitKey {           // compact_hits alone: hits first
    const int4* hits;
    __device__ __forceinline__ int operator()(int i) const { return __ldg(&hits[i].x) >= 0 ? 0 : 1; }
};

template <class KeyFn>
__global__ void __launch_bounds__(PART_THREADS) part_hist(KeyFn key, const int* count_ptr, int count_imm, int nbins, int* keys, int* block_hist) {
    extern __shared__ int sh[];
    const int n = count_ptr ? *count_ptr : count_imm;
    for (int b = threadIdx.x; b < nbins; b += PART_THREADS) sh[b] = 0;
    __syncthreads();
    const int base = blockIdx.x * PART_CHUNK;
    for (int k = threadIdx.x; k < PART_CHUNK; k += PART_THREADS) {
        const int i = base + k;
        if (i < n) {
            const int kk = key(i);
            keys[i] = kk;
            // warp-aggregated shared atomic: one add per distinct key in the warp
            const unsigned peers = __match_any_sync(__activemask(), kk);
            if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&sh[kk], __popc(peers));
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += PART_THREADS) block_hist[b * gridDim.x + blockIdx.x] = sh[b];
}

/* Single block: exclusive scan of block_hist (bin-major); bin_start[b] = first position of bin b, bin_start[nbins] = n. */
__global__ void __launch_bounds__(1024) part_scan(int nblocks, int nbins, int* block_hist, int* bin_start) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int total = nblocks * nbins;
    for (int base = 0; base < total; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < total ? block_hist[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_sums[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, w, o); if (threadIdx.x >= o) w += y; }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        const int excl = carry + (threadIdx.x >= 32 ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (i < total) {
            block_hist[i] = excl;
            if (i % nblocks == 0) bin_start[i / nblocks] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_start[nbins] = carry;
}

__global__ void __launch_bounds__(PART_THREADS) part_scatter(const int* keys, const int* count_ptr, int count_imm, int nbins, const int* block_off, int* out_idx) {
    extern __shared__ int bin_base[];
    const int n = count_ptr ? *count_ptr : count_imm;
    for (int b = threadIdx.x; b < nbins; b += PART_THREADS) bin_base[b] = block_off[b * gridDim.x + blockIdx.x];
    __syncthreads();
    const int base = blockIdx.x * PART_CHUNK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (base >= n) return;
    for (int k = 0; k < PART_CHUNK; k += PART_THREADS) {
        const int i = base + k + threadIdx.x;
        const bool valid = i < n;
        const int kk = valid ? keys[i] : -1;
        const unsigned peers = __match_any_sync(FULL, kk);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        int pos = 0;
        // warps take their turn in order, so positions grow with the element index (stability)
        for (int w = 0; w < PART_THREADS / 32; ++w) {
            if (warp == w && valid && lane == leader) { pos = bin_base[kk]; bin_base[kk] = pos + __popc(peers); }
            __syncthreads();
        }
        pos = __shfl_sync(FULL, pos, leader);
        if (valid) out_idx[pos + rank] = i;
    }
}

/* Same scatter without the per-warp turn taking: positions inside a (block, bin) range are handed out by shared-memory
 * atomics (one per distinct key per warp).  Grouping is identical, the order inside a group depends on warp timing --
 * what the reference's own scatter does (ray_queue.h:235-243); used by the integrators, where only the grouping matters. */
__global__ void __launch_bounds__(PART_THREADS) part_scatter_fast(const int* keys, const int* count_ptr, int count_imm, int nbins, const int* block_off, int* out_idx) {
    extern __shared__ int bin_base[];
    const int n = count_ptr ? *count_ptr : count_imm;
    const int base = blockIdx.x * PART_CHUNK;
    if (base >= n) return;
    for (int b = threadIdx.x; b < nbins; b += PART_THREADS) bin_base[b] = block_off[b * gridDim.x + blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < PART_CHUNK; k += PART_THREADS) {
        const int i = base + k;
        if (i < n) {
            const int kk = keys[i];
            const unsigned peers = __match_any_sync(__activemask(), kk);
            const int leader = __ffs(peers) - 1;
            int pos = 0;
            if (lane == leader) pos = atomicAdd(&bin_base[kk], __popc(peers));
            pos = __shfl_sync(peers, pos, leader);
            out_idx[pos + __popc(peers & ((1u << lane) - 1u))] = i;
        }
    }
}

/* compact_rays: out[j] = in[idx[j]] for j < *kept; 16-byte vector copies of the 32-byte ray, 8-byte copies of the state */
__global__ void gather_rays_states(const float4* rays, const char* states, int state_bytes, const int* idx, const int* kept_ptr,
                                   float4* rays_out, char* states_out, int* count_out) {
    const int kept = *kept_ptr;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0 && count_out) *count_out = kept;
    if (j >= kept) return;
    const int i = idx[j];
    rays_out[2 * j] = rays[2 * i]; rays_out[2 * j + 1] = rays[2 * i + 1];
    const uint2* s = reinterpret_cast<const uint2*>(states + (size_t)i * state_bytes);
    uint2* d = reinterpret_cast<uint2*>(states_out + (size_t)j * state_bytes);
    for (int k = 0; k < state_bytes / 8; ++k) d[k] = s[k];
}

__global__ void gather_queue(const float4* rays, const int4* hits, const char* states, int state_bytes, const int* idx, const int* n_ptr, int n_imm,
                             float4* rays_out, int4* hits_out, char* states_out) {
    const int n = n_ptr ? *n_ptr : n_imm;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int i = idx[j];
    rays_out[2 * j] = rays[2 * i]; rays_out[2 * j + 1] = rays[2 * i + 1];
    if (hits_out) hits_out[j] = hits[i];
    const uint2* s = reinterpret_cast<const uint2*>(states + (size_t)i * state_bytes);
    uint2* d = reinterpret_cast<uint2*>(states_out + (size_t)j * state_bytes);
    for (int k = 0; k < state_bytes / 8; ++k) d[k] = s[k];
}

template <class KeyFn>
cudaError_t run_partition(const PartitionBuffers& b, KeyFn key, const int* count_ptr, int count_imm, int capacity, int nbins, cudaStream_t st, bool stable = true) {
    if (capacity <= 0) return cudaSuccess;
    const int nblocks = (capacity + PART_CHUNK - 1) / PART_CHUNK;
    if (nblocks > b.max_blocks || nbins > b.max_bins) return cudaErrorInvalidValue;
    part_hist<KeyFn><<<nblocks, PART_THREADS, nbins * sizeof(int), st>>>(key, count_ptr, count_imm, nbins, b.keys, b.block_hist);
    part_scan<<<1, 1024, 0, st>>>(nblocks, nbins, b.block_hist, b.bin_start);
    if (stable) part_scatter<<<nblocks, PART_THREADS, nbins * sizeof(int), st>>>(b.keys, count_ptr, count_imm, nbins, b.block_hist, b.indices);
    else        part_scatter_fast<<<nblocks, PART_THREADS, nbins * sizeof(int), st>>>(b.keys, count_ptr, count_imm, nbins, b.block_hist, b.indices);
    return cudaGetLastError();
}

} // namespace

cudaError_t imba_partition_alloc(PartitionBuffers& b, int capacity, int max_bins) {
    b.max_blocks = (capacity + PART_CHUNK - 1) / PART_CHUNK; b.max_bins = max_bins;
    cudaError_t e;
    if ((e = cudaMalloc((void**)&b.keys, sizeof(int) * (size_t)std::max(capacity, 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&b.indices, sizeof(int) * (size_t)std::max(capacity, 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&b.block_hist, sizeof(int) * (size_t)std::max(b.max_blocks * max_bins, 1))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&b.bin_start, sizeof(int) * (size_t)(max_bins + 1))) != cudaSuccess) return e;
    return cudaSuccess;
}

void imba_partition_free(PartitionBuffers& b) {
    cudaFree(b.keys); cudaFree(b.indices); cudaFree(b.block_hist); cudaFree(b.bin_start);
    b = PartitionBuffers();
}

cudaError_t imba_sort_hits_by_material(const PartitionBuffers& b, const int4* hits, const int4* indices, int nmat, int* shadow_pixel, int shadow_slots,
                                       const int* count_ptr, int count_imm, int capacity, cudaStream_t st, bool stable) {
    return run_partition(b, HitMaterialKey{ hits, indices, nmat, shadow_pixel, shadow_slots }, count_ptr, count_imm, capacity, nmat + 1, st, stable);
}

cudaError_t imba_compact_alive(const PartitionBuffers& b, const int4* hits, const float4* rays, const void* states, int state_bytes,
                               const int* count_ptr, int count_imm, int capacity, float4* rays_out, void* states_out, int* count_out, cudaStream_t st) {
    cudaError_t e;
    if (hits) e = run_partition(b, AliveKey{ hits, (const char*)states, state_bytes }, count_ptr, count_imm, capacity, 2, st);
    else      e = run_partition(b, PixelAliveKey{ (const char*)states, state_bytes }, count_ptr, count_imm, capacity, 2, st);
    if (e != cudaSuccess || capacity <= 0) return e;
    gather_rays_states<<<(capacity + 255) / 256, 256, 0, st>>>(rays, (const char*)states, state_bytes, b.indices, b.bin_start + 1, rays_out, (char*)states_out, count_out);
    return cudaGetLastError();
}

cudaError_t imba_gather_queue(const int* idx, const int* n_ptr, int n_imm, int capacity, const float4* rays, const int4* hits, const void* states,
                              int state_bytes, float4* rays_out, int4* hits_out, void* states_out, cudaStream_t st) {
    if (capacity <= 0) return cudaSuccess;
    gather_queue<<<(capacity + 255) / 256, 256, 0, st>>>(rays, hits, (const char*)states, state_bytes, idx, n_ptr, n_imm, rays_out, hits_out, (char*)states_out);
    return cudaGetLastError();
}


/* ---- the same primitives on HOST arrays: RayQueue::{compact_hits, compact_rays, sort_by_material} keep their queues in host
 * memory (ray_queue.h:161-244); these entry points stage the arrays through device memory, run the kernels above and copy the
 * result back in place.  Used by the C++ mirror of RayQueue (include/imba_b200.hpp); the device integrators never come here. */

namespace {
struct DeviceBuf {
    void* p = nullptr;
    ~DeviceBuf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
};
}

extern "C" int imba_queue_compact_hits_host(imba_ray* h_rays, imba_hit* h_hits, void* h_states, int state_bytes, int count, int* hit_count) {
    if (count < 0 || state_bytes < 8 || state_bytes % 8 || !hit_count || (count > 0 && (!h_rays || !h_hits || !h_states))) { imba_set_error("invalid argument"); return IMBA_ERR_INVALID; }
    *hit_count = 0;
    if (count == 0) return IMBA_OK;
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { imba_set_error("no CUDA device (this library has no CPU path)"); return IMBA_ERR_NO_DEVICE; }
    DeviceBuf rays, hits, states, rays_o, hits_o, states_o;
    const size_t rb = (size_t)count * 32, hb = (size_t)count * 16, sb = (size_t)count * state_bytes;
    IMBA_CUDA(rays.alloc(rb), IMBA_ERR_CUDA); IMBA_CUDA(hits.alloc(hb), IMBA_ERR_CUDA); IMBA_CUDA(states.alloc(sb), IMBA_ERR_CUDA);
    IMBA_CUDA(rays_o.alloc(rb), IMBA_ERR_CUDA); IMBA_CUDA(hits_o.alloc(hb), IMBA_ERR_CUDA); IMBA_CUDA(states_o.alloc(sb), IMBA_ERR_CUDA);
    cudaStream_t st = cudaStreamPerThread;
    IMBA_CUDA(cudaMemcpyAsync(rays.p, h_rays, rb, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(hits.p, h_hits, hb, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(states.p, h_states, sb, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    PartitionBuffers b;
    IMBA_CUDA(imba_partition_alloc(b, count, 2), IMBA_ERR_CUDA);
    cudaError_t e = run_partition(b, HitKey{ (const int4*)hits.p }, nullptr, count, count, 2, st);
    if (e == cudaSuccess) e = imba_gather_queue(b.indices, nullptr, count, count, (const float4*)rays.p, (const int4*)hits.p, states.p, state_bytes,
                                                (float4*)rays_o.p, (int4*)hits_o.p, states_o.p, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hit_count, b.bin_start + 1, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_rays, rays_o.p, rb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_hits, hits_o.p, hb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_states, states_o.p, sb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    imba_partition_free(b);
    g_imba_launches.fetch_add(4);
    IMBA_CUDA(e, IMBA_ERR_CUDA);
    return IMBA_OK;
}

extern "C" int imba_queue_compact_rays_host(imba_ray* h_rays, void* h_states, int state_bytes, int count, int* kept) {
    if (count < 0 || state_bytes < 8 || state_bytes % 8 || !kept || (count > 0 && (!h_rays || !h_states))) { imba_set_error("invalid argument"); return IMBA_ERR_INVALID; }
    *kept = 0;
    if (count == 0) return IMBA_OK;
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { imba_set_error("no CUDA device (this library has no CPU path)"); return IMBA_ERR_NO_DEVICE; }
    DeviceBuf rays, states, rays_o, states_o, cnt;
    const size_t rb = (size_t)count * 32, sb = (size_t)count * state_bytes;
    IMBA_CUDA(rays.alloc(rb), IMBA_ERR_CUDA); IMBA_CUDA(states.alloc(sb), IMBA_ERR_CUDA); IMBA_CUDA(rays_o.alloc(rb), IMBA_ERR_CUDA);
    IMBA_CUDA(states_o.alloc(sb), IMBA_ERR_CUDA); IMBA_CUDA(cnt.alloc(sizeof(int)), IMBA_ERR_CUDA);
    cudaStream_t st = cudaStreamPerThread;
    IMBA_CUDA(cudaMemcpyAsync(rays.p, h_rays, rb, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    IMBA_CUDA(cudaMemcpyAsync(states.p, h_states, sb, cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    PartitionBuffers b;
    IMBA_CUDA(imba_partition_alloc(b, count, 2), IMBA_ERR_CUDA);
    cudaError_t e = imba_compact_alive(b, nullptr, (const float4*)rays.p, states.p, state_bytes, nullptr, count, count, (float4*)rays_o.p, states_o.p, (int*)cnt.p, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(kept, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && *kept > 0) {
        e = cudaMemcpyAsync(h_rays, rays_o.p, (size_t)*kept * 32, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_states, states_o.p, (size_t)*kept * state_bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    imba_partition_free(b);
    g_imba_launches.fetch_add(4);
    IMBA_CUDA(e, IMBA_ERR_CUDA);
    return IMBA_OK;
}

extern "C" int imba_queue_sort_keys_host(const int* h_keys, int count, int num_bins, int* h_sorted_indices) {
    if (count < 0 || num_bins <= 0 || (count > 0 && (!h_keys || !h_sorted_indices))) { imba_set_error("invalid argument"); return IMBA_ERR_INVALID; }
    if (count == 0) return IMBA_OK;
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { imba_set_error("no CUDA device (this library has no CPU path)"); return IMBA_ERR_NO_DEVICE; }
    DeviceBuf keys;
    IMBA_CUDA(keys.alloc((size_t)count * sizeof(int)), IMBA_ERR_CUDA);
    cudaStream_t st = cudaStreamPerThread;
    IMBA_CUDA(cudaMemcpyAsync(keys.p, h_keys, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, st), IMBA_ERR_CUDA);
    PartitionBuffers b;
    IMBA_CUDA(imba_partition_alloc(b, count, num_bins), IMBA_ERR_CUDA);
    cudaError_t e = run_partition(b, ArrayKey{ (const int*)keys.p, num_bins }, nullptr, count, count, num_bins, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_sorted_indices, b.indices, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    imba_partition_free(b);
    g_imba_launches.fetch_add(3);
    IMBA_CUDA(e, IMBA_ERR_CUDA);
    return IMBA_OK;
}
