/* lbvh.cuh -- device construction of the binary tree of one mesh (Karras 2012): Morton codes of the triangle centroids, one
 * radix sort, the parallel hierarchy and a bottom-up refit.  Shared by lbvh.cu (which emits the tree in the reference's BVH2
 * layout) and bvh8_gpu.cu (which collapses it into 8-wide nodes on the device).  Node ids: 0 .. n-2 internal (0 = root),
 * n-1 .. 2n-2 leaves (leaf n-1+p holds the triangle at sorted position p). */
#ifndef IMBA_LBVH_CUH
#define IMBA_LBVH_CUH
#include <cfloat>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include "api_common.h"

namespace {

__device__ __forceinline__ int ordered_int(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float from_ordered(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct MeshView { const float4* vertices; const int4* indices; int vbase, tbase, nt; };

__device__ __forceinline__ void tri_verts(const MeshView& m, int t, float4& a, float4& b, float4& c) {
    const int4 idx = m.indices[m.tbase + t];
    a = m.vertices[m.vbase + idx.x]; b = m.vertices[m.vbase + idx.y]; c = m.vertices[m.vbase + idx.z];
}

__global__ void lbvh_bounds_kernel(MeshView m, int* bbox /* 6 ordered ints: centroid bounds */) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    if (t < m.nt) {
        float4 a, b, c; tri_verts(m, t, a, b, c);
        const float cx = (a.x + b.x + c.x) * (1.0f / 3.0f), cy = (a.y + b.y + c.y) * (1.0f / 3.0f), cz = (a.z + b.z + c.z) * (1.0f / 3.0f);
        lo[0] = hi[0] = cx; lo[1] = hi[1] = cy; lo[2] = hi[2] = cz;
    }
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(bbox + k, ordered_int(lo[k])); atomicMax(bbox + 3 + k, ordered_int(hi[k])); }
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) { // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void lbvh_morton_kernel(MeshView m, const int* bbox, unsigned long long* keys, int* ids) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m.nt) return;
    float4 a, b, c; tri_verts(m, t, a, b, c);
    const float ctr[3] = { (a.x + b.x + c.x) * (1.0f / 3.0f), (a.y + b.y + c.y) * (1.0f / 3.0f), (a.z + b.z + c.z) * (1.0f / 3.0f) };
    unsigned long long code = 0;
    for (int k = 0; k < 3; ++k) {
        const float lo = from_ordered(bbox[k]), hi = from_ordered(bbox[3 + k]);
        const float ext = hi - lo;
        float f = ext > 0.0f ? (ctr[k] - lo) / ext : 0.0f;
        f = fminf(fmaxf(f, 0.0f), 1.0f);
        const unsigned long long q = (unsigned long long)fminf(f * 2097152.0f, 2097151.0f);
        code |= spread21(q) << (2 - k);
    }
    keys[t] = code; ids[t] = t;
}

/* Common-prefix length of the (key, position) pairs i and j; -1 outside the array (Karras 2012, section 4). */
__device__ __forceinline__ int delta(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll(a ^ b);
}

__global__ void lbvh_hierarchy_kernel(const unsigned long long* keys, int n, int2* children, int* parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do { t = (t + 1) / 2; if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t; } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int left = min(i, j) == gamma ? (n - 1) + gamma : gamma;          // ids >= n-1 are leaves
    const int right = max(i, j) == gamma + 1 ? (n - 1) + gamma + 1 : gamma + 1;
    children[i] = make_int2(left, right);
    parent[left] = i; parent[right] = i;
    if (i == 0) parent[0] = -1;
}

struct Box6 { float v[6]; };

__global__ void lbvh_refit_kernel(MeshView m, const int* ids, int n, const int2* children, const int* parent, int* flags, Box6* boxes) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float4 a, b, c; tri_verts(m, ids[p], a, b, c);
    Box6 bx;
    bx.v[0] = fminf(a.x, fminf(b.x, c.x)); bx.v[1] = fminf(a.y, fminf(b.y, c.y)); bx.v[2] = fminf(a.z, fminf(b.z, c.z));
    bx.v[3] = fmaxf(a.x, fmaxf(b.x, c.x)); bx.v[4] = fmaxf(a.y, fmaxf(b.y, c.y)); bx.v[5] = fmaxf(a.z, fmaxf(b.z, c.z));
    boxes[(n - 1) + p] = bx;
    __threadfence();
    int node = parent[(n - 1) + p];
    while (node >= 0) {
        if (atomicAdd(flags + node, 1) == 0) return; // the sibling subtree is not finished: its thread will continue
        __threadfence();
        const int2 ch = children[node];
        const Box6 l = boxes[ch.x], r = boxes[ch.y];
        Box6 u;
        for (int k = 0; k < 3; ++k) { u.v[k] = fminf(l.v[k], r.v[k]); u.v[3 + k] = fmaxf(l.v[3 + k], r.v[3 + k]); }
        boxes[node] = u;
        __threadfence();
        node = parent[node];
    }
}

struct DeviceBuf {
    void* p = nullptr;
    ~DeviceBuf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename T> T* as() { return (T*)p; }
};


struct LbvhTree {
    int n = 0;
    int root = 0;   // id of the root among the internal nodes (0 for the Karras hierarchy; the last merge for PLOC)
    DeviceBuf bbox, keys, keys2, ids, ids2, children, parent, flags, boxes, temp;
    const int*  sorted_ids() { return ids2.as<int>(); }
    const int2* kids() { return children.as<int2>(); }
    const int*  parents() { return parent.as<int>(); }
    const Box6* bounds() { return boxes.as<Box6>(); }
};

/* Allocates the buffers and enqueues the build of mesh `m` on the default stream (5 launches incl. the sort, 3 when n == 1). */
inline cudaError_t lbvh_build_tree(const MeshView& m, LbvhTree& t, cudaEvent_t after_alloc = nullptr) {
    const int n = m.nt; t.n = n;
    cudaError_t e;
#define LBT(x) do { if ((e = (x)) != cudaSuccess) return e; } while (0)
    LBT(t.bbox.alloc(32)); LBT(t.keys.alloc((size_t)n * 8)); LBT(t.keys2.alloc((size_t)n * 8)); LBT(t.ids.alloc((size_t)n * 4)); LBT(t.ids2.alloc((size_t)n * 4));
    LBT(t.children.alloc((size_t)n * 8)); LBT(t.parent.alloc((size_t)(2 * n) * 4)); LBT(t.flags.alloc((size_t)n * 4)); LBT(t.boxes.alloc((size_t)(2 * n) * sizeof(Box6)));
    size_t temp_bytes = 0;
    LBT(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, t.keys.as<unsigned long long>(), t.keys2.as<unsigned long long>(), t.ids.as<int>(), t.ids2.as<int>(), n, 0, 63));
    LBT(t.temp.alloc(temp_bytes));
    static const int init_bbox[6] = { 0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000u, (int)0x80800000u, (int)0x80800000u };
    LBT(cudaMemcpy(t.bbox.p, init_bbox, sizeof(init_bbox), cudaMemcpyHostToDevice));
    LBT(cudaMemset(t.flags.p, 0, (size_t)n * 4));
    const int blocks = (n + 255) / 256;
    if (after_alloc) cudaEventRecord(after_alloc);   // device time is measured from here: cudaMalloc is host work
    lbvh_bounds_kernel<<<blocks, 256>>>(m, t.bbox.as<int>());
    lbvh_morton_kernel<<<blocks, 256>>>(m, t.bbox.as<int>(), t.keys.as<unsigned long long>(), t.ids.as<int>());
    LBT(cub::DeviceRadixSort::SortPairs(t.temp.p, temp_bytes, t.keys.as<unsigned long long>(), t.keys2.as<unsigned long long>(), t.ids.as<int>(), t.ids2.as<int>(), n, 0, 63));
    if (n > 1) lbvh_hierarchy_kernel<<<blocks, 256>>>(t.keys2.as<unsigned long long>(), n, t.children.as<int2>(), t.parent.as<int>());
    // n == 1: the refit kernel still writes the leaf box (parent[0] must read as "no parent")
    else LBT(cudaMemset(t.parent.p, 0xff, 8));
    lbvh_refit_kernel<<<blocks, 256>>>(m, t.ids2.as<int>(), n, t.children.as<int2>(), t.parent.as<int>(), t.flags.as<int>(), t.boxes.as<Box6>());
#undef LBT
    g_imba_launches.fetch_add(n > 1 ? 5 : 4);
    return cudaGetLastError();
}


/* ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) ------------------------------------------------------
 * Bottom-up agglomeration over the Morton-sorted clusters: every cluster looks for the neighbour within PLOC_RADIUS positions
 * whose union box has the smallest area; mutual nearest neighbours merge; the cluster array is compacted; repeat until one
 * cluster is left.  Same outputs as the Karras hierarchy (children / parent / boxes, leaves n-1+p), trees close to a
 * binned-SAH build instead of a pure Morton split. */
constexpr int PLOC_RADIUS = 16;

__global__ void ploc_leaf_kernel(MeshView m, const int* ids, int n, Box6* boxes, int* cid) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float4 a, b, c; tri_verts(m, ids[p], a, b, c);
    Box6 bx;
    bx.v[0] = fminf(a.x, fminf(b.x, c.x)); bx.v[1] = fminf(a.y, fminf(b.y, c.y)); bx.v[2] = fminf(a.z, fminf(b.z, c.z));
    bx.v[3] = fmaxf(a.x, fmaxf(b.x, c.x)); bx.v[4] = fmaxf(a.y, fmaxf(b.y, c.y)); bx.v[5] = fmaxf(a.z, fmaxf(b.z, c.z));
    boxes[(n - 1) + p] = bx; cid[p] = (n - 1) + p;
}

__global__ void ploc_nearest_kernel(const int* cid, int count, const Box6* boxes, int* nearest) {
    __shared__ Box6 tile[256 + 2 * PLOC_RADIUS];
    const int base = blockIdx.x * blockDim.x - PLOC_RADIUS;
    for (int k = threadIdx.x; k < 256 + 2 * PLOC_RADIUS; k += blockDim.x) {
        const int j = base + k;
        if (j >= 0 && j < count) tile[k] = boxes[cid[j]];
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Box6 me = tile[threadIdx.x + PLOC_RADIUS];
    float best = FLT_MAX; int bj = -1;
    // candidates by increasing index distance; at equal distance even positions look right first, odd ones left first; a
    // candidate wins only when strictly better.  On ties (duplicated geometry: all distances equal) this pairs (0,1), (2,3), ...
    // instead of one pair per round, and a mutual pair always exists.
    for (int k = 1; k <= PLOC_RADIUS; ++k) {
        for (int side = 0; side < 2; ++side) {
            const int o = ((i & 1) == side) ? k : -k;
            const int j = i + o;
            if (j < 0 || j >= count) continue;
            const Box6& ot = tile[threadIdx.x + PLOC_RADIUS + o];
            const float dx = fmaxf(me.v[3], ot.v[3]) - fminf(me.v[0], ot.v[0]), dy = fmaxf(me.v[4], ot.v[4]) - fminf(me.v[1], ot.v[1]),
                        dz = fmaxf(me.v[5], ot.v[5]) - fminf(me.v[2], ot.v[2]);
            const float d = dx * dy + dy * dz + dz * dx;
            if (d < best) { best = d; bj = j; }
        }
    }
    if (bj < 0) bj = i > 0 ? i - 1 : 1;   // NaN boxes: still pair up
    nearest[i] = bj;
}

__global__ void ploc_merge_kernel(const int* cid, int count, const int* nearest, Box6* boxes, int2* children, int* parent, int* next_id, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int j = nearest[i];
    if (j < 0 || nearest[j] != i) { out[i] = cid[i]; return; }
    if (i > j) { out[i] = -1; return; }
    const int id = atomicAdd(next_id, 1);
    const int a = cid[i], b = cid[j];
    const Box6 ba = boxes[a], bb = boxes[b];
    Box6 u;
    for (int k = 0; k < 3; ++k) { u.v[k] = fminf(ba.v[k], bb.v[k]); u.v[3 + k] = fmaxf(ba.v[3 + k], bb.v[3 + k]); }
    boxes[id] = u; children[id] = make_int2(a, b); parent[a] = id; parent[b] = id;
    out[i] = id;
}

struct PlocValid { __device__ __forceinline__ bool operator()(const int& x) const { return x >= 0; } };

/* Same contract as lbvh_build_tree; t.root is the id of the last merge. */
inline cudaError_t ploc_build_tree(const MeshView& m, LbvhTree& t, cudaEvent_t after_alloc = nullptr) {
    const int n = m.nt; t.n = n;
    cudaError_t e;
#define LBT(x) do { if ((e = (x)) != cudaSuccess) return e; } while (0)
    LBT(t.bbox.alloc(32)); LBT(t.keys.alloc((size_t)n * 8)); LBT(t.keys2.alloc((size_t)n * 8)); LBT(t.ids.alloc((size_t)n * 4)); LBT(t.ids2.alloc((size_t)n * 4));
    LBT(t.children.alloc((size_t)n * 8)); LBT(t.parent.alloc((size_t)(2 * n) * 4)); LBT(t.flags.alloc((size_t)n * 4)); LBT(t.boxes.alloc((size_t)(2 * n) * sizeof(Box6)));
    DeviceBuf cidA, cidB, nearest, counters;
    LBT(cidA.alloc((size_t)n * 4)); LBT(cidB.alloc((size_t)n * 4)); LBT(nearest.alloc((size_t)n * 4)); LBT(counters.alloc(16));
    size_t sort_bytes = 0, select_bytes = 0;
    LBT(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, t.keys.as<unsigned long long>(), t.keys2.as<unsigned long long>(), t.ids.as<int>(), t.ids2.as<int>(), n, 0, 63));
    LBT(cub::DeviceSelect::If(nullptr, select_bytes, cidB.as<int>(), cidA.as<int>(), counters.as<int>() + 1, n, PlocValid()));
    LBT(t.temp.alloc(sort_bytes > select_bytes ? sort_bytes : select_bytes));
    size_t temp_bytes = sort_bytes > select_bytes ? sort_bytes : select_bytes;
    static const int init_bbox[6] = { 0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000u, (int)0x80800000u, (int)0x80800000u };
    LBT(cudaMemcpy(t.bbox.p, init_bbox, sizeof(init_bbox), cudaMemcpyHostToDevice));
    LBT(cudaMemset(counters.p, 0, 16));
    LBT(cudaMemset(t.parent.p, 0xff, (size_t)(2 * n) * 4));
    const int blocks = (n + 255) / 256;
    if (after_alloc) cudaEventRecord(after_alloc);   // device time is measured from here: cudaMalloc is host work
    lbvh_bounds_kernel<<<blocks, 256>>>(m, t.bbox.as<int>());
    lbvh_morton_kernel<<<blocks, 256>>>(m, t.bbox.as<int>(), t.keys.as<unsigned long long>(), t.ids.as<int>());
    LBT(cub::DeviceRadixSort::SortPairs(t.temp.p, temp_bytes, t.keys.as<unsigned long long>(), t.keys2.as<unsigned long long>(), t.ids.as<int>(), t.ids2.as<int>(), n, 0, 63));
    ploc_leaf_kernel<<<blocks, 256>>>(m, t.ids2.as<int>(), n, t.boxes.as<Box6>(), cidA.as<int>());
    int launches = 4, count = n, rounds = 0;
    while (count > 1) {
        if (++rounds > 4096) return cudaErrorUnknown;
        const int b = (count + 255) / 256;
        ploc_nearest_kernel<<<b, 256>>>(cidA.as<int>(), count, t.boxes.as<Box6>(), nearest.as<int>());
        ploc_merge_kernel<<<b, 256>>>(cidA.as<int>(), count, nearest.as<int>(), t.boxes.as<Box6>(), t.children.as<int2>(), t.parent.as<int>(), counters.as<int>(), cidB.as<int>());
        LBT(cub::DeviceSelect::If(t.temp.p, temp_bytes, cidB.as<int>(), cidA.as<int>(), counters.as<int>() + 1, count, PlocValid()));
        int c2[2];
        LBT(cudaMemcpy(c2, counters.p, 8, cudaMemcpyDeviceToHost));
        if (c2[1] >= count || c2[1] < 1) return cudaErrorUnknown;   // every round merges at least one pair
        count = c2[1]; launches += 3;
        if (count == 1) t.root = c2[0] - 1;
    }
    if (n == 1) t.root = 0;
#undef LBT
    g_imba_launches.fetch_add(launches);
    return cudaGetLastError();
}

} // namespace
#endif
