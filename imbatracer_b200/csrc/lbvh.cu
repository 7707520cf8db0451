/* lbvh.cu -- GPU builder of the per-mesh trees (row f1: replaces the single-threaded host SplitBvhBuilder, core/sbvh_builder.h,
 * as the start-up bottleneck of large scenes: 247 ms for 7 k triangles there).
 *
 * Linear BVH after Karras 2012: 63-bit Morton codes of the triangle centroids (21 bits per axis, made unique by the
 * triangle index), one radix sort (cub::DeviceRadixSort -- a library call, used for this one-off step only), the
 * fully parallel hierarchy construction (one thread per internal node), and a bottom-up refit with one atomic flag per
 * internal node.  The result is EMITTED IN THE REFERENCE'S BVH2 LAYOUT (core/gpu_adapter.cpp:62-143: two child boxes and
 * two child references per node, leaves = ~offset of (v0, v1|id, v2|sentinel) Vec4 triples), one triangle per leaf; it then
 * takes the normal path through the converter (bvh8_build.cpp), whose SAH collapse merges the single-triangle leaves into
 * leaves of up to three and builds the 8-wide nodes.  Top level: host (accel_build.cpp), it has few primitives.
 * Quality: LBVH trees are built without SAH; on the terrain scene traversal is slower than on the host binned-SAH tree
 * (numbers in DESIGN.md), the build is two orders of magnitude faster.
 */
#include <cfloat>
#include <cub/device/device_radix_sort.cuh>
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

__global__ void lbvh_emit_kernel(MeshView m, const int* ids, int n, const int2* children, const Box6* boxes, int node_base, int vec4_base,
                                 imba_node2* nodes, float4* tris) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { // leaf i (sorted position): one triangle, terminated by the sentinel
        float4 a, b, c; tri_verts(m, ids[i], a, b, c);
        a.w = 0.0f; b.w = __int_as_float(m.tbase + ids[i]); c.w = __int_as_float((int)0x80000000);
        tris[3 * i] = a; tris[3 * i + 1] = b; tris[3 * i + 2] = c;
    }
    if (i < n - 1) {
        const int2 ch = children[i];
        imba_node2 nd;
        const Box6 l = boxes[ch.x], r = boxes[ch.y];
        nd.left_bb.lo_x = l.v[0]; nd.left_bb.lo_y = l.v[1]; nd.left_bb.lo_z = l.v[2]; nd.left_bb.hi_x = l.v[3]; nd.left_bb.hi_y = l.v[4]; nd.left_bb.hi_z = l.v[5];
        nd.right_bb.lo_x = r.v[0]; nd.right_bb.lo_y = r.v[1]; nd.right_bb.lo_z = r.v[2]; nd.right_bb.hi_x = r.v[3]; nd.right_bb.hi_y = r.v[4]; nd.right_bb.hi_z = r.v[5];
        nd.left = ch.x >= n - 1 ? ~(vec4_base + 3 * (ch.x - (n - 1))) : node_base + ch.x;
        nd.right = ch.y >= n - 1 ? ~(vec4_base + 3 * (ch.y - (n - 1))) : node_base + ch.y;
        nodes[i] = nd;
    }
}

struct DeviceBuf {
    void* p = nullptr;
    ~DeviceBuf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename T> T* as() { return (T*)p; }
};

} // namespace

#define LB_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { imba_set_error("%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); delete out; return nullptr; } } while (0)

/* Like imba_accel_build, with the per-mesh trees built on the current CUDA device.  out_seconds (optional, 2 floats):
 * device time of the mesh builds (kernels + sort) and host time of the whole call. */
extern "C" imba_accel* imba_accel_build_gpu(const imba_scene_desc* d, float* out_seconds) {
    if (!d || d->mesh_count <= 0 || !d->mesh_vertex_base || !d->mesh_tri_base || !d->vertices || !d->indices || d->instance_count <= 0 || !d->instances) {
        imba_set_error("imba_accel_build_gpu needs meshes (vertices, indices, bases) and instances"); return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { imba_set_error("no CUDA device (this library has no CPU path)"); return nullptr; }
    imba_accel* out = new imba_accel;
    const int nv = d->mesh_vertex_base[d->mesh_count], nt_all = d->mesh_tri_base[d->mesh_count];
    DeviceBuf d_vertices, d_indices;
    LB_CUDA(d_vertices.alloc((size_t)nv * 16)); LB_CUDA(d_indices.alloc((size_t)nt_all * 16));
    LB_CUDA(cudaMemcpy(d_vertices.p, d->vertices, (size_t)nv * 16, cudaMemcpyHostToDevice));
    LB_CUDA(cudaMemcpy(d_indices.p, d->indices, (size_t)nt_all * 16, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float device_ms = 0.0f;
    std::vector<int> mesh_root(d->mesh_count); std::vector<float> mesh_box(6 * (size_t)d->mesh_count);
    for (int mi = 0; mi < d->mesh_count; ++mi) {
        MeshView m{ d_vertices.as<float4>(), d_indices.as<int4>(), d->mesh_vertex_base[mi], d->mesh_tri_base[mi], d->mesh_tri_base[mi + 1] - d->mesh_tri_base[mi] };
        const int n = m.nt;
        if (n <= 0) { imba_set_error("mesh %d has no triangles", mi); delete out; return nullptr; }
        const int node_base = (int)out->nodes.size(), vec4_base = (int)out->tris.size();
        DeviceBuf bbox, keys, keys2, ids, ids2, children, parent, flags, boxes, d_nodes, d_tris, temp;
        LB_CUDA(bbox.alloc(32)); LB_CUDA(keys.alloc((size_t)n * 8)); LB_CUDA(keys2.alloc((size_t)n * 8)); LB_CUDA(ids.alloc((size_t)n * 4)); LB_CUDA(ids2.alloc((size_t)n * 4));
        LB_CUDA(children.alloc((size_t)n * 8)); LB_CUDA(parent.alloc((size_t)(2 * n) * 4)); LB_CUDA(flags.alloc((size_t)n * 4)); LB_CUDA(boxes.alloc((size_t)(2 * n) * sizeof(Box6)));
        LB_CUDA(d_nodes.alloc((size_t)std::max(n - 1, 1) * sizeof(imba_node2))); LB_CUDA(d_tris.alloc((size_t)n * 48));
        size_t temp_bytes = 0;
        LB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), ids.as<int>(), ids2.as<int>(), n, 0, 63));
        LB_CUDA(temp.alloc(temp_bytes));
        static const int init_bbox[6] = { 0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000u, (int)0x80800000u, (int)0x80800000u };
        LB_CUDA(cudaMemcpy(bbox.p, init_bbox, sizeof(init_bbox), cudaMemcpyHostToDevice));
        LB_CUDA(cudaMemset(flags.p, 0, (size_t)n * 4));
        const int blocks = (n + 255) / 256;
        cudaEventRecord(e0);
        lbvh_bounds_kernel<<<blocks, 256>>>(m, bbox.as<int>());
        lbvh_morton_kernel<<<blocks, 256>>>(m, bbox.as<int>(), keys.as<unsigned long long>(), ids.as<int>());
        LB_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), ids.as<int>(), ids2.as<int>(), n, 0, 63));
        if (n > 1) {
            lbvh_hierarchy_kernel<<<blocks, 256>>>(keys2.as<unsigned long long>(), n, children.as<int2>(), parent.as<int>());
            lbvh_refit_kernel<<<blocks, 256>>>(m, ids2.as<int>(), n, children.as<int2>(), parent.as<int>(), flags.as<int>(), boxes.as<Box6>());
        }
        lbvh_emit_kernel<<<blocks, 256>>>(m, ids2.as<int>(), n, children.as<int2>(), boxes.as<Box6>(), node_base, vec4_base, d_nodes.as<imba_node2>(), d_tris.as<float4>());
        cudaEventRecord(e1);
        LB_CUDA(cudaGetLastError());
        LB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.0f; cudaEventElapsedTime(&ms, e0, e1); device_ms += ms;
        g_imba_launches.fetch_add(n > 1 ? 5 : 3);
        out->tris.resize((size_t)vec4_base + 3 * (size_t)n);
        LB_CUDA(cudaMemcpy(out->tris.data() + vec4_base, d_tris.p, (size_t)n * 48, cudaMemcpyDeviceToHost));
        if (n > 1) {
            out->nodes.resize((size_t)node_base + n - 1);
            LB_CUDA(cudaMemcpy(out->nodes.data() + node_base, d_nodes.p, (size_t)(n - 1) * sizeof(imba_node2), cudaMemcpyDeviceToHost));
            Box6 root;
            LB_CUDA(cudaMemcpy(&root, boxes.as<Box6>(), sizeof(Box6), cudaMemcpyDeviceToHost));
            memcpy(&mesh_box[6 * mi], root.v, 24);
        } else { // single triangle: dummy parent (gpu_adapter.cpp:14-31)
            imba_node2 nd; memset(&nd, 0, sizeof(nd));
            const imba_vec4* t = out->tris.data() + vec4_base;
            float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
            for (int k = 0; k < 3; ++k) { const float v[3] = { t[k].x, t[k].y, t[k].z }; for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], v[a]); hi[a] = std::max(hi[a], v[a]); } }
            nd.left_bb.lo_x = lo[0]; nd.left_bb.lo_y = lo[1]; nd.left_bb.lo_z = lo[2]; nd.left_bb.hi_x = hi[0]; nd.left_bb.hi_y = hi[1]; nd.left_bb.hi_z = hi[2];
            nd.right_bb.hi_x = nd.right_bb.hi_y = nd.right_bb.hi_z = -0.0f;
            nd.left = ~vec4_base; nd.right = 0x76543210;
            out->nodes.push_back(nd);
            for (int a = 0; a < 3; ++a) { mesh_box[6 * mi + a] = lo[a]; mesh_box[6 * mi + 3 + a] = hi[a]; }
        }
        mesh_root[mi] = node_base;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (imba_accel_finish_top(out, d, mesh_root.data(), mesh_box.data()) != IMBA_OK) { delete out; return nullptr; }
    if (out_seconds) out_seconds[0] = device_ms * 1e-3f;
    return out;
}
