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
#include "lbvh.cuh"

namespace {

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
    const bool use_ploc = getenv("IMBA_GPU_BUILDER") && !strcmp(getenv("IMBA_GPU_BUILDER"), "ploc"); // default here: the Karras hierarchy
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
        LbvhTree t; DeviceBuf d_nodes, d_tris;
        LB_CUDA(d_nodes.alloc((size_t)std::max(n - 1, 1) * sizeof(imba_node2))); LB_CUDA(d_tris.alloc((size_t)n * 48));
        const int blocks = (n + 255) / 256;
        LB_CUDA(use_ploc ? ploc_build_tree(m, t, e0) : lbvh_build_tree(m, t, e0));
        lbvh_emit_kernel<<<blocks, 256>>>(m, t.sorted_ids(), n, t.kids(), t.bounds(), node_base, vec4_base, d_nodes.as<imba_node2>(), d_tris.as<float4>());
        cudaEventRecord(e1);
        LB_CUDA(cudaGetLastError());
        LB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.0f; cudaEventElapsedTime(&ms, e0, e1); device_ms += ms;
        g_imba_launches.fetch_add(1);
        out->tris.resize((size_t)vec4_base + 3 * (size_t)n);
        LB_CUDA(cudaMemcpy(out->tris.data() + vec4_base, d_tris.p, (size_t)n * 48, cudaMemcpyDeviceToHost));
        if (n > 1) {
            out->nodes.resize((size_t)node_base + n - 1);
            LB_CUDA(cudaMemcpy(out->nodes.data() + node_base, d_nodes.p, (size_t)(n - 1) * sizeof(imba_node2), cudaMemcpyDeviceToHost));
            Box6 root;
            LB_CUDA(cudaMemcpy(&root, t.bounds() + t.root, sizeof(Box6), cudaMemcpyDeviceToHost));
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
        mesh_root[mi] = node_base + t.root;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (imba_accel_finish_top(out, d, mesh_root.data(), mesh_box.data()) != IMBA_OK) { delete out; return nullptr; }
    if (out_seconds) out_seconds[0] = device_ms * 1e-3f;
    return out;
}
