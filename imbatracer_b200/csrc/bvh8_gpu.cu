/* bvh8_gpu.cu -- device construction of the compressed 8-wide structure straight from triangle meshes (SURVEY.md 8f row f1:
 * "GPU BVH build (LBVH + 8-wide collapse)" replacing the host SplitBvhBuilder, core/sbvh_builder.h, and the host converter for
 * scenes that come without a reference-built structure).
 *
 *   1. binary tree per mesh on the device (lbvh.cuh): Morton codes and radix sort, then parallel locally-ordered clustering
 *      (PLOC, the default) or the Karras hierarchy + refit (IMBA_GPU_BUILDER=lbvh);
 *   2. collapse_cost_kernel: the SAH-optimal 8-wide collapse of Ylitie, Karras, Laine 2017 (section 3) as a bottom-up dynamic
 *      program -- the same recurrences as the host converter (bvh8_build.cpp: Collapse::run), one thread climbing from every
 *      leaf, the second thread to arrive at a node computes it (one atomic flag per node);
 *   3. emit_level_kernel, breadth first, one launch per level of the 8-wide tree, one thread per Node8: follows the stored
 *      decisions to the <= 8 children, assigns octant-ordered slots, quantises the boxes conservatively, allocates the inner
 *      children (contiguous) and the triangles of its leaf children with two atomic counters and writes Node8 / Tri48;
 *   4. the top level (few primitives: instances) is built and converted on the host (accel_build.cpp, bvh8_build.cpp) with
 *      node indices continuing after the mesh nodes.
 * Traversal results do not depend on the tree (tie rule): scenes built here give hits identical to the
 * oracle's on any other structure of the same triangles (tests/test_accel_build.py). */
#include "lbvh.cuh"
#include "bvh8_build.h"
#include <map>

int imba_accel_finish_top(imba_accel* out, const imba_scene_desc* d, const int* mesh_root, const float* mesh_box6);

namespace {

constexpr float COST_NODE = 1.0f;

__device__ __forceinline__ float box_area(const Box6& b) {
    if (!(b.v[0] <= b.v[3] && b.v[1] <= b.v[4] && b.v[2] <= b.v[5])) return 0.0f;
    const float dx = b.v[3] - b.v[0], dy = b.v[4] - b.v[1], dz = b.v[5] - b.v[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

/* Decisions per internal node, packed: bit 0 = as a single root the subtree is an 8-wide node (1) or a merged leaf (0);
 * bits 1..6: forest budget i = 2..7 is distributed over the children (1) or falls back to i-1 (0);
 * bits 8+3(j-2) .. : how many of a budget of j = 2..8 roots go to the left child. */
struct CostTables {
    float*    F;     // (2n-1) x 7: cheapest cost of the subtree as a forest of <= i roots, i = 1..7
    uint32_t* dec;   // n-1
    uint8_t*  ptot;  // 2n-1: min(triangles in the subtree, 4)
};

__global__ void collapse_cost_kernel(int n, const int2* children, const int* parent, const Box6* boxes, int* flags, CostTables T, float cost_prim) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int leaf = (n - 1) + p;
    const float a = box_area(boxes[leaf]);
    for (int i = 0; i < 7; ++i) T.F[(size_t)leaf * 7 + i] = a * cost_prim;
    T.ptot[leaf] = 1;
    __threadfence();
    int node = parent[leaf];
    while (node >= 0) {
        if (atomicAdd(flags + node, 1) == 0) return;
        __threadfence();
        const int2 ch = children[node];
        float Fl[8], Fr[8];
        for (int i = 1; i <= 7; ++i) { Fl[i] = T.F[(size_t)ch.x * 7 + i - 1]; Fr[i] = T.F[(size_t)ch.y * 7 + i - 1]; }
        const int pt = min(4, (int)T.ptot[ch.x] + (int)T.ptot[ch.y]);
        float D[9]; uint32_t dec = 0;
        for (int j = 2; j <= 8; ++j) {
            float best = FLT_MAX; int bk = 1;
            for (int k = 1; k < j; ++k) {
                if (k > 7 || j - k > 7) continue;
                const float c = Fl[k] + Fr[j - k];
                if (c < best) { best = c; bk = k; }
            }
            D[j] = best; dec |= (uint32_t)bk << (8 + 3 * (j - 2));
        }
        const float A = box_area(boxes[node]);
        const float inner = A * COST_NODE + D[8];
        const float leafc = pt <= 3 ? A * (float)pt * cost_prim : FLT_MAX;
        float F[8];
        if (leafc <= inner) F[1] = leafc; else { F[1] = inner; dec |= 1u; }
        for (int i = 2; i <= 7; ++i) {
            if (D[i] < F[i - 1]) { F[i] = D[i]; dec |= 1u << (i - 1); } else F[i] = F[i - 1];
        }
        for (int i = 1; i <= 7; ++i) T.F[(size_t)node * 7 + i - 1] = F[i];
        T.dec[node] = dec; T.ptot[node] = (uint8_t)pt;
        __threadfence();
        node = parent[node];
    }
}

struct EmitArgs {
    MeshView m; int n;
    const int* ids; const int2* children; const Box6* boxes; const uint32_t* dec;
    const int* work;         // work[k - work_base] = binary-tree node that becomes Node8 k
    int* work_out;
    uint32_t work_base;      // Node8 index of work[0]
    Node8* nodes; Tri48* tris;
    unsigned* node_counter; unsigned* tri_counter; // next free Node8 / Tri48 index (absolute)
    int* error;
};

__device__ __forceinline__ int ilog2_floor_d(double v) { int e; frexp(v, &e); return e - 1; }

__global__ void emit_level_kernel(EmitArgs A, uint32_t lo, uint32_t hi) {
    const uint32_t k = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= hi) return;
    const int n = A.n, first_leaf = n - 1;
    const int x = A.work[k - A.work_base];
    // ---- children of this 8-wide node: follow the decisions (bvh8_build.cpp: Collapse::distribute / collect) ----
    int ch[8]; bool ch_inner[8]; int nc = 0;
    if (x >= first_leaf) { ch[0] = x; ch_inner[0] = false; nc = 1; }
    else {
        int st_node[24]; int8_t st_i[24]; int8_t st_dist[24]; int sp = 0;
        st_node[0] = x; st_i[0] = 8; st_dist[0] = 1; sp = 1;
        while (sp > 0) {
            if (sp > 20 || nc >= 8) { atomicExch(A.error, 1); return; }
            --sp;
            const int y = st_node[sp]; int i = st_i[sp]; const bool dist = st_dist[sp] != 0;
            if (y >= first_leaf) { ch[nc] = y; ch_inner[nc] = false; ++nc; continue; }
            const uint32_t d = A.dec[y];
            if (dist) {
                const int kk = (int)((d >> (8 + 3 * (i - 2))) & 7u);
                const int2 c = A.children[y];
                st_node[sp] = c.y; st_i[sp] = (int8_t)(i - kk); st_dist[sp] = 0; ++sp;   // right is collected after left
                st_node[sp] = c.x; st_i[sp] = (int8_t)kk;       st_dist[sp] = 0; ++sp;
                continue;
            }
            if (i > 7) i = 7;
            if (i == 1) { ch[nc] = y; ch_inner[nc] = (d & 1u) != 0; ++nc; continue; }
            if ((d >> (i - 1)) & 1u) { st_node[sp] = y; st_i[sp] = (int8_t)i; st_dist[sp] = 1; ++sp; }
            else                     { st_node[sp] = y; st_i[sp] = (int8_t)(i - 1); st_dist[sp] = 0; ++sp; }
        }
    }
    if (nc > 8 || nc == 0) { atomicExch(A.error, 2); return; }
    Box6 cb[8]; float nlo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, nhi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int c = 0; c < nc; ++c) {
        cb[c] = A.boxes[ch[c]];
        for (int a = 0; a < 3; ++a) { nlo[a] = fminf(nlo[a], cb[c].v[a]); nhi[a] = fmaxf(nhi[a], cb[c].v[3 + a]); }
    }
    // ---- slot assignment: greedy on centroid . octant direction (Ylitie et al. 2017 section 4.2) ----
    int slot_of[8]; uint32_t slot_used = 0, child_done = 0;
    for (int it = 0; it < nc; ++it) {
        int bc = -1, bs = -1; float bv = FLT_MAX;
        for (int c = 0; c < nc; ++c) {
            if (child_done >> c & 1u) continue;
            float cc[3];
            for (int a = 0; a < 3; ++a) cc[a] = 0.5f * (cb[c].v[a] + cb[c].v[3 + a]) - 0.5f * (nlo[a] + nhi[a]);
            for (int s = 0; s < 8; ++s) {
                if (slot_used >> s & 1u) continue;
                float v = 0.0f;
                for (int a = 0; a < 3; ++a) v += cc[a] * (((s >> (2 - a)) & 1) ? -1.0f : 1.0f);
                if (v < bv) { bv = v; bc = c; bs = s; }
            }
        }
        if (bc < 0) { for (int c = 0; c < nc && bc < 0; ++c) if (!(child_done >> c & 1u)) bc = c; for (int s = 0; s < 8 && bs < 0; ++s) if (!(slot_used >> s & 1u)) bs = s; }
        slot_of[bc] = bs; child_done |= 1u << bc; slot_used |= 1u << bs;
    }
    int child_in_slot[8]; for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
    for (int c = 0; c < nc; ++c) child_in_slot[slot_of[c]] = c;
    // ---- quantisation grid (same rules as the host converter: >= 1/64 cell of slack, grid not finer than float resolution) ----
    Node8 nd; memset(&nd, 0, sizeof(nd));
    double scale[3];
    for (int a = 0; a < 3; ++a) {
        const float lo_ = nlo[a], hi_ = nhi[a];
        const double ext = (double)hi_ - (double)lo_;
        float p = (float)((double)lo_ - ext * (1.0 / 1024.0));
        while (p > lo_) p = nextafterf(p, -FLT_MAX);
        if (!(ext > 0.0)) p = lo_;
        const double span = (double)hi_ - (double)p;
        int e = span > 0.0 ? ilog2_floor_d(span / 254.0) : -120;
        if (e < -120) e = -120;
        for (;; ++e) {
            if (e >= 127) break;
            const double s = ldexp(1.0, e);
            if (span / s + 1.0 / 64.0 > 254.5) continue;
            const double mm = fmax(fabs((double)p), fabs((double)p + 255.0 * s));
            if (mm > 0.0 && e < ilog2_floor_d(mm) - 23) continue;
            break;
        }
        if (e > 127) e = 127;
        nd.p[a] = p; nd.e[a] = (uint8_t)(e + 127); scale[a] = ldexp(1.0, e);
    }
    // ---- children: counts, allocation, boxes, triangles ----
    int n_inner = 0, n_tris = 0; int leaf_tris[8][3]; int leaf_cnt[8];
    for (int s = 0; s < 8; ++s) {
        leaf_cnt[s] = 0;
        const int c = child_in_slot[s];
        if (c < 0) continue;
        if (ch_inner[c]) { ++n_inner; continue; }
        // triangles of the (merged) leaf: the <= 3 leaves below ch[c]
        int st[8]; int sp = 0; st[sp++] = ch[c];
        while (sp > 0) {
            const int y = st[--sp];
            if (y >= first_leaf) { if (leaf_cnt[s] < 3) leaf_tris[s][leaf_cnt[s]] = y - first_leaf; ++leaf_cnt[s]; continue; }
            const int2 cc = A.children[y];
            if (sp + 2 > 8) { atomicExch(A.error, 3); return; }
            st[sp++] = cc.y; st[sp++] = cc.x;
        }
        if (leaf_cnt[s] > 3 || leaf_cnt[s] == 0) { atomicExch(A.error, 4); return; }
        n_tris += leaf_cnt[s];
    }
    const uint32_t child_base = n_inner ? atomicAdd(A.node_counter, (unsigned)n_inner) : 0u;
    const uint32_t prim_base = n_tris ? atomicAdd(A.tri_counter, (unsigned)n_tris) : 0u;
    nd.child_base = child_base; nd.prim_base = prim_base;
    uint8_t* qlo[3] = { nd.qlo_x, nd.qlo_y, nd.qlo_z }; uint8_t* qhi[3] = { nd.qhi_x, nd.qhi_y, nd.qhi_z };
    int inner_rank = 0, prim_off = 0;
    for (int s = 0; s < 8; ++s) {
        const int c = child_in_slot[s];
        if (c < 0) { for (int a = 0; a < 3; ++a) { qlo[a][s] = 255; qhi[a][s] = 0; } continue; }
        for (int a = 0; a < 3; ++a) {
            const float blo = cb[c].v[a], bhi = cb[c].v[3 + a];
            double l = floor(((double)blo - (double)nd.p[a]) / scale[a] - 1.0 / 64.0);
            double h = ceil(((double)bhi - (double)nd.p[a]) / scale[a] + 1.0 / 64.0);
            if (!(l >= 0.0)) l = 0.0;
            if (!(h <= 255.0)) h = 255.0;
            if (l > 255.0) l = 255.0;
            if (h < 0.0) h = 0.0;
            const float sf = (float)scale[a];
            while (l > 0.0 && __fmaf_rn((float)l, sf, nd.p[a]) > blo) l -= 1.0;   // verify with the traversal kernel's own decode
            while (h < 255.0 && __fmaf_rn((float)h, sf, nd.p[a]) < bhi) h += 1.0;
            qlo[a][s] = (uint8_t)l; qhi[a][s] = (uint8_t)h;
        }
        if (ch_inner[c]) {
            nd.imask |= (uint8_t)(1u << s);
            nd.meta[s] = (uint8_t)((1u << 5) | (24u + s));
            A.work_out[child_base + inner_rank - A.work_base] = ch[c];
            ++inner_rank;
        } else {
            const int cnt = leaf_cnt[s];
            const uint8_t unary = cnt >= 3 ? 7 : (cnt == 2 ? 3 : 1);
            nd.meta[s] = (uint8_t)((unary << 5) | prim_off);
            for (int i = 0; i < cnt; ++i) {
                const int t = A.ids[leaf_tris[s][i]];
                float4 a, b, c3; tri_verts(A.m, t, a, b, c3);
                Tri48 o;
                o.v0[0] = a.x; o.v0[1] = a.y; o.v0[2] = a.z; o.id = A.m.tbase + t;
                o.e1[0] = a.x - b.x; o.e1[1] = a.y - b.y; o.e1[2] = a.z - b.z;
                o.e2[0] = c3.x - a.x; o.e2[1] = c3.y - a.y; o.e2[2] = c3.z - a.z;
                IMBA_TRI48_FINISH(o);
                A.tris[prim_base + prim_off + i] = o;
            }
            prim_off += cnt;
        }
    }
    A.nodes[k] = nd;
}

} // namespace

int imba_shading_upload(imba_scene* s, const imba_scene_desc* d); // shade_data.cu

#define BG_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { imba_set_error("%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); return IMBA_ERR_CUDA; } } while (0)

/* Builds the whole structure of `d` (meshes + instances; d->nodes / tris / inst_nodes are ignored) on the current device and
 * installs it in `s`.  seconds[0] = device time of the mesh builds, seconds[1] = levels of the deepest mesh tree. */
int imba_scene_build_accel_device(imba_scene* s, const imba_scene_desc* d, float* seconds) {
    if (d->mesh_count <= 0 || !d->mesh_vertex_base || !d->mesh_tri_base || !d->vertices || !d->indices || d->instance_count <= 0 || !d->instances) {
        imba_set_error("device build needs meshes (vertices, indices, bases) and instances"); return IMBA_ERR_INVALID;
    }
    const bool use_ploc = !(getenv("IMBA_GPU_BUILDER") && !strcmp(getenv("IMBA_GPU_BUILDER"), "lbvh")); // default: PLOC
    const float cost_prim = getenv("IMBA_COST_PRIM") ? (float)atof(getenv("IMBA_COST_PRIM")) : 0.5f;
    const int nv = d->mesh_vertex_base[d->mesh_count], nt_all = d->mesh_tri_base[d->mesh_count];
    size_t node_cap = 0;
    for (int mi = 0; mi < d->mesh_count; ++mi) {
        const int n = d->mesh_tri_base[mi + 1] - d->mesh_tri_base[mi];
        if (n <= 0) { imba_set_error("mesh %d has no triangles", mi); return IMBA_ERR_INVALID; }
        node_cap += (size_t)std::max(n - 1, 1);
    }
    // the kernels index the vertex array with what the index array says: reject anything out of range here
    for (int mi = 0; mi < d->mesh_count; ++mi) {
        const int nvm = d->mesh_vertex_base[mi + 1] - d->mesh_vertex_base[mi];
        for (int t = d->mesh_tri_base[mi]; t < d->mesh_tri_base[mi + 1]; ++t)
            for (int c = 0; c < 3; ++c) {
                const int v = d->indices[4 * (size_t)t + c];
                if (v < 0 || v >= nvm) { imba_set_error("triangle %d of mesh %d refers to vertex %d (mesh has %d)", t - d->mesh_tri_base[mi], mi, v, nvm); return IMBA_ERR_INVALID; }
            }
    }
    for (int i = 0; i < d->instance_count; ++i)
        if (d->instances[i].mesh_id < 0 || d->instances[i].mesh_id >= d->mesh_count) { imba_set_error("instance %d refers to mesh %d", i, d->instances[i].mesh_id); return IMBA_ERR_INVALID; }
    DeviceBuf d_vertices, d_indices, d_nodes, d_tris, d_work, d_ctr;
    BG_CUDA(d_vertices.alloc((size_t)nv * 16)); BG_CUDA(d_indices.alloc((size_t)nt_all * 16));
    BG_CUDA(cudaMemcpy(d_vertices.p, d->vertices, (size_t)nv * 16, cudaMemcpyHostToDevice));
    BG_CUDA(cudaMemcpy(d_indices.p, d->indices, (size_t)nt_all * 16, cudaMemcpyHostToDevice));
    BG_CUDA(d_nodes.alloc(node_cap * sizeof(Node8))); BG_CUDA(d_tris.alloc(((size_t)nt_all + 1) * sizeof(Tri48)));
    BG_CUDA(d_work.alloc(node_cap * 4)); BG_CUDA(d_ctr.alloc(16));
    BG_CUDA(cudaMemset(d_ctr.p, 0, 16));
    unsigned* ctr = d_ctr.as<unsigned>();  // [0] next Node8, [1] next Tri48, [2] error
    struct Events { cudaEvent_t a = nullptr, b = nullptr; Events() { cudaEventCreate(&a); cudaEventCreate(&b); } ~Events() { cudaEventDestroy(a); cudaEventDestroy(b); } } ev;
    const cudaEvent_t e0 = ev.a, e1 = ev.b;
    float device_ms = 0.0f; int mesh_depth = 0;
    std::vector<int> mesh_key(d->mesh_count); std::vector<float> mesh_box(6 * (size_t)d->mesh_count);
    std::map<int, uint32_t> mesh_roots;
    unsigned node_count = 0;
    for (int mi = 0; mi < d->mesh_count; ++mi) {
        MeshView m{ d_vertices.as<float4>(), d_indices.as<int4>(), d->mesh_vertex_base[mi], d->mesh_tri_base[mi], d->mesh_tri_base[mi + 1] - d->mesh_tri_base[mi] };
        const int n = m.nt;
        LbvhTree t; DeviceBuf F, dec, ptot, flags2;
        BG_CUDA(F.alloc((size_t)(2 * n) * 7 * 4)); BG_CUDA(dec.alloc((size_t)n * 4)); BG_CUDA(ptot.alloc((size_t)2 * n)); BG_CUDA(flags2.alloc((size_t)n * 4));
        BG_CUDA(cudaMemset(flags2.p, 0, (size_t)n * 4));
        BG_CUDA(use_ploc ? ploc_build_tree(m, t, e0) : lbvh_build_tree(m, t, e0));
        CostTables T{ F.as<float>(), dec.as<uint32_t>(), ptot.as<uint8_t>() };
        const int blocks = (n + 255) / 256;
        collapse_cost_kernel<<<blocks, 256>>>(n, t.kids(), t.parents(), t.bounds(), flags2.as<int>(), T, cost_prim);
        g_imba_launches.fetch_add(1);
        // breadth-first emission; the root of this mesh is Node8 `root_idx`
        const unsigned root_idx = node_count;
        const int root_node = t.root; // the root of the binary tree; for n == 1 the only leaf has id n-1 = 0 as well
        BG_CUDA(cudaMemcpy(d_work.as<int>() + root_idx, &root_node, 4, cudaMemcpyHostToDevice));
        node_count = root_idx + 1;
        BG_CUDA(cudaMemcpy(ctr, &node_count, 4, cudaMemcpyHostToDevice));
        EmitArgs A{ m, n, t.sorted_ids(), t.kids(), t.bounds(), dec.as<uint32_t>(), d_work.as<int>(), d_work.as<int>(), 0u,
                    d_nodes.as<Node8>(), d_tris.as<Tri48>(), ctr, ctr + 1, (int*)(ctr + 2) };
        unsigned lo = root_idx, hi = root_idx + 1; int levels = 0;
        while (lo < hi) {
            emit_level_kernel<<<(hi - lo + 127) / 128, 128>>>(A, lo, hi);
            g_imba_launches.fetch_add(1);
            unsigned c3[3];
            BG_CUDA(cudaMemcpy(c3, ctr, 12, cudaMemcpyDeviceToHost));
            if (c3[2]) { imba_set_error("device collapse failed (code %u)", c3[2]); return IMBA_ERR_INVALID; }
            if (c3[0] > node_cap) { imba_set_error("device collapse overflowed its node buffer"); return IMBA_ERR_OVERFLOW; }
            lo = hi; hi = c3[0]; ++levels;
        }
        node_count = hi;
        cudaEventRecord(e1);
        BG_CUDA(cudaEventSynchronize(e1));
        float ms = 0.0f; cudaEventElapsedTime(&ms, e0, e1); device_ms += ms;
        mesh_depth = std::max(mesh_depth, levels);
        Box6 root; BG_CUDA(cudaMemcpy(&root, t.bounds() + t.root, sizeof(Box6), cudaMemcpyDeviceToHost));
        memcpy(&mesh_box[6 * mi], root.v, 24);
        mesh_key[mi] = mi; mesh_roots[mi] = root_idx;
    }
    unsigned tri_count = 0; BG_CUDA(cudaMemcpy(&tri_count, ctr + 1, 4, cudaMemcpyDeviceToHost));
    // ---- top level on the host: BVH2 over the instance boxes, converted with node indices continuing after the meshes ----
    imba_accel top;
    if (imba_accel_finish_top(&top, d, mesh_key.data(), mesh_box.data()) != IMBA_OK) return IMBA_ERR_INVALID;
    Bvh8 tb; std::string err;
    if (!imba_build_bvh8_top(IMBA_ACCEL_BVH2, top.root, top.nodes.data(), top.inst_nodes.data(), mesh_roots, node_count, tb, err)) {
        imba_set_error("top-level conversion failed: %s", err.c_str()); return IMBA_ERR_INVALID;
    }
    const int depth = tb.max_depth + mesh_depth;
    if (depth + 6 > IMBA_TRAVERSAL_STACK) { imba_set_error("acceleration structure too deep for the traversal stack (depth %d, stack %d)", depth, IMBA_TRAVERSAL_STACK); return IMBA_ERR_OVERFLOW; }
    // ---- final arrays: exact size, mesh nodes + top-level nodes; triangles + one padding record ----
    const size_t total_nodes = (size_t)node_count + tb.nodes.size();
    BG_CUDA(cudaMalloc((void**)&s->d_nodes, total_nodes * sizeof(Node8)));
    BG_CUDA(cudaMemcpy(s->d_nodes, d_nodes.p, (size_t)node_count * sizeof(Node8), cudaMemcpyDeviceToDevice));
    BG_CUDA(cudaMemcpy((char*)s->d_nodes + (size_t)node_count * sizeof(Node8), tb.nodes.data(), tb.nodes.size() * sizeof(Node8), cudaMemcpyHostToDevice));
    BG_CUDA(cudaMalloc((void**)&s->d_tris, ((size_t)tri_count + 1) * sizeof(Tri48)));
    BG_CUDA(cudaMemcpy(s->d_tris, d_tris.p, (size_t)tri_count * sizeof(Tri48), cudaMemcpyDeviceToDevice));
    Tri48 pad; memset(&pad, 0, sizeof(pad)); pad.id = (int32_t)0x80000000;
    BG_CUDA(cudaMemcpy((char*)s->d_tris + (size_t)tri_count * sizeof(Tri48), &pad, sizeof(pad), cudaMemcpyHostToDevice));
    BG_CUDA(cudaMalloc((void**)&s->d_insts, std::max<size_t>(tb.insts.size(), 1) * sizeof(Inst64)));
    BG_CUDA(cudaMemcpy(s->d_insts, tb.insts.data(), tb.insts.size() * sizeof(Inst64), cudaMemcpyHostToDevice));
    BG_CUDA(cudaMalloc((void**)&s->d_counters, imba_scene::kCounters * sizeof(unsigned)));
    BG_CUDA(cudaMemset(s->d_counters, 0, imba_scene::kCounters * sizeof(unsigned)));
    s->top_root = tb.top_root; s->flat_inst = -1;
    if (tb.flat && !getenv("IMBA_NO_FLAT")) { s->top_root = tb.flat_root; s->flat_inst = tb.flat_inst; }
    s->stats[0] = (int64_t)total_nodes; s->stats[1] = (int64_t)tri_count; s->stats[2] = (int64_t)tb.insts.size();
    s->stats[3] = (int64_t)(total_nodes * sizeof(Node8) + ((size_t)tri_count + 1) * sizeof(Tri48) + tb.insts.size() * sizeof(Inst64));
    s->stats[4] = depth;
    if (seconds) { seconds[0] = device_ms * 1e-3f; seconds[1] = (float)mesh_depth; }
    return IMBA_OK;
}
