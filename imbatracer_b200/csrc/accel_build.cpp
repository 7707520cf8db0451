/* accel_build.cpp -- host builder of the two-level acceleration structure in the reference's BVH2 ("gpu") layout, for
 * scenes that do not come with one (procedurally generated scenes; row f1 "next").
 *
 * The reference builds its structures with a spatial-split BVH per mesh (core/sbvh_builder.h) and a binned-SAH BVH over
 * instance boxes (core/fast_bvh_builder.h); both are single-threaded host code.  This file is NOT a restatement of
 * those builders: it is a plain binned-SAH (16 bins, no spatial splits) top-down builder that EMITS THE SAME LAYOUT the
 * reference's GPU adapters write (core/gpu_adapter.cpp:62-143, 203-297; render/scene.cpp:168-176):
 *   - mesh nodes first, one tree per mesh, depth-first, children linked by index; leaves = ~offset into the Vec4 array of
 *     (v0, v1[w = global triangle id], v2) triples, last v2.w = 0x80000000;
 *   - top-level nodes after the mesh nodes, root = index of the first top-level node, leaves = ~index of an
 *     InstanceNode{inverse 3x4, instance id, next = mesh root, pad}, list terminated by pad[0] = -1;
 *   - a tree consisting of a single leaf gets the dummy parent with right = 0x76543210.
 * The output therefore goes through the same converter (bvh8_build.cpp) and can be traced by the CPU oracle as well.
 */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "api_common.h"

namespace {

struct Box3 {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
    void grow(const float* p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box3& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float half_area() const { const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2]; return x * y + y * z + z * x; }
};

inline float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
inline void set_bb(imba_bbox& d, const Box3& b) { d.lo_x = b.lo[0]; d.lo_y = b.lo[1]; d.lo_z = b.lo[2]; d.hi_x = b.hi[0]; d.hi_y = b.hi[1]; d.hi_z = b.hi[2]; }

// triangles per leaf of the binary tree (tuning knob IMBA_MAX_LEAF, 1..3; the 8-wide collapse may merge leaves again)
static const int MAX_LEAF = getenv("IMBA_MAX_LEAF") ? std::min(3, std::max(1, atoi(getenv("IMBA_MAX_LEAF")))) : 1;

struct Builder {
    const Box3* boxes; std::vector<int> prims; int max_leaf;
    std::vector<imba_node2>& nodes; int node_offset; // indices written = local index + node_offset
    // leaf emitter: returns the (positive) offset to store as ~offset
    struct LeafFn { virtual int emit(const int* prims, int n) = 0; virtual ~LeafFn() {} };
    LeafFn* leaf;

    Builder(const Box3* b, int n, int max_leaf, std::vector<imba_node2>& nodes, int node_offset, LeafFn* leaf)
        : boxes(b), prims(n), max_leaf(max_leaf), nodes(nodes), node_offset(node_offset), leaf(leaf) { for (int i = 0; i < n; ++i) prims[i] = i; }

    Box3 bounds(int first, int count) const { Box3 b; b.reset(); for (int i = first; i < first + count; ++i) b.grow(boxes[prims[i]]); return b; }

    // Splits [first, first+count) in place; returns the size of the left part (0 = make a leaf).
    int split(int first, int count, const Box3& bb) {
        if (count <= max_leaf) return 0;
        Box3 cb; cb.reset();
        for (int i = first; i < first + count; ++i) { const Box3& b = boxes[prims[i]]; const float c[3] = { 0.5f * (b.lo[0] + b.hi[0]), 0.5f * (b.lo[1] + b.hi[1]), 0.5f * (b.lo[2] + b.hi[2]) }; cb.grow(c); }
        constexpr int BINS = 16;
        float best_cost = FLT_MAX; int best_axis = -1, best_bin = -1;
        for (int a = 0; a < 3; ++a) {
            const float ext = cb.hi[a] - cb.lo[a];
            if (!(ext > 0.0f)) continue;
            Box3 bin_box[BINS]; int bin_n[BINS];
            for (int k = 0; k < BINS; ++k) { bin_box[k].reset(); bin_n[k] = 0; }
            const float scale = BINS / ext;
            for (int i = first; i < first + count; ++i) {
                const Box3& b = boxes[prims[i]];
                int k = (int)((0.5f * (b.lo[a] + b.hi[a]) - cb.lo[a]) * scale); k = std::min(std::max(k, 0), BINS - 1);
                bin_box[k].grow(b); ++bin_n[k];
            }
            float right_area[BINS]; Box3 acc; acc.reset(); int n_right = 0;
            for (int k = BINS - 1; k > 0; --k) { acc.grow(bin_box[k]); right_area[k] = acc.half_area(); }
            acc.reset(); int n_left = 0; (void)n_right;
            for (int k = 0; k < BINS - 1; ++k) {
                if (bin_n[k]) acc.grow(bin_box[k]);
                n_left += bin_n[k];
                if (n_left == 0 || n_left == count) continue;
                const float cost = acc.half_area() * n_left + right_area[k + 1] * (count - n_left);
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k; }
            }
        }
        int mid;
        if (best_axis < 0) { // all centroids coincide: split in the middle
            mid = first + count / 2;
        } else {
            const float ext = cb.hi[best_axis] - cb.lo[best_axis], scale = 16 / ext, lo = cb.lo[best_axis];
            const int a = best_axis, bin = best_bin;
            int* b = prims.data() + first; int* e = b + count;
            int* m = std::partition(b, e, [&](int p) {
                const Box3& bx = boxes[p];
                int k = (int)((0.5f * (bx.lo[a] + bx.hi[a]) - lo) * scale); k = std::min(std::max(k, 0), 15);
                return k <= bin; });
            mid = (int)(m - prims.data());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        (void)bb;
        return mid - first;
    }

    // Builds the subtree over [first, first+count); returns the child reference (node index or ~leaf offset).
    int build(int first, int count, const Box3& bb, int depth) {
        const int nl = depth > 60 ? 0 : split(first, count, bb);
        if (nl == 0) {
            int done = 0, ref = 0;
            // a leaf larger than max_leaf (degenerate input) is still emitted whole: the converter splits it
            ref = ~leaf->emit(prims.data() + first, count); (void)done;
            return ref;
        }
        const int me = (int)nodes.size();
        nodes.emplace_back();
        const Box3 lb = bounds(first, nl), rb = bounds(first + nl, count - nl);
        set_bb(nodes[me].left_bb, lb); set_bb(nodes[me].right_bb, rb);
        const int l = build(first, nl, lb, depth + 1);
        nodes[me].left = l;
        const int r = build(first + nl, count - nl, rb, depth + 1);
        nodes[me].right = r;
        return me + node_offset;
    }

    // Builds a whole tree; returns the root NODE index (a dummy parent is created around a single leaf).
    int run() {
        const int n = (int)prims.size();
        const Box3 bb = bounds(0, n);
        const int first_node = (int)nodes.size();
        const int ref = build(0, n, bb, 0);
        if (ref >= 0) return ref;
        nodes.emplace_back(); // fill_dummy_parent (gpu_adapter.cpp:14-31)
        imba_node2& d = nodes.back();
        d.left = ref; d.right = 0x76543210; set_bb(d.left_bb, bb);
        d.right_bb.lo_x = d.right_bb.lo_y = d.right_bb.lo_z = 0.0f; d.right_bb.hi_x = d.right_bb.hi_y = d.right_bb.hi_z = -0.0f;
        return first_node + node_offset;
    }
};

} // namespace

extern "C" void imba_accel_free(imba_accel* a) { delete a; }

extern "C" imba_accel* imba_accel_build(const imba_scene_desc* d) {
    if (!d || d->mesh_count <= 0 || !d->mesh_vertex_base || !d->mesh_tri_base || !d->vertices || !d->indices || d->instance_count <= 0 || !d->instances) {
        imba_set_error("imba_accel_build needs meshes (vertices, indices, bases) and instances"); return nullptr;
    }
    imba_accel* out = new imba_accel;
    std::vector<int> mesh_root(d->mesh_count); std::vector<Box3> mesh_box(d->mesh_count);
    // ---- one tree per mesh ------------------------------------------------------------------------------------------
    for (int m = 0; m < d->mesh_count; ++m) {
        const int vbase = d->mesh_vertex_base[m], tbase = d->mesh_tri_base[m], nt = d->mesh_tri_base[m + 1] - tbase;
        if (nt <= 0) { imba_set_error("mesh %d has no triangles", m); delete out; return nullptr; }
        std::vector<Box3> boxes(nt);
        for (int t = 0; t < nt; ++t) {
            boxes[t].reset();
            for (int k = 0; k < 3; ++k) boxes[t].grow(d->vertices + 4 * (size_t)(vbase + d->indices[4 * (size_t)(tbase + t) + k]));
        }
        struct TriLeaf : Builder::LeafFn {
            const imba_scene_desc* d; int vbase, tbase; std::vector<imba_vec4>* tris;
            int emit(const int* prims, int n) override {
                const int off = (int)tris->size();
                for (int i = 0; i < n; ++i) {
                    const int t = prims[i];
                    const float* v[3];
                    for (int k = 0; k < 3; ++k) v[k] = d->vertices + 4 * (size_t)(vbase + d->indices[4 * (size_t)(tbase + t) + k]);
                    tris->push_back(imba_vec4{ v[0][0], v[0][1], v[0][2], 0.0f });
                    tris->push_back(imba_vec4{ v[1][0], v[1][1], v[1][2], i2f(tbase + t) });
                    tris->push_back(imba_vec4{ v[2][0], v[2][1], v[2][2], 0.0f });
                }
                tris->back().w = i2f((int32_t)0x80000000);
                return off;
            }
        } leaf; leaf.d = d; leaf.vbase = vbase; leaf.tbase = tbase; leaf.tris = &out->tris;
        Builder b(boxes.data(), nt, MAX_LEAF, out->nodes, 0, &leaf);
        mesh_root[m] = b.run();
        mesh_box[m].reset(); for (int t = 0; t < nt; ++t) mesh_box[m].grow(boxes[t]);
    }
    std::vector<float> boxes6(6 * (size_t)d->mesh_count);
    for (int m = 0; m < d->mesh_count; ++m) for (int a = 0; a < 3; ++a) { boxes6[6 * m + a] = mesh_box[m].lo[a]; boxes6[6 * m + 3 + a] = mesh_box[m].hi[a]; }
    if (imba_accel_finish_top(out, d, mesh_root.data(), boxes6.data()) != IMBA_OK) { delete out; return nullptr; }
    return out;
}

/* Top level over the instance boxes, appended after the mesh nodes (shared by the host and the GPU mesh builders). */
int imba_accel_finish_top(imba_accel* out, const imba_scene_desc* d, const int* mesh_root_in, const float* mesh_box6) {
    std::vector<int> mesh_root(mesh_root_in, mesh_root_in + d->mesh_count);
    std::vector<Box3> mesh_box(d->mesh_count);
    for (int m = 0; m < d->mesh_count; ++m) for (int a = 0; a < 3; ++a) { mesh_box[m].lo[a] = mesh_box6[6 * m + a]; mesh_box[m].hi[a] = mesh_box6[6 * m + 3 + a]; }
    // ---- top level over instance boxes (cpu_adapter.cpp:233-249: centre +- |M| * half extent) ---------------------------
    std::vector<Box3> ibox(d->instance_count);
    for (int i = 0; i < d->instance_count; ++i) {
        const imba_instance& in = d->instances[i];
        if (in.mesh_id < 0 || in.mesh_id >= d->mesh_count) { imba_set_error("instance %d refers to mesh %d", i, in.mesh_id); return IMBA_ERR_INVALID; }
        const Box3& mb = mesh_box[in.mesh_id];
        for (int r = 0; r < 3; ++r) {
            float c = in.mat[4 * r + 3], e = 0.0f;
            for (int k = 0; k < 3; ++k) { c += in.mat[4 * r + k] * 0.5f * (mb.lo[k] + mb.hi[k]); e += fabsf(in.mat[4 * r + k]) * 0.5f * (mb.hi[k] - mb.lo[k]); }
            ibox[i].lo[r] = c - e; ibox[i].hi[r] = c + e;
        }
    }
    struct InstLeaf : Builder::LeafFn {
        const imba_scene_desc* d; const std::vector<int>* mesh_root; std::vector<imba_instance_node>* nodes;
        int emit(const int* prims, int n) override {
            const int off = (int)nodes->size();
            for (int i = 0; i < n; ++i) {
                const imba_instance& in = d->instances[prims[i]];
                imba_instance_node o; memcpy(o.transf, in.inv_mat, 48); o.id = prims[i]; o.next = (*mesh_root)[in.mesh_id]; o.pad[0] = o.pad[1] = 0;
                nodes->push_back(o);
            }
            nodes->back().pad[0] = -1; nodes->back().pad[1] = -1;
            return off;
        }
    } ileaf; ileaf.d = d; ileaf.mesh_root = &mesh_root; ileaf.nodes = &out->inst_nodes;
    // top-level nodes live in the same array, after the mesh nodes (scene.cpp:168-176); indices are absolute already
    Builder top(ibox.data(), d->instance_count, 1, out->nodes, 0, &ileaf);
    out->root = top.run();
    return IMBA_OK;
}

/* Points the acceleration-structure fields of `d` at the arrays owned by `a` (valid until imba_accel_free). */
extern "C" int imba_accel_fill(const imba_accel* a, imba_scene_desc* d) {
    if (!a || !d) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    d->accel_layout = IMBA_ACCEL_BVH2; d->root = a->root; d->nodes = a->nodes.data(); d->node_count = (int32_t)a->nodes.size();
    d->tris = a->tris.data(); d->tri_vec4_count = (int64_t)a->tris.size();
    d->inst_nodes = a->inst_nodes.data(); d->inst_node_count = (int32_t)a->inst_nodes.size();
    return IMBA_OK;
}

extern "C" int imba_accel_arrays(const imba_accel* a, const void** nodes, int64_t* node_count, const void** tris, int64_t* tri_vec4_count,
                                 const void** inst_nodes, int64_t* inst_node_count, int32_t* root) {
    if (!a) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    if (nodes) *nodes = a->nodes.data();
    if (node_count) *node_count = (int64_t)a->nodes.size();
    if (tris) *tris = a->tris.data();
    if (tri_vec4_count) *tri_vec4_count = (int64_t)a->tris.size();
    if (inst_nodes) *inst_nodes = a->inst_nodes.data();
    if (inst_node_count) *inst_node_count = (int64_t)a->inst_nodes.size();
    if (root) *root = a->root;
    return IMBA_OK;
}
