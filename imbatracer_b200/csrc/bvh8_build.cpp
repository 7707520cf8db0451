/* bvh8_build.cpp -- converts the reference's MBVH4 / BVH2 two-level structure into the compressed 8-wide
 * device layout of bvh8.h.  Host code, runs once per scene (replaces nothing in the reference: it consumes
 * what core/cpu_adapter.cpp / core/gpu_adapter.cpp wrote and render/scene.cpp:137-188 would upload).
 *
 * Steps per tree (the top level and every mesh reachable through InstanceNode::next):
 *   1. parse the reference nodes into a generic tree (child boxes come from the parent's slots);
 *   2. split leaves holding more than 3 primitives (median split on the longest centroid axis) -- a Node8
 *      leaf child addresses at most 3 primitives;
 *   3. collapse to 8-wide nodes with the SAH-optimal dynamic program of Ylitie, Karras, Laine 2017 (section 3): for every
 *      node of the (binarised) tree the cheapest representation as a leaf of <= 3 primitives, as one 8-wide node, or as a
 *      forest of i <= 7 roots handed to the parent; small sibling leaves are merged (node cost 1, primitive cost 0.5);
 *   4. assign children to slots so that `slot ^ ray octant` approximates front-to-back order (greedy
 *      assignment on centroid . octant-direction, Ylitie et al. 2017 section 4.2);
 *   5. quantise child boxes conservatively (always at least 1/64 cell of slack) to 8 bits on a power-of-two grid.
 */
#include "bvh8_build.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
    void grow(const Box& o) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], o.lo[a]); hi[a] = std::max(hi[a], o.hi[a]); } }
    void grow(const float* p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void clip(const Box& o) { for (int a = 0; a < 3; ++a) { lo[a] = std::max(lo[a], o.lo[a]); hi[a] = std::min(hi[a], o.hi[a]); if (hi[a] < lo[a]) hi[a] = lo[a]; } }
    bool valid() const { return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]; }
    float area() const {
        if (!valid()) return 0.0f;
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.0f * (dx * dy + dy * dz + dz * dx);
    }
};

struct TNode {
    Box b;
    int kids[4]; int nkids = 0;       // inner when nkids > 0
    int prim_first = 0, prim_count = 0; // leaf: range in Tree::prims
};

struct Tree {
    std::vector<TNode> n;
    std::vector<int>   prims;     // indices into the per-tree primitive store
    std::vector<Box>   prim_box;  // bounds per entry of `prims` (clipped to the reference leaf box)
    int root = -1;
};

inline int f2i(float f) { int i; memcpy(&i, &f, 4); return i; }

struct Parser {
    int layout; const void* nodes; const imba_vec4* tris; const imba_instance_node* inst_nodes;
    std::string* err; int visited = 0;
    // outputs of the primitive stores
    std::vector<Tri48>* tri_store = nullptr;                      // mesh trees
    std::vector<const imba_instance_node*>* inst_store = nullptr; // top-level tree
    bool top_level = false;

    bool fail(const char* m) { if (err->empty()) *err = m; return false; }

    static Box tri_bounds(const Tri48& t) {
        Box b; b.reset();
        float v1[3], v2[3];
        for (int a = 0; a < 3; ++a) { v1[a] = t.v0[a] - t.e1[a]; v2[a] = t.v0[a] + t.e2[a]; }
        b.grow(t.v0); b.grow(v1); b.grow(v2);
        // v1/v2 are reconstructed (one rounding each): widen by one ulp so the box certainly holds the stored vertices
        for (int a = 0; a < 3; ++a) { b.lo[a] = nextafterf(b.lo[a], -FLT_MAX); b.hi[a] = nextafterf(b.hi[a], FLT_MAX); }
        return b;
    }

    bool leaf(Tree& t, int tn, long first, const Box& box) {
        if (first < 0) return fail("negative leaf offset");
        t.n[tn].prim_first = (int)t.prims.size();
        int count = 0;
        if (top_level) {
            const imba_instance_node* in = inst_nodes + first;
            for (;; ++in) {
                t.prims.push_back((int)inst_store->size()); inst_store->push_back(in); t.prim_box.push_back(box); ++count;
                if (in->pad[0] == -1) break;
                if (count > (1 << 24)) return fail("unterminated instance list");
            }
        } else if (layout == IMBA_ACCEL_MBVH4) {
            const imba_vec4* p = tris + first;
            for (;;) {
                const float* f = (const float*)p;
                for (int l = 0; l < 4; ++l) {
                    const int id = f2i(f[48 + l]);
                    if (id < 0) continue;
                    Tri48 tr; memset(&tr, 0, sizeof(tr));
                    for (int a = 0; a < 3; ++a) { tr.v0[a] = f[4 * a + l]; tr.e1[a] = f[12 + 4 * a + l]; tr.e2[a] = f[24 + 4 * a + l]; }
                    tr.id = id; IMBA_TRI48_FINISH(tr);
                    Box pb = tri_bounds(tr); pb.clip(box);
                    t.prims.push_back((int)tri_store->size()); tri_store->push_back(tr); t.prim_box.push_back(pb); ++count;
                }
                p += 13;
                if ((uint32_t)f2i(p->x) == 0x80000000u && (uint32_t)f2i(p->y) == 0x80000000u &&
                    (uint32_t)f2i(p->z) == 0x80000000u && (uint32_t)f2i(p->w) == 0x80000000u) break;
                if (count > (1 << 24)) return fail("unterminated triangle packet list");
            }
        } else {
            const imba_vec4* p = tris + first;
            for (;; p += 3) {
                Tri48 tr; memset(&tr, 0, sizeof(tr));
                const float v0[3] = { p[0].x, p[0].y, p[0].z }, v1[3] = { p[1].x, p[1].y, p[1].z }, v2[3] = { p[2].x, p[2].y, p[2].z };
                for (int a = 0; a < 3; ++a) { tr.v0[a] = v0[a]; tr.e1[a] = v0[a] - v1[a]; tr.e2[a] = v2[a] - v0[a]; }
                tr.id = f2i(p[1].w); IMBA_TRI48_FINISH(tr);
                Box pb; pb.reset(); pb.grow(v0); pb.grow(v1); pb.grow(v2); pb.clip(box);
                t.prims.push_back((int)tri_store->size()); tri_store->push_back(tr); t.prim_box.push_back(pb); ++count;
                if ((uint32_t)f2i(p[2].w) == 0x80000000u) break;
                if (count > (1 << 24)) return fail("unterminated triangle list");
            }
        }
        t.n[tn].prim_count = count;
        return true;
    }

    // Parses reference node `idx`; returns the TNode index or -1.
    int inner(Tree& t, int idx, int depth) {
        if (depth > 256) { fail("reference tree deeper than 256"); return -1; }
        if (idx < 0) { fail("negative node index"); return -1; }
        ++visited;
        const int me = (int)t.n.size(); t.n.emplace_back();
        int ck[4]; Box cb[4]; int nc = 0;
        if (layout == IMBA_ACCEL_MBVH4) {
            const imba_node4& nd = ((const imba_node4*)nodes)[idx];
            for (int k = 0; k < 4; ++k) {
                if (nd.children[k] == 0) break;
                ck[nc] = nd.children[k];
                cb[nc].lo[0] = nd.min_x[k]; cb[nc].lo[1] = nd.min_y[k]; cb[nc].lo[2] = nd.min_z[k];
                cb[nc].hi[0] = nd.max_x[k]; cb[nc].hi[1] = nd.max_y[k]; cb[nc].hi[2] = nd.max_z[k];
                ++nc;
            }
        } else {
            const imba_node2& nd = ((const imba_node2*)nodes)[idx];
            for (int k = 0; k < 2; ++k) {
                const int c = k ? nd.right : nd.left;
                if (k == 1 && c == 0x76543210) break; // dummy parent of a single leaf (gpu_adapter.cpp:14-31)
                const imba_bbox& bb = k ? nd.right_bb : nd.left_bb;
                ck[nc] = c;
                cb[nc].lo[0] = bb.lo_x; cb[nc].lo[1] = bb.lo_y; cb[nc].lo[2] = bb.lo_z;
                cb[nc].hi[0] = bb.hi_x; cb[nc].hi[1] = bb.hi_y; cb[nc].hi[2] = bb.hi_z;
                ++nc;
            }
        }
        if (nc == 0) { fail("reference node without children"); return -1; }
        Box u; u.reset();
        for (int k = 0; k < nc; ++k) {
            int child;
            if (ck[k] < 0) {
                child = (int)t.n.size(); t.n.emplace_back();
                if (!leaf(t, child, (long)~ck[k], cb[k])) return -1;
            } else {
                child = inner(t, ck[k], depth + 1);
                if (child < 0) return -1;
            }
            t.n[child].b = cb[k];
            t.n[me].kids[k] = child;
            u.grow(cb[k]);
        }
        t.n[me].nkids = nc; t.n[me].b = u;
        return me;
    }
};

// Splits leaf `tn` (more than 3 primitives) into a small binary subtree with leaves of at most 3.
void split_leaf(Tree& t, int tn) {
    const int first = t.n[tn].prim_first, count = t.n[tn].prim_count;
    if (count <= 3) return;
    Box cb; cb.reset();
    for (int i = first; i < first + count; ++i) {
        float c[3]; for (int a = 0; a < 3; ++a) c[a] = 0.5f * (t.prim_box[i].lo[a] + t.prim_box[i].hi[a]);
        cb.grow(c);
    }
    int axis = 0; float ext = -1.0f;
    for (int a = 0; a < 3; ++a) if (cb.hi[a] - cb.lo[a] > ext) { ext = cb.hi[a] - cb.lo[a]; axis = a; }
    std::vector<int> order(count);
    for (int i = 0; i < count; ++i) order[i] = first + i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        return t.prim_box[x].lo[axis] + t.prim_box[x].hi[axis] < t.prim_box[y].lo[axis] + t.prim_box[y].hi[axis]; });
    std::vector<int> pr(count); std::vector<Box> pb(count);
    for (int i = 0; i < count; ++i) { pr[i] = t.prims[order[i]]; pb[i] = t.prim_box[order[i]]; }
    for (int i = 0; i < count; ++i) { t.prims[first + i] = pr[i]; t.prim_box[first + i] = pb[i]; }
    // left part: a multiple of 3 close to the median keeps leaves full
    int nl = ((count / 2 + 2) / 3) * 3; if (nl >= count) nl = count / 2; if (nl < 1) nl = 1;
    const int l = (int)t.n.size(); t.n.emplace_back(); const int r = (int)t.n.size(); t.n.emplace_back();
    t.n[l].prim_first = first;      t.n[l].prim_count = nl;
    t.n[r].prim_first = first + nl; t.n[r].prim_count = count - nl;
    for (int side = 0; side < 2; ++side) {
        TNode& c = t.n[side ? r : l]; c.b.reset();
        for (int i = c.prim_first; i < c.prim_first + c.prim_count; ++i) c.b.grow(t.prim_box[i]);
        c.b.clip(t.n[tn].b);
    }
    t.n[tn].nkids = 2; t.n[tn].kids[0] = l; t.n[tn].kids[1] = r; t.n[tn].prim_count = 0;
    split_leaf(t, l); split_leaf(t, r);
}


constexpr float COST_NODE = 1.0f;
// cost of one primitive test relative to one node visit; tuning knob IMBA_COST_PRIM
static const float COST_PRIM = getenv("IMBA_COST_PRIM") ? (float)atof(getenv("IMBA_COST_PRIM")) : 0.8f;

/* SAH-optimal collapse.  Works on a binarised copy of the tree's topology: bin[n] = {left, right} for inner nodes. */
struct Collapse {
    Tree& t;
    std::vector<int> left, right;          // binary children (-1 for leaves)
    std::vector<int> prims_total;          // primitives in the subtree
    std::vector<float> F, D;               // F[n*8 + i], i = 1..7; D[n*9 + j], j = 2..8
    std::vector<unsigned char> Fkind, Dsplit; // Fkind: 0 leaf, 1 inner (i == 1) / 2 distribute, 3 fall back to i-1 (i >= 2)

    explicit Collapse(Tree& tree) : t(tree) {}

    int make_binary(int n) { // returns the id of a node with exactly two children (or a leaf), adding helper nodes for 3-4 children
        TNode& nd = t.n[n];
        if (nd.nkids <= 2) return n;
        // fold: ((k0, k1), k2[, k3]) -- helper nodes get the union boxes
        int a = nd.kids[0];
        for (int k = 1; k + 1 < nd.nkids; ++k) {
            const int h = (int)t.n.size(); t.n.emplace_back();
            TNode& hn = t.n[h]; hn.nkids = 2; hn.kids[0] = a; hn.kids[1] = t.n[n].kids[k];
            hn.b = t.n[a].b; hn.b.grow(t.n[t.n[n].kids[k]].b);
            a = h;
        }
        TNode& nd2 = t.n[n];
        const int last = nd2.kids[nd2.nkids - 1];
        nd2.kids[0] = a; nd2.kids[1] = last; nd2.nkids = 2;
        return n;
    }

    void run(int root) {
        // binarise (iteratively; helper nodes are appended and already binary)
        for (size_t i = 0; i < t.n.size(); ++i) if (t.n[i].nkids > 2) make_binary((int)i);
        const size_t N = t.n.size();
        left.assign(N, -1); right.assign(N, -1); prims_total.assign(N, 0);
        F.assign(N * 8, 0.0f); D.assign(N * 9, 0.0f); Fkind.assign(N * 8, 0); Dsplit.assign(N * 9, 0);
        // post-order over the tree rooted at `root`
        std::vector<int> order, stack{ root };
        while (!stack.empty()) { const int n = stack.back(); stack.pop_back(); order.push_back(n); for (int k = 0; k < t.n[n].nkids; ++k) stack.push_back(t.n[n].kids[k]); }
        for (size_t oi = order.size(); oi-- > 0;) {
            const int n = order[oi]; const TNode& nd = t.n[n];
            const float A = nd.b.area();
            if (nd.nkids == 0) {
                prims_total[n] = nd.prim_count;
                for (int i = 1; i <= 7; ++i) { F[n * 8 + i] = A * nd.prim_count * COST_PRIM; Fkind[n * 8 + i] = 0; }
                continue;
            }
            if (nd.nkids == 1) { // dummy parent of a single leaf: behaves like its child
                const int c = nd.kids[0]; left[n] = c; right[n] = -1; prims_total[n] = prims_total[c];
                for (int i = 1; i <= 7; ++i) { F[n * 8 + i] = F[c * 8 + i]; Fkind[n * 8 + i] = 4; }
                for (int j = 2; j <= 8; ++j) D[n * 9 + j] = F[c * 8 + std::min(j, 7)];
                continue;
            }
            const int l = nd.kids[0], r = nd.kids[1];
            left[n] = l; right[n] = r; prims_total[n] = prims_total[l] + prims_total[r];
            for (int j = 2; j <= 8; ++j) {
                float best = FLT_MAX; int bk = 1;
                for (int k = 1; k < j; ++k) {
                    if (k > 7 || j - k > 7) continue;
                    const float c = F[l * 8 + k] + F[r * 8 + (j - k)];
                    if (c < best) { best = c; bk = k; }
                }
                D[n * 9 + j] = best; Dsplit[n * 9 + j] = (unsigned char)bk;
            }
            const float inner = A * COST_NODE + D[n * 9 + 8];
            const float leaf = prims_total[n] <= 3 ? A * prims_total[n] * COST_PRIM : FLT_MAX;
            if (leaf <= inner) { F[n * 8 + 1] = leaf; Fkind[n * 8 + 1] = 0; } else { F[n * 8 + 1] = inner; Fkind[n * 8 + 1] = 1; }
            for (int i = 2; i <= 7; ++i) {
                if (D[n * 9 + i] < F[n * 8 + i - 1]) { F[n * 8 + i] = D[n * 9 + i]; Fkind[n * 8 + i] = 2; }
                else { F[n * 8 + i] = F[n * 8 + i - 1]; Fkind[n * 8 + i] = 3; }
            }
        }
    }

    void gather_prims(int n, std::vector<int>& prims, std::vector<Box>& boxes) {
        const TNode& nd = t.n[n];
        if (nd.nkids == 0) { for (int i = 0; i < nd.prim_count; ++i) { prims.push_back(t.prims[nd.prim_first + i]); boxes.push_back(t.prim_box[nd.prim_first + i]); } return; }
        for (int k = 0; k < nd.nkids; ++k) gather_prims(nd.kids[k], prims, boxes);
    }

    // Children (TNode ids: leaves or nodes that become 8-wide nodes) of the forest of <= i roots representing subtree n.
    void collect(int n, int i, std::vector<int>& out) {
        const TNode& nd = t.n[n];
        if (nd.nkids == 0) { out.push_back(n); return; }
        if (nd.nkids == 1) { collect(nd.kids[0], i, out); return; }
        if (i > 7) i = 7;
        const unsigned char kind = Fkind[n * 8 + i];
        if (i == 1) {
            if (kind == 1) { out.push_back(n); return; }
            // merged leaf: a new leaf node holding all primitives of the subtree
            std::vector<int> prims; std::vector<Box> boxes;
            gather_prims(n, prims, boxes);
            const int m = (int)t.n.size(); t.n.emplace_back();
            TNode& leaf = t.n[m]; leaf.b = t.n[n].b; leaf.nkids = 0; leaf.prim_first = (int)t.prims.size(); leaf.prim_count = (int)prims.size();
            t.prims.insert(t.prims.end(), prims.begin(), prims.end()); t.prim_box.insert(t.prim_box.end(), boxes.begin(), boxes.end());
            out.push_back(m);
            return;
        }
        if (kind == 3) { collect(n, i - 1, out); return; }
        distribute(n, i, out);
    }
    // Children of subtree n when n itself is NOT a root: split the budget j between its two children.
    void distribute(int n, int j, std::vector<int>& out) {
        const TNode& nd = t.n[n];
        if (nd.nkids == 1) { collect(nd.kids[0], std::min(j, 7), out); return; }
        const int k = Dsplit[n * 9 + j];
        const int l = nd.kids[0], r = nd.kids[1];
        collect(l, k, out); collect(r, j - k, out);
    }
};

inline int ilog2_floor(double v) { int e; frexp(v, &e); return e - 1; } // v in [2^(e-1), 2^e)

struct Emitter {
    Bvh8& out; Tree& t; Collapse* collapse; bool top_level;
    const std::vector<Tri48>* tri_store; const std::vector<const imba_instance_node*>* inst_store;
    const std::map<int, uint32_t>* mesh_roots; // reference root index -> Node8 index (top level only)
    int max_depth = 0;
    uint32_t node_offset = 0; // Node8 index of out.nodes[0] in the final device array (device-built mesh trees come first)

    size_t prim_out_size() const { return top_level ? out.insts.size() : out.tris.size(); }
    void emit_prim(int store_idx) {
        if (top_level) {
            const imba_instance_node* in = (*inst_store)[store_idx];
            Inst64 o; memcpy(o.m, in->transf, 48); o.inst_id = in->id; o.root = (int32_t)mesh_roots->at(in->next); o.pad[0] = o.pad[1] = 0;
            out.insts.push_back(o);
        } else out.tris.push_back((*tri_store)[store_idx]);
    }

    struct Work { int tn; uint32_t idx; int depth; };

    uint32_t run() {
        const uint32_t root_idx = node_offset + (uint32_t)out.nodes.size();
        out.nodes.emplace_back();
        std::vector<Work> q; q.push_back({ t.root, root_idx, 1 });
        for (size_t qi = 0; qi < q.size(); ++qi) {
            const Work w = q[qi];
            max_depth = std::max(max_depth, w.depth);
            // 3. children of this 8-wide node from the collapse decisions
            std::vector<int> ch;
            collapse->distribute(w.tn, 8, ch);
            const int nc = (int)ch.size();
            Box nb; nb.reset();
            for (int c : ch) nb.grow(t.n[c].b);
            // 4. slot assignment
            float cost[8][8]; int slot_of[8]; bool slot_used[8] = { false }, child_done[8] = { false };
            for (int c = 0; c < nc; ++c) for (int s = 0; s < 8; ++s) {
                float v = 0.0f;
                for (int a = 0; a < 3; ++a) {
                    const float cc = 0.5f * (t.n[ch[c]].b.lo[a] + t.n[ch[c]].b.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
                    const float ds = ((s >> (2 - a)) & 1) ? -1.0f : 1.0f; // slot bit 4 = x, 2 = y, 1 = z
                    v += cc * ds;
                }
                cost[c][s] = v;
            }
            for (int it = 0; it < nc; ++it) {
                int bc = -1, bs = -1; float bv = FLT_MAX;
                for (int c = 0; c < nc; ++c) if (!child_done[c]) for (int s = 0; s < 8; ++s) if (!slot_used[s] && cost[c][s] < bv) { bv = cost[c][s]; bc = c; bs = s; }
                if (bc < 0) { for (int c = 0; c < nc && bc < 0; ++c) if (!child_done[c]) bc = c; for (int s = 0; s < 8 && bs < 0; ++s) if (!slot_used[s]) bs = s; } // NaN boxes
                slot_of[bc] = bs; child_done[bc] = true; slot_used[bs] = true;
            }
            int child_in_slot[8]; for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
            for (int c = 0; c < nc; ++c) child_in_slot[slot_of[c]] = ch[c];
            // 5. quantisation grid
            Node8 nd; memset(&nd, 0, sizeof(nd));
            double scale[3];
            for (int a = 0; a < 3; ++a) {
                const float lo = nb.lo[a], hi = nb.hi[a];
                const double ext = (double)hi - (double)lo;
                float p = (float)((double)lo - ext * (1.0 / 1024.0));
                while (p > lo) p = nextafterf(p, -FLT_MAX);
                if (!(ext > 0.0)) p = lo;
                const double span = (double)hi - (double)p;
                int e = span > 0.0 ? ilog2_floor(span / 254.0) : -120;
                if (e < -120) e = -120;
                for (;; ++e) {
                    if (e >= 127) break;
                    const double s = ldexp(1.0, e);
                    if (span / s + 1.0 / 64.0 > 254.5) continue;
                    // keep the grid at or above float resolution: 2^e >= ulp of the largest magnitude on the grid
                    const double m = std::max(fabs((double)p), fabs((double)p + 255.0 * s));
                    if (m > 0.0 && e < ilog2_floor(m) - 23) continue;
                    break;
                }
                if (e > 127) e = 127;
                nd.p[a] = p; nd.e[a] = (uint8_t)(e + 127); scale[a] = ldexp(1.0, e);
            }
            // children
            int n_inner = 0; for (int s = 0; s < 8; ++s) if (child_in_slot[s] >= 0 && t.n[child_in_slot[s]].nkids > 0) ++n_inner;
            nd.child_base = node_offset + (uint32_t)out.nodes.size();
            out.nodes.resize(out.nodes.size() + n_inner);
            nd.prim_base = (uint32_t)prim_out_size();
            int inner_rank = 0, prim_off = 0;
            uint8_t* qlo[3] = { nd.qlo_x, nd.qlo_y, nd.qlo_z }; uint8_t* qhi[3] = { nd.qhi_x, nd.qhi_y, nd.qhi_z };
            for (int s = 0; s < 8; ++s) {
                const int c = child_in_slot[s];
                if (c < 0) { for (int a = 0; a < 3; ++a) { qlo[a][s] = 255; qhi[a][s] = 0; } continue; }
                const TNode& cn = t.n[c];
                for (int a = 0; a < 3; ++a) {
                    double l = floor(((double)cn.b.lo[a] - (double)nd.p[a]) / scale[a] - 1.0 / 64.0);
                    double h = ceil(((double)cn.b.hi[a] - (double)nd.p[a]) / scale[a] + 1.0 / 64.0);
                    if (!(l >= 0.0)) l = 0.0;
                    if (!(h <= 255.0)) h = 255.0;
                    if (l > 255.0) l = 255.0;
                    if (h < 0.0) h = 0.0;
                    // verify with the device's own float decode
                    const float sf = (float)scale[a];
                    while (l > 0.0 && fmaf((float)l, sf, nd.p[a]) > cn.b.lo[a]) l -= 1.0;
                    while (h < 255.0 && fmaf((float)h, sf, nd.p[a]) < cn.b.hi[a]) h += 1.0;
                    qlo[a][s] = (uint8_t)l; qhi[a][s] = (uint8_t)h;
                }
                if (cn.nkids > 0) {
                    nd.imask |= (uint8_t)(1u << s);
                    nd.meta[s] = (uint8_t)((1u << 5) | (24u + s));
                    q.push_back({ c, nd.child_base + (uint32_t)inner_rank, w.depth + 1 });
                    ++inner_rank;
                } else {
                    const int cnt = cn.prim_count; // 1..3 after split_leaf
                    const uint8_t unary = cnt >= 3 ? 7 : (cnt == 2 ? 3 : 1);
                    nd.meta[s] = (uint8_t)((unary << 5) | prim_off);
                    for (int i = 0; i < cnt && i < 3; ++i) emit_prim(t.prims[cn.prim_first + i]);
                    prim_off += std::min(cnt, 3);
                }
            }
            out.nodes[w.idx - node_offset] = nd;
        }
        return root_idx;
    }
};

bool build_tree(Parser& ps, int ref_root, Tree& t) {
    t.root = ps.inner(t, ref_root, 0);
    if (t.root < 0) return false;
    const size_t n0 = t.n.size();
    for (size_t i = 0; i < n0; ++i) if (t.n[i].nkids == 0 && t.n[i].prim_count > 3) split_leaf(t, (int)i);
    return true;
}

} // namespace

bool imba_build_bvh8(int layout, int root, const void* nodes, const imba_vec4* tris,
                     const imba_instance_node* inst_nodes, Bvh8& out, std::string& err) {
    out = Bvh8();
    if (layout != IMBA_ACCEL_MBVH4 && layout != IMBA_ACCEL_BVH2) { err = "unknown accel layout"; return false; }
    if (!nodes || !tris || !inst_nodes) { err = "null acceleration structure pointer"; return false; }
    // top level
    std::vector<const imba_instance_node*> inst_store;
    Parser top{ layout, nodes, tris, inst_nodes, &err };
    top.top_level = true; top.inst_store = &inst_store;
    Tree tt;
    if (!build_tree(top, root, tt)) return false;
    int visited = top.visited;
    // meshes
    std::map<int, uint32_t> mesh_roots; int mesh_depth = 0;
    for (const imba_instance_node* in : inst_store) {
        if (mesh_roots.count(in->next)) continue;
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<Tri48> tri_store;
        Parser mp{ layout, nodes, tris, inst_nodes, &err };
        mp.tri_store = &tri_store;
        Tree mt;
        if (!build_tree(mp, in->next, mt)) return false;
        visited += mp.visited;
        const auto t1 = std::chrono::steady_clock::now();
        Collapse mc(mt); mc.run(mt.root);
        const auto t2 = std::chrono::steady_clock::now();
        Emitter em{ out, mt, &mc, false, &tri_store, nullptr, nullptr };
        mesh_roots[in->next] = em.run();
        mesh_depth = std::max(mesh_depth, em.max_depth);
        if (getenv("IMBA_VERBOSE")) {
            const auto t3 = std::chrono::steady_clock::now();
            auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[imba] bvh8 mesh: %zu tree nodes; parse %.0f ms, collapse %.0f ms, emit %.0f ms\n", mt.n.size(), ms(t0, t1), ms(t1, t2), ms(t2, t3));
        }
    }
    Collapse tc(tt); tc.run(tt.root);
    Emitter em{ out, tt, &tc, true, nullptr, &inst_store, &mesh_roots };
    out.top_root = em.run();
    out.max_depth = em.max_depth + mesh_depth;
    out.ref_node_count = visited;
    if (out.insts.size() == 1) {
        static const float ident[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
        if (memcmp(out.insts[0].m, ident, sizeof(ident)) == 0) { out.flat = true; out.flat_root = (uint32_t)out.insts[0].root; out.flat_inst = out.insts[0].inst_id; }
    }
    // one padding triangle so that 3-wide speculative loads past the last leaf stay in bounds
    Tri48 pad; memset(&pad, 0, sizeof(pad)); pad.id = (int32_t)0x80000000; out.tris.push_back(pad);
    return true;
}

/* Top level only: the mesh trees already exist (built on the device, bvh8_gpu.cu).  `mesh_roots` maps InstanceNode::next to the
 * Node8 index of that mesh's root; the top-level nodes produced here will be stored from index `node_offset` on. */
bool imba_build_bvh8_top(int layout, int root, const void* nodes, const imba_instance_node* inst_nodes,
                         const std::map<int, uint32_t>& mesh_roots, uint32_t node_offset, Bvh8& out, std::string& err) {
    out = Bvh8();
    if (layout != IMBA_ACCEL_MBVH4 && layout != IMBA_ACCEL_BVH2) { err = "unknown accel layout"; return false; }
    if (!nodes || !inst_nodes) { err = "null acceleration structure pointer"; return false; }
    std::vector<const imba_instance_node*> inst_store;
    Parser top{ layout, nodes, nullptr, inst_nodes, &err };
    top.top_level = true; top.inst_store = &inst_store;
    Tree tt;
    if (!build_tree(top, root, tt)) return false;
    for (const imba_instance_node* in : inst_store) if (!mesh_roots.count(in->next)) { err = "instance refers to an unknown mesh"; return false; }
    Collapse tc(tt); tc.run(tt.root);
    Emitter em{ out, tt, &tc, true, nullptr, &inst_store, &mesh_roots };
    em.node_offset = node_offset;
    out.top_root = em.run();
    out.max_depth = em.max_depth;
    out.ref_node_count = top.visited;
    if (out.insts.size() == 1) {
        static const float ident[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
        if (memcmp(out.insts[0].m, ident, sizeof(ident)) == 0) { out.flat = true; out.flat_root = (uint32_t)out.insts[0].root; out.flat_inst = out.insts[0].inst_id; }
    }
    return true;
}
