/* ORACLE / TEST INFRASTRUCTURE -- never linked into or called by the product path.
 *
 * CPU restatement of the four traversal entry points the reference binds from the
 * external AnyDSL/traversal library (core/traversal_interface.h:4-10; call sites
 * render/scheduling/ray_queue.h:254-265, 279-290, 304-315, 329-340).
 *
 * PARITY UNPINNED BY SOURCE: the kernels' source is not in /root/reference (unpinned
 * third-party dependency AnyDSL/traversal, found via CMakeLists.txt:35-37). What IS
 * pinned, and what this file follows:
 *   - MBVH4 node / 4-triangle SoA packet / sentinel encoding:   core/cpu_adapter.cpp:56-208
 *   - BVH2 node / 3-Vec4 triangle / sentinel encoding:           core/gpu_adapter.cpp:62-143
 *   - top-level nodes, InstanceNode lists, pad[0]==-1 sentinel:  core/cpu_adapter.cpp:268-367
 *   - root = node_count, mesh roots in InstanceNode::next:       render/scene.cpp:168-176
 *   - consumer conventions (tri_id<0 miss, tmax=t, u = weight of v1): render/integrators/integrator.h:49-101
 *   - opacity masks (1 byte/texel, desc (1,1,0) = opaque):       core/mask.h:9-40
 * The arithmetic that is NOT pinned is fixed here and mirrored instruction for
 * instruction by the CUDA kernels (imbatracer_b200/csrc/traverse.cu):
 *   - instance transform and triangle test use explicit fmaf in a fixed order
 *     (file must be compiled with -ffp-contract=off so nothing else is fused);
 *   - triangle test = Embree/AnyDSL style: c=v0-o, r=d x c, det=n.d, U=(r.e2)*sgn, V=(r.e1)*sgn,
 *     T=(n.c)*sgn, accept U>=0, V>=0, U+V<=|det|, |det|*tmin<=T<=|det|*tmax, det!=0;
 *     t=T*(1/|det|), u=U*(1/|det|);
 *   - a candidate replaces the current best iff (t, tri_id, inst_id) is lexicographically smaller, with t the
 *     ROUNDED value T*(1/|det|); the test T <= |det|*best.t*(1+1e-6) is only a pre-filter.  The result therefore
 *     does not depend on tree shape or visit order (the reference scenes contain exactly duplicated triangles,
 *     e.g. CornellBox-Original triangles 30/32 and 31/33);
 *   - box test is the robust slab test (Ize 2013): tfar is scaled by (1+2^-21) before the
 *     compare, zero direction components are replaced by +-2^-80.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>

namespace {

struct Vec4 { float x, y, z, w; };
struct Vec2 { float x, y; };
struct Ray  { Vec4 org; Vec4 dir; };
struct Hit  { int tri_id; int inst_id; float tmax; float u; };
struct InstanceNode { float transf[3][4]; int id; int next; int pad[2]; };
struct TransparencyMask { int width; int height; int offset; };
struct Node4 { float min_x[4], min_y[4], min_z[4], max_x[4], max_y[4], max_z[4]; int children[4]; };
struct BBox2 { float lo_x, lo_y, lo_z, hi_x, hi_y, hi_z; };
struct Node2 { BBox2 left_bb, right_bb; int left, right; };

static_assert(sizeof(Node4) == 112 && sizeof(Node2) == 56 && sizeof(InstanceNode) == 64 && sizeof(Ray) == 32 && sizeof(Hit) == 16, "ABI sizes");

inline int   f2i(float f) { int i; memcpy(&i, &f, 4); return i; }

struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
// cross and dot with the fused pattern shared with the device code
inline V3 crossf(V3 a, V3 b) {
    return { fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)) };
}
inline float dotf(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
inline float safe_rcp(float d) {
    const float eps = 8.2718061e-25f; // 2^-80
    return 1.0f / (fabsf(d) > eps ? d : copysignf(eps, d));
}

struct Scene {
    int root; const void* nodes; const InstanceNode* instances; const Vec4* tris;
    const int* indices; const Vec2* texcoords; const TransparencyMask* masks; const char* mask_buffer;
};

struct Counters { uint32_t nodes = 0, tri_blocks = 0, tris = 0, instances = 0; };

struct LocalRay {
    V3 o, d, id; float tmin;
};

struct Best { float t; float u; int tri; int inst; };
constexpr int NO_HIT = 0x7fffffff; // initial (tri, inst): any candidate with t == tmax still wins the tie
constexpr float PREFILTER = 1.000001f;

inline bool better(float t, int tri, int inst, const Best& b) {
    return t < b.t || (t == b.t && (tri < b.tri || (tri == b.tri && inst < b.inst)));
}

// Alpha test. (u, v) are the weights of v1, v2. indices: 4 ints per global triangle (i0,i1,i2,material),
// render/scene.cpp:89-93.
inline bool mask_opaque(const Scene& s, int tri_id, float u, float v) {
    if (!s.masks || !s.indices) return true;
    const int mat = s.indices[tri_id * 4 + 3];
    const TransparencyMask m = s.masks[mat];
    if (m.width == 1 && m.height == 1 && m.offset == 0) return true;
    const Vec2 a = s.texcoords[s.indices[tri_id * 4 + 0]];
    const Vec2 b = s.texcoords[s.indices[tri_id * 4 + 1]];
    const Vec2 c = s.texcoords[s.indices[tri_id * 4 + 2]];
    const float w = 1.0f - u - v;
    float tu = a.x * w + b.x * u + c.x * v;
    float tv = a.y * w + b.y * u + c.y * v;
    tu = tu - floorf(tu);
    tv = 1.0f - (tv - floorf(tv)); // images are stored top-down, texture_sampler.h:23
    int x = (int)(tu * m.width);  if (x > m.width  - 1) x = m.width  - 1; if (x < 0) x = 0;
    int y = (int)(tv * m.height); if (y > m.height - 1) y = m.height - 1; if (y < 0) y = 0;
    return s.mask_buffer[m.offset + y * m.width + x] != 0;
}

// Returns true if the hit was accepted (best updated).
inline bool tri_test(const Scene& s, const LocalRay& r, V3 v0, V3 e1, V3 e2, V3 n, int tri_id, int inst_id, Best& best) {
    const V3 c = sub(v0, r.o);
    const V3 rr = crossf(r.d, c);
    const float det = dotf(n, r.d);
    const float adet = fabsf(det);
    const uint32_t sgn = (uint32_t)f2i(det) & 0x80000000u;
    auto xs = [sgn](float f) { uint32_t i; memcpy(&i, &f, 4); i ^= sgn; float o; memcpy(&o, &i, 4); return o; };
    const float U = xs(dotf(rr, e2));
    const float V = xs(dotf(rr, e1));
    if (!(U >= 0.0f) || !(V >= 0.0f) || !(U + V <= adet) || det == 0.0f) return false;
    const float T = xs(dotf(n, c));
    // upper bound: a pre-filter only (never rejects a candidate with t <= best.t, see DESIGN.md "tie rule");
    // the decision is taken on the rounded t below so that it does not depend on the visiting order
    if (!(T >= adet * r.tmin) || !(T <= adet * best.t * PREFILTER)) return false;
    const float inv = 1.0f / adet;
    const float t = T * inv, u = U * inv, v = V * inv;
    if (!better(t, tri_id, inst_id, best)) return false;
    if (!mask_opaque(s, tri_id, u, v)) return false;
    best.t = t; best.u = u; best.tri = tri_id; best.inst = inst_id;
    return true;
}

constexpr float ROBUST = 1.00000048f; // 1 + 2^-21, applied to tfar

// ---------------------------------------------------------------------------------------------
// MBVH4 ("cpu") layout
// ---------------------------------------------------------------------------------------------
template <bool ANY>
bool traverse_mesh4(const Scene& s, const Node4* nodes, int root, const LocalRay& r, int inst_id, Best& best, Counters* cnt) {
    int stack[128]; int sp = 0;
    stack[sp++] = root;
    while (sp) {
        const int id = stack[--sp];
        if (id > 0 || id == root) { // inner (root may be index 0)
            const Node4& n = nodes[id];
            if (cnt) cnt->nodes++;
            float dist[4]; int kid[4]; int k = 0;
            for (int i = 0; i < 4; ++i) {
                const int c = n.children[i];
                if (c == 0) break;
                const float t0x = (n.min_x[i] - r.o.x) * r.id.x, t1x = (n.max_x[i] - r.o.x) * r.id.x;
                const float t0y = (n.min_y[i] - r.o.y) * r.id.y, t1y = (n.max_y[i] - r.o.y) * r.id.y;
                const float t0z = (n.min_z[i] - r.o.z) * r.id.z, t1z = (n.max_z[i] - r.o.z) * r.id.z;
                const float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), r.tmin));
                const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), best.t));
                if (tn <= tf * ROBUST) { dist[k] = tn; kid[k] = c; ++k; }
            }
            // push far-to-near so that the nearest child is popped first
            for (int a = 1; a < k; ++a) { float d = dist[a]; int c = kid[a]; int b = a - 1; while (b >= 0 && dist[b] < d) { dist[b + 1] = dist[b]; kid[b + 1] = kid[b]; --b; } dist[b + 1] = d; kid[b + 1] = c; }
            for (int a = 0; a < k; ++a) stack[sp++] = kid[a];
        } else {
            // leaf: 13-Vec4 packets of 4 triangles (SoA) until a sentinel Vec4 (cpu_adapter.cpp:130-206)
            const Vec4* p = s.tris + ~id;
            for (;;) {
                if (cnt) cnt->tri_blocks++;
                const float* f = (const float*)p;
                for (int l = 0; l < 4; ++l) {
                    const int tid = f2i(f[48 + l]);
                    if (tid < 0) continue; // 0x80000000 = padding lane
                    if (cnt) cnt->tris++;
                    const V3 v0{ f[0 + l], f[4 + l], f[8 + l] }, e1{ f[12 + l], f[16 + l], f[20 + l] };
                    const V3 e2{ f[24 + l], f[28 + l], f[32 + l] }, nn{ f[36 + l], f[40 + l], f[44 + l] };
                    if (tri_test(s, r, v0, e1, e2, nn, tid, inst_id, best) && ANY) return true;
                }
                p += 13;
                if ((uint32_t)f2i(p->x) == 0x80000000u && (uint32_t)f2i(p->y) == 0x80000000u &&
                    (uint32_t)f2i(p->z) == 0x80000000u && (uint32_t)f2i(p->w) == 0x80000000u) break;
            }
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// BVH2 ("gpu") layout
// ---------------------------------------------------------------------------------------------
template <bool ANY>
bool traverse_mesh2(const Scene& s, const Node2* nodes, int root, const LocalRay& r, int inst_id, Best& best, Counters* cnt) {
    int stack[128]; int sp = 0;
    stack[sp++] = root;
    while (sp) {
        const int id = stack[--sp];
        if (id > 0 || id == root) {
            const Node2& n = nodes[id];
            if (cnt) cnt->nodes++;
            float dist[2]; int kid[2]; int k = 0;
            for (int i = 0; i < 2; ++i) {
                const BBox2& b = i ? n.right_bb : n.left_bb;
                const int c = i ? n.right : n.left;
                if (i == 1 && c == 0x76543210) break; // dummy parent of a single-leaf mesh (gpu_adapter.cpp:14-31)
                const float t0x = (b.lo_x - r.o.x) * r.id.x, t1x = (b.hi_x - r.o.x) * r.id.x;
                const float t0y = (b.lo_y - r.o.y) * r.id.y, t1y = (b.hi_y - r.o.y) * r.id.y;
                const float t0z = (b.lo_z - r.o.z) * r.id.z, t1z = (b.hi_z - r.o.z) * r.id.z;
                const float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), r.tmin));
                const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), best.t));
                if (tn <= tf * ROBUST) { dist[k] = tn; kid[k] = c; ++k; }
            }
            if (k == 2 && dist[0] < dist[1]) std::swap(kid[0], kid[1]);
            for (int a = 0; a < k; ++a) stack[sp++] = kid[a];
        } else {
            // leaf: (v0, v1[w = id], v2) triples; last v2.w == 0x80000000 ends the leaf (gpu_adapter.cpp:121-135)
            const Vec4* p = s.tris + ~id;
            for (;;) {
                if (cnt) { cnt->tri_blocks++; cnt->tris++; }
                const V3 v0{ p[0].x, p[0].y, p[0].z }, v1{ p[1].x, p[1].y, p[1].z }, v2{ p[2].x, p[2].y, p[2].z };
                const int tid = f2i(p[1].w);
                // same derived quantities as the CPU layout stores (cpu_adapter.cpp:146-149); unfused cross
                const V3 e1 = sub(v0, v1), e2 = sub(v2, v0);
                const V3 nn{ e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x };
                if (tri_test(s, r, v0, e1, e2, nn, tid, inst_id, best) && ANY) return true;
                if ((uint32_t)f2i(p[2].w) == 0x80000000u) break;
                p += 3;
            }
        }
    }
    return false;
}

inline LocalRay to_instance(const InstanceNode& in, const Ray& ray) {
    LocalRay r;
    const float (*m)[4] = in.transf;
    r.o = { fmaf(m[0][2], ray.org.z, fmaf(m[0][1], ray.org.y, fmaf(m[0][0], ray.org.x, m[0][3]))),
            fmaf(m[1][2], ray.org.z, fmaf(m[1][1], ray.org.y, fmaf(m[1][0], ray.org.x, m[1][3]))),
            fmaf(m[2][2], ray.org.z, fmaf(m[2][1], ray.org.y, fmaf(m[2][0], ray.org.x, m[2][3]))) };
    r.d = { fmaf(m[0][2], ray.dir.z, fmaf(m[0][1], ray.dir.y, m[0][0] * ray.dir.x)),
            fmaf(m[1][2], ray.dir.z, fmaf(m[1][1], ray.dir.y, m[1][0] * ray.dir.x)),
            fmaf(m[2][2], ray.dir.z, fmaf(m[2][1], ray.dir.y, m[2][0] * ray.dir.x)) };
    r.id = { safe_rcp(r.d.x), safe_rcp(r.d.y), safe_rcp(r.d.z) };
    r.tmin = ray.org.w;
    return r;
}

template <bool ANY, bool WIDE4>
void trace_one(const Scene& s, const Ray& ray, Hit& hit, Counters* cnt) {
    Best best{ ray.dir.w, 0.0f, NO_HIT, NO_HIT };
    LocalRay w;
    w.o = { ray.org.x, ray.org.y, ray.org.z }; w.d = { ray.dir.x, ray.dir.y, ray.dir.z };
    w.id = { safe_rcp(w.d.x), safe_rcp(w.d.y), safe_rcp(w.d.z) }; w.tmin = ray.org.w;
    bool found = false;

    int stack[128]; int sp = 0;
    stack[sp++] = s.root;
    while (sp && !found) {
        const int id = stack[--sp];
        if (id > 0 || id == s.root) {
            if (cnt) cnt->nodes++;
            float dist[4]; int kid[4]; int k = 0;
            const int width = WIDE4 ? 4 : 2;
            for (int i = 0; i < width; ++i) {
                float lo[3], hi[3]; int c;
                if (WIDE4) {
                    const Node4& n = ((const Node4*)s.nodes)[id];
                    c = n.children[i]; if (c == 0) break;
                    lo[0] = n.min_x[i]; lo[1] = n.min_y[i]; lo[2] = n.min_z[i]; hi[0] = n.max_x[i]; hi[1] = n.max_y[i]; hi[2] = n.max_z[i];
                } else {
                    const Node2& n = ((const Node2*)s.nodes)[id];
                    const BBox2& b = i ? n.right_bb : n.left_bb; c = i ? n.right : n.left;
                    if (i == 1 && c == 0x76543210) break;
                    lo[0] = b.lo_x; lo[1] = b.lo_y; lo[2] = b.lo_z; hi[0] = b.hi_x; hi[1] = b.hi_y; hi[2] = b.hi_z;
                }
                const float t0x = (lo[0] - w.o.x) * w.id.x, t1x = (hi[0] - w.o.x) * w.id.x;
                const float t0y = (lo[1] - w.o.y) * w.id.y, t1y = (hi[1] - w.o.y) * w.id.y;
                const float t0z = (lo[2] - w.o.z) * w.id.z, t1z = (hi[2] - w.o.z) * w.id.z;
                const float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), w.tmin));
                const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), best.t));
                if (tn <= tf * ROBUST) { dist[k] = tn; kid[k] = c; ++k; }
            }
            for (int a = 1; a < k; ++a) { float d = dist[a]; int c = kid[a]; int b = a - 1; while (b >= 0 && dist[b] < d) { dist[b + 1] = dist[b]; kid[b + 1] = kid[b]; --b; } dist[b + 1] = d; kid[b + 1] = c; }
            for (int a = 0; a < k; ++a) stack[sp++] = kid[a];
        } else {
            // top-level leaf: list of InstanceNodes ending with pad[0] == -1 (cpu_adapter.cpp:343-365)
            const InstanceNode* in = s.instances + ~id;
            for (;;) {
                if (cnt) cnt->instances++;
                const LocalRay r = to_instance(*in, ray);
                if (WIDE4) found = traverse_mesh4<ANY>(s, (const Node4*)s.nodes, in->next, r, in->id, best, cnt);
                else       found = traverse_mesh2<ANY>(s, (const Node2*)s.nodes, in->next, r, in->id, best, cnt);
                if (found || in->pad[0] == -1) break;
                ++in;
            }
        }
    }
    if (best.tri != NO_HIT) { hit.tri_id = best.tri; hit.inst_id = best.inst; hit.tmax = best.t; hit.u = best.u; }
    else               { hit.tri_id = -1; hit.inst_id = -1; hit.tmax = ray.dir.w; hit.u = 0.0f; }
}

int g_threads = 0; // 0 = hardware_concurrency

template <bool ANY, bool WIDE4>
void trace_all(const Scene& s, const Ray* rays, Hit* hits, int count, uint32_t* counters /* 4 per ray or null */) {
    int nt = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (count < 4096) nt = 1;
    std::atomic<int> next(0);
    const int chunk = 1024;
    auto work = [&]() {
        for (;;) {
            const int b = next.fetch_add(chunk);
            if (b >= count) break;
            const int e = std::min(count, b + chunk);
            for (int i = b; i < e; ++i) {
                Counters c;
                trace_one<ANY, WIDE4>(s, rays[i], hits[i], counters ? &c : nullptr);
                if (counters) { counters[4 * i] = c.nodes; counters[4 * i + 1] = c.tri_blocks; counters[4 * i + 2] = c.tris; counters[4 * i + 3] = c.instances; }
            }
        }
    };
    if (nt == 1) { work(); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < nt - 1; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

} // namespace

extern "C" {

void oracle_set_threads(int n) { g_threads = n; }
int  oracle_get_threads() { int nt = g_threads > 0 ? g_threads : (int)std::thread::hardware_concurrency(); return nt < 1 ? 1 : nt; }

#define ENTRY(NAME, ANY, WIDE4)                                                                                          \
    void NAME(int root, void* nodes, void* instances, void* tris, void* rays, void* hits, int* indices, void* texcoords, \
              void* masks, char* mask_buffer, int count) {                                                               \
        Scene s{ root, nodes, (const InstanceNode*)instances, (const Vec4*)tris, indices, (const Vec2*)texcoords,        \
                 (const TransparencyMask*)masks, mask_buffer };                                                          \
        trace_all<ANY, WIDE4>(s, (const Ray*)rays, (Hit*)hits, count, nullptr);                                          \
    }                                                                                                                    \
    void NAME##_counted(int root, void* nodes, void* instances, void* tris, void* rays, void* hits, int* indices,        \
                        void* texcoords, void* masks, char* mask_buffer, int count, uint32_t* counters) {                \
        Scene s{ root, nodes, (const InstanceNode*)instances, (const Vec4*)tris, indices, (const Vec2*)texcoords,        \
                 (const TransparencyMask*)masks, mask_buffer };                                                          \
        trace_all<ANY, WIDE4>(s, (const Ray*)rays, (Hit*)hits, count, counters);                                         \
    }

ENTRY(intersect_cpu_masked_instanced, false, true)
ENTRY(occluded_cpu_masked_instanced,  true,  true)
ENTRY(intersect_gpu_masked_instanced, false, false)
ENTRY(occluded_gpu_masked_instanced,  true,  false)

} // extern "C"
