/* ORACLE / TEST INFRASTRUCTURE -- never linked into or called by the product path.
 *
 * Scalar CPU walker over the product's converted device layout (imbatracer_b200/csrc/bvh8.h), i.e. a restatement
 * of the control flow of imbatracer_b200/csrc/traverse.cu.  It lets the `-m "not gpu"` tests check the HOST logic
 * (the MBVH4/BVH2 -> Node8 converter: conservative quantisation, completeness, slot/meta encoding, instance
 * records) against oracle/traversal_oracle.cpp without a GPU.  Compile with -ffp-contract=off.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include "../imbatracer_b200/csrc/bvh8.h"

namespace {
struct Ray { float ox, oy, oz, tmin, dx, dy, dz, tmax; };
struct Hit { int tri, inst; float t, u; };
inline float safe_rcp(float d) { return 1.0f / (fabsf(d) > IMBA_RCP_EPS ? d : copysignf(IMBA_RCP_EPS, d)); }
inline float dotf(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(az, bz, fmaf(ay, by, ax * bx)); }
inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline float xorf(float f, uint32_t s) { return bitsf(fbits(f) ^ s); }
struct R { float ox, oy, oz, dx, dy, dz, ix, iy, iz; uint32_t octinv;
    void finish() { ix = safe_rcp(dx); iy = safe_rcp(dy); iz = safe_rcp(dz);
        octinv = 7u - ((ix < 0 ? 4u : 0u) | (iy < 0 ? 2u : 0u) | (iz < 0 ? 1u : 0u)); } };
}

extern "C" int bvh8_walk(const Node8* nodes, const Tri48* tris, const Inst64* insts, uint32_t top_root,
                         const float* rays, int32_t* hits, int count, int any_hit, uint32_t* stats /* 4 per ray or null */) {
    for (int i = 0; i < count; ++i) {
        Ray ray; memcpy(&ray, rays + 8 * i, 32);
        R r{ ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz }; r.finish();
        const float tmin = ray.tmin; float tbest = ray.tmax, ubest = 0.0f; int tri_best = 0x7fffffff, inst_best = 0x7fffffff, cur_inst = -1;
        uint32_t st[4] = { 0, 0, 0, 0 };
        struct G { uint32_t x, y; } stack[128]; int sp = 0;
        G ng{ top_root, 0x80000000u }, tg{ 0, 0 };
        bool done = false;
        while (!done) {
            if (ng.y > 0x00ffffffu) {
                const uint32_t hitsm = ng.y; const uint32_t bit = 31u - (uint32_t)__builtin_clz(hitsm);
                ng.y &= ~(1u << bit);
                if (ng.y > 0x00ffffffu) stack[sp++] = ng;
                const uint32_t slot = (bit - 24u) ^ r.octinv;
                const uint32_t rel = (uint32_t)__builtin_popcount(hitsm & 0xffu & ~(0xffffffffu << slot));
                const Node8& n = nodes[ng.x + rel]; ++st[0];
                const float adj[3] = { bitsf((uint32_t)n.e[0] << 23) * r.ix, bitsf((uint32_t)n.e[1] << 23) * r.iy, bitsf((uint32_t)n.e[2] << 23) * r.iz };
                const float org[3] = { (n.p[0] - r.ox) * r.ix, (n.p[1] - r.oy) * r.iy, (n.p[2] - r.oz) * r.iz };
                uint32_t hitmask = 0;
                for (int s = 0; s < 8; ++s) {
                    const uint32_t meta = n.meta[s];
                    const bool inner = (meta & 0x18u) == 0x18u;
                    const uint32_t bit_index = (inner ? (meta ^ r.octinv) : meta) & 0x1fu, child_bits = meta >> 5;
                    const float lx = fmaf((float)n.qlo_x[s], adj[0], org[0]), hx = fmaf((float)n.qhi_x[s], adj[0], org[0]);
                    const float ly = fmaf((float)n.qlo_y[s], adj[1], org[1]), hy = fmaf((float)n.qhi_y[s], adj[1], org[1]);
                    const float lz = fmaf((float)n.qlo_z[s], adj[2], org[2]), hz = fmaf((float)n.qhi_z[s], adj[2], org[2]);
                    const float t0x = r.ix < 0 ? hx : lx, t1x = r.ix < 0 ? lx : hx;
                    const float t0y = r.iy < 0 ? hy : ly, t1y = r.iy < 0 ? ly : hy;
                    const float t0z = r.iz < 0 ? hz : lz, t1z = r.iz < 0 ? lz : hz;
                    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
                    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tbest));
                    if (tn <= tf * IMBA_ROBUST_TFAR) hitmask |= child_bits << bit_index;
                }
                ng = G{ n.child_base, (hitmask & 0xff000000u) | n.imask };
                tg = G{ n.prim_base, hitmask & 0x00ffffffu };
            } else { tg = ng; ng = G{ 0, 0 }; }
            while (tg.y) {
                const uint32_t bit = 31u - (uint32_t)__builtin_clz(tg.y); tg.y &= ~(1u << bit);
                if (cur_inst < 0) {
                    if (ng.y > 0x00ffffffu) stack[sp++] = ng;
                    if (tg.y) stack[sp++] = tg;
                    stack[sp++] = G{ 0xffffffffu, 0 }; ++st[2];
                    if ((uint32_t)sp > st[3]) st[3] = sp;
                    const Inst64& in = insts[tg.x + bit]; const float* m = in.m;
                    R w{ ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz };
                    r.ox = fmaf(m[2], w.oz, fmaf(m[1], w.oy, fmaf(m[0], w.ox, m[3])));
                    r.oy = fmaf(m[6], w.oz, fmaf(m[5], w.oy, fmaf(m[4], w.ox, m[7])));
                    r.oz = fmaf(m[10], w.oz, fmaf(m[9], w.oy, fmaf(m[8], w.ox, m[11])));
                    r.dx = fmaf(m[2], w.dz, fmaf(m[1], w.dy, m[0] * w.dx));
                    r.dy = fmaf(m[6], w.dz, fmaf(m[5], w.dy, m[4] * w.dx));
                    r.dz = fmaf(m[10], w.dz, fmaf(m[9], w.dy, m[8] * w.dx));
                    r.finish(); cur_inst = in.inst_id;
                    ng = G{ (uint32_t)in.root, 0x80000000u }; tg = G{ 0, 0 };
                    break;
                }
                const Tri48& tr = tris[tg.x + bit]; ++st[1];
                const float* b = tr.e1; const float* c = tr.e2;
                const float nx = b[1] * c[2] - b[2] * c[1], ny = b[2] * c[0] - b[0] * c[2], nz = b[0] * c[1] - b[1] * c[0];
                const float cx = tr.v0[0] - r.ox, cy = tr.v0[1] - r.oy, cz = tr.v0[2] - r.oz;
                const float rx = fmaf(r.dy, cz, -(r.dz * cy)), ry = fmaf(r.dz, cx, -(r.dx * cz)), rz = fmaf(r.dx, cy, -(r.dy * cx));
                const float det = dotf(nx, ny, nz, r.dx, r.dy, r.dz), adet = fabsf(det);
                const uint32_t sgn = fbits(det) & 0x80000000u;
                const float U = xorf(dotf(rx, ry, rz, c[0], c[1], c[2]), sgn), V = xorf(dotf(rx, ry, rz, b[0], b[1], b[2]), sgn);
                if (!(U >= 0.0f) || !(V >= 0.0f) || !(U + V <= adet) || det == 0.0f) continue;
                const float T = xorf(dotf(nx, ny, nz, cx, cy, cz), sgn);
                if (!(T >= adet * tmin) || !(T <= adet * tbest * 1.000001f)) continue;
                const float inv = 1.0f / adet, t = T * inv, u = U * inv;
                if (!(t < tbest || (t == tbest && (tr.id < tri_best || (tr.id == tri_best && cur_inst < inst_best))))) continue;
                tbest = t; ubest = u; tri_best = tr.id; inst_best = cur_inst;
                if (any_hit) { done = true; break; }
            }
            if (!done && ng.y <= 0x00ffffffu && tg.y == 0) {
                for (;;) {
                    if (sp == 0) { done = true; break; }
                    const G e = stack[--sp];
                    if (e.y == 0) { r = R{ ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz }; r.finish(); cur_inst = -1; continue; }
                    ng = e; break;
                }
            }
            if ((uint32_t)sp > st[3]) st[3] = sp;
        }
        int32_t* h = hits + 4 * i;
        if (tri_best != 0x7fffffff) { h[0] = tri_best; h[1] = inst_best; memcpy(h + 2, &tbest, 4); memcpy(h + 3, &ubest, 4); }
        else { h[0] = -1; h[1] = -1; memcpy(h + 2, &tbest, 4); h[3] = 0; }
        if (stats) memcpy(stats + 4 * i, st, 16);
    }
    return 0;
}
