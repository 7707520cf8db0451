/* traverse_core.cuh -- the arithmetic of one traversal step, shared by the persistent wavefront kernel (traverse.cu) and
 * by kernels that trace single rays inline (the light-path tail kernel, vcm.cu).  Every operation mirrors
 * oracle/traversal_oracle.cpp; translation units including this header are compiled with -fmad=false. */
#ifndef IMBA_TRAVERSE_CORE_CUH
#define IMBA_TRAVERSE_CORE_CUH
#include <cuda_runtime.h>
#include <stdint.h>
#include "bvh8.h"
#include "kernels.h"

/* Reciprocal of a direction component.  It feeds the node tests only (conservative, IMBA_ROBUST_TFAR_DEV), never the triangle test,
 * so the 1-ulp MUFU.RCP replaces the IEEE division (about 12 instructions each, three per ray and per instance entered). */
#ifndef IMBA_FAST_RCP
#define IMBA_FAST_RCP 1
#endif
__device__ __forceinline__ float safe_rcp(float d) {
    const float c = fabsf(d) > IMBA_RCP_EPS ? d : copysignf(IMBA_RCP_EPS, d);
#if IMBA_FAST_RCP
    float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(c)); return r;
#else
    return 1.0f / c;
#endif
}
#if IMBA_FAST_RCP
#define IMBA_ROBUST_TFAR_DEV 1.00000095f /* 1 + 2^-20: one more ulp of slack for the approximate reciprocal */
#else
#define IMBA_ROBUST_TFAR_DEV IMBA_ROBUST_TFAR
#endif
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
    uint32_t r; asm("prmt.b32 %0, %1, 0, 0xBA98;" : "=r"(r) : "r"(x)); return r;
}
__device__ __forceinline__ float byte_f(uint32_t x, int j) { return (float)((x >> (8 * j)) & 0xffu); }
__device__ __forceinline__ float dot_f(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(az, bz, __fmaf_rn(ay, by, ax * bx));
}
__device__ __forceinline__ float xorf(float f, uint32_t s) { return __uint_as_float(__float_as_uint(f) ^ s); }

struct RayRegs {
    float ox, oy, oz, dx, dy, dz, ix, iy, iz;
    uint32_t octinv4;
    __device__ __forceinline__ void finish() {
        ix = safe_rcp(dx); iy = safe_rcp(dy); iz = safe_rcp(dz);
        const uint32_t oct = (ix < 0.0f ? 4u : 0u) | (iy < 0.0f ? 2u : 0u) | (iz < 0.0f ? 1u : 0u);
        octinv4 = (7u - oct) * 0x01010101u;
    }
};

template <bool MASKS>
__device__ __forceinline__ bool mask_opaque(const TraceArgs& P, int tri_id, float u, float v) {
    if (!MASKS) return true;
    const int mat = __ldg(P.indices + tri_id * 4 + 3);
    const int mw = __ldg(P.masks + mat * 3), mh = __ldg(P.masks + mat * 3 + 1), mo = __ldg(P.masks + mat * 3 + 2);
    if (mw == 1 && mh == 1 && mo == 0) return true;
    const float2 a = __ldg(P.texcoords + __ldg(P.indices + tri_id * 4 + 0));
    const float2 b = __ldg(P.texcoords + __ldg(P.indices + tri_id * 4 + 1));
    const float2 c = __ldg(P.texcoords + __ldg(P.indices + tri_id * 4 + 2));
    const float w = 1.0f - u - v;
    float tu = a.x * w + b.x * u + c.x * v;
    float tv = a.y * w + b.y * u + c.y * v;
    tu = tu - floorf(tu);
    tv = 1.0f - (tv - floorf(tv));
    int x = (int)(tu * mw); if (x > mw - 1) x = mw - 1; if (x < 0) x = 0;
    int y = (int)(tv * mh); if (y > mh - 1) y = mh - 1; if (y < 0) y = 0;
    return P.mask_buffer[mo + y * mw + x] != 0;
}


struct HitRegs { float tbest, ubest; int tri_best, inst_best; };

/* Byte j of x as 1 + byte * 2^-15: ONE permute on the ALU pipe instead of an I2F on the XU pipe (the busiest pipe of the
 * traversal kernel, profiles/r1f_trace_full.csv).  The slab test then runs as fma(1 + q 2^-15, adj 2^15, org - adj 2^15): one
 * extra rounding of 2^-9 cell per node axis, far inside the 1/64 cell of slack the builder leaves (bvh8.h). */
#ifndef IMBA_PRMT_AXES
#define IMBA_PRMT_AXES 0   /* bit a set: axis a converts by permute (tuning switch) */
#endif
template <int J> __device__ __forceinline__ float byte_m(uint32_t x) { return __uint_as_float(__byte_perm(x, 0x3F800000u, 0x7604u | (J << 4))); }
template <int AXIS, int J> __device__ __forceinline__ float byte_cvt(uint32_t x) {
    if ((IMBA_PRMT_AXES >> AXIS) & 1) return byte_m<J>(x);
    return (float)((x >> (8 * J)) & 0xffu);
}
template <int AXIS> __device__ __forceinline__ void slab_setup(float& adj, float& org) {
    if ((IMBA_PRMT_AXES >> AXIS) & 1) { adj *= 32768.0f; org = org - adj; }
}

template <int J>
__device__ __forceinline__ uint32_t child_test(uint32_t nx, uint32_t fx, uint32_t ny, uint32_t fy, uint32_t nz, uint32_t fz, float adjx, float adjy, float adjz,
                                               float orgx, float orgy, float orgz, float tmin, float tbest, uint32_t child_bits4, uint32_t bit_index4) {
    const float t0x = __fmaf_rn(byte_cvt<0, J>(nx), adjx, orgx), t1x = __fmaf_rn(byte_cvt<0, J>(fx), adjx, orgx);
    const float t0y = __fmaf_rn(byte_cvt<1, J>(ny), adjy, orgy), t1y = __fmaf_rn(byte_cvt<1, J>(fy), adjy, orgy);
    const float t0z = __fmaf_rn(byte_cvt<2, J>(nz), adjz, orgz), t1z = __fmaf_rn(byte_cvt<2, J>(fz), adjz, orgz);
    const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
    const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tbest));
    if (tn <= tf * IMBA_ROBUST_TFAR_DEV) {
        const uint32_t cb = (child_bits4 >> (8 * J)) & 0xffu, bi = (bit_index4 >> (8 * J)) & 0xffu;
        return cb << bi;
    }
    return 0u;
}

/* Intersects the 8 quantised child boxes of one node: returns the 8-bit inner / 24-bit primitive hit mask. */
__device__ __forceinline__ uint32_t node_step(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4, const RayRegs& r, float tmin, float tbest) {
    const uint32_t e_imask = n0.w;
    float adjx = __uint_as_float((e_imask & 0xffu) << 23) * r.ix;
    float adjy = __uint_as_float(((e_imask >> 8) & 0xffu) << 23) * r.iy;
    float adjz = __uint_as_float(((e_imask >> 16) & 0xffu) << 23) * r.iz;
    float orgx = (__uint_as_float(n0.x) - r.ox) * r.ix;
    float orgy = (__uint_as_float(n0.y) - r.oy) * r.iy;
    float orgz = (__uint_as_float(n0.z) - r.oz) * r.iz;
    slab_setup<0>(adjx, orgx); slab_setup<1>(adjy, orgy); slab_setup<2>(adjz, orgz);
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (r.octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
        const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
        const uint32_t nx = r.ix < 0.0f ? qhix : qlox, fx = r.ix < 0.0f ? qlox : qhix;
        const uint32_t ny = r.iy < 0.0f ? qhiy : qloy, fy = r.iy < 0.0f ? qloy : qhiy;
        const uint32_t nz = r.iz < 0.0f ? qhiz : qloz, fz = r.iz < 0.0f ? qloz : qhiz;
        hitmask |= child_test<0>(nx, fx, ny, fy, nz, fz, adjx, adjy, adjz, orgx, orgy, orgz, tmin, tbest, child_bits4, bit_index4);
        hitmask |= child_test<1>(nx, fx, ny, fy, nz, fz, adjx, adjy, adjz, orgx, orgy, orgz, tmin, tbest, child_bits4, bit_index4);
        hitmask |= child_test<2>(nx, fx, ny, fy, nz, fz, adjx, adjy, adjz, orgx, orgy, orgz, tmin, tbest, child_bits4, bit_index4);
        hitmask |= child_test<3>(nx, fx, ny, fy, nz, fz, adjx, adjy, adjz, orgx, orgy, orgz, tmin, tbest, child_bits4, bit_index4);
    }
    return hitmask;
}

/* Transforms the ray into instance space (cpu_adapter.cpp:356-365): t is shared between the levels (no renormalisation). */
__device__ __forceinline__ void enter_instance(RayRegs& r, const float4 m0, const float4 m1, const float4 m2) {
    const float ox = r.ox, oy = r.oy, oz = r.oz, dx = r.dx, dy = r.dy, dz = r.dz;
    r.ox = __fmaf_rn(m0.z, oz, __fmaf_rn(m0.y, oy, __fmaf_rn(m0.x, ox, m0.w)));
    r.oy = __fmaf_rn(m1.z, oz, __fmaf_rn(m1.y, oy, __fmaf_rn(m1.x, ox, m1.w)));
    r.oz = __fmaf_rn(m2.z, oz, __fmaf_rn(m2.y, oy, __fmaf_rn(m2.x, ox, m2.w)));
    r.dx = __fmaf_rn(m0.z, dz, __fmaf_rn(m0.y, dy, m0.x * dx));
    r.dy = __fmaf_rn(m1.z, dz, __fmaf_rn(m1.y, dy, m1.x * dx));
    r.dz = __fmaf_rn(m2.z, dz, __fmaf_rn(m2.y, dy, m2.x * dx));
    r.finish();
}

/* Tests one triangle (a = v0|id, b = e1, c = e2) and updates the hit if (t, tri, inst) is lexicographically smaller. */
/* ONE_INST: every triangle belongs to the same instance (flat scenes), so the instance id neither takes part in the tie rule
 * nor needs a register; the caller fills it in when it writes the hit. */
template <bool MASKS, bool ONE_INST = false>
__device__ __forceinline__ bool tri_test(const TraceArgs& P, const float4 a, const float4 b, const float4 c, const RayRegs& r, float tmin, int cur_inst, HitRegs& h) {
    float& tbest = h.tbest; float& ubest = h.ubest; int& tri_best = h.tri_best; int& inst_best = h.inst_best;
    const int tri_id = __float_as_int(a.w);
    // n = e1 x e2, unfused (what cpu_adapter.cpp:149 stores): x and y travel in the record's padding (bvh8.h), z is computed
#ifndef IMBA_TRI_RECOMPUTE_N
    const float nx = b.w, ny = c.w, nz = b.x * c.y - b.y * c.x;
#else
    const float nx = b.y * c.z - b.z * c.y, ny = b.z * c.x - b.x * c.z, nz = b.x * c.y - b.y * c.x;
#endif
    const float cx = a.x - r.ox, cy = a.y - r.oy, cz = a.z - r.oz;
    const float rx = __fmaf_rn(r.dy, cz, -(r.dz * cy));
    const float ry = __fmaf_rn(r.dz, cx, -(r.dx * cz));
    const float rz = __fmaf_rn(r.dx, cy, -(r.dy * cx));
    const float det = dot_f(nx, ny, nz, r.dx, r.dy, r.dz);
    const float adet = fabsf(det);
    const uint32_t sgn = __float_as_uint(det) & 0x80000000u;
    const float U = xorf(dot_f(rx, ry, rz, c.x, c.y, c.z), sgn);
    const float V = xorf(dot_f(rx, ry, rz, b.x, b.y, b.z), sgn);
    // (det == 0 is decided with the second group of tests: a degenerate candidate passes the first group only with U = V = 0)
    if (!(U >= 0.0f) || !(V >= 0.0f) || !(U + V <= adet)) return false;
    const float T = xorf(dot_f(nx, ny, nz, cx, cy, cz), sgn);
    if (!(T >= adet * tmin) || !(T <= adet * tbest * IMBA_T_PREFILTER) || det == 0.0f) return false;
    const float inv = 1.0f / adet;
    const float t = T * inv, u = U * inv;
    if (ONE_INST) { if (!(t < tbest || (t == tbest && tri_id < tri_best))) return false; }
    else if (!(t < tbest || (t == tbest && (tri_id < tri_best || (tri_id == tri_best && cur_inst < inst_best))))) return false;
    if (MASKS && !mask_opaque<MASKS>(P, tri_id, u, V * inv)) return false;
    tbest = t; ubest = u; tri_best = tri_id; if (!ONE_INST) inst_best = cur_inst;
    return true;
}

/* One ray, one thread, no work sharing: used where a handful of rays is traced inside another kernel.
 * Same results as trace_kernel (same steps, same tie rule); returns true if something was hit (any hit: first found). */
template <bool ANY, bool MASKS>
__device__ __noinline__ bool trace_single(const TraceArgs& P, const float4 o, const float4 d, HitRegs& h) {
    uint2 stack[IMBA_TRAVERSAL_STACK];
    int sp = 0;
    RayRegs r; r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
    if (P.flat_inst >= 0) { r.ox += 0.0f; r.oy += 0.0f; r.oz += 0.0f; r.dx += 0.0f; r.dy += 0.0f; r.dz += 0.0f; }
    r.finish();
    const float tmin = o.w;
    h.tbest = d.w; h.ubest = 0.0f; h.tri_best = IMBA_NO_HIT; h.inst_best = IMBA_NO_HIT;
    int cur_inst = P.flat_inst;
    uint2 ngroup = make_uint2(P.top_root, 0x80000000u), tgroup = make_uint2(0, 0);
    for (;;) {
        if (ngroup.y > 0x00ffffffu) {
            const uint32_t hits = ngroup.y;
            const uint32_t bit = 31u - (uint32_t)__clz(hits);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
            const uint32_t slot = (bit - 24u) ^ (r.octinv4 & 0xffu);
            const uint32_t rel = __popc(hits & 0xffu & ~(0xffffffffu << slot));
            const uint4* np = P.nodes + (size_t)(ngroup.x + rel) * 5;
            const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            const uint32_t hitmask = node_step(n0, n1, n2, n3, n4, r, tmin, h.tbest);
            ngroup = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
            tgroup = make_uint2(n1.y, hitmask & 0x00ffffffu);
        } else { tgroup = ngroup; ngroup = make_uint2(0, 0); }
        while (tgroup.y != 0) {
            const uint32_t bit = 31u - (uint32_t)__clz(tgroup.y);
            tgroup.y &= ~(1u << bit);
            if (cur_inst < 0) {
                if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
                if (tgroup.y != 0) stack[sp++] = tgroup;
                stack[sp++] = make_uint2(0xffffffffu, 0u);
                const float4* ip = P.insts + (size_t)(tgroup.x + bit) * 4;
                const float4 m0 = __ldg(ip), m1 = __ldg(ip + 1), m2 = __ldg(ip + 2);
                const int4 id = __ldg((const int4*)(ip + 3));
                enter_instance(r, m0, m1, m2);
                cur_inst = id.x;
                ngroup = make_uint2((uint32_t)id.y, 0x80000000u);
                tgroup = make_uint2(0, 0);
                break;
            }
            const float4* tp = P.tris + (size_t)(tgroup.x + bit) * 3;
            if (tri_test<MASKS>(P, __ldg(tp), __ldg(tp + 1), __ldg(tp + 2), r, tmin, cur_inst, h) && ANY) return true;
        }
        if (ngroup.y <= 0x00ffffffu && tgroup.y == 0) {
            for (;;) {
                if (sp == 0) return h.tri_best != IMBA_NO_HIT;
                const uint2 e = stack[--sp];
                if (e.y == 0u) { r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z; r.finish(); cur_inst = -1; continue; }
                ngroup = e;
                break;
            }
        }
    }
}

/* ---- one ray traced by a whole warp ------------------------------------------------------------------------------------------
 * For the few very long light paths at the end of a light pass (vcm.cu: vcm_light_tail_warp_kernel) latency matters, not
 * throughput: the 8 child boxes of a node are tested by 8 lanes and the triangles of a leaf group by one lane each, so one
 * traversal step costs ~1/6 of the dependent instructions of the one-thread version.  All lanes hold the same ray, stack and
 * hit; every box / triangle test is the very same arithmetic as in node_step / tri_test, child results are OR-ed and triangle
 * candidates reduced with the (t, tri, inst) order, so the result is bit-identical to trace_single and trace_kernel. */
__device__ __forceinline__ uint32_t node_step_coop(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4, const RayRegs& r, float tmin, float tbest, unsigned lane) {
    const uint32_t e_imask = n0.w;
    const float adjx = __uint_as_float((e_imask & 0xffu) << 23) * r.ix;
    const float adjy = __uint_as_float(((e_imask >> 8) & 0xffu) << 23) * r.iy;
    const float adjz = __uint_as_float(((e_imask >> 16) & 0xffu) << 23) * r.iz;
    const float orgx = (__uint_as_float(n0.x) - r.ox) * r.ix;
    const float orgy = (__uint_as_float(n0.y) - r.oy) * r.iy;
    const float orgz = (__uint_as_float(n0.z) - r.oz) * r.iz;
    uint32_t contrib = 0;
    if (lane < 8u) {
        const bool half = lane >= 4u; const int j = (int)(lane & 3u);
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (r.octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
        const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
        const uint32_t nx = r.ix < 0.0f ? qhix : qlox, fx = r.ix < 0.0f ? qlox : qhix;
        const uint32_t ny = r.iy < 0.0f ? qhiy : qloy, fy = r.iy < 0.0f ? qloy : qhiy;
        const uint32_t nz = r.iz < 0.0f ? qhiz : qloz, fz = r.iz < 0.0f ? qloz : qhiz;
        const float t0x = __fmaf_rn(byte_f(nx, j), adjx, orgx), t1x = __fmaf_rn(byte_f(fx, j), adjx, orgx);
        const float t0y = __fmaf_rn(byte_f(ny, j), adjy, orgy), t1y = __fmaf_rn(byte_f(fy, j), adjy, orgy);
        const float t0z = __fmaf_rn(byte_f(nz, j), adjz, orgz), t1z = __fmaf_rn(byte_f(fz, j), adjz, orgz);
        const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
        const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tbest));
        if (tn <= tf * IMBA_ROBUST_TFAR_DEV) {
            const uint32_t cb = (child_bits4 >> (8 * j)) & 0xffu, bi = (bit_index4 >> (8 * j)) & 0xffu;
            contrib = cb << bi;
        }
    }
    return __reduce_or_sync(0xffffffffu, contrib);
}

/* Must be called by all 32 lanes with identical arguments; returns the same result in every lane. */
template <bool ANY, bool MASKS>
__device__ __noinline__ bool trace_warp(const TraceArgs& P, const float4 o, const float4 d, HitRegs& h) {
    const unsigned lane = threadIdx.x & 31u;
    uint2 stack[IMBA_TRAVERSAL_STACK];
    int sp = 0;
    RayRegs r; r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
    if (P.flat_inst >= 0) { r.ox += 0.0f; r.oy += 0.0f; r.oz += 0.0f; r.dx += 0.0f; r.dy += 0.0f; r.dz += 0.0f; }
    r.finish();
    const float tmin = o.w;
    h.tbest = d.w; h.ubest = 0.0f; h.tri_best = IMBA_NO_HIT; h.inst_best = IMBA_NO_HIT;
    int cur_inst = P.flat_inst;
    uint2 ngroup = make_uint2(P.top_root, 0x80000000u), tgroup = make_uint2(0, 0);
    for (;;) {
        if (ngroup.y > 0x00ffffffu) {
            const uint32_t hits = ngroup.y;
            const uint32_t bit = 31u - (uint32_t)__clz(hits);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
            const uint32_t slot = (bit - 24u) ^ (r.octinv4 & 0xffu);
            const uint32_t rel = __popc(hits & 0xffu & ~(0xffffffffu << slot));
            const uint4* np = P.nodes + (size_t)(ngroup.x + rel) * 5;
            const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            const uint32_t hitmask = node_step_coop(n0, n1, n2, n3, n4, r, tmin, h.tbest, lane);
            ngroup = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
            tgroup = make_uint2(n1.y, hitmask & 0x00ffffffu);
        } else { tgroup = ngroup; ngroup = make_uint2(0, 0); }
        if (tgroup.y != 0) {
            if (cur_inst < 0) { // top level: instances one at a time, as in trace_single
                const uint32_t bit = 31u - (uint32_t)__clz(tgroup.y);
                tgroup.y &= ~(1u << bit);
                if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
                if (tgroup.y != 0) stack[sp++] = tgroup;
                stack[sp++] = make_uint2(0xffffffffu, 0u);
                const float4* ip = P.insts + (size_t)(tgroup.x + bit) * 4;
                const float4 m0 = __ldg(ip), m1 = __ldg(ip + 1), m2 = __ldg(ip + 2);
                const int4 id = __ldg((const int4*)(ip + 3));
                enter_instance(r, m0, m1, m2);
                cur_inst = id.x;
                ngroup = make_uint2((uint32_t)id.y, 0x80000000u);
                tgroup = make_uint2(0, 0);
            } else {            // one triangle of the group per lane
                const unsigned pos = __fns(tgroup.y, 0u, (int)lane + 1);
                HitRegs mine = h; bool got = false;
                if (pos != 0xffffffffu) {
                    const float4* tp = P.tris + (size_t)(tgroup.x + pos) * 3;
                    got = tri_test<MASKS>(P, __ldg(tp), __ldg(tp + 1), __ldg(tp + 2), r, tmin, cur_inst, mine);
                }
                tgroup.y = 0;
                if (ANY) {
                    if (__any_sync(0xffffffffu, got)) { h = mine; return true; }
                } else {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const float ot = __shfl_xor_sync(0xffffffffu, mine.tbest, off), ou = __shfl_xor_sync(0xffffffffu, mine.ubest, off);
                        const int otri = __shfl_xor_sync(0xffffffffu, mine.tri_best, off), oinst = __shfl_xor_sync(0xffffffffu, mine.inst_best, off);
                        if (ot < mine.tbest || (ot == mine.tbest && (otri < mine.tri_best || (otri == mine.tri_best && oinst < mine.inst_best)))) {
                            mine.tbest = ot; mine.ubest = ou; mine.tri_best = otri; mine.inst_best = oinst;
                        }
                    }
                    h = mine;
                }
            }
        }
        if (ngroup.y <= 0x00ffffffu && tgroup.y == 0) {
            for (;;) {
                if (sp == 0) return h.tri_best != IMBA_NO_HIT;
                const uint2 e = stack[--sp];
                if (e.y == 0u) { r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z; r.finish(); cur_inst = -1; continue; }
                ngroup = e;
                break;
            }
        }
    }
}

#endif
