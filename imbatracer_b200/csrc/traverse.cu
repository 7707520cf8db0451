/* traverse.cu -- persistent-warp closest-hit / any-hit traversal of ray queues (sm_100a).
 *
 * Replaces the external AnyDSL kernels intersect_*_masked_instanced / occluded_*_masked_instanced
 * (call sites render/scheduling/ray_queue.h:254-265, 279-290, 304-315, 329-340).
 *
 * Design
 *   - one ray per lane; a warp pulls rays from a global atomic counter and REFILLS idle lanes as soon as fewer than
 *     REFILL_LANES lanes are still traversing (Aila/Laine persistent threads with dynamic fetch);
 *   - compressed 8-wide nodes (bvh8.h), five 128-bit loads per node, child boxes tested in the quantised frame
 *     with one FMA per slab; hits are kept as an 8+24-bit mask pair (inner children / primitives), the inner part
 *     ordered by `slot ^ octant`, so the traversal stack holds 8-byte node groups and stays short;
 *   - two-level: a top-level leaf is an instance record; entering one pushes a sentinel, transforms the ray
 *     (no renormalisation: t is shared between levels) and continues in the same node array;
 *   - triangle test, instance transform and tie-break are the oracle's, operation for operation
 *     (oracle/traversal_oracle.cpp): this file is compiled with -fmad=false, every fused multiply-add is explicit.
 *     Equal t resolves to the smaller (tri_id, inst_id), so results do not depend on the tree shape.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "bvh8.h"
#include "kernels.h"

namespace {

constexpr int      BLOCK_THREADS = 128;
constexpr int      STACK_ENTRIES = IMBA_TRAVERSAL_STACK;
constexpr unsigned FULL = 0xffffffffu;
constexpr int      REFILL_LANES = 20; // refill when fewer lanes than this are active

__device__ __forceinline__ float safe_rcp(float d) {
    return 1.0f / (fabsf(d) > IMBA_RCP_EPS ? d : copysignf(IMBA_RCP_EPS, d));
}
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

/* MODE 0 = closest hit, 1 = any hit, 2 = any hit fused with process_shadow_rays (integrator.h:49-63): no hit record is
 * written; an unoccluded ray adds its ShadowState throughput to the accumulation buffer; slots whose pixel_id < 0 are skipped. */
template <int MODE, bool MASKS, bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS) trace_kernel(const TraceArgs P) {
    constexpr bool ANY = MODE != 0;
    constexpr bool RESOLVE = MODE == 2;
    const int ray_count = P.count_ptr ? *P.count_ptr * P.count_mul : P.count;
    const unsigned lane = threadIdx.x & 31u;
    const uint4*  __restrict__ nodes = P.nodes;
    const float4* __restrict__ tris  = P.tris;

    uint2 stack[STACK_ENTRIES];
    int sp = 0;
    RayRegs r;
    float tmin = 0.0f, tbest = 0.0f, ubest = 0.0f;
    int tri_best = IMBA_NO_HIT, inst_best = IMBA_NO_HIT, cur_inst = -1, ray_idx = -1;
    uint2 ngroup = make_uint2(0, 0), tgroup = make_uint2(0, 0);
    uint32_t st_nodes = 0, st_tris = 0, st_insts = 0, st_sp = 0;
    bool active = false, exhausted = false;

    for (;;) {
        // ---- refill idle lanes from the global work counter --------------------------------------
        // (in resolve mode empty shadow slots are skipped, so a few rounds may be needed to fill the warp)
        for (int round = 0; round < (RESOLVE ? 4 : 1) && !exhausted; ++round) {
            const unsigned need = __ballot_sync(FULL, !active);
            if (!need) break;
            const int leader = __ffs(need) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(P.counter, (unsigned)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (base + (unsigned)__popc(need) >= (unsigned)ray_count) exhausted = true; // warp-uniform
            if (!active) {
                const unsigned i = base + (unsigned)__popc(need & ((1u << lane) - 1u));
                bool take = i < (unsigned)ray_count;
                if (RESOLVE && take) take = __ldg(reinterpret_cast<const int*>(P.shadow_states) + 4 * i) >= 0;
                if (take) {
                    ray_idx = (int)i;
                    const float4 o = __ldcs(P.rays + 2 * i), d = __ldcs(P.rays + 2 * i + 1);
                    r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                    if (P.flat_inst >= 0) { // the identity transform of the skipped instance turns -0 into +0 (1 * x + 0)
                        r.ox += 0.0f; r.oy += 0.0f; r.oz += 0.0f; r.dx += 0.0f; r.dy += 0.0f; r.dz += 0.0f;
                    }
                    r.finish();
                    tmin = o.w; tbest = d.w; ubest = 0.0f; tri_best = IMBA_NO_HIT; inst_best = IMBA_NO_HIT; cur_inst = P.flat_inst;
                    sp = 0; ngroup = make_uint2(P.top_root, 0x80000000u); tgroup = make_uint2(0, 0);
                    if (STATS) { st_nodes = st_tris = st_insts = st_sp = 0; }
                    active = true;
                }
            }
        }
        if (__ballot_sync(FULL, active) == 0) { if (exhausted) break; continue; }

        // ---- traverse ----------------------------------------------------------------------------
        while (active) {
            if (ngroup.y > 0x00ffffffu) {
                // pick the nearest pending inner child of the current node group
                const uint32_t hits = ngroup.y;
                const uint32_t bit = 31u - (uint32_t)__clz(hits);
                ngroup.y &= ~(1u << bit);
                if (ngroup.y > 0x00ffffffu) { stack[sp++] = ngroup; if (STATS) st_sp = max(st_sp, (uint32_t)sp); }
                const uint32_t slot = (bit - 24u) ^ (r.octinv4 & 0xffu);
                const uint32_t rel  = __popc(hits & 0xffu & ~(0xffffffffu << slot));
                const uint4* np = nodes + (size_t)(ngroup.x + rel) * 5;
                const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
                if (STATS) ++st_nodes;

                const uint32_t e_imask = n0.w;
                const float adjx = __uint_as_float((e_imask & 0xffu) << 23) * r.ix;
                const float adjy = __uint_as_float(((e_imask >> 8) & 0xffu) << 23) * r.iy;
                const float adjz = __uint_as_float(((e_imask >> 16) & 0xffu) << 23) * r.iz;
                const float orgx = (__uint_as_float(n0.x) - r.ox) * r.ix;
                const float orgy = (__uint_as_float(n0.y) - r.oy) * r.iy;
                const float orgz = (__uint_as_float(n0.z) - r.oz) * r.iz;
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
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float t0x = __fmaf_rn(byte_f(nx, j), adjx, orgx), t1x = __fmaf_rn(byte_f(fx, j), adjx, orgx);
                        const float t0y = __fmaf_rn(byte_f(ny, j), adjy, orgy), t1y = __fmaf_rn(byte_f(fy, j), adjy, orgy);
                        const float t0z = __fmaf_rn(byte_f(nz, j), adjz, orgz), t1z = __fmaf_rn(byte_f(fz, j), adjz, orgz);
                        const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
                        const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tbest));
                        if (tn <= tf * IMBA_ROBUST_TFAR) {
                            const uint32_t cb = (child_bits4 >> (8 * j)) & 0xffu, bi = (bit_index4 >> (8 * j)) & 0xffu;
                            hitmask |= cb << bi;
                        }
                    }
                }
                ngroup = make_uint2(n1.x, (hitmask & 0xff000000u) | (e_imask >> 24));
                tgroup = make_uint2(n1.y, hitmask & 0x00ffffffu);
            } else {
                tgroup = ngroup; ngroup = make_uint2(0, 0);
            }

            // ---- primitives of the current group -----------------------------------------------
            bool done = false;
            while (tgroup.y != 0) {
                const uint32_t bit = 31u - (uint32_t)__clz(tgroup.y);
                tgroup.y &= ~(1u << bit);
                if (cur_inst < 0) {
                    // top level: enter the instance (ray * inverse transform, cpu_adapter.cpp:356-365)
                    if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
                    if (tgroup.y != 0) stack[sp++] = tgroup;
                    stack[sp++] = make_uint2(0xffffffffu, 0u);
                    if (STATS) { st_sp = max(st_sp, (uint32_t)sp); ++st_insts; }
                    const float4* ip = P.insts + (size_t)(tgroup.x + bit) * 4;
                    const float4 m0 = __ldg(ip), m1 = __ldg(ip + 1), m2 = __ldg(ip + 2);
                    const int4 id = __ldg((const int4*)(ip + 3));
                    const float ox = r.ox, oy = r.oy, oz = r.oz, dx = r.dx, dy = r.dy, dz = r.dz;
                    r.ox = __fmaf_rn(m0.z, oz, __fmaf_rn(m0.y, oy, __fmaf_rn(m0.x, ox, m0.w)));
                    r.oy = __fmaf_rn(m1.z, oz, __fmaf_rn(m1.y, oy, __fmaf_rn(m1.x, ox, m1.w)));
                    r.oz = __fmaf_rn(m2.z, oz, __fmaf_rn(m2.y, oy, __fmaf_rn(m2.x, ox, m2.w)));
                    r.dx = __fmaf_rn(m0.z, dz, __fmaf_rn(m0.y, dy, m0.x * dx));
                    r.dy = __fmaf_rn(m1.z, dz, __fmaf_rn(m1.y, dy, m1.x * dx));
                    r.dz = __fmaf_rn(m2.z, dz, __fmaf_rn(m2.y, dy, m2.x * dx));
                    r.finish();
                    cur_inst = id.x;
                    ngroup = make_uint2((uint32_t)id.y, 0x80000000u);
                    tgroup = make_uint2(0, 0);
                    break;
                }
                const float4* tp = tris + (size_t)(tgroup.x + bit) * 3;
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                if (STATS) ++st_tris;
                const int tri_id = __float_as_int(a.w);
                // n = e1 x e2, unfused (what cpu_adapter.cpp:149 stores)
                const float nx = b.y * c.z - b.z * c.y, ny = b.z * c.x - b.x * c.z, nz = b.x * c.y - b.y * c.x;
                const float cx = a.x - r.ox, cy = a.y - r.oy, cz = a.z - r.oz;
                const float rx = __fmaf_rn(r.dy, cz, -(r.dz * cy));
                const float ry = __fmaf_rn(r.dz, cx, -(r.dx * cz));
                const float rz = __fmaf_rn(r.dx, cy, -(r.dy * cx));
                const float det = dot_f(nx, ny, nz, r.dx, r.dy, r.dz);
                const float adet = fabsf(det);
                const uint32_t sgn = __float_as_uint(det) & 0x80000000u;
                const float U = xorf(dot_f(rx, ry, rz, c.x, c.y, c.z), sgn);
                const float V = xorf(dot_f(rx, ry, rz, b.x, b.y, b.z), sgn);
                if (!(U >= 0.0f) || !(V >= 0.0f) || !(U + V <= adet) || det == 0.0f) continue;
                const float T = xorf(dot_f(nx, ny, nz, cx, cy, cz), sgn);
                if (!(T >= adet * tmin) || !(T <= adet * tbest * IMBA_T_PREFILTER)) continue;
                const float inv = 1.0f / adet;
                const float t = T * inv, u = U * inv;
                if (!(t < tbest || (t == tbest && (tri_id < tri_best || (tri_id == tri_best && cur_inst < inst_best))))) continue;
                if (MASKS && !mask_opaque<MASKS>(P, tri_id, u, V * inv)) continue;
                tbest = t; ubest = u; tri_best = tri_id; inst_best = cur_inst;
                if (ANY) { done = true; break; }
            }

            // ---- pop --------------------------------------------------------------------------------
            if (!done && ngroup.y <= 0x00ffffffu && tgroup.y == 0) {
                for (;;) {
                    if (sp == 0) { done = true; break; }
                    const uint2 e = stack[--sp];
                    if (e.y == 0u) { // sentinel: leave the instance, restore the world-space ray
                        const float4 o = __ldg(P.rays + 2 * ray_idx), d = __ldg(P.rays + 2 * ray_idx + 1);
                        r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z; r.finish();
                        cur_inst = -1;
                        continue;
                    }
                    ngroup = e;
                    break;
                }
            }
            if (done) {
                if (RESOLVE) {
                    if (tri_best == IMBA_NO_HIT) {
                        const float4 st = __ldg(reinterpret_cast<const float4*>(P.shadow_states) + ray_idx);
                        float* px = P.framebuffer + 3 * (size_t)__float_as_int(st.x);
                        atomicAdd(px, st.y); atomicAdd(px + 1, st.z); atomicAdd(px + 2, st.w);
                    }
                } else {
                    int4 h;
                    if (tri_best != IMBA_NO_HIT) h = make_int4(tri_best, inst_best, __float_as_int(tbest), __float_as_int(ubest));
                    else               h = make_int4(-1, -1, __float_as_int(tbest), 0);
                    __stcs(P.hits + ray_idx, h);
                }
                if (STATS) {
                    reinterpret_cast<uint4*>(P.stats)[ray_idx] = make_uint4(st_nodes, st_tris, st_insts, st_sp);
                }
                active = false;
                break;
            }
            if (!exhausted && __popc(__activemask()) < REFILL_LANES) break;
        }
    }
}

template <int ANY, bool MASKS, bool STATS>
cudaError_t launch(const TraceArgs& a, int grid, cudaStream_t s) {
    trace_kernel<ANY, MASKS, STATS><<<grid, BLOCK_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

template <int ANY, bool MASKS, bool STATS>
int occupancy() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_kernel<ANY, MASKS, STATS>, BLOCK_THREADS, 0);
    return n;
}

} // namespace

int imba_trace_block_threads() { return BLOCK_THREADS; }

#define FOR_ALL_VARIANTS(X) \
    X(0, false, false) X(0, false, true) X(0, true, false) X(0, true, true) \
    X(1, false, false) X(1, false, true) X(1, true, false) X(1, true, true) \
    X(2, false, false) X(2, true, false)

int imba_trace_blocks_per_sm(int mode, bool masks, bool stats) {
#define OCC(A, M, S) if (mode == A && masks == M && stats == S) return occupancy<A, M, S>();
    FOR_ALL_VARIANTS(OCC)
#undef OCC
    return 0;
}

cudaError_t imba_launch_trace(const TraceArgs& a, int mode, bool masks, bool stats, int grid, cudaStream_t s) {
#define RUN(A, M, S) if (mode == A && masks == M && stats == S) return launch<A, M, S>(a, grid, s);
    FOR_ALL_VARIANTS(RUN)
#undef RUN
    return cudaErrorInvalidValue;
}
