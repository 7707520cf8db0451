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
#include "traverse_core.cuh"

namespace {

#ifndef IMBA_BLOCK_THREADS
#define IMBA_BLOCK_THREADS 128
#endif
constexpr int      BLOCK_THREADS = IMBA_BLOCK_THREADS;
constexpr int      STACK_ENTRIES = IMBA_TRAVERSAL_STACK;
constexpr unsigned FULL = 0xffffffffu;
#ifndef IMBA_REFILL_LANES
#define IMBA_REFILL_LANES 20
#endif
constexpr int      REFILL_LANES = IMBA_REFILL_LANES; // refill when fewer lanes than this are active
#ifndef IMBA_PREFETCH_NEXT
#define IMBA_PREFETCH_NEXT 0  /* 1 / 2: prefetch the next inner child's node into L1 / L2 before the triangle tests (tuning switch) */
#endif
#ifndef IMBA_SMEM_STACK
#define IMBA_SMEM_STACK 0  /* traversal-stack entries per thread kept in shared memory (tuning switch) */
#endif

/* Index of the highest set bit (x != 0): one FLO; `31 - __clz(x)` costs a second instruction that the compiler does not fold. */
__device__ __forceinline__ uint32_t top_bit(uint32_t x) {
#ifndef IMBA_NO_BFIND
    uint32_t b; asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(x)); return b;
#else
    return 31u - (uint32_t)__clz(x);
#endif
}

/* MODE 0 = closest hit, 1 = any hit, 2 = any hit fused with process_shadow_rays (integrator.h:49-63): no hit record is
 * written; an unoccluded ray adds its ShadowState throughput to the accumulation buffer; slots whose pixel_id < 0 are skipped. */
/* FLAT = the scene is one identity instance and traversal starts at its mesh root (P.flat_inst >= 0): the instance level --
 * entering, the sentinel, the saved world-space ray, the instance id of the best hit -- is compiled out. */
template <int MODE, bool MASKS, bool STATS, bool FLAT>
__device__ __forceinline__ void trace_body(const TraceArgs& P) {
    constexpr bool ANY = MODE != 0;
    constexpr bool RESOLVE = MODE == 2;
    const int ray_count = P.count_ptr ? *P.count_ptr * P.count_mul : P.count;
    const unsigned lane = threadIdx.x & 31u;
    const uint4*  __restrict__ nodes = P.nodes;
    const float4* __restrict__ tris  = P.tris;

#if IMBA_SMEM_STACK > 0
    // the top IMBA_SMEM_STACK entries of the traversal stack live in shared memory ([entry][thread]: conflict-free 64-bit accesses),
    // deeper ones in the local-memory frame
    __shared__ uint2 s_stack[IMBA_SMEM_STACK * BLOCK_THREADS];
    uint2 stack[STACK_ENTRIES > IMBA_SMEM_STACK ? STACK_ENTRIES - IMBA_SMEM_STACK : 1];
#define STACK_PUSH(e) do { if (sp < IMBA_SMEM_STACK) s_stack[sp * BLOCK_THREADS + threadIdx.x] = (e); else stack[sp - IMBA_SMEM_STACK] = (e); ++sp; } while (0)
#define STACK_POP()   (--sp, sp < IMBA_SMEM_STACK ? s_stack[sp * BLOCK_THREADS + threadIdx.x] : stack[sp - IMBA_SMEM_STACK])
#else
    uint2 stack[STACK_ENTRIES];
#define STACK_PUSH(e) (stack[sp++] = (e))
#define STACK_POP()   (stack[--sp])
#endif
    int sp = 0;
    RayRegs r;
    float tmin = 0.0f;
    HitRegs hr; hr.tbest = 0.0f; hr.ubest = 0.0f; hr.tri_best = IMBA_NO_HIT; hr.inst_best = IMBA_NO_HIT;
    int cur_inst = -1, ray_idx = -1;
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
                    if (FLAT || P.flat_inst >= 0) { // the identity transform of the skipped instance turns -0 into +0 (1 * x + 0)
                        r.ox += 0.0f; r.oy += 0.0f; r.oz += 0.0f; r.dx += 0.0f; r.dy += 0.0f; r.dz += 0.0f;
                    }
                    r.finish();
                    tmin = o.w; hr.tbest = d.w; hr.ubest = 0.0f; hr.tri_best = IMBA_NO_HIT; hr.inst_best = IMBA_NO_HIT; cur_inst = P.flat_inst;
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
                const uint32_t bit = top_bit(hits);
                ngroup.y &= ~(1u << bit);
                if (ngroup.y > 0x00ffffffu) { STACK_PUSH(ngroup); if (STATS) st_sp = max(st_sp, (uint32_t)sp); }
                const uint32_t slot = (bit - 24u) ^ (r.octinv4 & 0xffu);
                const uint32_t rel  = __popc(hits & 0xffu & ~(0xffffffffu << slot));
                const uint4* np = nodes + (size_t)(ngroup.x + rel) * 5;
                const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
                if (STATS) ++st_nodes;

                const uint32_t e_imask = n0.w;
                const uint32_t hitmask = node_step(n0, n1, n2, n3, n4, r, tmin, hr.tbest);
                ngroup = make_uint2(n1.x, (hitmask & 0xff000000u) | (e_imask >> 24));
                tgroup = make_uint2(n1.y, hitmask & 0x00ffffffu);
#if IMBA_PREFETCH_NEXT
                // triangles are tested first and an inner child is visited next: start fetching that node now (tuning switch)
                if (tgroup.y != 0 && ngroup.y > 0x00ffffffu) {
                    const uint32_t nbit = 31u - (uint32_t)__clz(ngroup.y);
                    const uint32_t nslot = (nbit - 24u) ^ (r.octinv4 & 0xffu);
                    const uint32_t nrel = __popc(ngroup.y & 0xffu & ~(0xffffffffu << nslot));
                    const char* pf = reinterpret_cast<const char*>(nodes + (size_t)(ngroup.x + nrel) * 5);
#if IMBA_PREFETCH_NEXT == 2
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(pf)); asm volatile("prefetch.global.L2 [%0];" :: "l"(pf + 64));
#else
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(pf)); asm volatile("prefetch.global.L1 [%0];" :: "l"(pf + 64));
#endif
                }
#endif
            } else {
                tgroup = ngroup; ngroup = make_uint2(0, 0);
            }

            // ---- primitives of the current group -----------------------------------------------
            bool done = false;
            while (tgroup.y != 0) {
                const uint32_t bit = top_bit(tgroup.y);
                tgroup.y &= ~(1u << bit);
                if (!FLAT && cur_inst < 0) {
                    // top level: enter the instance (ray * inverse transform, cpu_adapter.cpp:356-365)
                    if (ngroup.y > 0x00ffffffu) STACK_PUSH(ngroup);
                    if (tgroup.y != 0) STACK_PUSH(tgroup);
                    // the world-space ray (with its reciprocals and octant) goes onto the stack under the sentinel: leaving the instance
                    // pops it back instead of reloading the ray from the queue and dividing three times again
#ifndef IMBA_RELOAD_WORLD_RAY
                    STACK_PUSH(make_uint2(__float_as_uint(r.ox), __float_as_uint(r.oy)));
                    STACK_PUSH(make_uint2(__float_as_uint(r.oz), __float_as_uint(r.dx)));
                    STACK_PUSH(make_uint2(__float_as_uint(r.dy), __float_as_uint(r.dz)));
                    STACK_PUSH(make_uint2(__float_as_uint(r.ix), __float_as_uint(r.iy)));
                    STACK_PUSH(make_uint2(__float_as_uint(r.iz), r.octinv4));
#endif
                    STACK_PUSH(make_uint2(0xffffffffu, 0u));
                    if (STATS) { st_sp = max(st_sp, (uint32_t)sp); ++st_insts; }
                    const float4* ip = P.insts + (size_t)(tgroup.x + bit) * 4;
                    const float4 m0 = __ldg(ip), m1 = __ldg(ip + 1), m2 = __ldg(ip + 2);
                    const int4 id = __ldg((const int4*)(ip + 3));
                    enter_instance(r, m0, m1, m2);
                    cur_inst = id.x;
                    ngroup = make_uint2((uint32_t)id.y, 0x80000000u);
                    tgroup = make_uint2(0, 0);
                    break;
                }
                const float4* tp = tris + (size_t)(tgroup.x + bit) * 3;
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                if (STATS) ++st_tris;
                if (!tri_test<MASKS, FLAT>(P, a, b, c, r, tmin, cur_inst, hr)) continue;
                if (ANY) { done = true; break; }
            }

            // ---- pop --------------------------------------------------------------------------------
            if (!done && ngroup.y <= 0x00ffffffu && tgroup.y == 0) {
                for (;;) {
                    if (sp == 0) { done = true; break; }
                    const uint2 e = STACK_POP();
                    if (!FLAT && e.y == 0u) { // sentinel: leave the instance, restore the world-space ray saved under it
#ifndef IMBA_RELOAD_WORLD_RAY
                        const uint2 s4 = STACK_POP(), s3 = STACK_POP(), s2 = STACK_POP(), s1 = STACK_POP(), s0 = STACK_POP();
                        r.ox = __uint_as_float(s0.x); r.oy = __uint_as_float(s0.y); r.oz = __uint_as_float(s1.x); r.dx = __uint_as_float(s1.y);
                        r.dy = __uint_as_float(s2.x); r.dz = __uint_as_float(s2.y); r.ix = __uint_as_float(s3.x); r.iy = __uint_as_float(s3.y);
                        r.iz = __uint_as_float(s4.x); r.octinv4 = s4.y;
#else   // round-1 behaviour (A/B switch): reload the ray from the queue and divide again
                        const float4 o = __ldg(P.rays + 2 * ray_idx), d = __ldg(P.rays + 2 * ray_idx + 1);
                        r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z; r.finish();
#endif
                        cur_inst = -1;
                        continue;
                    }
                    ngroup = e;
                    break;
                }
            }
            if (done) {
                if (RESOLVE) {
                    if (hr.tri_best == IMBA_NO_HIT) {
                        const float4 st = __ldg(reinterpret_cast<const float4*>(P.shadow_states) + ray_idx);
                        float* px = P.framebuffer + 3 * (size_t)__float_as_int(st.x);
                        atomicAdd(px, st.y); atomicAdd(px + 1, st.z); atomicAdd(px + 2, st.w);
                    }
                } else {
                    int4 h;
                    if (hr.tri_best != IMBA_NO_HIT) h = make_int4(hr.tri_best, FLAT ? P.flat_inst : hr.inst_best, __float_as_int(hr.tbest), __float_as_int(hr.ubest));
                    else                           h = make_int4(-1, -1, __float_as_int(hr.tbest), 0);
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

#undef STACK_PUSH
#undef STACK_POP

/* Register budgets.  Two-level kernel: no explicit minimum (an explicit 1 makes ptxas spend 92-96 registers; open it takes 56 =
 * 9 resident blocks; 48 spills).  Flat kernel: without the instance level it fits 48 registers without spilling, 10 resident
 * blocks = 40 warps per SM: +3.4 % on the bench batches (11 / 12 blocks spill: -3 %). */
#ifndef IMBA_FLAT_MIN_BLOCKS
#define IMBA_FLAT_MIN_BLOCKS 10
#endif
#ifndef IMBA_FLAT_RESOLVE_TIGHT
#define IMBA_FLAT_RESOLVE_TIGHT 0   // the fused resolve mode and the counting variants spill at 48 registers: they keep 56
#endif
template <int MODE, bool MASKS, bool STATS>
#ifdef IMBA_MIN_BLOCKS
__global__ void __launch_bounds__(BLOCK_THREADS, IMBA_MIN_BLOCKS) trace_kernel(const TraceArgs P) { trace_body<MODE, MASKS, STATS, false>(P); }
#else
__global__ void __launch_bounds__(BLOCK_THREADS) trace_kernel(const TraceArgs P) { trace_body<MODE, MASKS, STATS, false>(P); }
#endif
template <int MODE, bool MASKS, bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS, (STATS || (MODE == 2 && !IMBA_FLAT_RESOLVE_TIGHT)) ? 9 : IMBA_FLAT_MIN_BLOCKS) trace_flat_kernel(const TraceArgs P) { trace_body<MODE, MASKS, STATS, true>(P); }

#ifdef IMBA_COOP_TRACE
#include "traverse_coop.cuh"
#endif

#ifndef IMBA_NO_FLAT_KERNEL
#define IMBA_NO_FLAT_KERNEL 0   // A/B switch: 1 = flat scenes run the two-level kernel (round-2a behaviour)
#endif

template <int ANY, bool MASKS, bool STATS>
cudaError_t launch(const TraceArgs& a, int grid, cudaStream_t s) {
#ifndef IMBA_COOP_TRACE
    if (a.flat_inst >= 0 && !IMBA_NO_FLAT_KERNEL) trace_flat_kernel<ANY, MASKS, STATS><<<grid, BLOCK_THREADS, 0, s>>>(a);
    else                  trace_kernel<ANY, MASKS, STATS><<<grid, BLOCK_THREADS, 0, s>>>(a);
#else
    const int smem = (BLOCK_THREADS / 32) * (a.flat_inst >= 0 ? COOP_WARP_BYTES_FLAT : COOP_WARP_BYTES_INST);
    trace_coop_kernel<ANY, MASKS, STATS><<<grid, BLOCK_THREADS, smem, s>>>(a);
#endif
    return cudaGetLastError();
}

template <int ANY, bool MASKS, bool STATS>
int occupancy(bool flat) {
    int n = 0;
#ifndef IMBA_COOP_TRACE
    if (flat && !IMBA_NO_FLAT_KERNEL) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_flat_kernel<ANY, MASKS, STATS>, BLOCK_THREADS, 0);
    else      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_kernel<ANY, MASKS, STATS>, BLOCK_THREADS, 0);
#else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, trace_coop_kernel<ANY, MASKS, STATS>, BLOCK_THREADS,
                                                  (BLOCK_THREADS / 32) * (flat ? COOP_WARP_BYTES_FLAT : COOP_WARP_BYTES_INST));
#endif
    return n;
}

} // namespace

int imba_trace_block_threads() { return BLOCK_THREADS; }

#define FOR_ALL_VARIANTS(X) \
    X(0, false, false) X(0, false, true) X(0, true, false) X(0, true, true) \
    X(1, false, false) X(1, false, true) X(1, true, false) X(1, true, true) \
    X(2, false, false) X(2, true, false)

int imba_trace_blocks_per_sm(int mode, bool masks, bool stats, bool flat) {
#define OCC(A, M, S) if (mode == A && masks == M && stats == S) return occupancy<A, M, S>(flat);
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
