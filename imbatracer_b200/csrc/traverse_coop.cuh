/* traverse_coop.cuh -- EXPERIMENT (round 2), not in the default build: a lock-step, warp-cooperative variant of trace_kernel.
 * Included by traverse.cu only with -DIMBA_COOP_TRACE (tools/build_variants.sh coop:"-DIMBA_COOP_TRACE").  Results are
 * bit-identical to trace_kernel's (tests/test_traverse_gpu.py passes with it).  Measured on the bench batches (DESIGN.md
 * section 3, profiles/r2a_coop_*): lanes per instruction 16.3 -> 24.1 on the incoherent closest-hit batch, but 4 230 instead of
 * 2 715 thread instructions per ray (ring management, prefix sums, ray shuffles, shared-memory atomics), so 4.45 instead of
 * 4.70 Grays/s -- the divergent triangle loop of trace_kernel is cheaper than the machinery that removes the divergence. */
/* ---- warp-cooperative variant ------------------------------------------------------------------------------------------------
 * The loop above spends ~40 % of its warp instructions in the triangle loop at ~5 of 32 lanes (profiles/r1f_trace_lines_*):
 * only the lanes whose current node has hit a leaf are in it.  Here the warp runs in lock step instead: every iteration each
 * lane that has node work pops / intersects ONE node; the triangles its leaves produce are not tested by the lane itself but
 * DEPOSITED as one (owner lane | first triangle, 24-bit hit mask) group in a per-warp shared-memory ring, and when enough have
 * accumulated all 32 lanes test one triangle each: the groups at the head of the ring are expanded by a warp prefix sum, the
 * owner's ray comes by shuffle, its current best t from shared memory.  Candidates return through a 64-bit shared-memory
 * atomicMin on (ordered t, tri_id) -- the (t, tri, inst) order of tri_test, so the hit records stay bit-identical whatever
 * the order of the tests -- followed by the winner writing its exact t, u (and, in instanced scenes, the instance id, ties
 * going to the smaller one).
 *   shared memory per warp: key[32] u64 | tu[32] float2 | ring[RING] uint2 | (instanced scenes only) inst[32] i32 | wray[6][32] f32
 *   a lane may neither finish nor change space (leave / enter an instance) while groups of its are pending: the shuffled ray
 *   registers must be the ones the groups were deposited with.  Such lanes wait ("want"), and their number forces a flush. */
#ifndef IMBA_RING
#define IMBA_RING 64
#endif
#ifndef IMBA_FLUSH_MIN
#define IMBA_FLUSH_MIN 28    /* pending triangles that trigger a flush */
#endif
#ifndef IMBA_WANT_MIN
#define IMBA_WANT_MIN 3      /* waiting lanes that trigger a flush */
#endif
#ifndef IMBA_REFILL_IDLE
#define IMBA_REFILL_IDLE 8   /* idle lanes that trigger a refill */
#endif
constexpr uint32_t RING = IMBA_RING, RING_MASK = RING - 1;
constexpr int      COOP_WARP_BYTES_FLAT = 256 + 256 + 8 * (int)RING;
constexpr int      COOP_WARP_BYTES_INST = COOP_WARP_BYTES_FLAT + 128 + 6 * 128;
static_assert((RING & RING_MASK) == 0 && RING >= 64, "ring size");

__device__ __forceinline__ unsigned long long pack_key(float t, int tri) {
    const uint32_t b = __float_as_uint(t + 0.0f);                    // -0 -> +0: equal t must compare equal
    const uint32_t k = b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u); // total order of the floats
    return ((unsigned long long)k << 32) | (uint32_t)tri;
}
__device__ __forceinline__ unsigned fetch_add_u32(unsigned* p, unsigned v) { // (plain atomicAdd under `lane == leader` is rewritten into a warp scan)
    unsigned old; asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory"); return old;
}

template <int MODE, bool MASKS, bool STATS>
__global__ void __launch_bounds__(BLOCK_THREADS) trace_coop_kernel(const TraceArgs P) {
    constexpr bool ANY = MODE != 0;
    constexpr bool RESOLVE = MODE == 2;
    extern __shared__ uint4 coop_smem[];
    const int ray_count = P.count_ptr ? *P.count_ptr * P.count_mul : P.count;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool flat = P.flat_inst >= 0;
    const uint4*  __restrict__ nodes = P.nodes;
    const float4* __restrict__ tris  = P.tris;
    char* wbase = reinterpret_cast<char*>(coop_smem) + (threadIdx.x >> 5) * (flat ? COOP_WARP_BYTES_FLAT : COOP_WARP_BYTES_INST);
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(wbase);
    float2*   s_tu   = reinterpret_cast<float2*>(wbase + 256);
    uint2*    s_ring = reinterpret_cast<uint2*>(wbase + 512);
    int*      s_inst = reinterpret_cast<int*>(wbase + COOP_WARP_BYTES_FLAT);
    float*    s_wray = reinterpret_cast<float*>(wbase + COOP_WARP_BYTES_FLAT + 128);

    uint2 stack[STACK_ENTRIES];
    int sp = 0;
    RayRegs r; r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = r.ix = r.iy = r.iz = 0.0f; r.octinv4 = 0;
    float tmin = 0.0f;
    int cur_inst = -1, ray_idx = -1;
    uint2 ngroup = make_uint2(0, 0), tgroup = make_uint2(0, 0);
    uint32_t st_nodes = 0, st_tris = 0, st_insts = 0, st_sp = 0;
    uint32_t head = 0, tail = 0, ntri = 0; // ring positions in groups and pending triangle count (warp-uniform)
    uint32_t my_last = 0;                  // ring position after this lane's last group
    bool active = false, exhausted = false;

    // expands the groups at the head of the ring into (at most) 32 triangles and tests one per lane
    auto flush_trip = [&]() {
        __syncwarp();
        const uint32_t pos = head + lane;
        const bool have = (int32_t)(tail - pos) > 0;
        const uint2 g = have ? s_ring[pos & RING_MASK] : make_uint2(0, 0);
        const uint32_t c = __popc(g.y);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, d); if ((int)lane >= d) incl += v; }
        const bool fits = have && incl <= 32u;                 // a prefix of the groups
        const unsigned groups = __popc(__ballot_sync(FULL, fits));
        const uint32_t excl = incl - c;
        const uint32_t starts = __reduce_or_sync(FULL, fits ? (1u << excl) : 0u);
        const uint32_t total = __shfl_sync(FULL, incl, groups - 1u);
        bool valid = lane < total;
        const unsigned j = __popc(starts & (0xffffffffu >> (31u - lane))) - 1u;   // the group this lane's triangle belongs to
        const uint32_t gx = __shfl_sync(FULL, g.x, j), gy = __shfl_sync(FULL, g.y, j), gfirst = __shfl_sync(FULL, excl, j);
        uint32_t m = gy;
        for (uint32_t q = valid ? lane - gfirst : 0u; q != 0; --q) m &= m - 1u;
        const uint32_t tri_index = (gx & 0x07ffffffu) + (uint32_t)__ffs(m) - 1u;
        const unsigned owner = gx >> 27;
        const float ox = __shfl_sync(FULL, r.ox, owner), oy = __shfl_sync(FULL, r.oy, owner), oz = __shfl_sync(FULL, r.oz, owner);
        const float dx = __shfl_sync(FULL, r.dx, owner), dy = __shfl_sync(FULL, r.dy, owner), dz = __shfl_sync(FULL, r.dz, owner);
        const float otmin = __shfl_sync(FULL, tmin, owner);
        const int oinst = flat ? P.flat_inst : __shfl_sync(FULL, cur_inst, owner);
        const float tb = s_tu[owner].x;
        if (ANY && reinterpret_cast<const uint32_t*>(s_key)[2 * owner] != (uint32_t)IMBA_NO_HIT) valid = false;   // the owner is occluded already
        bool pass = false; float t = 0.0f, u = 0.0f; int tri_id = 0;
        if (valid) {
            const float4* tp = tris + (size_t)tri_index * 3;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c2 = __ldg(tp + 2);
            tri_id = __float_as_int(a.w);
            const float nx = b.y * c2.z - b.z * c2.y, ny = b.z * c2.x - b.x * c2.z, nz = b.x * c2.y - b.y * c2.x;
            const float cx = a.x - ox, cy = a.y - oy, cz = a.z - oz;
            const float rx = __fmaf_rn(dy, cz, -(dz * cy));
            const float ry = __fmaf_rn(dz, cx, -(dx * cz));
            const float rz = __fmaf_rn(dx, cy, -(dy * cx));
            const float det = dot_f(nx, ny, nz, dx, dy, dz);
            const float adet = fabsf(det);
            const uint32_t sgn = __float_as_uint(det) & 0x80000000u;
            const float U = xorf(dot_f(rx, ry, rz, c2.x, c2.y, c2.z), sgn);
            const float V = xorf(dot_f(rx, ry, rz, b.x, b.y, b.z), sgn);
            if (U >= 0.0f && V >= 0.0f && U + V <= adet && det != 0.0f) {
                const float T = xorf(dot_f(nx, ny, nz, cx, cy, cz), sgn);
                if (T >= adet * otmin && T <= adet * tb * IMBA_T_PREFILTER) {
                    const float inv = 1.0f / adet;
                    t = T * inv; u = U * inv;
                    pass = t <= tb;
                    if (MASKS && pass) pass = mask_opaque<MASKS>(P, tri_id, u, V * inv);
                }
            }
        }
        const unsigned long long key = pack_key(t, tri_id);
        if (pass) {
            const unsigned long long old = atomicMin(&s_key[owner], key);
            if (!flat && key < old) s_inst[owner] = IMBA_NO_HIT;   // a new best (t, tri): instance ids of the old one do not count
        }
        __syncwarp();
        if (!flat) {
            if (pass && s_key[owner] == key) atomicMin(&s_inst[owner], oinst);
            __syncwarp();
        }
        if (pass && s_key[owner] == key && (flat || s_inst[owner] == oinst)) s_tu[owner] = make_float2(t, u);
        head += groups; ntri -= total;
        __syncwarp();
    };

    for (;;) {
        // ---- refill idle lanes from the global work counter --------------------------------------
        {
            unsigned idle = __ballot_sync(FULL, !active);
            if (!exhausted && __popc(idle) >= IMBA_REFILL_IDLE) {
                for (int round = 0; round < (RESOLVE ? 4 : 1) && !exhausted; ++round) {
                    const unsigned need = round ? __ballot_sync(FULL, !active) : idle;
                    if (!need) break;
                    const int leader = __ffs(need) - 1;
                    unsigned base = 0;
                    if ((int)lane == leader) base = fetch_add_u32(P.counter, (unsigned)__popc(need));
                    base = __shfl_sync(FULL, base, leader);
                    if (base + (unsigned)__popc(need) >= (unsigned)ray_count) exhausted = true; // warp-uniform
                    if (!active) {
                        const unsigned i = base + (unsigned)__popc(need & lt_mask);
                        bool take = i < (unsigned)ray_count;
                        if (RESOLVE && take) take = __ldg(reinterpret_cast<const int*>(P.shadow_states) + 4 * i) >= 0;
                        if (take) {
                            ray_idx = (int)i;
                            const float4 o = __ldcs(P.rays + 2 * i), d = __ldcs(P.rays + 2 * i + 1);
                            r.ox = o.x; r.oy = o.y; r.oz = o.z; r.dx = d.x; r.dy = d.y; r.dz = d.z;
                            if (flat) { // the identity transform of the skipped instance turns -0 into +0 (1 * x + 0)
                                r.ox += 0.0f; r.oy += 0.0f; r.oz += 0.0f; r.dx += 0.0f; r.dy += 0.0f; r.dz += 0.0f;
                            }
                            r.finish();
                            tmin = o.w; cur_inst = P.flat_inst;
                            s_key[lane] = pack_key(d.w, IMBA_NO_HIT); s_tu[lane] = make_float2(d.w, 0.0f);
                            if (!flat) s_inst[lane] = IMBA_NO_HIT;
                            sp = 0; ngroup = make_uint2(P.top_root, 0x80000000u); tgroup = make_uint2(0, 0);
                            my_last = head;
                            if (STATS) { st_nodes = st_tris = st_insts = st_sp = 0; }
                            active = true;
                        }
                    }
                }
                idle = __ballot_sync(FULL, !active);
            }
            if (idle == FULL) { if (exhausted) break; continue; }
        }

        // ---- pop: lanes whose current node group has no inner children left ----------------------------
        bool want = false;      // blocked until this lane's pending groups have been tested
        if (active && ngroup.y <= 0x00ffffffu) {
            if (sp == 0) want = true;                           // (a finished lane without pending groups was retired below)
            else {
                const uint2 e = stack[--sp];
                if (e.y != 0u) ngroup = e;
                else if ((int32_t)(my_last - head) > 0) { ++sp; want = true; }
                else { // sentinel: leave the instance, restore the world-space ray; the next iteration pops again (or retires the lane)
                    r.ox = s_wray[lane]; r.oy = s_wray[32 + lane]; r.oz = s_wray[64 + lane];
                    r.dx = s_wray[96 + lane]; r.dy = s_wray[128 + lane]; r.dz = s_wray[160 + lane]; r.finish();
                    cur_inst = -1; ngroup = make_uint2(0, 0);
                }
            }
        }
        __syncwarp();

        // ---- one node per lane ---------------------------------------------------------------------
        if (active && !want && ngroup.y > 0x00ffffffu) {
            // the nearest pending inner child of the current node group
            const uint32_t hits = ngroup.y;
            const uint32_t bit = 31u - (uint32_t)__clz(hits);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00ffffffu) { stack[sp++] = ngroup; if (STATS) st_sp = max(st_sp, (uint32_t)sp); }
            const uint32_t slot = (bit - 24u) ^ (r.octinv4 & 0xffu);
            const uint32_t rel  = __popc(hits & 0xffu & ~(0xffffffffu << slot));
            const uint4* np = nodes + (size_t)(ngroup.x + rel) * 5;
            const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (STATS) ++st_nodes;
            const uint32_t hitmask = node_step(n0, n1, n2, n3, n4, r, tmin, s_tu[lane].x);
            ngroup = make_uint2(n1.x, (hitmask & 0xff000000u) | (n0.w >> 24));
            tgroup = make_uint2(n1.y, hitmask & 0x00ffffffu);
        } else if (active && !want && ngroup.y != 0u) {
            tgroup = ngroup; ngroup = make_uint2(0, 0);           // a popped group of top-level leaves
        }
        __syncwarp();
        if (!flat && active && !want && cur_inst < 0 && tgroup.y != 0) {
            // top level: enter the instance (ray * inverse transform, cpu_adapter.cpp:356-365); the lane has no pending groups here
            const uint32_t bit = 31u - (uint32_t)__clz(tgroup.y);
            tgroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
            if (tgroup.y != 0) stack[sp++] = tgroup;
            stack[sp++] = make_uint2(0xffffffffu, 0u);
            if (STATS) { st_sp = max(st_sp, (uint32_t)sp); ++st_insts; }
            const float4* ip = P.insts + (size_t)(tgroup.x + bit) * 4;
            const float4 m0 = __ldg(ip), m1 = __ldg(ip + 1), m2 = __ldg(ip + 2);
            const int4 id = __ldg((const int4*)(ip + 3));
            s_wray[lane] = r.ox; s_wray[32 + lane] = r.oy; s_wray[64 + lane] = r.oz;
            s_wray[96 + lane] = r.dx; s_wray[128 + lane] = r.dy; s_wray[160 + lane] = r.dz;
            enter_instance(r, m0, m1, m2);
            cur_inst = id.x;
            ngroup = make_uint2((uint32_t)id.y, 0x80000000u);
            tgroup = make_uint2(0, 0);
        }

        // ---- deposit the triangle group found -------------------------------------------------------
        {
            while (tail - head > RING - 32u) flush_trip();         // room for one group per lane
            const uint32_t c = active && !want ? (uint32_t)__popc(tgroup.y) : 0u;
            const unsigned m = __ballot_sync(FULL, c != 0u);
            if (c != 0u) {
                const uint32_t at = tail + __popc(m & lt_mask);
                s_ring[at & RING_MASK] = make_uint2((lane << 27) | tgroup.x, tgroup.y);
                my_last = at + 1u;
                if (STATS) st_tris += c;
                tgroup.y = 0;
            }
            tail += __popc(m); ntri += __reduce_add_sync(FULL, c);
        }

        // ---- test the triangles when there are enough, or when lanes wait for theirs ---------------------
        {
            const bool last = active && ngroup.y <= 0x00ffffffu && sp == 0;     // this lane finishes once its groups are tested
            const bool pending = (int32_t)(my_last - head) > 0;
            const unsigned waiting = __ballot_sync(FULL, (want || last) && pending);
            const unsigned moving  = __ballot_sync(FULL, active && !want && !last);
            if (ntri != 0 && (__popc(waiting) >= IMBA_WANT_MIN || moving == 0)) {
                while (head != tail) flush_trip();
            } else {
                while (ntri >= IMBA_FLUSH_MIN) flush_trip();
            }
        }

        // ---- retire finished lanes --------------------------------------------------------------------
        if (active) {
            const bool pending = (int32_t)(my_last - head) > 0;
            const bool hit = reinterpret_cast<const uint32_t*>(s_key)[2 * lane] != (uint32_t)IMBA_NO_HIT;
            if (ANY && hit) { sp = 0; ngroup.y = 0; }             // occluded: nothing more to find
            if (!pending && ngroup.y <= 0x00ffffffu && sp == 0) {
                if (RESOLVE) {
                    if (!hit) {
                        const float4 st = __ldg(reinterpret_cast<const float4*>(P.shadow_states) + ray_idx);
                        float* px = P.framebuffer + 3 * (size_t)__float_as_int(st.x);
                        atomicAdd(px, st.y); atomicAdd(px + 1, st.z); atomicAdd(px + 2, st.w);
                    }
                } else {
                    const float2 tu = s_tu[lane];
                    int4 h;
                    if (hit) h = make_int4((int)(uint32_t)s_key[lane], flat ? P.flat_inst : s_inst[lane], __float_as_int(tu.x), __float_as_int(tu.y));
                    else     h = make_int4(-1, -1, __float_as_int(tu.x), 0);
                    __stcs(P.hits + ray_idx, h);
                }
                if (STATS) reinterpret_cast<uint4*>(P.stats)[ray_idx] = make_uint4(st_nodes, st_tris, st_insts, st_sp);
                active = false;
            }
        }
    }
}

