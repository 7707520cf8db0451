/* vcm.cu -- device stages of VCMIntegrator<algo> for algo = VCM, BPT, PPM, LT and the VCM-style PT
 * (render/integrators/vcm.cpp, vcm.h), of the light-vertex cache (light_vertices.h:59-123) and of the photon hash grid
 * (rangesearch/rangesearch.h:25-140).
 *
 *   vcm_light_raygen    trace_light_paths' sample function + LightRayGen / UniformLightTileGen   vcm.cpp:64-96, ray_gen.h:122-152, tile_gen.h:102-158
 *   vcm_light_shade     process_light_rays, connect_to_camera, bounce(adjoint)                  vcm.cpp:193-258, 261-320, 133-190
 *   vcm_camera_raygen   trace_camera_paths' sample function                                       vcm.cpp:109-129
 *   vcm_camera_shade    process_camera_rays, direct_illum, connect, vertex_merging, bounce        vcm.cpp:323-439, 442-486, 489-569, 572-611
 *   grid_*              HashGrid::build (bbox, count, scan, scatter) and ::query (8 cells, k-NN)  rangesearch.h:25-129
 *   vcm_merge_query     the k-NN search of a camera wavefront as its own pass (lists in registers for k = 4, 8, 10, 16)
 *   vcm_light_tail*     the last <= 128 Ki light paths: one thread, then one warp per path (trace_single / trace_warp)
 * mis_pow(a) = powf(a, 1) = a (vcm.h:112-115) is applied as the identity.
 */
#include "integrator.cuh"
#include "traverse_core.cuh"
#include "vcm.cuh"

#ifndef IMBA_VCM_CAM_MIN_BLOCKS
#define IMBA_VCM_CAM_MIN_BLOCKS 1   /* tuning switch: resident blocks per SM the register allocation of the camera shading kernel aims at */
#endif
#ifndef IMBA_VCM_LIGHT_MIN_BLOCKS
#define IMBA_VCM_LIGHT_MIN_BLOCKS 1
#endif
namespace {

__device__ __forceinline__ void add_contribution(float* fb, int pixel, V3 c) {
    float* p = fb + 3 * (size_t)pixel;
    atomicAdd(p, c.x); atomicAdd(p + 1, c.y); atomicAdd(p + 2, c.z);
}

__device__ __forceinline__ float shading_normal_adjoint(V3 normal, V3 geom_normal, V3 out_dir, V3 in_dir) {
    return dot(out_dir, normal) * dot(in_dir, geom_normal) / dot(out_dir, geom_normal);
}

__device__ __forceinline__ Bsdf vertex_bsdf(const DeviceShading& S, const DevLightVertex& v, Isect& is) {
    is.pos = v3(v.pos_len); is.out_dir = v3(v.out_mat); is.normal = v3(v.n_dvc); is.geom_normal = v3(v.gn_dvcm);
    is.u_tangent = v3(v.ut_dvm); is.v_tangent = v3(v.vt_u); is.uv = make_float2(v.vt_u.w, v.thr_v.w); is.mat = __float_as_int(v.out_mat.w);
    return make_bsdf(S, is, true);
}

/* ---- light pass ------------------------------------------------------------------------------------------- */
__global__ void vcm_light_raygen_kernel(const DeviceShading S, const VcmParams* Pd, PtSeeds seeds, int per_iteration, int total, float4* rays, VCMState* states, int* count_out) {
    const VcmParams P = *Pd;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot == 0) *count_out = total;
    if (slot >= total) return;
    // the light tracer may trace the light paths of several iterations as one pass (iterations_per_pass): slot -> (iteration of
    // the group, path of that iteration); everything below is per iteration, as in a pass of its own
    const int group = total / per_iteration, count = per_iteration;
    const int p = slot / group;
    const uint32_t seed_base = seeds.base[slot % group];
    // UniformLightTileGen: path_count / light_count paths per light, the remainder goes to light 0
    const int per = count / S.light_count, first = per + count % S.light_count;
    const int light = (p < first || per == 0) ? 0 : 1 + (p - first) / per;
    Rng rng; rng.s = seed_primary(seed_base, p, p);
    const float pdf_lightpick = 1.0f / S.light_count;
    const EmitSample s = light_sample_emit(S, light, rng);
    rays[2 * slot] = make_float4(s.pos.x, s.pos.y, s.pos.z, 1e-3f);
    rays[2 * slot + 1] = make_float4(s.dir.x, s.dir.y, s.dir.z, FLT_MAX);
    VCMState st;
    st.pixel_id = p; st.light_id = light; st.pad = 0;
    const V3 thr = s.radiance / pdf_lightpick;
    st.tr = thr.x; st.tg = thr.y; st.tb = thr.z;
    st.len_finite = 1u | (s.finite ? 0x80000000u : 0u);
    st.dVCM = s.pdf_direct_a / s.pdf_emit_w;
    st.dVC = s.delta ? 0.0f : s.cos_out / (s.pdf_emit_w * pdf_lightpick);
    st.dVM = st.dVC * P.mis_eta_vc;
    st.rng = rng.s;
    states[slot] = st;
}

/* bounce (vcm.cpp:133-190); returns false when the path ends */
template <int ALGO>
__device__ __forceinline__ bool vcm_bounce(const VcmParams& P, VCMState& st, Rng& rng, const Isect& is, const Bsdf& bsdf, bool adjoint, float offset,
                                           float4& ray_o, float4& ray_d) {
    V3 thr = v3(st.tr, st.tg, st.tb);
    const float kill_prob = luminance(thr) * 2.0f;
    const float rr_pdf = fminf(1.0f, kill_prob);
    if (!(rng.uniform() < rr_pdf)) return false;
    const bool specular_only = ALGO == IMBA_ALGO_PPM && !adjoint;
    float pdf_dir_w; V3 dir; int flags;
    const V3 value = bsdf.sample(is.out_dir, dir, rng, flags, pdf_dir_w, specular_only);
    const bool is_specular = (flags & BSDF_SPECULAR) != 0;
    if (flags == 0 || pdf_dir_w == 0.0f || is_black(value)) return false;
    float pdf_rev_w = pdf_dir_w;
    if (!is_specular) pdf_rev_w = bsdf.pdf(dir, is.out_dir);
    const float cos_theta_i = adjoint ? fabsf(shading_normal_adjoint(is.normal, is.geom_normal, is.out_dir, dir)) : fabsf(dot(dir, is.normal));
    if (is_specular) {
        st.dVCM = 0.0f; st.dVC *= cos_theta_i; st.dVM *= cos_theta_i;
    } else {
        st.dVC = (cos_theta_i / pdf_dir_w) * (st.dVC * pdf_rev_w + st.dVCM + P.mis_eta_vm);
        st.dVM = (cos_theta_i / pdf_dir_w) * (st.dVM * pdf_rev_w + st.dVCM * P.mis_eta_vc + 1.0f);
        st.dVCM = 1.0f / pdf_dir_w;
    }
    thr *= value * cos_theta_i / (rr_pdf * pdf_dir_w);
    st.tr = thr.x; st.tg = thr.y; st.tb = thr.z;
    st.len_finite = ((st.len_finite & 0x7fffffffu) + 1u) | (st.len_finite & 0x80000000u);
    ray_o = make_float4(is.pos.x, is.pos.y, is.pos.z, offset);
    ray_d = make_float4(dir.x, dir.y, dir.z, FLT_MAX);
    return true;
}

/* One light-path vertex (process_light_rays, vcm.cpp:193-258): completes the MIS partials, appends to the vertex cache,
 * builds the camera connection (shadow ray + contribution for the target pixel) and samples the next direction.
 * Returns false when the path ends; otherwise (ro, rd, st) hold the continuation. */
template <int ALGO>
__device__ __forceinline__ bool light_vertex(const DeviceShading& S, const VcmParams& P, const DevCamera& cam, float4& ro, float4& rd, const int4 h, VCMState& st,
                                             DevLightVertex* cache, int cache_capacity, int* cache_last, bool& has_shadow, float4& sh_o, float4& sh_d, ShadowState& ss) {
    has_shadow = false;
    const float t = __int_as_float(h.z);
    const Isect is = calc_intersection(S, ro, rd, h.x, h.y, t, __int_as_float(h.w));
    const float cos_theta_o = fabsf(dot(is.out_dir, is.normal));
    if (cos_theta_o == 0.0f) return false;
    const int path_length = (int)(st.len_finite & 0x7fffffffu);
    if (path_length > 1 || (st.len_finite >> 31)) st.dVCM *= sqr(t);
    st.dVCM *= 1.0f / cos_theta_o; st.dVC *= 1.0f / cos_theta_o; st.dVM *= 1.0f / cos_theta_o;
    const Bsdf bsdf = make_bsdf(S, is, true);
    const V3 thr = v3(st.tr, st.tg, st.tb);
    if (!bsdf.is_specular()) {
        if (ALGO != IMBA_ALGO_LT) { // add_vertex_to_cache (light_vertices.h:77-84)
            const int slot = atomicAdd(cache_last, 1);
            if (slot < cache_capacity) {
                DevLightVertex v;
                v.pos_len = make_float4(is.pos.x, is.pos.y, is.pos.z, __int_as_float(path_length + 1));
                v.out_mat = make_float4(is.out_dir.x, is.out_dir.y, is.out_dir.z, __int_as_float(is.mat));
                v.n_dvc = make_float4(is.normal.x, is.normal.y, is.normal.z, st.dVC);
                v.gn_dvcm = make_float4(is.geom_normal.x, is.geom_normal.y, is.geom_normal.z, st.dVCM);
                v.ut_dvm = make_float4(is.u_tangent.x, is.u_tangent.y, is.u_tangent.z, st.dVM);
                v.vt_u = make_float4(is.v_tangent.x, is.v_tangent.y, is.v_tangent.z, is.uv.x);
                v.thr_v = make_float4(thr.x, thr.y, thr.z, is.uv.y);
                cache[slot] = v;
            }
        }
        if (ALGO != IMBA_ALGO_PPM) { // connect_to_camera (vcm.cpp:261-320)
            V3 dir_to_cam = v3(cam.px, cam.py, cam.pz) - is.pos;
            const V3 cam_dir = v3(cam.dx, cam.dy, cam.dz);
            if (!(dot(-dir_to_cam, cam_dir) < 0.0f)) {
                float depth;
                const float2 raster = cam.to_raster(is.pos, depth);
                const int px = (int)floorf(raster.x), py = (int)floorf(raster.y);
                const int pixel = (px < 0 || px >= cam.width || py < 0 || py >= cam.height) ? -1 : py * cam.width + px;
                if (pixel >= 0 && pixel < cam.width * cam.height && depth > 0.0f) {
                    const float dist_sqr = dot(dir_to_cam, dir_to_cam), dist = sqrtf(dist_sqr);
                    dir_to_cam = dir_to_cam / dist;
                    const float cos_theta_cam = fabsf(dot(cam_dir, -dir_to_cam));
                    const float cos_theta_surf = fabsf(shading_normal_adjoint(is.normal, is.geom_normal, is.out_dir, dir_to_cam));
                    const V3 bsdf_value = bsdf.eval(is.out_dir, dir_to_cam);
                    const float pdf_rev_w = bsdf.pdf(dir_to_cam, is.out_dir);
                    if (pdf_rev_w != 0.0f) {
                        const float img_to_surf = (sqr(cam.img_plane_dist) * cos_theta_surf) / (dist_sqr * cos_theta_cam * sqr(cos_theta_cam));
                        const float pdf_cam = img_to_surf;
                        const float mis_weight_light = (pdf_cam / P.light_path_count) * (P.mis_eta_vm + st.dVCM + st.dVC * pdf_rev_w);
                        const float mis_weight = ALGO == IMBA_ALGO_LT ? 1.0f : (1.0f / (mis_weight_light + 1.0f));
                        const V3 c = thr * (mis_weight * bsdf_value * img_to_surf / (float)P.light_path_count);
                        const float off = dist * 1e-3f;
                        sh_o = make_float4(is.pos.x, is.pos.y, is.pos.z, off);
                        sh_d = make_float4(dir_to_cam.x, dir_to_cam.y, dir_to_cam.z, dist - off);
                        ss.pixel_id = pixel; ss.r = c.x; ss.g = c.y; ss.b = c.z;
                        has_shadow = true;
                    }
                }
            }
        }
    }
    Rng rng; rng.s = st.rng;
    float4 no, nd;
    if (!vcm_bounce<ALGO>(P, st, rng, is, bsdf, true, t * 1e-4f, no, nd)) return false;
    st.rng = rng.s; ro = no; rd = nd;
    return true;
}

template <int ALGO>
__global__ void __launch_bounds__(128, IMBA_VCM_LIGHT_MIN_BLOCKS) vcm_light_shade_kernel(const DeviceShading S, const VcmParams* Pd, const DevCamera cam, float4* rays, const int4* hits, VCMState* states,
                                                              const int* sorted, const int* hit_count_ptr, const int* queue_count_ptr, float4* shadow_rays,
                                                              ShadowState* shadow_states, DevLightVertex* cache, int cache_capacity, int* cache_last,
                                                              unsigned long long* ray_stats, float4* out_rays, VCMState* out_states, int* out_count) {
    const VcmParams P = *Pd;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && ray_stats) atomicAdd(ray_stats, (unsigned long long)*queue_count_ptr);
    // sorted == null: queue order, every slot (the integrator's default, see pt_shade_kernel); else the reference's sorted hits
    if (i >= (sorted ? *hit_count_ptr : *queue_count_ptr)) return;
    const int idx = sorted ? sorted[i] : i;
    if (!sorted) {
        shadow_states[idx].pixel_id = -1;
        if (hits[idx].x < 0) { states[idx].pixel_id = -1; return; }
    }
    float4 ro = rays[2 * idx], rd = rays[2 * idx + 1];
    VCMState st = states[idx];
    bool has_shadow; float4 sh_o, sh_d; ShadowState ss;
    const bool alive = light_vertex<ALGO>(S, P, cam, ro, rd, hits[idx], st, cache, cache_capacity, cache_last, has_shadow, sh_o, sh_d, ss);
    if (has_shadow) {
        shadow_rays[2 * idx] = sh_o; shadow_rays[2 * idx + 1] = sh_d; shadow_states[idx] = ss;
        if (ray_stats) count_rays(ray_stats + 1);
    }
    if (alive) {
        const int o = out_count ? queue_push_slot(out_count) : idx;
        (out_count ? out_states : states)[o] = st;
        float4* dst = (out_count ? out_rays : rays) + 2 * (size_t)o;
        dst[0] = ro; dst[1] = rd;
    } else states[idx].pixel_id = -1;
}

/* Tail of the light pass: one thread follows one light path through ALL its remaining bounces (closest-hit traversal,
 * vertex processing, camera-connection shadow ray and splat inline).  Light paths are not length-capped (vcm.cpp:133-139);
 * in scenes with glass a few thousand paths bounce hundreds to thousands of times, which as wavefronts costs one
 * launch-bound round per bounce.  Results are those of the wavefront kernels (same functions, same RNG streams). */
template <int ALGO, bool MASKS>
__global__ void __launch_bounds__(64) vcm_light_tail_kernel(const DeviceShading S, const VcmParams* Pd, const DevCamera cam, const TraceArgs T, const float4* rays,
                                                            const VCMState* states, const int* count_ptr, DevLightVertex* cache, int cache_capacity, int* cache_last,
                                                            float* fb, unsigned long long* ray_stats, int max_bounces,
                                                            float4* out_rays, VCMState* out_states, int* out_count) {
    const VcmParams P = *Pd;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_ptr) return;
    float4 ro = rays[2 * i], rd = rays[2 * i + 1];
    VCMState st = states[i];
    unsigned long long traced = 0, shadows = 0;
    bool alive = true;
    for (int bounce = 0; bounce < max_bounces; ++bounce) {
        HitRegs hr;
        ++traced;
        if (!trace_single<false, MASKS>(T, ro, rd, hr)) { alive = false; break; }
        const int4 h = make_int4(hr.tri_best, hr.inst_best, __float_as_int(hr.tbest), __float_as_int(hr.ubest));
        bool has_shadow; float4 sh_o, sh_d; ShadowState ss;
        alive = light_vertex<ALGO>(S, P, cam, ro, rd, h, st, cache, cache_capacity, cache_last, has_shadow, sh_o, sh_d, ss);
        if (has_shadow) {
            ++shadows;
            HitRegs occ;
            if (!trace_single<true, MASKS>(T, sh_o, sh_d, occ)) add_contribution(fb, ss.pixel_id, v3(ss.r, ss.g, ss.b));
        }
        if (!alive) break;
    }
    if (alive && out_count) { // bounce budget used up: hand the path to the warp-per-path kernel
        const int slot = atomicAdd(out_count, 1);
        out_rays[2 * slot] = ro; out_rays[2 * slot + 1] = rd; out_states[slot] = st;
    }
    if (ray_stats) { atomicAdd(ray_stats, traced); atomicAdd(ray_stats + 1, shadows); }
}

/* The longest light paths (thousands of bounces inside glass): one WARP per path, traversal by trace_warp (traverse_core.cuh),
 * vertex processing by lane 0.  Same functions, same RNG stream, same results as the kernels above; only the latency of one
 * bounce differs. */
template <int ALGO, bool MASKS>
__global__ void __launch_bounds__(128) vcm_light_tail_warp_kernel(const DeviceShading S, const VcmParams* Pd, const DevCamera cam, const TraceArgs T, const float4* rays,
                                                                  const VCMState* states, const int* count_ptr, DevLightVertex* cache, int cache_capacity, int* cache_last,
                                                                  float* fb, unsigned long long* ray_stats, int max_bounces) {
    const VcmParams P = *Pd;
    const unsigned lane = threadIdx.x & 31u;
    const int warps = (gridDim.x * blockDim.x) >> 5, count = *count_ptr;
    unsigned long long traced = 0, shadows = 0;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < count; i += warps) {
        float4 ro = rays[2 * i], rd = rays[2 * i + 1];
        VCMState st = states[i];
        for (int bounce = 0; bounce < max_bounces; ++bounce) {
            HitRegs hr;
            ++traced;
            if (!trace_warp<false, MASKS>(T, ro, rd, hr)) break;
            const int4 h = make_int4(hr.tri_best, hr.inst_best, __float_as_int(hr.tbest), __float_as_int(hr.ubest));
            int flags = 0; float4 sh_o = make_float4(0, 0, 0, 0), sh_d = sh_o; ShadowState ss;
            if (lane == 0) {
                bool has_shadow;
                const bool alive = light_vertex<ALGO>(S, P, cam, ro, rd, h, st, cache, cache_capacity, cache_last, has_shadow, sh_o, sh_d, ss);
                flags = (alive ? 1 : 0) | (has_shadow ? 2 : 0);
            }
            __syncwarp();
            flags = __shfl_sync(0xffffffffu, flags, 0);
            ro.x = __shfl_sync(0xffffffffu, ro.x, 0); ro.y = __shfl_sync(0xffffffffu, ro.y, 0); ro.z = __shfl_sync(0xffffffffu, ro.z, 0); ro.w = __shfl_sync(0xffffffffu, ro.w, 0);
            rd.x = __shfl_sync(0xffffffffu, rd.x, 0); rd.y = __shfl_sync(0xffffffffu, rd.y, 0); rd.z = __shfl_sync(0xffffffffu, rd.z, 0); rd.w = __shfl_sync(0xffffffffu, rd.w, 0);
            if (flags & 2) {
                sh_o.x = __shfl_sync(0xffffffffu, sh_o.x, 0); sh_o.y = __shfl_sync(0xffffffffu, sh_o.y, 0); sh_o.z = __shfl_sync(0xffffffffu, sh_o.z, 0); sh_o.w = __shfl_sync(0xffffffffu, sh_o.w, 0);
                sh_d.x = __shfl_sync(0xffffffffu, sh_d.x, 0); sh_d.y = __shfl_sync(0xffffffffu, sh_d.y, 0); sh_d.z = __shfl_sync(0xffffffffu, sh_d.z, 0); sh_d.w = __shfl_sync(0xffffffffu, sh_d.w, 0);
                ++shadows;
                HitRegs occ;
                const bool blocked = trace_warp<true, MASKS>(T, sh_o, sh_d, occ);
                if (!blocked && lane == 0) add_contribution(fb, ss.pixel_id, v3(ss.r, ss.g, ss.b));
            }
            if (!(flags & 1)) break;
        }
    }
    if (ray_stats && lane == 0) { atomicAdd(ray_stats, traced); atomicAdd(ray_stats + 1, shadows); }
}

/* ---- photon hash grid (rangesearch.h) ------------------------------------------------------------------------- */
__device__ __forceinline__ int ordered_int(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float from_ordered(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256) grid_bbox_kernel(const DevLightVertex* cache, const int* last, int capacity, int* bbox /* 6 ordered ints */) {
    // grid-stride over the vertices, then warp and block reductions: 6 atomics per BLOCK (one per warp made 1.6 M atomics on six
    // addresses out of 8.8 M vertices: 1.7 ms)
    __shared__ float red[6][8];
    const int n = min(*last, capacity);
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = cache[i].pos_len;
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o)); }
        if (lane == 0) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][w]) : fmaxf(v, red[threadIdx.x][w]);
        if (threadIdx.x < 3) { if (v != FLT_MAX) atomicMin(bbox + threadIdx.x, ordered_int(v)); }
        else if (v != -FLT_MAX) atomicMax(bbox + threadIdx.x, ordered_int(v));
    }
}

/* bbox_.max += extents * 0.001f; bbox_.min -= ... (rangesearch.h:45-47); also publishes count_ = min(size, last_) */
__global__ void grid_prepare_kernel(const int* bbox, const int* last, int capacity, GridParams* g, float radius, unsigned* table_state) {
    const int n = min(*last, capacity);
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = from_ordered(bbox[a]); hi[a] = from_ordered(bbox[3 + a]); }
    if (n == 0) for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
    for (int a = 0; a < 3; ++a) { const float e = hi[a] - lo[a]; hi[a] += e * 0.001f; lo[a] -= e * 0.001f; }
    g->lo[0] = lo[0]; g->lo[1] = lo[1]; g->lo[2] = lo[2]; g->hi[0] = hi[0]; g->hi[1] = hi[1]; g->hi[2] = hi[2];
    g->radius_sqr = radius * radius; g->inv_cell_size = 1.f / (radius * 2.f); g->count = n;
    // the reference's table has 2 * photon_count cells and is only re-allocated when it must grow; the size matters for
    // results, not just speed: two of the 8 query cells that collide make their photons count twice (rangesearch.h:96-127)
    // (the size lives in *table_state, shared by the two double-buffered grids of the pipelined loop: it must follow the history
    // of ALL iterations, as the reference's single grid does)
    const unsigned want = (unsigned)n * 2u;
    unsigned t = *table_state;
    if (t < want) t = want;
    if (t == 0) t = 1;
    *table_state = t; g->table = t;
}

__device__ __forceinline__ unsigned cell_hash(unsigned x, unsigned y, unsigned z, unsigned table) {
    return ((x * 73856093u) ^ (y * 19349663u) ^ (z * 83492791u)) % table;
}
__device__ __forceinline__ unsigned cell_of(const GridParams& g, V3 p, unsigned table) {
    const int cx = (int)floorf(g.inv_cell_size * (p.x - g.lo[0])), cy = (int)floorf(g.inv_cell_size * (p.y - g.lo[1])), cz = (int)floorf(g.inv_cell_size * (p.z - g.lo[2]));
    return cell_hash((unsigned)cx, (unsigned)cy, (unsigned)cz, table);
}

__global__ void grid_count_kernel(const DevLightVertex* cache, const GridParams* gp, int* cells, unsigned) {
    const GridParams g = *gp; const unsigned table = g.table;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.count) return;
    atomicAdd(cells + cell_of(g, v3(cache[i].pos_len), table), 1);
}

/* multi-block exclusive scan of `cells` (3 kernels); block_sums has one int per 2048-element chunk */
constexpr int SCAN_CHUNK = 2048;
__global__ void __launch_bounds__(256) scan_local_kernel(int* data, int n, int* block_sums) {
    __shared__ int warp_sums[8];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int base = blockIdx.x * SCAN_CHUNK;
    for (int k = 0; k < SCAN_CHUNK; k += 256) {
        const int i = base + k + threadIdx.x;
        const int v = i < n ? data[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_sums[w];
        const int excl = carry + woff + x - v;
        if (i < n) data[i] = excl;
        __syncthreads();
        if (threadIdx.x == 255) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = carry;
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(int* sums, int n) { // single block, n <= 1024 * 64
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? sums[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_sums[w];
        const int excl = carry + woff + x - v;
        if (i < n) sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void scan_add_kernel(int* data, int n, const int* block_sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += block_sums[i / SCAN_CHUNK];
}

__global__ void grid_scatter_kernel(const DevLightVertex* cache, const GridParams* gp, int* cells, unsigned, DevPhoton* photons, float4* photon_pos) {
    const GridParams g = *gp; const unsigned table = g.table;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.count) return;
    const DevLightVertex v = cache[i];
    const int slot = atomicAdd(cells + cell_of(g, v3(v.pos_len), table), 1); // cells[c] ends up as the END of cell c
    DevPhoton p;
    p.pos_dvcm = make_float4(v.pos_len.x, v.pos_len.y, v.pos_len.z, v.gn_dvcm.w);
    p.dir_dvm = make_float4(v.out_mat.x, v.out_mat.y, v.out_mat.z, v.ut_dvm.w);
    p.thr = make_float4(v.thr_v.x, v.thr_v.y, v.thr_v.z, 0.0f);
    photons[slot] = p; photon_pos[slot] = p.pos_dvcm;
}

/* HashGrid::query (rangesearch.h:78-129): k nearest photons within the radius; returns their indices sorted by distance */
__device__ __forceinline__ int grid_query(const GridParams& g, const int* cells, unsigned table, const float4* photon_pos, V3 q, int k, int* out, float* dist) {
    if (!(q.x >= g.lo[0] && q.x <= g.hi[0] && q.y >= g.lo[1] && q.y <= g.hi[1] && q.z >= g.lo[2] && q.z <= g.hi[2])) return 0;
    const V3 cell = g.inv_cell_size * (q - v3(g.lo[0], g.lo[1], g.lo[2]));
    const V3 coord = v3(floorf(cell.x), floorf(cell.y), floorf(cell.z));
    const int px = (int)coord.x, py = (int)coord.y, pz = (int)coord.z;
    const V3 fr = cell - coord;
    const int pxo = px + (fr.x < 0.5f ? -1 : 1), pyo = py + (fr.y < 0.5f ? -1 : 1), pzo = pz + (fr.z < 0.5f ? -1 : 1);
    int count = 0;
    for (int j = 0; j < 8; ++j) {
        const unsigned c = cell_hash((unsigned)((j & 4) ? pxo : px), (unsigned)((j & 2) ? pyo : py), (unsigned)((j & 1) ? pzo : pz), table);
        int b = c == 0 ? 0 : cells[c - 1];
        const int e = cells[c];
        for (; b < e; ++b) {
            const float4 pp = __ldg(photon_pos + b);
            const V3 d = q - v3(pp);
            const float dist_sqr = dot(d, d);
            if (dist_sqr <= g.radius_sqr) {
                if (count == k) { if (dist[count - 1] < dist_sqr) continue; }
                else count++;
                dist[count - 1] = dist_sqr; out[count - 1] = b;
                for (int l = count - 2; l >= 0; --l) {
                    if (dist[l] > dist_sqr) { const float td = dist[l]; dist[l] = dist[l + 1]; dist[l + 1] = td; const int ti = out[l]; out[l] = out[l + 1]; out[l + 1] = ti; }
                    else break;
                }
            }
        }
    }
    return count;
}

/* ---- camera pass ---------------------------------------------------------------------------------------------- */
__global__ void vcm_camera_raygen_kernel(DevCamera cam, const VcmParams* Pd, int spp, uint32_t seed_base, int count, float4* rays, VCMState* states, int* count_out) {
    const VcmParams P = *Pd;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count_out = count;
    if (i >= count) return;
    const int pixel_idx = i / spp;
    const int y = pixel_idx / cam.width, x = pixel_idx % cam.width;
    Rng rng; rng.s = seed_primary(seed_base, i, pixel_idx);
    const float sx = (float)x + rng.uniform();
    const float sy = (float)y + rng.uniform();
    const V3 d = cam.ray_dir(sx, sy);
    rays[2 * i] = make_float4(cam.px, cam.py, cam.pz, 0.0f);
    rays[2 * i + 1] = make_float4(d.x, d.y, d.z, FLT_MAX);
    VCMState st;
    st.pixel_id = pixel_idx; st.light_id = i % spp; st.rng = rng.s; st.pad = 0;
    st.tr = st.tg = st.tb = 1.0f; st.len_finite = 1u;
    const float cos_theta_o = dot(d, v3(cam.dx, cam.dy, cam.dz));
    const float pdf_cam_w = sqr(cam.img_plane_dist / cos_theta_o) / cos_theta_o;
    st.dVC = 0.0f; st.dVM = 0.0f; st.dVCM = (float)P.light_path_count / pdf_cam_w;
    states[i] = st;
}

/* grid_query for a compile-time k with the candidate list in registers: same visiting order, same insertion rule (a new
 * candidate goes behind entries of equal distance; a full list takes it if it is not farther than the last entry), so the
 * result is the list grid_query returns -- without the local-memory traffic of the dynamically indexed insertion sort, which
 * was 60 % of the stall samples of this pass (ncu, stilllife-proxy: ~430 photons scanned and ~40 insertions per query). */
template <int K>
__device__ __forceinline__ int grid_query_fixed(const GridParams& g, const int* cells, unsigned table, const float4* photon_pos, V3 q, int* out) {
    if (!(q.x >= g.lo[0] && q.x <= g.hi[0] && q.y >= g.lo[1] && q.y <= g.hi[1] && q.z >= g.lo[2] && q.z <= g.hi[2])) return 0;
    const V3 cell = g.inv_cell_size * (q - v3(g.lo[0], g.lo[1], g.lo[2]));
    const V3 coord = v3(floorf(cell.x), floorf(cell.y), floorf(cell.z));
    const int px = (int)coord.x, py = (int)coord.y, pz = (int)coord.z;
    const V3 fr = cell - coord;
    const int pxo = px + (fr.x < 0.5f ? -1 : 1), pyo = py + (fr.y < 0.5f ? -1 : 1), pzo = pz + (fr.z < 0.5f ? -1 : 1);
    float dist[K]; int idx[K];
#pragma unroll
    for (int l = 0; l < K; ++l) { dist[l] = FLT_MAX; idx[l] = -1; }   // unused slots: farther than anything inside the radius
    int count = 0;
    for (int j = 0; j < 8; ++j) {
        const unsigned c = cell_hash((unsigned)((j & 4) ? pxo : px), (unsigned)((j & 2) ? pyo : py), (unsigned)((j & 1) ? pzo : pz), table);
        int b = c == 0 ? 0 : cells[c - 1];
        const int e = cells[c];
        for (; b < e; ++b) {
            const float4 pp = __ldg(photon_pos + b);
            const V3 d = q - v3(pp);
            const float ds = dot(d, d);
            if (ds <= g.radius_sqr && !(dist[K - 1] < ds)) {
                if (count < K) ++count;
                bool placed = false;
#pragma unroll
                for (int l = K - 1; l >= 1; --l) {
                    if (dist[l - 1] > ds) { dist[l] = dist[l - 1]; idx[l] = idx[l - 1]; }
                    else if (!placed) { dist[l] = ds; idx[l] = b; placed = true; }
                }
                if (!placed) { dist[0] = ds; idx[0] = b; }
            }
        }
    }
#pragma unroll
    for (int l = 0; l < K; ++l) if (l < count) out[l] = idx[l];
    return count;
}

/* k-NN search of the camera wavefront as its own pass: light on registers (full occupancy hides the latency of the scattered
 * cell reads), streaming through 16-byte positions; the shading kernel picks the lists up.  Searches are issued for every hit on
 * a non-specular, non-emitting material below the length cap -- a superset of the vertices that merge (vcm.cpp:572-611). */
template <int K>   // K > 0: candidate list of exactly K entries in registers; 0: any k, list in local memory
__global__ void __launch_bounds__(256) vcm_merge_query_kernel(const DeviceShading S, const VcmParams* Pd, const float4* rays, const int4* hits, const VCMState* states,
                                                              const int* sorted, const int* hit_count_ptr, const int* queue_count_ptr, const GridParams* gp, const int* cells,
                                                              const float4* photon_pos, int* found_out, int stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (sorted ? *hit_count_ptr : *queue_count_ptr)) return;
    const int idx = sorted ? sorted[i] : i;
    const int4 h = hits[idx];
    if (h.x < 0) { found_out[idx] = 0; return; }
    const int mat = __ldg(S.indices + h.x).w;   // hits carry global triangle ids (= mesh_tri_base[mesh] + local id, as calc_intersection resolves them)
    const float4 mh = __ldg(S.materials + 5 * mat);
    const int kind = __float_as_int(mh.x);
    const int path_length = (int)(states[idx].len_finite & 0x7fffffffu);
    int count = 0;
    const int k = Pd->num_knn;
    int found[IMBA_MAX_KNN]; float dist[IMBA_MAX_KNN];
    if (__float_as_int(mh.w) == 0 && kind != MAT_MIRROR && kind != MAT_GLASS && path_length < Pd->max_path_len) {
        const float4 ro = rays[2 * idx], rd = rays[2 * idx + 1];
        const V3 pos = v3(ro) + __int_as_float(h.z) * v3(rd);      // calc_intersection: pos = org + t * dir
        const GridParams g = *gp;
        count = K > 0 ? grid_query_fixed<(K > 0 ? K : 1)>(g, cells, g.table, photon_pos, pos, found)
                      : grid_query(g, cells, g.table, photon_pos, pos, k, found, dist);
    }
    found_out[idx] = count;
    for (int j = 0; j < count; ++j) found_out[(size_t)(j + 1) * stride + idx] = found[j];
}

template <int ALGO>
__global__ void __launch_bounds__(128, IMBA_VCM_CAM_MIN_BLOCKS) vcm_camera_shade_kernel(const DeviceShading S, const VcmParams* Pd, float4* rays, const int4* hits, VCMState* states, const int* sorted,
                                                               const int* hit_count_ptr, const int* queue_count_ptr, float4* shadow_rays, ShadowState* shadow_states,
                                                               const DevLightVertex* cache, const GridParams* gp, const int* cells, unsigned table,
                                                               const DevPhoton* photons, const int* merge_found, int merge_stride, float* fb, unsigned long long* ray_stats,
                                                               float4* out_rays, VCMState* out_states, int* out_count) {
    const VcmParams P = *Pd;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && ray_stats) atomicAdd(ray_stats, (unsigned long long)*queue_count_ptr);
    if (i >= (sorted ? *hit_count_ptr : *queue_count_ptr)) return;
    const int idx = sorted ? sorted[i] : i;
    const int slots = P.num_connections + 1;
    const int4 h = hits[idx];
    if (!sorted) { // queue order: this kernel sees every slot, also the misses, and resets the shadow slots itself
        for (int s = 0; s < slots; ++s) shadow_states[(size_t)idx * slots + s].pixel_id = -1;
        if (h.x < 0) { states[idx].pixel_id = -1; return; }
    }
    const float4 ro = rays[2 * idx], rd = rays[2 * idx + 1];
    VCMState st = states[idx];
    const float t = __int_as_float(h.z);
    const Isect is = calc_intersection(S, ro, rd, h.x, h.y, t, __int_as_float(h.w));
    const float cos_theta_o = fabsf(dot(is.out_dir, is.normal));
    const Bsdf bsdf = make_bsdf(S, is, false);
    st.dVCM *= sqr(t) / cos_theta_o; st.dVC *= 1.0f / cos_theta_o; st.dVM *= 1.0f / cos_theta_o;
    if (cos_theta_o == 0.0f) { states[idx].pixel_id = -1; return; }
    const V3 thr = v3(st.tr, st.tg, st.tb);
    const int path_length = (int)(st.len_finite & 0x7fffffffu);
    const float4 mh = __ldg(S.materials + 5 * is.mat);
    if (__float_as_int(mh.w) != 0) { // emitter hit (vcm.cpp:391-409)
        const float4 e = __ldg(S.materials + 5 * is.mat + 4);
        const float pdf_lightpick = 1.0f / S.light_count;
        float pdf_direct_a, pdf_emit_w;
        const V3 radiance = emitter_radiance(v3(e), e.w, is.out_dir, is.geom_normal, pdf_direct_a, pdf_emit_w);
        const float pdf_di = pdf_direct_a * pdf_lightpick, pdf_e = pdf_emit_w * pdf_lightpick;
        const float mis_weight_camera = pdf_di * st.dVCM + pdf_e * st.dVC;
        const float mis_weight = (ALGO == IMBA_ALGO_PPM || path_length == 1) ? 1.0f : (1.0f / (mis_weight_camera + 1.0f));
        add_contribution(fb, st.pixel_id, thr * radiance * mis_weight);
        states[idx].pixel_id = -1;
        return;
    }
    if (!(path_length < P.max_path_len)) { states[idx].pixel_id = -1; return; }
    Rng rng; rng.s = st.rng;
    if (ALGO != IMBA_ALGO_PPM) { // direct_illum (vcm.cpp:442-486)
        const int li = rng.range(0, S.light_count);
        const float pdf_lightpick_inv = (float)S.light_count;
        const DirectSample s = light_sample_direct(S, li, is.pos, rng);
        const float cos_theta_i = fabsf(dot(is.normal, s.dir));
        const V3 bsdf_value = bsdf.eval(is.out_dir, s.dir);
        const float pdf_dir_w = bsdf.pdf(is.out_dir, s.dir), pdf_rev_w = bsdf.pdf(s.dir, is.out_dir);
        if (!(pdf_dir_w == 0.0f || pdf_rev_w == 0.0f)) {
            const float pdf_forward = s.delta ? 0.0f : pdf_dir_w;
            const float mis_weight_light = pdf_forward * pdf_lightpick_inv / s.pdf_direct_w;
            const float mis_weight_camera = (s.pdf_emit_w * cos_theta_i / (s.pdf_direct_w * s.cos_out)) * (P.mis_eta_vm + st.dVCM + st.dVC * pdf_rev_w);
            const float mis_weight = ALGO == IMBA_ALGO_VCM_PT ? 1.0f : (1.0f / (mis_weight_camera + 1.0f + mis_weight_light));
            const V3 c = thr * mis_weight * bsdf_value * cos_theta_i * s.radiance * pdf_lightpick_inv;
            const float off = 1e-3f * (s.distance == FLT_MAX ? 1.0f : s.distance);
            const size_t slot = (size_t)idx * slots;
            shadow_rays[2 * slot] = make_float4(is.pos.x, is.pos.y, is.pos.z, off);
            shadow_rays[2 * slot + 1] = make_float4(s.dir.x, s.dir.y, s.dir.z, s.distance - off);
            ShadowState ss; ss.pixel_id = st.pixel_id; ss.r = c.x; ss.g = c.y; ss.b = c.z;
            shadow_states[slot] = ss;
            if (ray_stats) count_rays(ray_stats + 1);
        }
    }
    const GridParams g = *gp; // g.count = number of cached light vertices of this iteration
    if (ALGO != IMBA_ALGO_VCM_PT && ALGO != IMBA_ALGO_PPM && !bsdf.is_specular() && g.count > 0) { // connect (vcm.cpp:489-569)
        const float vc_weight = g.count / ((float)P.light_path_count * (float)P.num_connections);
        for (int c = 0; c < P.num_connections; ++c) {
            const DevLightVertex lv = cache[rng.range(0, g.count)];
            const int light_len = __float_as_int(lv.pos_len.w);
            if (light_len + path_length > P.max_path_len) continue;
            Isect lis;
            const Bsdf light_bsdf = vertex_bsdf(S, lv, lis);
            V3 connect_dir = lis.pos - is.pos;
            const float dist_sq = dot(connect_dir, connect_dir), dist = sqrtf(dist_sq);
            connect_dir *= 1.0f / dist;
            if (dist < P.base_radius) continue;
            const V3 bsdf_value_cam = bsdf.eval(is.out_dir, connect_dir);
            const float pdf_dir_cam_w = bsdf.pdf(is.out_dir, connect_dir), pdf_rev_cam_w = bsdf.pdf(connect_dir, is.out_dir);
            const V3 bsdf_value_light = light_bsdf.eval(lis.out_dir, -connect_dir);
            const float pdf_dir_light_w = light_bsdf.pdf(lis.out_dir, -connect_dir), pdf_rev_light_w = light_bsdf.pdf(-connect_dir, lis.out_dir);
            if (pdf_dir_cam_w == 0.0f || pdf_dir_light_w == 0.0f || pdf_rev_cam_w == 0.0f || pdf_rev_light_w == 0.0f) continue;
            const float cos_theta_cam = fabsf(dot(is.normal, connect_dir));
            const float cos_theta_light = fabsf(shading_normal_adjoint(lis.normal, lis.geom_normal, lis.out_dir, -connect_dir));
            const float geom_term = cos_theta_cam * cos_theta_light / dist_sq;
            if (geom_term <= 0.0f) continue;
            const float pdf_cam_a = pdf_dir_cam_w * cos_theta_light / dist_sq, pdf_light_a = pdf_dir_light_w * cos_theta_cam / dist_sq;
            const float mis_weight_light = pdf_cam_a * (P.mis_eta_vm + lv.gn_dvcm.w + lv.n_dvc.w * pdf_rev_light_w);
            const float mis_weight_camera = pdf_light_a * (P.mis_eta_vm + st.dVCM + st.dVC * pdf_rev_cam_w);
            const float mis_weight = 1.0f / (mis_weight_camera + 1.0f + mis_weight_light);
            const V3 cc = thr * vc_weight * mis_weight * geom_term * bsdf_value_cam * bsdf_value_light * v3(lv.thr_v);
            const float off = 1e-3f * dist;
            const size_t slot = (size_t)idx * slots + 1 + c;
            shadow_rays[2 * slot] = make_float4(is.pos.x, is.pos.y, is.pos.z, off);
            shadow_rays[2 * slot + 1] = make_float4(connect_dir.x, connect_dir.y, connect_dir.z, dist - off);
            ShadowState ss; ss.pixel_id = st.pixel_id; ss.r = cc.x; ss.g = cc.y; ss.b = cc.z;
            shadow_states[slot] = ss;
            if (ray_stats) count_rays(ray_stats + 1);
        }
    }
    if (ALGO != IMBA_ALGO_BPT && ALGO != IMBA_ALGO_VCM_PT && !bsdf.is_specular()) { // vertex_merging (vcm.cpp:572-611)
        // the k nearest photons were found by vcm_merge_query_kernel (sorted by distance, nearest first)
        const int k = P.num_knn;
        const int count = merge_found[idx];
        float radius_sqr = P.pm_radius * P.pm_radius;
        if (count == k) { const V3 dk = is.pos - v3(photons[merge_found[(size_t)k * merge_stride + idx]].pos_dvcm); radius_sqr = dot(dk, dk); }
        V3 contrib = v3(0.0f);
        for (int j = 0; j < count; ++j) {
            const DevPhoton p = photons[merge_found[(size_t)(j + 1) * merge_stride + idx]];
            const V3 in_dir = v3(p.dir_dvm);
            const V3 bsdf_value = bsdf.eval(is.out_dir, in_dir);
            const float pdf_dir_w = bsdf.pdf(is.out_dir, in_dir), pdf_rev_w = bsdf.pdf(in_dir, is.out_dir);
            if (pdf_dir_w == 0.0f || pdf_rev_w == 0.0f || is_black(bsdf_value)) continue;
            const float mis_weight_light = p.pos_dvcm.w * P.mis_eta_vc + p.dir_dvm.w * pdf_dir_w;
            const float mis_weight_camera = st.dVCM * P.mis_eta_vc + st.dVM * pdf_rev_w;
            const float mis_weight = ALGO == IMBA_ALGO_PPM ? 1.0f : (1.0f / (mis_weight_light + 1.0f + mis_weight_camera));
            const V3 dd = v3(p.pos_dvcm) - is.pos;
            const float kernel = 1.0f - dot(dd, dd) / radius_sqr;
            contrib = contrib + mis_weight * bsdf_value * kernel * v3(p.thr);
        }
        contrib *= 2.0f / (IMBA_PI * radius_sqr * P.light_path_count);
        add_contribution(fb, st.pixel_id, thr * contrib);
    }
    float4 no, nd;
    if (vcm_bounce<ALGO>(P, st, rng, is, bsdf, false, t * 1e-4f, no, nd)) {
        st.rng = rng.s;
        const int o = out_count ? queue_push_slot(out_count) : idx;
        (out_count ? out_states : states)[o] = st;
        float4* dst = (out_count ? out_rays : rays) + 2 * (size_t)o;
        dst[0] = no; dst[1] = nd;
    } else states[idx].pixel_id = -1;
}

/* Camera paths that leave the scene pick up the environment map (vcm.cpp:345-366); not for the VCM-style PT. */
template <int ALGO>
__global__ void vcm_camera_miss_kernel(const DeviceShading S, const float4* rays, const int4* hits, const VCMState* states, const int* sorted, const int* hit_count_ptr,
                                       const int* queue_count_ptr, float* fb) {
    const int i = (sorted ? *hit_count_ptr : 0) + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *queue_count_ptr) return;
    const int idx = sorted ? sorted[i] : i;
    if (!sorted && hits[idx].x >= 0) return;
    const VCMState st = states[idx];
    const V3 out_dir = normalize(v3(rays[2 * idx + 1]));
    float pdf_direct_w, pdf_emit_w;
    const V3 li = env_radiance(S, out_dir, pdf_direct_w, pdf_emit_w);
    const float pdf_lightpick = 1.0f / S.light_count;
    const float pdf_di = pdf_direct_w * pdf_lightpick, pdf_e = pdf_emit_w * pdf_lightpick;
    const float mis_weight_camera = pdf_di * st.dVCM + pdf_e * st.dVC;
    const float mis_weight = ALGO == IMBA_ALGO_PPM ? 1.0f : (1.0f / (mis_weight_camera + 1.0f));
    add_contribution(fb, st.pixel_id, v3(st.tr, st.tg, st.tb) * li * mis_weight);
}

template <int ALGO>
cudaError_t light_shade(const VcmLaunch& a, cudaStream_t st) {
    vcm_light_shade_kernel<ALGO><<<(a.capacity + 127) / 128, 128, 0, st>>>(*a.S, a.P, a.cam, a.rays, a.hits, a.states, a.sorted, a.hit_count_ptr, a.queue_count_ptr,
                                                                          a.shadow_rays, a.shadow_states, a.cache, a.cache_capacity, a.cache_last, a.ray_stats, a.out_rays, a.out_states, a.out_count);
    return cudaGetLastError();
}
template <int ALGO>
cudaError_t camera_shade(const VcmLaunch& a, cudaStream_t st) {
    if (a.S->env_tex >= 0 && ALGO != IMBA_ALGO_VCM_PT)
        vcm_camera_miss_kernel<ALGO><<<(a.capacity + 255) / 256, 256, 0, st>>>(*a.S, a.rays, a.hits, a.states, a.sorted, a.hit_count_ptr, a.queue_count_ptr, a.fb);
    if (ALGO == IMBA_ALGO_VCM || ALGO == IMBA_ALGO_PPM) {
        const int blocks = (a.capacity + 255) / 256;
#define MERGE_QUERY(KK) vcm_merge_query_kernel<KK><<<blocks, 256, 0, st>>>(*a.S, a.P, a.rays, a.hits, a.states, a.sorted, a.hit_count_ptr, a.queue_count_ptr, a.grid, a.cells, \
                                                                          a.photon_pos, a.merge_found, a.merge_stride)
        switch (getenv("IMBA_KNN_GENERIC") ? 0 : a.knn) { // register-resident lists for the usual sizes (10 = the reference's default, cmd_line.h)
            case 10: MERGE_QUERY(10); break;
            case 4:  MERGE_QUERY(4); break;
            case 8:  MERGE_QUERY(8); break;
            case 16: MERGE_QUERY(16); break;
            default: MERGE_QUERY(0);
        }
#undef MERGE_QUERY
    }
    vcm_camera_shade_kernel<ALGO><<<(a.capacity + 127) / 128, 128, 0, st>>>(*a.S, a.P, a.rays, a.hits, a.states, a.sorted, a.hit_count_ptr, a.queue_count_ptr, a.shadow_rays,
                                                                           a.shadow_states, a.cache, a.grid, a.cells, a.table, a.photons, a.merge_found, a.merge_stride, a.fb,
                                                                           a.ray_stats, a.out_rays, a.out_states, a.out_count);
    return cudaGetLastError();
}

/* ---- LightVertices::compute_cache_size (light_vertices.cpp:15-129): the reference's sizing probe, restated ----------------------
 * NOT the light pass above: 10 000 light paths with a plain throughput (radiance / pdf_lightpick, no MIS quantities), Russian
 * roulette on that throughput, BSDF_ALL sampling with adjoint BSDFs, one count per hit on a non-specular material; the host loop
 * stops as soon as 256 or fewer paths are left.  Its average is what sizes the vertex cache, so it is this estimator -- not the
 * mean vertex count of the real light pass, which is up to 10 % higher -- that has to be reproduced. */
__global__ void vcm_probe_raygen_kernel(const DeviceShading S, uint32_t seed_base, int count, float4* rays, PTState* states, int* count_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count_out = count;
    if (i >= count) return;
    // state_out.rng = RNG(rd()): a fresh 32-bit value per path as the 64-bit generator state (light_vertices.cpp:34-35)
    uint32_t z = seed_base + 0x9E3779B9u * (uint32_t)(i + 1);
    z = (z ^ (z >> 16)) * 0x85EBCA6Bu; z = (z ^ (z >> 13)) * 0xC2B2AE35u; z ^= z >> 16;
    Rng rng; rng.s = (uint64_t)z;
    const int light = rng.range(0, S.light_count);
    const float pdf_lightpick = 1.0f / S.light_count;
    const EmitSample s = light_sample_emit(S, light, rng);
    rays[2 * i] = make_float4(s.pos.x, s.pos.y, s.pos.z, 1e-3f);
    rays[2 * i + 1] = make_float4(s.dir.x, s.dir.y, s.dir.z, FLT_MAX);
    const V3 thr = s.radiance / pdf_lightpick;
    PTState st; st.pixel_id = i; st.sample_id = 0; st.rng = rng.s; st.tr = thr.x; st.tg = thr.y; st.tb = thr.z; st.bounces_spec = 0; st.last_pdf = 0.0f; st.pad = 0;
    states[i] = st;
}

__global__ void vcm_probe_shade_kernel(const DeviceShading S, const float4* rays, const int4* hits, const PTState* states, const int* count_ptr,
                                       float4* out_rays, PTState* out_states, int* out_count, int* vertex_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count_ptr) return;
    const int4 h = hits[i];
    if (h.x < 0) return;                                   // `tri_id < 0 || !russian_roulette(...)`: no random number for a miss
    PTState st = states[i];
    Rng rng; rng.s = st.rng;
    V3 thr = v3(st.tr, st.tg, st.tb);
    const float rr_pdf = fminf(1.0f, luminance(thr) * 2.0f);
    if (!(rng.uniform() < rr_pdf)) return;
    const float4 ro = rays[2 * i], rd = rays[2 * i + 1];
    const float t = __int_as_float(h.z);
    const Isect is = calc_intersection(S, ro, rd, h.x, h.y, t, __int_as_float(h.w));
    const Bsdf bsdf = make_bsdf(S, is, true);
    if (!bsdf.is_specular()) atomicAdd(vertex_count, 1);   // "we would store a vertex at this position"
    float pdf_dir_w; V3 dir; int flags;
    const V3 value = bsdf.sample(is.out_dir, dir, rng, flags, pdf_dir_w);
    if (flags == 0 || pdf_dir_w == 0.0f || is_black(value)) return;
    const float cos_theta_i = fabsf(dot(is.normal, dir));
    thr *= value * cos_theta_i / (rr_pdf * pdf_dir_w);
    st.tr = thr.x; st.tg = thr.y; st.tb = thr.z; st.rng = rng.s;
    const int o = queue_push_slot(out_count);
    out_states[o] = st;
    out_rays[2 * (size_t)o] = make_float4(is.pos.x, is.pos.y, is.pos.z, t * 1e-3f);
    out_rays[2 * (size_t)o + 1] = make_float4(dir.x, dir.y, dir.z, FLT_MAX);
}

} // namespace

cudaError_t imba_vcm_probe_raygen(const DeviceShading& S, uint32_t seed_base, int count, float4* rays, PTState* states, int* count_out, cudaStream_t st) {
    vcm_probe_raygen_kernel<<<(std::max(count, 1) + 255) / 256, 256, 0, st>>>(S, seed_base, count, rays, states, count_out);
    return cudaGetLastError();
}
cudaError_t imba_vcm_probe_shade(const DeviceShading& S, const float4* rays, const int4* hits, const PTState* states, const int* count_ptr, int capacity,
                                 float4* out_rays, PTState* out_states, int* out_count, int* vertex_count, cudaStream_t st) {
    if (capacity <= 0) return cudaSuccess;
    vcm_probe_shade_kernel<<<(capacity + 127) / 128, 128, 0, st>>>(S, rays, hits, states, count_ptr, out_rays, out_states, out_count, vertex_count);
    return cudaGetLastError();
}

cudaError_t imba_vcm_light_raygen(const DeviceShading& S, const VcmParams* P, const PtSeeds& seeds, int per_iteration, int total, float4* rays, VCMState* states, int* count_out, cudaStream_t st) {
    vcm_light_raygen_kernel<<<(std::max(total, 1) + 255) / 256, 256, 0, st>>>(S, P, seeds, per_iteration, total, rays, states, count_out);
    return cudaGetLastError();
}
cudaError_t imba_vcm_camera_raygen(const DevCamera& cam, const VcmParams* P, int spp, uint32_t seed_base, int count, float4* rays, VCMState* states, int* count_out, cudaStream_t st) {
    vcm_camera_raygen_kernel<<<(std::max(count, 1) + 255) / 256, 256, 0, st>>>(cam, P, spp, seed_base, count, rays, states, count_out);
    return cudaGetLastError();
}
cudaError_t imba_vcm_light_shade(int algo, const VcmLaunch& a, cudaStream_t st) {
    if (a.capacity <= 0) return cudaSuccess;
    switch (algo) {
        case IMBA_ALGO_BPT: return light_shade<IMBA_ALGO_BPT>(a, st);
        case IMBA_ALGO_PPM: return light_shade<IMBA_ALGO_PPM>(a, st);
        case IMBA_ALGO_LT:  return light_shade<IMBA_ALGO_LT>(a, st);
        default:            return light_shade<IMBA_ALGO_VCM>(a, st);
    }
}
cudaError_t imba_vcm_camera_shade(int algo, const VcmLaunch& a, cudaStream_t st) {
    if (a.capacity <= 0) return cudaSuccess;
    switch (algo) {
        case IMBA_ALGO_BPT:    return camera_shade<IMBA_ALGO_BPT>(a, st);
        case IMBA_ALGO_PPM:    return camera_shade<IMBA_ALGO_PPM>(a, st);
        case IMBA_ALGO_VCM_PT: return camera_shade<IMBA_ALGO_VCM_PT>(a, st);
        default:               return camera_shade<IMBA_ALGO_VCM>(a, st);
    }
}

cudaError_t imba_vcm_light_tail(int algo, bool masks, const VcmLaunch& a, const TraceArgs& T, int capacity, cudaStream_t st) {
    if (capacity <= 0) return cudaSuccess;
    // stage 1: one thread per path for up to `budget` bounces (most paths end here); stage 2: the survivors, one warp per path
    const int budget = getenv("IMBA_TAIL_BUDGET") ? atoi(getenv("IMBA_TAIL_BUDGET")) : 8;   // sweep 8 / 16 / 24 / 48 on stilllife-proxy LT: 15.2 / 15.5 / 15.8 / 16.6 ms
    const bool two_stage = budget > 0 && a.out_rays && a.out_states && a.out_count;
    const int grid = (capacity + 63) / 64, unlimited = 1 << 20;
    if (two_stage) { const cudaError_t e = cudaMemsetAsync(a.out_count, 0, sizeof(int), st); if (e != cudaSuccess) return e; }
    const int grid2 = std::min((capacity + 3) / 4, 148 * 8);
#define TAIL(A, M) do { \
        vcm_light_tail_kernel<A, M><<<grid, 64, 0, st>>>(*a.S, a.P, a.cam, T, a.rays, a.states, a.queue_count_ptr, a.cache, a.cache_capacity, a.cache_last, a.fb, a.ray_stats, \
                                                         two_stage ? budget : unlimited, two_stage ? a.out_rays : nullptr, two_stage ? a.out_states : nullptr, two_stage ? a.out_count : nullptr); \
        if (two_stage) vcm_light_tail_warp_kernel<A, M><<<grid2, 128, 0, st>>>(*a.S, a.P, a.cam, T, a.out_rays, a.out_states, a.out_count, a.cache, a.cache_capacity, a.cache_last, \
                                                                             a.fb, a.ray_stats, unlimited); } while (0)
    switch (algo) {
        case IMBA_ALGO_BPT: if (masks) TAIL(IMBA_ALGO_BPT, true); else TAIL(IMBA_ALGO_BPT, false); break;
        case IMBA_ALGO_PPM: if (masks) TAIL(IMBA_ALGO_PPM, true); else TAIL(IMBA_ALGO_PPM, false); break;
        case IMBA_ALGO_LT:  if (masks) TAIL(IMBA_ALGO_LT, true);  else TAIL(IMBA_ALGO_LT, false); break;
        default:            if (masks) TAIL(IMBA_ALGO_VCM, true); else TAIL(IMBA_ALGO_VCM, false); break;
    }
#undef TAIL
    return cudaGetLastError();
}

/* HashGrid::build on the vertex cache; `cells` has `table` + 1 ints, block_sums (table / 2048 + 1) ints */
cudaError_t imba_grid_build(const DevLightVertex* cache, const int* cache_last, int cache_capacity, float radius, GridParams* grid, int* bbox, int* cells, unsigned table,
                            int* block_sums, DevPhoton* photons, float4* photon_pos, bool build_cells, unsigned* table_state, cudaStream_t st) {
    static const int init_bbox[6] = { 0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)(0xff7fffffu ^ 0x7fffffffu), (int)(0xff7fffffu ^ 0x7fffffffu), (int)(0xff7fffffu ^ 0x7fffffffu) };
    cudaError_t e = cudaMemcpyAsync(bbox, init_bbox, sizeof(init_bbox), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    const int blocks = (std::max(cache_capacity, 1) + 255) / 256;
    if (build_cells) grid_bbox_kernel<<<std::min(blocks, 148 * 8), 256, 0, st>>>(cache, cache_last, cache_capacity, bbox);
    grid_prepare_kernel<<<1, 1, 0, st>>>(bbox, cache_last, cache_capacity, grid, radius, table_state);
    if (build_cells) {
        if ((e = cudaMemsetAsync(cells, 0, sizeof(int) * (size_t)(table + 1), st)) != cudaSuccess) return e;
        grid_count_kernel<<<blocks, 256, 0, st>>>(cache, grid, cells, table);
        const int nchunks = ((int)table + SCAN_CHUNK - 1) / SCAN_CHUNK;
        scan_local_kernel<<<nchunks, 256, 0, st>>>(cells, (int)table, block_sums);
        scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nchunks);
        scan_add_kernel<<<((int)table + 255) / 256, 256, 0, st>>>(cells, (int)table, block_sums);
        grid_scatter_kernel<<<blocks, 256, 0, st>>>(cache, grid, cells, table, photons, photon_pos);
    }
    return cudaGetLastError();
}
