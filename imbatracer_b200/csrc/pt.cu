/* pt.cu -- path-tracer wavefront stages on the device.
 *   pt_raygen_kernel   PixelRayGen::fill_queue + PathTracer::render's sample function (ray_gen.h:36-80, pt.cpp:168-177,
 *                      camera.h:57-65); the paths of several progressive iterations may share one pass (iterations_per_pass)
 *   pt_shade_kernel    PathTracer::process_primary_rays / compute_direct_illum / bounce (pt.cpp:82-160, 17-44, 46-80),
 *                      in queue order (the integrator's default) or on material-sorted hit indices (the reference's order:
 *                      parity harness, IMBA_PT_SORT=1); the shadow ray of slot i is written to slot i of the shadow queue (no
 *                      atomic push), survivors are pushed into the next queue (or overwrite their slot in the harness path)
 *   pt_miss_kernel     environment map for the rays that left the scene (pt.cpp:95-116)
 *   pixel-size probe   Integrator::estimate_pixel_size (integrator.cpp:14-71)
 */
#include "integrator.cuh"

namespace {

__device__ __forceinline__ void add_contribution(float* fb, int pixel, V3 c) {
    float* p = fb + 3 * (size_t)pixel;
    atomicAdd(p, c.x); atomicAdd(p + 1, c.y); atomicAdd(p + 2, c.z);
}

__global__ void pt_raygen_kernel(DevCamera cam, int spp, PtSeeds seeds, int per_iteration, int first, int count, float4* rays, PTState* states, int* count_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) *count_out = count;
    if (k >= count) return;
    // several progressive iterations may share one pass (imba_render_settings::iterations_per_pass): slot k belongs to
    // iteration k / per_iteration of the group and is seeded exactly as in a pass of its own
    // interleaved: the paths of one pixel in the `group` iterations sit next to each other in the queue, like the samples of one
    // pixel with --spp > 1 (more coherent warps than one iteration after the other: +4 % at group 4)
    const int group = count / per_iteration;
    const int sub = k % group;
    const int i = first + k / group;
    const uint32_t seed_base = seeds.base[sub];
    const int pixel_idx = i / spp, sample_idx = i % spp;
    const int y = pixel_idx / cam.width, x = pixel_idx % cam.width;
    Rng rng; rng.s = seed_primary(seed_base, i, pixel_idx);
    const float sx = (float)x + rng.uniform();
    const float sy = (float)y + rng.uniform();
    const V3 d = cam.ray_dir(sx, sy);
    rays[2 * k] = make_float4(cam.px, cam.py, cam.pz, 0.0f);
    rays[2 * k + 1] = make_float4(d.x, d.y, d.z, FLT_MAX);
    PTState st;
    st.pixel_id = pixel_idx; st.sample_id = sample_idx; st.rng = rng.s;
    st.tr = st.tg = st.tb = 1.0f; st.bounces_spec = 0; st.last_pdf = 0.0f; st.pad = 0;
    states[k] = st;
}

#ifndef IMBA_PT_MIN_BLOCKS
#define IMBA_PT_MIN_BLOCKS 8   /* tuning switch: resident blocks per SM the register allocation aims at: 95 / 80 / 64 registers at 1 / 6 / 8 give 4.18 / 4.14 / 4.12 ms per PT iteration */
#endif
__global__ void __launch_bounds__(128, IMBA_PT_MIN_BLOCKS) pt_shade_kernel(const DeviceShading S, float4* rays, const int4* hits, PTState* states, const int* sorted,
                                                       const int* hit_count_ptr, const int* queue_count_ptr, float4* shadow_rays, ShadowState* shadow_states,
                                                       float* fb, int max_path_len, unsigned long long* ray_stats, float4* out_rays, PTState* out_states,
                                                       int* out_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && ray_stats) atomicAdd(ray_stats, (unsigned long long)*queue_count_ptr); // rays traced in this wavefront (tile_scheduler.h:117-127)
    // sorted != null: the reference's order (compact_hits + sort_by_material, pt.cpp:85-93), hits only.  sorted == null: queue
    // order, every slot -- the integrator's default: the per-path results are the same, and on the GPU the sort costs more
    // (14 % of an iteration) than the divergence between four material kinds it removes.
    if (i >= (sorted ? *hit_count_ptr : *queue_count_ptr)) return;
    const int idx = sorted ? sorted[i] : i;
    const int4 h = hits[idx];
    if (!sorted) {
        shadow_states[idx].pixel_id = -1;             // what the sort pass does for every slot: no shadow ray yet
        if (h.x < 0) { states[idx].pixel_id = -1; return; }
    }
    const float4 ro = rays[2 * idx], rd = rays[2 * idx + 1];
    PTState st = states[idx];
    const float t = __int_as_float(h.z);
    const Isect is = calc_intersection(S, ro, rd, h.x, h.y, t, __int_as_float(h.w));
    const float offset = 1e-3f * t;
    V3 thr = v3(st.tr, st.tg, st.tb);
    const int bounces = (int)(st.bounces_spec & 0x7fffffffu);
    const bool last_specular = (st.bounces_spec >> 31) != 0;

    const float4 mh = __ldg(S.materials + 5 * is.mat);
    if (__float_as_int(mh.w) != 0) { // emitter hit (pt.cpp:133-151)
        const float4 e = __ldg(S.materials + 5 * is.mat + 4);
        float pdf_direct_a, pdf_emit_w;
        const V3 li = emitter_radiance(v3(e), e.w, is.out_dir, is.geom_normal, pdf_direct_a, pdf_emit_w);
        float pdf_di = pdf_direct_a / S.light_count;
        const float d_sqr = t * t;
        const float cos_light = dot(is.normal, is.out_dir);
        pdf_di *= d_sqr / cos_light;
        const float mis = (bounces == 0 || last_specular) ? 1.0f : st.last_pdf / (st.last_pdf + pdf_di);
        add_contribution(fb, st.pixel_id, thr * li * mis);
        states[idx].pixel_id = -1;
        return;
    }

    const Bsdf bsdf = make_bsdf(S, is, false);
    Rng rng; rng.s = st.rng;

    // compute_direct_illum (pt.cpp:17-44)
    {
        const int li = rng.range(0, S.light_count);
        const float pdf_lightpick = 1.0f / S.light_count;
        const DirectSample s = light_sample_direct(S, li, is.pos, rng);
        const V3 bsdf_value = bsdf.eval(is.out_dir, s.dir);
        const float pdf_hit = bsdf.pdf(is.out_dir, s.dir);
        const float pdf_di = pdf_lightpick * s.pdf_direct_w;
        const float mis = s.delta ? 1.0f : pdf_di / (pdf_di + pdf_hit);
        if (!(pdf_hit == 0.0f || pdf_di == 0.0f)) {
            const V3 c = thr * bsdf_value * fabsf(dot(is.normal, s.dir)) * s.radiance * mis / pdf_lightpick;
            const float off = 1e-3f * (s.distance == FLT_MAX ? 1.0f : s.distance);
            shadow_rays[2 * idx] = make_float4(is.pos.x, is.pos.y, is.pos.z, off);
            shadow_rays[2 * idx + 1] = make_float4(s.dir.x, s.dir.y, s.dir.z, s.distance - off);
            ShadowState ss; ss.pixel_id = st.pixel_id; ss.r = c.x; ss.g = c.y; ss.b = c.z;
            shadow_states[idx] = ss;
            if (ray_stats) count_rays(ray_stats + 1);
        }
    }

    // bounce (pt.cpp:46-80)
    bool alive = bounces + 1 < max_path_len;
    float rr_pdf = 1.0f;
    if (alive) {
        const float kill_prob = luminance(thr) * 2.0f;
        rr_pdf = fminf(1.0f, kill_prob);
        alive = rng.uniform() < rr_pdf;
    }
    if (alive) {
        float pdf; V3 dir; int flags;
        const V3 value = bsdf.sample(is.out_dir, dir, rng, flags, pdf);
        if (pdf == 0.0f || flags == BSDF_NONE || is_black(value)) alive = false;
        else {
            const float cos_term = fabsf(dot(is.normal, dir));
            thr *= value * cos_term / (pdf * rr_pdf);
            st.tr = thr.x; st.tg = thr.y; st.tb = thr.z;
            st.bounces_spec = (uint32_t)(bounces + 1) | ((flags & BSDF_SPECULAR) ? 0x80000000u : 0u);
            st.last_pdf = pdf; st.rng = rng.s;
            // continuation: in place like the reference (then compact_rays), or straight into the next wavefront's queue
            const int o = out_count ? queue_push_slot(out_count) : idx;
            (out_count ? out_states : states)[o] = st;
            float4* dst = (out_count ? out_rays : rays) + 2 * (size_t)o;
            dst[0] = make_float4(is.pos.x, is.pos.y, is.pos.z, offset);
            dst[1] = make_float4(dir.x, dir.y, dir.z, FLT_MAX);
        }
    }
    if (!alive) states[idx].pixel_id = -1;
}

/* Rays that left the scene pick up the environment map (pt.cpp:95-116); misses are the last bin of the material sort. */
__global__ void pt_miss_kernel(const DeviceShading S, const float4* rays, const int4* hits, const PTState* states, const int* sorted, const int* hit_count_ptr,
                               const int* queue_count_ptr, float* fb) {
    const int i = (sorted ? *hit_count_ptr : 0) + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *queue_count_ptr) return;
    const int idx = sorted ? sorted[i] : i;
    if (!sorted && hits[idx].x >= 0) return;          // queue order: pick the misses out
    const PTState st = states[idx];
    const V3 out_dir = normalize(v3(rays[2 * idx + 1]));
    float pdf_direct_w, pdf_emit_w;
    const V3 li = env_radiance(S, out_dir, pdf_direct_w, pdf_emit_w);
    const float pdf_lightpick = 1.0f / S.light_count;
    const float pdf_di = pdf_direct_w * pdf_lightpick;
    const float pdf_hit = st.last_pdf;
    const bool first_or_specular = (st.bounces_spec & 0x7fffffffu) == 0 || (st.bounces_spec >> 31) != 0;
    const float mis = first_or_specular ? 1.0f : (pdf_hit / (pdf_hit + pdf_di));
    add_contribution(fb, st.pixel_id, v3(st.tr, st.tg, st.tb) * li * mis);
}

/* Integrator::estimate_pixel_size (integrator.cpp:14-71): corner rays of every 8th pixel */
__global__ void probe_raygen_kernel(DevCamera cam, float4* rays, int count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int quad = k / 4, corner = k % 4;
    const int ny = (cam.height + 7) / 8;
    const int x = (quad / ny) * 8 + (corner & 1), y = (quad % ny) * 8 + (corner >> 1);
    const V3 d = cam.ray_dir((float)x, (float)y);
    rays[2 * k] = make_float4(cam.px, cam.py, cam.pz, 0.0f);
    rays[2 * k + 1] = make_float4(d.x, d.y, d.z, FLT_MAX);
}

__global__ void probe_reduce_kernel(const float4* rays, const int4* hits, int quads, float* total, int* count) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= quads) return;
    V3 p[4];
    for (int c = 0; c < 4; ++c) {
        const int4 h = hits[4 * q + c];
        if (h.x < 0) return;
        p[c] = __int_as_float(h.z) * v3(rays[2 * (4 * q + c) + 1]) + v3(rays[2 * (4 * q + c)]);
    }
    const float s = length(p[0] - p[1]) + length(p[1] - p[3]) + length(p[3] - p[2]) + length(p[2] - p[0]);
    atomicAdd(total, s); atomicAdd(count, 4);
}

__global__ void clear_shadow_kernel(ShadowState* s, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s[i].pixel_id = -1;
}

} // namespace

cudaError_t imba_pt_raygen(const DevCamera& cam, int spp, const PtSeeds& seeds, int per_iteration, int first, int count, float4* rays, PTState* states, int* count_out, cudaStream_t st) {
    pt_raygen_kernel<<<(std::max(count, 1) + 255) / 256, 256, 0, st>>>(cam, spp, seeds, per_iteration, first, count, rays, states, count_out);
    return cudaGetLastError();
}

cudaError_t imba_pt_shade(const DeviceShading& S, float4* rays, const int4* hits, PTState* states, const int* sorted, const int* hit_count_ptr, const int* queue_count_ptr, int capacity,
                          float4* shadow_rays, ShadowState* shadow_states, float* fb, int max_path_len, unsigned long long* ray_stats, float4* out_rays, PTState* out_states,
                          int* out_count, cudaStream_t st) {
    if (capacity <= 0) return cudaSuccess;
    pt_shade_kernel<<<(capacity + 127) / 128, 128, 0, st>>>(S, rays, hits, states, sorted, hit_count_ptr, queue_count_ptr, shadow_rays, shadow_states, fb, max_path_len, ray_stats, out_rays, out_states, out_count);
    return cudaGetLastError();
}

cudaError_t imba_pt_miss(const DeviceShading& S, const float4* rays, const int4* hits, const PTState* states, const int* sorted, const int* hit_count_ptr,
                         const int* queue_count_ptr, int capacity, float* fb, cudaStream_t st) {
    if (capacity <= 0 || S.env_tex < 0) return cudaSuccess;
    pt_miss_kernel<<<(capacity + 255) / 256, 256, 0, st>>>(S, rays, hits, states, sorted, hit_count_ptr, queue_count_ptr, fb);
    return cudaGetLastError();
}

cudaError_t imba_probe_raygen(const DevCamera& cam, float4* rays, int count, cudaStream_t st) {
    probe_raygen_kernel<<<(count + 255) / 256, 256, 0, st>>>(cam, rays, count);
    return cudaGetLastError();
}

cudaError_t imba_probe_reduce(const float4* rays, const int4* hits, int quads, float* total, int* count, cudaStream_t st) {
    probe_reduce_kernel<<<(quads + 255) / 256, 256, 0, st>>>(rays, hits, quads, total, count);
    return cudaGetLastError();
}

cudaError_t imba_clear_shadow(ShadowState* s, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    clear_shadow_kernel<<<(n + 255) / 256, 256, 0, st>>>(s, n);
    return cudaGetLastError();
}
