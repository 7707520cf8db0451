/* shade.cuh -- device-side shading library: RNG, surface reconstruction, BSDFs, lights.
 *
 * Restates, for a flat tagged material / light table instead of virtual classes + a placement-new arena:
 *   RNG (MWC64X)                         render/random.h:15-47
 *   calculate_intersection               render/integrators/integrator.h:71-123
 *   Material::bump, get_bsdf             render/materials/materials.h:34-55, 76-182
 *   BSDF::{eval,sample,pdf}              render/materials/bsdf.h:118-238
 *   Lambertian, SpecularReflection, CookTorrance(Blinn), CombineBxDF   render/materials/brdfs.h, bsdf.h:66-90
 *   SpecularTransmission<adjoint>        render/materials/btdfs.h:6-65
 *   FresnelConductor / FresnelDielectric render/materials/fresnel.h:7-75
 *   TextureSampler::sample               render/texture_sampler.h:18-46
 *   lights                               render/light.h
 * The order of random-number draws follows the reference AS COMPILED BY g++ (function arguments are evaluated right to
 * left), so that a path continued from the same 8-byte RNG state takes the same decisions; this is checked against the
 * reference's own PathTracer on injected batches (tests/test_shading_gpu.py).
 */
#ifndef IMBA_SHADE_CUH
#define IMBA_SHADE_CUH
#include <cfloat>
#include <stdint.h>
#include "vecmath.cuh"

enum { BSDF_NONE = 0, BSDF_REFLECTION = 1, BSDF_TRANSMISSION = 2, BSDF_DIFFUSE = 4, BSDF_GLOSSY = 8, BSDF_SPECULAR = 16 };
enum { MAT_DIFFUSE = 0, MAT_MIRROR = 1, MAT_GLASS = 2, MAT_GLOSSY = 3, MAT_PHONG = 4, MAT_OREN_NAYAR = 5 };
enum { LIGHT_TRIANGLE = 0, LIGHT_DIRECTIONAL = 1, LIGHT_POINT = 2, LIGHT_SPOT = 3, LIGHT_ENV = 4 };

struct DevLight {            // 7 x float4
    float4 intensity;        // w = kind (int bits)
    float4 p0;               // triangle v0 | position | direction (directional);  w = area
    float4 p1;               // triangle v1 | spot axis;                            w = cos(angle)
    float4 p2;               // triangle v2
    float4 normal, tangent, binormal;
};

struct DeviceShading {
    const int*    mesh_vertex_base; const int* mesh_tri_base;
    const float4* vertices; const float4* normals; const float2* texcoords; const float4* geom_normals; const int4* indices;
    const float4* shade_tris;   // 8 x float4 (one 128-byte line) per global triangle: everything calculate_intersection reads of it, gathered
                                // once at upload -- q0 v0|material id, q1 v1|uv0.x, q2 v2|uv0.y, q3 n0|uv1.x, q4 n1|uv1.y, q5 n2|uv2.x, q6 gn|uv2.y, q7 unused
    const int*    inst_mesh; const float4* inst_mat; const float4* inst_inv; // 3 float4 per instance each
    const float4* materials;  // imba_material as 5 x float4
    const DevLight* lights;
    const int4*   textures;   // width, height, offset (texels)
    const float4* texels;
    int material_count, light_count, instance_count, texture_count;
    float bs_cx, bs_cy, bs_cz, bs_radius, bs_inv_radius_sqr;
    int   env_tex; float env_intensity;   // environment map (light.h:361-485): texture index or -1
};

/* ---- RNG ------------------------------------------------------------------------------------------------- */
struct Rng {
    uint64_t s;
    __device__ __forceinline__ uint32_t next() {
        const uint32_t c = (uint32_t)(s >> 32), x = (uint32_t)(s & 0xffffffffu);
        s = (uint64_t)x * 4294883355ull + c;
        return x ^ c;
    }
    __device__ __forceinline__ float uniform() { return (float)next() / (float)0xFFFFFFFFu; }
    __device__ __forceinline__ int range(int lo, int hi) { return hi == lo ? lo : (int)(next() % (uint32_t)(hi - lo)) + lo; }
    __device__ __forceinline__ void discard(int n) { for (int i = 0; i < n; ++i) next(); }
};

/* ---- surface point -------------------------------------------------------------------------------------- */
struct Isect {
    V3 pos, out_dir, normal, geom_normal, u_tangent, v_tangent;
    float2 uv;
    int mat;
};

__device__ __forceinline__ V3 tex_sample(const DeviceShading& S, int tex, float2 uv) {
    const int4 d = __ldg(S.textures + tex);
    const int w = d.x, h = d.y;
    float u = clampf(uv.x - (int)uv.x, -1.0f, 1.0f);
    float v = clampf(uv.y - (int)uv.y, -1.0f, 1.0f);
    u += u < 0.0f ? 1.0f : 0.0f;
    v += v < 0.0f ? 1.0f : 0.0f;
    v = 1.0f - v;
    const float kx = u * (w - 1), ky = v * (h - 1);
    const int x0 = (int)kx, y0 = (int)ky, x1 = (x0 + 1) % w, y1 = (y0 + 1) % h;
    const float gx = kx - floorf(kx), gy = ky - floorf(ky), hx = 1.0f - gx, hy = 1.0f - gy;
    const float4* t = S.texels + d.z;
    const V3 i00 = v3(__ldg(t + y0 * w + x0)), i10 = v3(__ldg(t + y0 * w + x1));
    const V3 i01 = v3(__ldg(t + y1 * w + x0)), i11 = v3(__ldg(t + y1 * w + x1));
    return hy * (hx * i00 + gx * i10) + gy * (hx * i01 + gx * i11);
}

__device__ __forceinline__ float2 lerp2(float2 a, float2 b, float2 c, float u, float v) {
    const float w = 1 - u - v;
    return make_float2(a.x * w + b.x * u + c.x * v, a.y * w + b.y * u + c.y * v);
}

/* integrator.h:71-123 */
__device__ __forceinline__ Isect calc_intersection(const DeviceShading& S, float4 ro, float4 rd, int tri_id, int inst_id, float t, float u) {
    // The reference follows instance -> mesh -> index record -> three vertices / normals / texture coordinates -> geometric normal
    // (integrator.h:72-107): four dependent levels and about ten cache lines per hit.  The hit carries the GLOBAL triangle id, so the
    // device reads one 128-byte record per triangle instead (shade_data.cu: same floats, same arithmetic below).
    const float4* q = S.shade_tris + 8 * (size_t)tri_id;
    const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4), q5 = __ldg(q + 5), q6 = __ldg(q + 6);
    const float4 m0 = __ldg(S.inst_inv + 3 * inst_id), m1 = __ldg(S.inst_inv + 3 * inst_id + 1), m2 = __ldg(S.inst_inv + 3 * inst_id + 2);
    const V3 org = v3(ro), out_dir = v3(rd);
    const V3 pos = org + t * out_dir;
    const V3 local_pos = v3(m0.x * pos.x + m0.y * pos.y + m0.z * pos.z + m0.w * 1.0f,
                            m1.x * pos.x + m1.y * pos.y + m1.z * pos.z + m1.w * 1.0f,
                            m2.x * pos.x + m2.y * pos.y + m2.z * pos.z + m2.w * 1.0f);
    const V3 v0 = v3(q0);
    const V3 e1 = v3(q1) - v0, e2 = v3(q2) - v0;
    const float v = dot(local_pos - v0 - u * e1, e2) / dot(e2, e2);
    Isect r;
    r.uv = lerp2(make_float2(q1.w, q2.w), make_float2(q3.w, q4.w), make_float2(q5.w, q6.w), u, v);
    const V3 n0 = v3(q3), n1 = v3(q4), n2 = v3(q5);
    const V3 ln = n0 * (1 - u - v) + n1 * u + n2 * v;
    // row vector * 3x4 matrix (float3x4.h:41-46): transforms normals by the inverse transpose
    const V3 nw = v3(m0.x * ln.x + m1.x * ln.y + m2.x * ln.z, m0.y * ln.x + m1.y * ln.y + m2.y * ln.z, m0.z * ln.x + m1.z * ln.y + m2.z * ln.z);
    const V3 g = v3(q6);
    const V3 gw = v3(m0.x * g.x + m1.x * g.y + m2.x * g.z, m0.y * g.x + m1.y * g.y + m2.y * g.z, m0.z * g.x + m1.z * g.y + m2.z * g.z);
    r.pos = pos; r.out_dir = -normalize(out_dir);
    r.normal = normalize(nw); r.geom_normal = normalize(gw);
    local_coordinates(r.normal, r.u_tangent, r.v_tangent);
    r.mat = __float_as_int(q0.w);
    // Material::bump (materials.h:34-55)
    const float4 mh = __ldg(S.materials + 5 * r.mat);
    const int bump = __float_as_int(mh.z);
    if (bump >= 0) {
        const float du = 0.001f, dv = 0.001f, vscale = 0.02f;
        const float ud = tex_sample(S, bump, make_float2(r.uv.x + du, r.uv.y)).x;
        const float vd = tex_sample(S, bump, make_float2(r.uv.x, r.uv.y + dv)).x;
        const float d0 = tex_sample(S, bump, r.uv).x;
        const float diff_u = vscale * (ud - d0) / du, diff_v = vscale * (vd - d0) / dv;
        const V3 n = cross(r.v_tangent, r.u_tangent);
        r.u_tangent = normalize(r.u_tangent + diff_u * n);
        r.v_tangent = normalize(r.v_tangent + diff_v * n);
        r.normal = cross(r.u_tangent, r.v_tangent);
    }
    if (dot(r.geom_normal, r.normal) < 0.0f) r.normal = -r.normal;
    return r;
}

/* ---- Fresnel (fresnel.h) --------------------------------------------------------------------------------- */
__device__ __forceinline__ float fresnel_conductor(float cosi, float eta, float kappa) {
    const float ekc = (eta * eta + kappa * kappa) * cosi * cosi;
    const float par = (ekc - (2.f * eta * cosi) + 1) / (ekc + (2.f * eta * cosi) + 1);
    const float ek = eta * eta + kappa * kappa;
    const float perp = (ek - (2.f * eta * cosi) + cosi * cosi) / (ek + (2.f * eta * cosi) + cosi * cosi);
    return (par + perp) / 2.f;
}
__device__ __forceinline__ float fresnel_dielectric_eval(float cosi, float eta_outside, float eta_inside) {
    float eta_in = eta_outside, eta_trans = eta_inside;
    if (cosi <= 0.0f) { const float t = eta_in; eta_in = eta_trans; eta_trans = t; }
    const float sin_trans = eta_in / eta_trans * sqrtf(fmaxf(0.0f, 1.0f - sqr(cosi)));
    if (sin_trans >= 1.0f) return 1.0f;
    const float cos_trans = sqrtf(fmaxf(0.0f, 1.0f - sqr(sin_trans)));
    const float ci = fabsf(cosi);
    const float par = (eta_trans * ci - eta_in * cos_trans) / (eta_trans * ci + eta_in * cos_trans);
    const float perp = (eta_in * ci - eta_trans * cos_trans) / (eta_in * ci + eta_trans * cos_trans);
    return (par * par + perp * perp) / 2.f;
}

/* ---- BSDF ------------------------------------------------------------------------------------------------ */
struct Bsdf {
    int kind; bool adjoint;
    V3 kd, ks, tf; float eta, kappa, exponent;
    V3 normal, geom_normal, u_tangent, v_tangent;

    __device__ __forceinline__ V3 to_local(V3 d) const { return v3(dot(v_tangent, d), dot(u_tangent, d), dot(normal, d)); }
    __device__ __forceinline__ V3 to_world(V3 d) const {
        return v3(v_tangent.x * d.x + u_tangent.x * d.y + normal.x * d.z, v_tangent.y * d.x + u_tangent.y * d.y + normal.y * d.z,
                  v_tangent.z * d.x + u_tangent.z * d.y + normal.z * d.z);
    }
    __device__ __forceinline__ bool has_btdf() const { return kind == MAT_GLASS; }
    __device__ __forceinline__ bool is_specular() const { return kind == MAT_MIRROR || kind == MAT_GLASS; } // Material::is_specular
    __device__ __forceinline__ int brdf_flags() const {
        return (kind == MAT_DIFFUSE || kind == MAT_OREN_NAYAR) ? (BSDF_DIFFUSE | BSDF_REFLECTION)
             : kind == MAT_GLOSSY  ? (BSDF_GLOSSY | BSDF_DIFFUSE | BSDF_REFLECTION)
             : kind == MAT_PHONG   ? (BSDF_GLOSSY | BSDF_REFLECTION) : (BSDF_SPECULAR | BSDF_REFLECTION);
    }
    // Phong (brdfs.h:48-92): lobe of power `exponent` about the mirrored direction, coefficient ks.  Quirks kept: eval mirrors the
    // INCOMING direction, pdf() is the power-cosine density of in_dir.z (about the normal, not about the mirrored direction).
    __device__ __forceinline__ V3 phong_eval(V3 o, V3 i) const {
        const V3 reflected_in = v3(-i.x, -i.y, i.z);
        float cos_r_o = fmaxf(0.0f, dot(reflected_in, o));
        cos_r_o = fminf(cos_r_o, 1.0f);
        return o.z * i.z > 0.0f ? ((exponent + 2.0f) / (2.0f * IMBA_PI)) * ks * powf(cos_r_o, exponent) : v3(0.0f);
    }
    __device__ __forceinline__ float phong_pdf(V3 o, V3 i) const {
        return i.z > 0.0f ? ((exponent + 1.0f) * powf(i.z, exponent) * 1.0f / (2.0f * IMBA_PI)) : 0.0f;
    }
    __device__ __forceinline__ V3 phong_sample(V3 o, V3& i, Rng& rng, float& pdf_out) const {
        const float u2 = rng.uniform(), u1 = rng.uniform();   // sample_power_cos_hemisphere(exponent, u1, u2): arguments right to left
        const float phi = IMBA_PI * 2.0f * u1;
        const float cos_t = powf(u2, 1.0f / (exponent + 1.0f));
        const float sin_t = sqrtf(1.0f - sqr(cos_t));
        const V3 dir = v3(sin_t * cosf(phi), sin_t * sinf(phi), cos_t);
        pdf_out = (exponent + 1.0f) * powf(cos_t, exponent) * 1.0f / (2.0f * IMBA_PI);
        const V3 reflected_in = v3(-o.x, -o.y, o.z);
        V3 t, b; local_coordinates(reflected_in, t, b);
        i = v3(b.x * dir.x + t.x * dir.y + reflected_in.x * dir.z, b.y * dir.x + t.y * dir.y + reflected_in.y * dir.z,
               b.z * dir.x + t.z * dir.y + reflected_in.z * dir.z);
        return o.z * i.z > 0.0f ? phong_eval(o, i) : v3(0.0f);
    }
    // CookTorrance with a Blinn distribution (brdfs.h:136-226), fresnel = conductor(1, exponent) (materials.h:174)
    __device__ __forceinline__ V3 cook_eval(V3 o, V3 i) const {
        if (fabsf(o.z) == 0.0f || fabsf(i.z) == 0.0f) return v3(0.0f);
        const V3 h = normalize(i + o);
        const float cos_half = dot(i, h);
        const float fr = fresnel_conductor(cos_half, 1.0f, exponent);
        if (!(o.z * i.z > 0.0f)) return v3(0.0f);
        const float d = (exponent + 2.0f) / (2.0f * IMBA_PI) * powf(fabsf(h.z), exponent);
        const float odh = dot(o, h);
        const float g = fminf(1.0f, fminf(2.0f * fabsf(h.z) * fabsf(o.z) / odh, 2.0f * fabsf(h.z) * fabsf(i.z) / odh));
        return (ks * d * g * fr) / (4.0f * fabsf(i.z) * fabsf(o.z));
    }
    __device__ __forceinline__ float cook_pdf(V3 o, V3 i) const {
        if (!(o.z * i.z > 0.0f)) return 0.0f;
        const V3 h = normalize(i + o);
        if (dot(o, h) <= 0.0f) return 0.0f;
        return (exponent + 1.0f) * powf(fabsf(h.z), exponent) / (2.0f * IMBA_PI * 4.0f * dot(o, h));
    }
    __device__ __forceinline__ V3 lambert_eval(V3 o, V3 i) const { return o.z * i.z > 0.0f ? kd * (1.0f / IMBA_PI) : v3(0.0f); }
    // OrenNayar (brdfs.h:94-134): reflectance kd, roughness sigma (degrees) in `exponent`; sample / pdf are BxDF's cosine-hemisphere defaults
    static __device__ __forceinline__ float on_sin_theta(V3 d) { return sqrtf(fmaxf(0.0f, 1.0f - sqr(d.z))); }
    static __device__ __forceinline__ float on_cos_phi(V3 d) { const float st = on_sin_theta(d); return st == 0.0f ? 1.0f : fminf(fmaxf(d.x / st, -1.0f), 1.0f); }
    static __device__ __forceinline__ float on_sin_phi(V3 d) { const float st = on_sin_theta(d); return st == 0.0f ? 0.0f : fminf(fmaxf(d.y / st, -1.0f), 1.0f); }
    __device__ __forceinline__ V3 oren_nayar_eval(V3 o, V3 i) const {
        const float sigma = exponent * IMBA_PI / 180.0f;
        const float sigma_sqr = sigma * sigma;
        const float pa = 1.0f - sigma_sqr / (2.0f * (sigma_sqr + 0.33f));
        const float pb = (float)(0.45 * (double)sigma_sqr / (double)(sigma_sqr + 0.09f));   // `0.45 * ...` is a double expression in the reference
        const float sin_theta_in = on_sin_theta(i), sin_theta_out = on_sin_theta(o);
        float max_cos = 0.0f;
        if (sin_theta_in > 0.0001f && sin_theta_out > 0.0001f)
            max_cos = fmaxf(0.0f, on_cos_phi(i) * on_cos_phi(o) + on_sin_phi(i) * on_sin_phi(o));
        float sin_alpha, tan_beta;
        if (fabsf(i.z) > fabsf(o.z)) { sin_alpha = sin_theta_out; tan_beta = sin_theta_in / fabsf(i.z); }
        else                         { sin_alpha = sin_theta_in;  tan_beta = sin_theta_out / fabsf(o.z); }
        return o.z * i.z > 0.0f ? kd * (1.0f / IMBA_PI) * (pa + pb * max_cos * sin_alpha * tan_beta) : v3(0.0f);
    }
    __device__ __forceinline__ float lambert_pdf(V3 o, V3 i) const { return o.z * i.z > 0.0f ? fabsf(i.z) / IMBA_PI : 0.0f; }

    __device__ __forceinline__ V3 brdf_eval(V3 o, V3 i) const {
        if (kind == MAT_DIFFUSE) return lambert_eval(o, i);
        if (kind == MAT_GLOSSY) return 0.5f * (cook_eval(o, i) + lambert_eval(o, i));
        if (kind == MAT_PHONG) return phong_eval(o, i);
        if (kind == MAT_OREN_NAYAR) return oren_nayar_eval(o, i);
        return v3(0.0f);
    }
    __device__ __forceinline__ float brdf_pdf(V3 o, V3 i) const {
        if (kind == MAT_DIFFUSE || kind == MAT_OREN_NAYAR) return lambert_pdf(o, i);
        if (kind == MAT_GLOSSY) return (cook_pdf(o, i) + lambert_pdf(o, i)) * 0.5f;
        if (kind == MAT_PHONG) return phong_pdf(o, i);
        return 0.0f;
    }
    /* bsdf.h:126-147, 207-222 with flags = BSDF_ALL */
    __device__ __forceinline__ V3 eval(V3 out_dir, V3 in_dir) const {
        if (dot(in_dir, geom_normal) * dot(out_dir, geom_normal) <= 0.0f) return v3(0.0f); // btdf (specular) or none
        return brdf_eval(to_local(out_dir), to_local(in_dir));
    }
    __device__ __forceinline__ float pdf(V3 out_dir, V3 in_dir) const {
        if (dot(in_dir, geom_normal) * dot(out_dir, geom_normal) <= 0.0f) return 0.0f;
        return brdf_pdf(to_local(out_dir), to_local(in_dir));
    }
    __device__ __forceinline__ V3 lambert_sample(V3 o, V3& i, Rng& rng, float& pdf_out) const {
        const float u2 = rng.uniform(), u1 = rng.uniform(); // g++ evaluates call arguments right to left
        const float st = sqrtf(1 - u2), ct = sqrtf(u2), phi = 2.f * IMBA_PI * u1;
        i = v3(st * cosf(phi), st * sinf(phi), ct);
        pdf_out = i.z / IMBA_PI;
        if (o.z < 0.0f) i.z = -i.z;
        return kind == MAT_OREN_NAYAR ? oren_nayar_eval(o, i) : lambert_eval(o, i);
    }
    __device__ __forceinline__ V3 cook_sample(V3 o, V3& i, Rng& rng, float& pdf_out) const {
        const float r2 = rng.uniform(), r1 = rng.uniform();
        const float c_theta = powf(r1, 1.0f / (exponent + 1.0f));
        const float s_theta = sqrtf(fmaxf(0.0f, 1.0f - sqr(c_theta)));
        const float phi = r2 * 2.0f * IMBA_PI;
        V3 h = v3(s_theta * cosf(phi), s_theta * sinf(phi), c_theta);
        if (!(o.z * h.z > 0.0f)) h = -h;
        i = -o + 2.0f * dot(o, h) * h;
        if (dot(o, h) <= 0.0f) pdf_out = 1.0f;
        else pdf_out = (exponent + 1.0f) * powf(c_theta, exponent) / (2.0f * IMBA_PI * 4.0f * dot(o, h));
        return o.z * i.z > 0.0f ? cook_eval(o, i) : v3(0.0f);
    }
    __device__ __forceinline__ V3 brdf_sample(V3 o, V3& i, Rng& rng, float& pdf_out) const {
        if (kind == MAT_DIFFUSE || kind == MAT_OREN_NAYAR) return lambert_sample(o, i, rng, pdf_out);
        if (kind == MAT_GLOSSY) return rng.uniform() < 0.5f ? cook_sample(o, i, rng, pdf_out) : lambert_sample(o, i, rng, pdf_out);
        if (kind == MAT_PHONG) return phong_sample(o, i, rng, pdf_out);
        // SpecularReflection (brdfs.h:25-45): conductor fresnel for mirrors, dielectric(1, eta) for glass
        i = v3(-o.x, -o.y, o.z);
        pdf_out = 1.0f;
        const float fr = kind == MAT_MIRROR ? fresnel_conductor(o.z, eta, kappa) : fresnel_dielectric_eval(o.z, 1.0f, eta);
        return fr * ks / fabsf(i.z);
    }
    /* SpecularTransmission(eta_inside = eta, eta_outside = 1) (btdfs.h:6-65, materials.h:137-141) */
    __device__ __forceinline__ V3 btdf_sample(V3 o, V3& i, float& pdf_out) const {
        pdf_out = 1.0f;
        float eta_in = 1.0f, eta_trans = eta;
        const bool entering = o.z > 0.0f;
        if (!entering) { eta_in = eta; eta_trans = 1.0f; }
        const float sin_in_sqr = fmaxf(0.0f, 1.0f - sqr(o.z));
        const float eta_frac = eta_in / eta_trans;
        const float sin_trans_sqr = sqr(eta_frac) * sin_in_sqr;
        if (sin_trans_sqr >= 1.0f) { i = v3(-o.x, -o.y, o.z); return v3(0.0f); }
        float cos_trans = sqrtf(fmaxf(0.0f, 1.0f - sin_trans_sqr));
        if (entering) cos_trans = -cos_trans;
        i = v3(eta_frac * -o.x, eta_frac * -o.y, cos_trans);
        const float fr = fresnel_dielectric_eval(o.z, 1.0f, eta);
        const float factor = adjoint ? 1.0f : sqr(eta_in / eta_trans);
        return factor * (1.0f - fr) * tf / fabsf(i.z);
    }
    /* bsdf.h:149-204 with flags = BSDF_ALL, or (specular_only) SPECULAR | REFLECTION | TRANSMISSION as the PPM camera
     * pass requests (vcm.cpp:141-143): non-specular BRDFs then do not match and nothing is drawn from the RNG */
    __device__ __forceinline__ V3 sample(V3 out_dir, V3& in_dir, Rng& rng, int& sampled_flags, float& pdf_out, bool specular_only = false) const {
        if (specular_only && (kind == MAT_DIFFUSE || kind == MAT_GLOSSY || kind == MAT_PHONG || kind == MAT_OREN_NAYAR)) { pdf_out = 0.0f; sampled_flags = 0; return v3(0.0f); }
        const V3 lo = to_local(out_dir);
        const float btdf_prob = has_btdf() ? 1.0f - fresnel_dielectric_eval(lo.z, 1.0f, eta) : 0.0f;
        const float rnd = rng.uniform();
        const bool use_btdf = rnd < btdf_prob;
        const float component_pdf = use_btdf ? btdf_prob : 1.0f - btdf_prob;
        V3 li;
        const V3 value = use_btdf ? btdf_sample(lo, li, pdf_out) : brdf_sample(lo, li, rng, pdf_out);
        if (pdf_out == 0.0f) { sampled_flags = 0; return v3(0.0f); }
        sampled_flags = use_btdf ? (BSDF_TRANSMISSION | BSDF_SPECULAR) : brdf_flags();
        in_dir = to_world(li);
        pdf_out *= component_pdf;
        const float side = dot(in_dir, geom_normal) * dot(out_dir, geom_normal);
        if ((!use_btdf && side <= 0.0f) || (use_btdf && side >= 0.0f)) { sampled_flags = 0; return v3(0.0f); }
        return value;
    }
};

/* Material::get_bsdf (materials.h:76-182) */
__device__ __forceinline__ Bsdf make_bsdf(const DeviceShading& S, const Isect& is, bool adjoint) {
    const float4* m = S.materials + 5 * is.mat;
    const float4 a = __ldg(m), b = __ldg(m + 1), c = __ldg(m + 2), d = __ldg(m + 3);
    Bsdf f;
    f.kind = __float_as_int(a.x); f.adjoint = adjoint;
    const int dtex = __float_as_int(a.y);
    f.kd = v3(b.x, b.y, b.z); f.eta = b.w;
    f.ks = v3(c.x, c.y, c.z); f.kappa = c.w;
    f.tf = v3(d.x, d.y, d.z); f.exponent = d.w;
    if (dtex >= 0) f.kd = tex_sample(S, dtex, is.uv);
    f.normal = is.normal; f.geom_normal = is.geom_normal; f.u_tangent = is.u_tangent; f.v_tangent = is.v_tangent;
    return f;
}

/* AreaEmitter::radiance (light.h:20-44): returns intensity or 0; pdfs by reference */
__device__ __forceinline__ V3 emitter_radiance(V3 intensity, float area, V3 out_dir, V3 normal, float& pdf_direct_a, float& pdf_emit_w) {
    const float cos_theta_o = dot(normal, out_dir);
    if (cos_theta_o <= 0.0f) { pdf_direct_a = 1.0f; pdf_emit_w = 1.0f; return v3(0.0f); }
    pdf_direct_a = 1.0f / area;
    V3 t, bn; local_coordinates(normal, t, bn);
    const float lz = dot(normal, out_dir);
    pdf_emit_w = 1.0f / area * (fabsf(lz) / IMBA_PI);
    return intensity;
}

/* ---- environment map (EnvMap, light.h:361-485; lat-long image, uniform sampling: the importance tables built by the
 * reference's constructor are never used, light.h:433-435,466) ----------------------------------------------------------- */
__device__ __forceinline__ V3 env_lookup(const DeviceShading& S, float s, float t) {
    const int4 d = __ldg(S.textures + S.env_tex);
    const int x = (int)(s * (d.x - 1)), y = (int)(t * (d.y - 1));
    return S.env_intensity * v3(__ldg(S.texels + d.z + y * d.x + x));
}
/* EnvMap::radiance (light.h:443-453); note the reference divides by sinf(t) with t in [0, 1], not by sin(theta) */
__device__ __forceinline__ V3 env_radiance(const DeviceShading& S, V3 out_dir, float& pdf_direct_w, float& pdf_emit_w) {
    float phi = atan2f(out_dir.x, out_dir.z);
    phi = (phi < 0.0f) ? (phi + 2.0f * IMBA_PI) : phi;
    const float s = phi / (2.0f * IMBA_PI);
    const float t = acosf(out_dir.y) / IMBA_PI;
    pdf_direct_w = 1.0f / (2.0f * IMBA_PI * IMBA_PI * sinf(t));
    pdf_emit_w = (1.0f / IMBA_PI) * S.bs_inv_radius_sqr * pdf_direct_w;
    return env_lookup(S, s, t);
}
/* EnvMap::sample_dir / sample_uv (light.h:458-485) */
__device__ __forceinline__ V3 env_sample_dir(const DeviceShading& S, Rng& rng, V3& dir, float& pdf) {
    const int4 d = __ldg(S.textures + S.env_tex);
    pdf = 1.0f / float(d.x * d.y);
    const float ux = rng.uniform();
    const float uy = rng.uniform();
    pdf *= float(d.x * d.y);
    const V3 color = env_lookup(S, ux, uy);
    const float theta = IMBA_PI * uy, sin_theta = sinf(theta), phi = 2.0f * IMBA_PI * ux;
    dir = v3(sin_theta * sinf(phi), cosf(theta), sin_theta * cosf(phi));
    if (sin_theta != 0.0f) pdf /= 2.0f * IMBA_PI * IMBA_PI * sin_theta; else pdf = 0.0f;
    return color;
}

/* ---- lights (light.h) --------------------------------------------------------------------------------- */
struct DirectSample { V3 dir; float distance; V3 radiance; float cos_out, pdf_emit_w, pdf_direct_w; bool delta; };
struct EmitSample { V3 pos, dir, radiance; float cos_out, pdf_emit_w, pdf_direct_a; bool delta, finite; };

__device__ __forceinline__ DirectSample light_sample_direct(const DeviceShading& S, int li, V3 from, Rng& rng) {
    const DevLight* L = S.lights + li;
    const float4 i4 = __ldg(&L->intensity), p0 = __ldg(&L->p0), p1 = __ldg(&L->p1);
    const int kind = __float_as_int(i4.w);
    const V3 intensity = v3(i4);
    DirectSample s;
    s.delta = kind != LIGHT_TRIANGLE && kind != LIGHT_ENV;
    if (kind == LIGHT_ENV) { // EnvLight::sample_direct (light.h:516-530)
        float pdf; V3 dir;
        const V3 radiance = env_sample_dir(S, rng, dir, pdf);
        s.dir = dir; s.distance = FLT_MAX; s.radiance = radiance / pdf; s.pdf_direct_w = pdf;
        s.pdf_emit_w = (1.0f / IMBA_PI) * S.bs_inv_radius_sqr * pdf; s.cos_out = 1.0f;
    } else if (kind == LIGHT_TRIANGLE) {
        const float area = p0.w;
        const float r2 = rng.uniform(), r1 = rng.uniform();
        const float sq = sqrtf(r1), u = 1.0f - sq, v = r2 * sq;
        const V3 pos = u * v3(p0) + v * v3(p1) + (1.0f - u - v) * v3(__ldg(&L->p2));
        s.dir = pos - from;
        const float distsq = dot(s.dir, s.dir);
        s.distance = sqrtf(distsq);
        s.dir = s.dir * (1.0f / s.distance);
        const float cos_out = dot(v3(__ldg(&L->normal)), -1.0f * s.dir);
        if (cos_out > 0.0f && cos_out < 1.0f) {
            s.radiance = intensity * cos_out * (area / distsq);
            s.cos_out = cos_out; s.pdf_emit_w = (fabsf(cos_out) / IMBA_PI) / area; s.pdf_direct_w = 1.0f / area * distsq / cos_out;
        } else { s.radiance = v3(0.0f); s.cos_out = 1.0f; s.pdf_emit_w = 1.0f; s.pdf_direct_w = 1.0f; }
    } else if (kind == LIGHT_DIRECTIONAL) {
        s.dir = -v3(p0); s.distance = FLT_MAX; s.radiance = intensity; s.pdf_direct_w = 1.0f;
        s.pdf_emit_w = (1.0f / IMBA_PI) * S.bs_inv_radius_sqr; s.cos_out = 1.0f;
    } else if (kind == LIGHT_POINT) {
        V3 dir = v3(p0) - from;
        const float sqdist = dot(dir, dir), dist = sqrtf(sqdist);
        dir *= 1.0f / dist;
        s.dir = dir; s.distance = dist; s.radiance = intensity / (4.0f * IMBA_PI * sqdist);
        s.pdf_direct_w = sqdist; s.pdf_emit_w = 1.0f / (4.0f * IMBA_PI); s.cos_out = 1.0f;
    } else { // spot
        V3 dir = v3(p0) - from;
        const float sqdist = dot(dir, dir), dist = sqrtf(sqdist);
        dir *= 1.0f / dist;
        const float cos_angle = p1.w;
        const float cos_o = -dot(dir, v3(p1));
        s.dir = dir; s.distance = dist; s.pdf_direct_w = sqdist;
        if (cos_o < cos_angle) { s.radiance = v3(0.0f); s.pdf_emit_w = 0.0f; }
        else { s.radiance = intensity / sqdist; s.pdf_emit_w = cos_o < cos_angle ? 0.0f : 1.0f / (2.0f * IMBA_PI * (1 - cos_angle)); }
        s.cos_out = 1.0f;
    }
    return s;
}

__device__ __forceinline__ float2 concentric_disc(float u1, float u2) {
    float phi, r;
    const float a = 2 * u1 - 1, b = 2 * u2 - 1;
    if (a > -b) {
        if (a > b) { r = a; phi = (IMBA_PI * 0.25f) * (b / a); }
        else       { r = b; phi = (IMBA_PI * 0.25f) * (2.f - (a / b)); }
    } else {
        if (a < b) { r = -a; phi = (IMBA_PI * 0.25f) * (4.f + (b / a)); }
        else       { r = -b; phi = b != 0 ? (IMBA_PI * 0.25f) * (6.f - (a / b)) : 0; }
    }
    return make_float2(r * cosf(phi), r * sinf(phi));
}

__device__ __forceinline__ EmitSample light_sample_emit(const DeviceShading& S, int li, Rng& rng) {
    const DevLight* L = S.lights + li;
    const float4 i4 = __ldg(&L->intensity), p0 = __ldg(&L->p0), p1 = __ldg(&L->p1);
    const int kind = __float_as_int(i4.w);
    const V3 intensity = v3(i4);
    EmitSample s;
    s.delta = kind != LIGHT_TRIANGLE && kind != LIGHT_ENV; s.finite = kind != LIGHT_DIRECTIONAL && kind != LIGHT_ENV;
    if (kind == LIGHT_ENV) { // EnvLight::sample_emit (light.h:494-513)
        float pdf; V3 dir;
        const V3 radiance = env_sample_dir(S, rng, dir, pdf);
        dir *= -1.0f;
        const float d2 = rng.uniform(), d1 = rng.uniform();
        const float2 disc = concentric_disc(d1, d2);
        V3 tangent, binormal; local_coordinates(dir, tangent, binormal);
        s.pos = v3(S.bs_cx, S.bs_cy, S.bs_cz) + S.bs_radius * (-dir + binormal * disc.x + tangent * disc.y);
        s.dir = dir; s.pdf_direct_a = pdf; s.pdf_emit_w = (1.0f / IMBA_PI) * S.bs_inv_radius_sqr * pdf; s.cos_out = 1.0f;
        s.radiance = radiance / s.pdf_emit_w;
    } else if (kind == LIGHT_TRIANGLE) {
        const float area = p0.w;
        const float r2 = rng.uniform(), r1 = rng.uniform();
        const float sq = sqrtf(r1), u = 1.0f - sq, v = r2 * sq;
        s.pos = u * v3(p0) + v * v3(p1) + (1.0f - u - v) * v3(__ldg(&L->p2));
        const float c2 = rng.uniform(), c1 = rng.uniform();
        const float st = sqrtf(1 - c2), ct = sqrtf(c2), phi = 2.f * IMBA_PI * c1;
        const V3 d = v3(st * cosf(phi), st * sinf(phi), ct);
        const float pdf = d.z / IMBA_PI;
        s.dir = d.x * v3(__ldg(&L->binormal)) + d.y * v3(__ldg(&L->tangent)) + d.z * v3(__ldg(&L->normal));
        if (pdf <= 0.0f) { s.radiance = v3(0.0f); s.cos_out = 0.0f; s.pdf_emit_w = 1.0f; s.pdf_direct_a = 1.0f; return s; }
        s.radiance = intensity * area * IMBA_PI;
        s.cos_out = d.z; s.pdf_emit_w = pdf / area; s.pdf_direct_a = 1.0f / area;
    } else if (kind == LIGHT_DIRECTIONAL) {
        const float d2 = rng.uniform(), d1 = rng.uniform();
        const float2 disc = concentric_disc(d1, d2);
        const V3 dir = v3(p0), c = v3(S.bs_cx, S.bs_cy, S.bs_cz);
        s.pos = c + S.bs_radius * (-dir + v3(__ldg(&L->binormal)) * disc.x + v3(__ldg(&L->tangent)) * disc.y);
        s.dir = dir; s.pdf_direct_a = 1.0f; s.pdf_emit_w = (1.0f / IMBA_PI) * S.bs_inv_radius_sqr; s.cos_out = 1.0f;
        s.radiance = intensity / s.pdf_emit_w;
    } else if (kind == LIGHT_POINT) {
        s.pos = v3(p0); s.radiance = intensity;
        const float u2 = rng.uniform(), u1 = rng.uniform();
        const float a = 2.0f * IMBA_PI * u1, b = 2.0f * sqrtf(u2 - sqr(u2));
        s.dir = v3(b * cosf(a), b * sinf(a), 1.0f - 2.0f * u2);
        s.pdf_direct_a = 1.0f; s.pdf_emit_w = 1.0f / (4.0f * IMBA_PI); s.cos_out = 1.0f;
    } else {
        s.pos = v3(p0);
        const float cos_angle = p1.w;
        const float u2 = rng.uniform(), u1 = rng.uniform();
        const float phi = 2.0f * IMBA_PI * u1, cos_t = 1.0f - u2 * (1.0f - cos_angle), sin_t = sqrtf(1.0f - cos_t * cos_t);
        const V3 d = v3(sin_t * cosf(phi), sin_t * sinf(phi), cos_t);
        const float pdf = 1.0f / (2.0f * IMBA_PI * (1 - cos_angle));
        s.cos_out = 1.0f;
        s.dir = d.x * v3(__ldg(&L->binormal)) + d.y * v3(__ldg(&L->tangent)) + d.z * v3(p1);
        s.radiance = intensity / pdf; s.pdf_direct_a = 1.0f; s.pdf_emit_w = pdf;
    }
    return s;
}

/* ---- states (render/scheduling/ray_queue.h:21-37, render/integrators/pt.h:10-15) ------------------------------- */
struct PTState {     // 40 bytes, layout of the reference's struct
    int pixel_id; int sample_id; uint64_t rng;
    float tr, tg, tb;
    uint32_t bounces_spec; // bounces : 31, last_specular : 1
    float last_pdf; uint32_t pad;
};
struct ShadowState { int pixel_id; float r, g, b; }; // 16 bytes
struct VCMState {    // 48 bytes (render/integrators/vcm.h:26-36)
    int pixel_id; int light_id; uint64_t rng;
    float tr, tg, tb;
    uint32_t len_finite; // path_length : 31, finite_light : 1
    float dVC, dVCM, dVM; uint32_t pad;
};
static_assert(sizeof(VCMState) == 48, "state layouts");
/* LightPathVertex (light_vertices.h:15-35) packed to 7 x float4 */
struct DevLightVertex {
    float4 pos_len;   // position, path_length (int bits)
    float4 out_mat;   // out_dir, material id (int bits)
    float4 n_dvc;     // shading normal, dVC
    float4 gn_dvcm;   // geometric normal, dVCM
    float4 ut_dvm;    // u_tangent, dVM
    float4 vt_u;      // v_tangent, uv.x
    float4 thr_v;     // throughput, uv.y
};
/* VCMPhoton (light_vertices.h:39-56) padded to 3 x float4 */
struct DevPhoton { float4 pos_dvcm; float4 dir_dvm; float4 thr; };
static_assert(sizeof(PTState) == 40 && sizeof(ShadowState) == 16, "state layouts");

#endif
