/* shade_data.cu -- device copies of the data the shading stage reads.  The reference keeps meshes, materials,
 * lights and textures host-side only (render/scene.h:65-91); a device-resident integrator needs flattened copies. */
#include "api_common.h"
#include "shade.cuh"
#include <cmath>
#include <cstring>

struct ShadingStorage {
    DeviceShading view;          // host copy of the pointer table (passed to kernels by value)
    std::vector<void*> allocs;
};

template <typename T>
static const T* up(ShadingStorage* st, const void* src, size_t count, bool& ok) {
    void* d = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    if (cudaMalloc(&d, bytes) != cudaSuccess) { ok = false; return nullptr; }
    st->allocs.push_back(d);
    if (count && src && cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
    return (const T*)d;
}

int imba_shading_upload(imba_scene* s, const imba_scene_desc* d) {
    // the shading kernels follow these indices on the device: refuse anything out of range here (the reference trusts its loader)
    for (int mi = 0; mi < d->mesh_count; ++mi) {
        const int nvm = d->mesh_vertex_base[mi + 1] - d->mesh_vertex_base[mi];
        for (int t = d->mesh_tri_base[mi]; t < d->mesh_tri_base[mi + 1]; ++t) {
            const int32_t* q = d->indices + 4 * (size_t)t;
            if (q[0] < 0 || q[0] >= nvm || q[1] < 0 || q[1] >= nvm || q[2] < 0 || q[2] >= nvm) { imba_set_error("triangle %d of mesh %d refers to a vertex outside the mesh (%d vertices)", t - d->mesh_tri_base[mi], mi, nvm); return IMBA_ERR_INVALID; }
            if (d->materials && (q[3] < 0 || q[3] >= d->material_count)) { imba_set_error("triangle %d of mesh %d refers to material %d (%d materials)", t - d->mesh_tri_base[mi], mi, q[3], d->material_count); return IMBA_ERR_INVALID; }
        }
    }
    for (int i = 0; i < d->instance_count; ++i)
        if (d->instances[i].mesh_id < 0 || d->instances[i].mesh_id >= d->mesh_count) { imba_set_error("instance %d refers to mesh %d", i, d->instances[i].mesh_id); return IMBA_ERR_INVALID; }
    ShadingStorage* st = new ShadingStorage;
    memset(&st->view, 0, sizeof(st->view));
    s->shading = reinterpret_cast<DeviceShading*>(st);
    DeviceShading& v = st->view;
    bool ok = true;
    const int nv = d->mesh_vertex_base[d->mesh_count], nt = d->mesh_tri_base[d->mesh_count];
    v.mesh_vertex_base = up<int>(st, d->mesh_vertex_base, d->mesh_count + 1, ok);
    v.mesh_tri_base = up<int>(st, d->mesh_tri_base, d->mesh_count + 1, ok);
    v.vertices = up<float4>(st, d->vertices, nv, ok);
    v.normals = up<float4>(st, d->normals, nv, ok);
    v.texcoords = up<float2>(st, d->texcoords, nv, ok);
    v.geom_normals = up<float4>(st, d->geom_normals, nt, ok);
    v.indices = up<int4>(st, d->indices, nt, ok);
    {   // one 128-byte record per triangle with everything calc_intersection reads of it (shade.cuh: DeviceShading::shade_tris)
        std::vector<float> rec(32 * (size_t)nt, 0.0f);
        for (int mi = 0; mi < d->mesh_count; ++mi) {
            const int vb = d->mesh_vertex_base[mi];
            for (int t = d->mesh_tri_base[mi]; t < d->mesh_tri_base[mi + 1]; ++t) {
                const int32_t* q = d->indices + 4 * (size_t)t;
                float* o = rec.data() + 32 * (size_t)t;
                for (int k = 0; k < 3; ++k) {
                    const size_t vi = (size_t)vb + q[k];
                    memcpy(o + 4 * k, d->vertices + 4 * vi, 12);            // q0..q2: positions
                    memcpy(o + 12 + 4 * k, d->normals + 4 * vi, 12);        // q3..q5: shading normals
                    o[4 * (1 + 2 * k) + 3] = d->texcoords[2 * vi];          // q1.w q3.w q5.w: u of corner k
                    o[4 * (2 + 2 * k) + 3] = d->texcoords[2 * vi + 1];      // q2.w q4.w q6.w: v of corner k
                }
                memcpy(o + 3, q + 3, 4);                                    // q0.w: material id
                memcpy(o + 24, d->geom_normals + 4 * (size_t)t, 12);        // q6: geometric normal
            }
        }
        v.shade_tris = up<float4>(st, rec.data(), 8 * (size_t)nt, ok);
    }
    std::vector<int> im(d->instance_count); std::vector<float> mat(12 * (size_t)d->instance_count), inv(12 * (size_t)d->instance_count);
    for (int i = 0; i < d->instance_count; ++i) {
        im[i] = d->instances[i].mesh_id;
        memcpy(&mat[12 * i], d->instances[i].mat, 48); memcpy(&inv[12 * i], d->instances[i].inv_mat, 48);
    }
    v.inst_mesh = up<int>(st, im.data(), im.size(), ok);
    v.inst_mat = up<float4>(st, mat.data(), 3 * (size_t)d->instance_count, ok);
    v.inst_inv = up<float4>(st, inv.data(), 3 * (size_t)d->instance_count, ok);
    static_assert(sizeof(imba_material) == 80, "material record");
    v.materials = up<float4>(st, d->materials, 5 * (size_t)d->material_count, ok);
    // lights: derived quantities of TriangleLight / DirectionalLight / SpotLight constructors (light.h:96-108, 196-200, 296-304)
    std::vector<DevLight> lights(d->light_count);
    for (int i = 0; i < d->light_count; ++i) {
        const imba_light& L = d->lights[i]; DevLight& o = lights[i];
        memset(&o, 0, sizeof(o));
        o.intensity = make_float4(L.intensity[0], L.intensity[1], L.intensity[2], 0.0f);
        memcpy(&o.intensity.w, &L.kind, 4);
        o.p0 = make_float4(L.p0[0], L.p0[1], L.p0[2], 0.0f); o.p1 = make_float4(L.p1[0], L.p1[1], L.p1[2], 0.0f);
        o.p2 = make_float4(L.p2[0], L.p2[1], L.p2[2], 0.0f);
        V3 n = v3(0.0f, 0.0f, 1.0f), t, b;
        if (L.kind == IMBA_LIGHT_TRIANGLE) {
            n = cross(v3(o.p1) - v3(o.p0), v3(o.p2) - v3(o.p0));
            o.p0.w = length(n) * 0.5f;
            n = normalize(n);
        } else if (L.kind == IMBA_LIGHT_DIRECTIONAL) n = v3(o.p0);
        else if (L.kind == IMBA_LIGHT_SPOT) { n = v3(o.p1); o.p1.w = cosf(L.angle); }
        local_coordinates(n, t, b);
        o.normal = make_float4(n.x, n.y, n.z, 0.0f); o.tangent = make_float4(t.x, t.y, t.z, 0.0f); o.binormal = make_float4(b.x, b.y, b.z, 0.0f);
    }
    v.lights = up<DevLight>(st, lights.data(), lights.size(), ok);
    std::vector<int4> tex(d->texture_count);
    for (int i = 0; i < d->texture_count; ++i) tex[i] = make_int4(d->textures[i].width, d->textures[i].height, (int)d->textures[i].offset, 0);
    v.textures = up<int4>(st, tex.data(), tex.size(), ok);
    v.texels = up<float4>(st, d->texels, (size_t)d->texel_count, ok);
    v.material_count = d->material_count; v.light_count = d->light_count; v.instance_count = d->instance_count; v.texture_count = d->texture_count;
    v.env_tex = d->env_tex; v.env_intensity = d->env_intensity;
    if (v.env_tex >= d->texture_count) { imba_set_error("env_tex out of range"); return IMBA_ERR_INVALID; }
    v.bs_cx = d->bsphere_center[0]; v.bs_cy = d->bsphere_center[1]; v.bs_cz = d->bsphere_center[2];
    v.bs_radius = d->bsphere_radius; v.bs_inv_radius_sqr = 1.0f / (d->bsphere_radius * d->bsphere_radius);
    if (!ok) { imba_set_error("shading data upload failed: %s", cudaGetErrorString(cudaGetLastError())); return IMBA_ERR_CUDA; }
    return IMBA_OK;
}

void imba_shading_free(imba_scene* s) {
    if (!s->shading) return;
    ShadingStorage* st = reinterpret_cast<ShadingStorage*>(s->shading);
    for (void* p : st->allocs) cudaFree(p);
    delete st;
    s->shading = nullptr;
}

const DeviceShading* imba_shading_view(const imba_scene* s) {
    return s->shading ? &reinterpret_cast<const ShadingStorage*>(s->shading)->view : nullptr;
}
