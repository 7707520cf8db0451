/* ORACLE / TEST INFRASTRUCTURE -- never linked into or called by the product path.
 *
 * Headless C-API driver around the UNMODIFIED reference host code, compiled in place from
 * /root/reference/src (see oracle/Makefile) against the shim headers in oracle/shim and the
 * traversal restatement in oracle/traversal_oracle.cpp.  It replaces frontend/main.cpp +
 * frontend/render_window.cpp (SDL) and exposes:
 *   - scene loading (frontend/build_scene.cpp:548) and export of the flattened scene in the
 *     product's imba_scene_desc form (what a maintainer's binding would fill);
 *   - full renders through the reference's own TileScheduler + PathTracer / VCMIntegrator;
 *   - PathTracer::process_primary_rays on an injected batch, via a RayScheduler subclass
 *     (the reference's own plug-in point, ray_scheduler.h:15-38);
 *   - RNG known-answer values.
 * Output lives in oracle/_ref/ (git-ignored, travels to the GPU box).
 */
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <memory>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <array>
#include <numeric>
#include <condition_variable>
#include <deque>

// Test infrastructure needs to read private parameters of the reference's materials and lights
// (there are no getters).  Standard headers are all included above, before the keyword hack.
#define private public
#define protected public
#include "imbatracer/frontend/cmd_line.h"
#include "imbatracer/frontend/build_scene.h"
#include "imbatracer/loaders/loaders.h"
#include "imbatracer/render/scene.h"
#include "imbatracer/render/ray_gen/ray_gen.h"
#include "imbatracer/render/ray_gen/camera.h"
#include "imbatracer/render/scheduling/tile_scheduler.h"
#include "imbatracer/render/integrators/pt.h"
#include "imbatracer/render/integrators/vcm.h"
#undef private
#undef protected

#define IMBA_B200_NO_L1_DECLS /* the shim headers declare the same symbols inside namespaces */
#include "../include/imba_b200.h"
#ifdef IMBA_REF_ON_B200
/* oracle/Makefile target ref_b200: this build links libimba_b200.so instead of the traversal restatement, keeps the scene's
 * "gpu" arrays in device memory (oracle/shim_cuda) and also exercises the product's plug-ins for the reference's interfaces. */
#include "../include/imba_b200_ref.hpp"
#endif

using namespace imba;

// load_png.cpp needs libpng (absent); the Cornell scenes have no textures.
namespace imba { bool load_png(const Path&, Image&) { return false; } }

// TEST HOOK for the Phong BxDF (render/materials/brdfs.h:48-92).  The reference defines the class but none of its materials builds
// one (build_scene.cpp:125-161), so parity needs a material on the reference side that does: BSDF(isect, Phong, nullptr), built in
// the per-thread arena exactly like DiffuseMaterial builds its Lambertian (materials.h:76-84).
class PhongTestMaterial : public Material {
public:
    PhongTestMaterial(const rgb& coefficient, float exponent) : coefficient_(coefficient), exponent_(exponent) {}
    Material* duplicate() const override final { return new PhongTestMaterial(coefficient_, exponent_); }
    BSDF* get_bsdf(const Intersection& isect, MemoryArena& mem_arena, bool adjoint) const override {
        auto brdf = mem_arena.alloc<Phong>(coefficient_, exponent_);
        return mem_arena.alloc<BSDF>(isect, brdf, nullptr);
    }
    rgb coefficient_; float exponent_;
};
// ... and for OrenNayar (brdfs.h:94-134), equally never instantiated by the reference's materials
class OrenNayarTestMaterial : public Material {
public:
    OrenNayarTestMaterial(const rgb& reflectance, float roughness_degrees) : reflectance_(reflectance), roughness_(roughness_degrees) {}
    Material* duplicate() const override final { return new OrenNayarTestMaterial(reflectance_, roughness_); }
    BSDF* get_bsdf(const Intersection& isect, MemoryArena& mem_arena, bool adjoint) const override {
        auto brdf = mem_arena.alloc<OrenNayar>(reflectance_, roughness_);
        return mem_arena.alloc<BSDF>(isect, brdf, nullptr);
    }
    rgb reflectance_; float roughness_;
};


namespace {

struct RefScene {
    std::unique_ptr<Scene> scene;
    float3 cam_pos, cam_dir, cam_up;
    bool gpu_layout;
    // storage for the exported flat scene
    std::vector<int32_t> mesh_vertex_base, mesh_tri_base, indices;
    std::vector<float> vertices, normals, texcoords, geom_normals;
    std::vector<imba_instance> instances;
    std::vector<imba_material> materials;
    std::vector<imba_light> lights;
    std::vector<imba_texture> textures;
    std::vector<float> texels;
    int node_count = 0;
};

struct Quiet {
    std::streambuf* old;
    std::ostringstream sink;
    Quiet() { old = std::cout.rdbuf(sink.rdbuf()); }
    ~Quiet() { std::cout.rdbuf(old); }
};

template <typename Node> int count_nodes(const Node* nodes, int root);
template <> int count_nodes<traversal_cpu::Node>(const traversal_cpu::Node* nodes, int root) {
    int maxi = root; std::vector<int> st{ root };
    while (!st.empty()) { int i = st.back(); st.pop_back(); maxi = std::max(maxi, i);
        for (int k = 0; k < 4; ++k) { int c = nodes[i].children[k]; if (c == 0) break; if (c > 0) st.push_back(c); } }
    return maxi + 1;
}
template <> int count_nodes<traversal_gpu::Node>(const traversal_gpu::Node* nodes, int root) {
    int maxi = root; std::vector<int> st{ root };
    while (!st.empty()) { int i = st.back(); st.pop_back(); maxi = std::max(maxi, i);
        if (nodes[i].left > 0) st.push_back(nodes[i].left);
        if (nodes[i].right > 0 && nodes[i].right != 0x76543210) st.push_back(nodes[i].right); }
    return maxi + 1;
}

void setup_camera(PerspectiveCamera& cam, const float3& pos, const float3& dir_in, const float3& up_in) {
    // CameraControl::setup, frontend/main.cpp:64-71
    float3 up = normalize(normalize(up_in));
    float3 dir = normalize(normalize(dir_in));
    float3 right = normalize(cross(dir, up));
    up = normalize(cross(right, dir));
    cam.move(pos, dir, up);
}

/// Scheduler that feeds one injected batch through the integrator's callbacks.
template <typename State, typename Shadow>
class BatchScheduler : public RayScheduler<State, Shadow> {
    using Base = RayScheduler<State, Shadow>;
public:
    BatchScheduler(Scene& s) : Base(s, false) {}
    RayQueue<State>* prim = nullptr; RayQueue<Shadow>* shadow = nullptr;
    void run_iteration(AtomicImage& out, typename Base::ProcessShadowFn, typename Base::ProcessPrimaryFn process_primary,
                       typename Base::SamplePixelFn) override {
        process_primary(*prim, *shadow, out);
    }
};

} // namespace

namespace {
struct ExportVertex { float pos[3]; int path_length; float out_dir[3]; int mat; float normal[3]; float dVC; float geom_normal[3]; float dVCM;
                      float u_tangent[3]; float dVM; float v_tangent[3]; float uv_x; float throughput[3]; float uv_y; };
static_assert(sizeof(ExportVertex) == 112, "vertex export record");

template <typename Integ>
int run_vcm_stage(Scene& scene, PerspectiveCamera& cam, const UserSettings& s, float base_radius, int iteration, int stage, int count, const imba_ray* rays,
                  const imba_hit* hits, const void* states, const ExportVertex* cache_in, int cache_in_count, imba_ray* rays_out, void* states_out, int* out_count,
                  imba_ray* shadow_rays, void* shadow_states, int* shadow_count, float* fb_rgb, ExportVertex* cache_out, int* cache_out_count) {
    static_assert(sizeof(VCMState) == 48 && sizeof(VCMShadowState) == 16, "state sizes");
    BatchScheduler<VCMState, VCMShadowState> sched(scene);
    Integ integ(scene, cam, sched, s);
    // what VCMIntegrator::preprocess / render would have set (vcm.h:74-81, vcm.cpp:28-46)
    const float radius_alpha = 0.75f;
    integ.base_radius_ = base_radius;
    integ.cur_iteration_ = iteration;
    integ.pm_radius_ = std::max(base_radius / powf(static_cast<float>(iteration), 0.5f * (1.0f - radius_alpha)), 1e-7f);
    const float eta_vcm = pi * sqr(integ.pm_radius_) * s.light_path_count;
    integ.mis_eta_vc_ = 1.0f / eta_vcm;
    integ.mis_eta_vm_ = std::is_same<Integ, BPT>::value ? 0.0f : eta_vcm;
    auto& lv = integ.light_vertices_;
    const int cache_size = std::max(count * 2, cache_in_count) + 16;
    lv.cache_.resize(cache_size); lv.last_.store(0); lv.count_ = 0;
    RayQueue<VCMState> prim(count, false); RayQueue<VCMShadowState> shadow(count * (s.num_connections + 1), false);
    const VCMState* st = (const VCMState*)states;
    for (int i = 0; i < count; ++i) { ::Ray r; memcpy(&r, &rays[i], 32); prim.push(r, st[i]); }
    memcpy(prim.hits(), hits, sizeof(Hit) * count);
    AtomicImage img(s.width, s.height); img.clear();
    if (stage == 0) {
        integ.process_light_rays(prim, shadow, img);
        const int n = std::min<int>(lv.cache_.size(), lv.last_.load());
        for (int i = 0; i < n; ++i) {
            const LightPathVertex& v = lv.cache_[i]; ExportVertex& o = cache_out[i];
            for (int k = 0; k < 3; ++k) { o.pos[k] = v.isect.pos[k]; o.out_dir[k] = v.isect.out_dir[k]; o.normal[k] = v.isect.normal[k]; o.geom_normal[k] = v.isect.geom_normal[k];
                                          o.u_tangent[k] = v.isect.u_tangent[k]; o.v_tangent[k] = v.isect.v_tangent[k]; o.throughput[k] = v.throughput[k]; }
            o.path_length = v.path_length; o.dVC = v.dVC; o.dVCM = v.dVCM; o.dVM = v.dVM; o.uv_x = v.isect.uv.x; o.uv_y = v.isect.uv.y;
            o.mat = -1; for (size_t m = 0; m < scene.material_count(); ++m) if (scene.material(m).get() == v.isect.mat) { o.mat = (int)m; break; }
        }
        *cache_out_count = n;
    } else {
        for (int i = 0; i < cache_in_count; ++i) {
            const ExportVertex& o = cache_in[i]; LightPathVertex v;
            for (int k = 0; k < 3; ++k) { v.isect.pos[k] = o.pos[k]; v.isect.out_dir[k] = o.out_dir[k]; v.isect.normal[k] = o.normal[k]; v.isect.geom_normal[k] = o.geom_normal[k];
                                          v.isect.u_tangent[k] = o.u_tangent[k]; v.isect.v_tangent[k] = o.v_tangent[k]; v.throughput[k] = o.throughput[k]; }
            v.isect.uv.x = o.uv_x; v.isect.uv.y = o.uv_y; v.isect.mat = scene.material(o.mat).get();
            v.path_length = o.path_length; v.dVC = o.dVC; v.dVCM = o.dVCM; v.dVM = o.dVM;
            lv.cache_[i] = v;
        }
        lv.last_.store(cache_in_count);
        lv.build(integ.pm_radius_, !std::is_same<Integ, BPT>::value && !std::is_same<Integ, VCM_PT>::value);
        integ.process_camera_rays(prim, shadow, img);
        if (cache_out_count) *cache_out_count = 0;
    }
    *out_count = prim.size(); *shadow_count = shadow.size();
    memcpy(rays_out, prim.rays(), 32 * prim.size()); memcpy(states_out, prim.states(), sizeof(VCMState) * prim.size());
    memcpy(shadow_rays, shadow.rays(), 32 * shadow.size()); memcpy(shadow_states, shadow.states(), sizeof(VCMShadowState) * shadow.size());
    if (fb_rgb) for (int i = 0; i < s.width * s.height; ++i) { fb_rgb[3 * i] = img.pixels()[i][0]; fb_rgb[3 * i + 1] = img.pixels()[i][1]; fb_rgb[3 * i + 2] = img.pixels()[i][2]; }
    return 0;
}
} // namespace

extern "C" {

void* ref_scene_load(const char* path, int gpu_layout) {
    Quiet q;
    auto rs = new RefScene;
    rs->gpu_layout = gpu_layout != 0;
    rs->scene.reset(new Scene(!rs->gpu_layout, rs->gpu_layout));
    if (!build_scene(Path(path), *rs->scene, rs->cam_pos, rs->cam_dir, rs->cam_up)) { delete rs; return nullptr; }
    return rs;
}

void ref_scene_free(void* h) { delete (RefScene*)h; }

void ref_scene_camera(void* h, float* pos, float* dir, float* up) {
    auto rs = (RefScene*)h;
    for (int i = 0; i < 3; ++i) { pos[i] = rs->cam_pos[i]; dir[i] = rs->cam_dir[i]; up[i] = rs->cam_up[i]; }
}

/// LightVertices::compute_cache_size (light_vertices.cpp:15-129) of the reference for `light_path_count` paths: 10 000 probe light
/// paths traced until fewer than 256 are left; returns the number of vertex slots it allocates (margin 1.1, or 10 for < 1000 paths).
int ref_vcm_cache_size(void* h, int light_path_count) {
    Quiet q;
    auto rs = (RefScene*)h;
    LightVertices lv(light_path_count);
    lv.compute_cache_size(*rs->scene, rs->gpu_layout);
    return (int)lv.cache_.size();
}

/// Makes the reference's std::random_device reproducible (oracle/shim/det_random.h); 0 = the real device again.
void ref_set_seed(uint64_t seed) { imba_det::seed(seed); }

/// 1 when this library is the reference linked against libimba_b200.so (Makefile target ref_b200), 0 for the oracle build.
int ref_is_linked_to_product(void) {
#ifdef IMBA_REF_ON_B200
    return 1;
#else
    return 0;
#endif
}

/// Fills `d` with pointers into storage owned by the handle.
int ref_scene_export(void* h, imba_scene_desc* d) {
    auto rs = (RefScene*)h;
#ifdef IMBA_REF_ON_B200
    if (rs->gpu_layout) return -4;   // the traversal arrays live in device memory in this build
#endif
    Scene& sc = *rs->scene;
    memset(d, 0, sizeof(*d));
    rs->mesh_vertex_base.assign(1, 0); rs->mesh_tri_base.assign(1, 0);
    rs->vertices.clear(); rs->normals.clear(); rs->texcoords.clear(); rs->geom_normals.clear(); rs->indices.clear();
    for (size_t m = 0; m < sc.mesh_count(); ++m) {
        const Mesh& mesh = sc.mesh(m);
        auto tc = mesh.attribute<float2>(MeshAttributes::TEXCOORDS);
        auto nr = mesh.attribute<float3>(MeshAttributes::NORMALS);
        auto gn = mesh.attribute<float3>(MeshAttributes::GEOM_NORMALS);
        for (size_t v = 0; v < mesh.vertex_count(); ++v) {
            const float4& p = mesh.vertices()[v];
            rs->vertices.insert(rs->vertices.end(), { p.x, p.y, p.z, 1.0f });
            rs->normals.insert(rs->normals.end(), { nr[v].x, nr[v].y, nr[v].z, 0.0f });
            rs->texcoords.insert(rs->texcoords.end(), { tc[v].x, tc[v].y });
        }
        for (size_t t = 0; t < mesh.triangle_count(); ++t) {
            rs->geom_normals.insert(rs->geom_normals.end(), { gn[t].x, gn[t].y, gn[t].z, 0.0f });
            for (int k = 0; k < 4; ++k) rs->indices.push_back((int32_t)mesh.indices()[t * 4 + k]);
        }
        rs->mesh_vertex_base.push_back(rs->mesh_vertex_base.back() + (int)mesh.vertex_count());
        rs->mesh_tri_base.push_back(rs->mesh_tri_base.back() + (int)mesh.triangle_count());
        if (sc.local_tri_id(0, (int)m) != -rs->mesh_tri_base[m]) return -1;
    }
    rs->instances.clear();
    for (auto& inst : sc.instances()) {
        imba_instance o; o.mesh_id = inst.id;
        memcpy(o.mat, &inst.mat, 48); memcpy(o.inv_mat, &inst.inv_mat, 48);
        rs->instances.push_back(o);
    }
    // textures
    rs->textures.clear(); rs->texels.clear();
    std::unordered_map<const TextureSampler*, int> tex_id;
    for (size_t t = 0; t < sc.texture_count(); ++t) {
        const Image& img = sc.texture(t)->image();
        imba_texture o; o.width = img.width(); o.height = img.height(); o.offset = (int64_t)rs->texels.size() / 4;
        for (int i = 0; i < img.width() * img.height(); ++i) { const rgba& p = img.pixels()[i]; rs->texels.insert(rs->texels.end(), { p.x, p.y, p.z, p.w }); }
        rs->textures.push_back(o); tex_id[sc.texture(t).get()] = (int)t;
    }
    auto tid = [&](const TextureSampler* s) { return s ? tex_id.at(s) : -1; };
    rs->materials.clear();
    for (size_t m = 0; m < sc.material_count(); ++m) {
        Material* mat = sc.material(m).get();
        imba_material o; memset(&o, 0, sizeof(o)); o.diffuse_tex = -1; o.bump_tex = tid(mat->bump_);
        if (auto p = dynamic_cast<DiffuseMaterial*>(mat)) {
            o.kind = IMBA_MAT_DIFFUSE; o.diffuse_tex = tid(p->sampler_);
            if (!p->sampler_) { o.kd[0] = p->color_.x; o.kd[1] = p->color_.y; o.kd[2] = p->color_.z; }
        } else if (auto p = dynamic_cast<MirrorMaterial*>(mat)) {
            o.kind = IMBA_MAT_MIRROR; o.eta = p->fresnel_.eta_; o.kappa = p->fresnel_.kappa_;
            o.ks[0] = p->scale_.x; o.ks[1] = p->scale_.y; o.ks[2] = p->scale_.z;
        } else if (auto p = dynamic_cast<GlassMaterial*>(mat)) {
            o.kind = IMBA_MAT_GLASS; o.eta = p->eta_;
            o.ks[0] = p->reflectance_.x; o.ks[1] = p->reflectance_.y; o.ks[2] = p->reflectance_.z;
            o.tf[0] = p->transmittance_.x; o.tf[1] = p->transmittance_.y; o.tf[2] = p->transmittance_.z;
        } else if (auto p = dynamic_cast<OrenNayarTestMaterial*>(mat)) {
            o.kind = IMBA_MAT_OREN_NAYAR; o.exponent = p->roughness_;
            o.kd[0] = p->reflectance_.x; o.kd[1] = p->reflectance_.y; o.kd[2] = p->reflectance_.z;
        } else if (auto p = dynamic_cast<PhongTestMaterial*>(mat)) {
            o.kind = IMBA_MAT_PHONG; o.exponent = p->exponent_;
            o.ks[0] = p->coefficient_.x; o.ks[1] = p->coefficient_.y; o.ks[2] = p->coefficient_.z;
        } else if (auto p = dynamic_cast<GlossyMaterial*>(mat)) {
            o.kind = IMBA_MAT_GLOSSY; o.exponent = p->exponent_; o.diffuse_tex = tid(p->diff_sampler_);
            o.ks[0] = p->specular_color_.x; o.ks[1] = p->specular_color_.y; o.ks[2] = p->specular_color_.z;
            o.kd[0] = p->diffuse_color_.x; o.kd[1] = p->diffuse_color_.y; o.kd[2] = p->diffuse_color_.z;
        } else return -2;
        if (auto e = mat->emitter()) {
            o.is_emitter = 1; o.emit[0] = e->intensity.x; o.emit[1] = e->intensity.y; o.emit[2] = e->intensity.z; o.emit_area = e->area;
        }
        rs->materials.push_back(o);
    }
    rs->lights.clear();
    for (size_t l = 0; l < sc.light_count(); ++l) {
        Light* li = sc.light(l).get();
        imba_light o; memset(&o, 0, sizeof(o));
        auto set3 = [](float* d, const float3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
        if (auto p = dynamic_cast<TriangleLight*>(li)) {
            o.kind = IMBA_LIGHT_TRIANGLE; set3(o.intensity, p->emit_.intensity);
            set3(o.p0, p->verts_[0]); set3(o.p1, p->verts_[1]); set3(o.p2, p->verts_[2]);
        } else if (auto p = dynamic_cast<DirectionalLight*>(li)) {
            o.kind = IMBA_LIGHT_DIRECTIONAL; set3(o.intensity, p->intensity_); set3(o.p0, p->dir_);
        } else if (auto p = dynamic_cast<PointLight*>(li)) {
            o.kind = IMBA_LIGHT_POINT; set3(o.intensity, p->intensity_); set3(o.p0, p->pos_);
        } else if (auto p = dynamic_cast<SpotLight*>(li)) {
            o.kind = IMBA_LIGHT_SPOT; set3(o.intensity, p->intensity_); set3(o.p0, p->pos_); set3(o.p1, p->normal_); o.angle = p->angle_;
        } else if (dynamic_cast<EnvLight*>(li)) {
            o.kind = IMBA_LIGHT_ENV;
        } else return -3;
        rs->lights.push_back(o);
    }
    d->mesh_count = (int)sc.mesh_count();
    d->mesh_vertex_base = rs->mesh_vertex_base.data(); d->mesh_tri_base = rs->mesh_tri_base.data();
    d->vertices = rs->vertices.data(); d->normals = rs->normals.data(); d->texcoords = rs->texcoords.data();
    d->geom_normals = rs->geom_normals.data(); d->indices = rs->indices.data();
    d->instance_count = (int)rs->instances.size(); d->instances = rs->instances.data();
    d->material_count = (int)rs->materials.size(); d->materials = rs->materials.data();
    d->light_count = (int)rs->lights.size(); d->lights = rs->lights.data();
    d->texture_count = (int)rs->textures.size(); d->textures = rs->textures.data();
    d->texels = rs->texels.data(); d->texel_count = (int64_t)rs->texels.size() / 4;
    d->env_tex = -1; d->env_intensity = 0.0f;
    if (sc.env_map()) { // EnvMap keeps its own image (light.h:476): exported as one more texture
        const Image& img = sc.env_map()->img_;
        imba_texture o; o.width = img.width(); o.height = img.height(); o.offset = (int64_t)rs->texels.size() / 4;
        for (int i = 0; i < img.width() * img.height(); ++i) { const rgba& p = img.pixels()[i]; rs->texels.insert(rs->texels.end(), { p.x, p.y, p.z, p.w }); }
        rs->textures.push_back(o);
        d->env_tex = (int)rs->textures.size() - 1; d->env_intensity = sc.env_map()->intensity_;
        d->texture_count = (int)rs->textures.size(); d->textures = rs->textures.data();
        d->texels = rs->texels.data(); d->texel_count = (int64_t)rs->texels.size() / 4;
    }
    d->bsphere_center[0] = sc.bounding_sphere().center.x; d->bsphere_center[1] = sc.bounding_sphere().center.y;
    d->bsphere_center[2] = sc.bounding_sphere().center.z; d->bsphere_radius = sc.bounding_sphere().radius;
    if (rs->gpu_layout) {
        auto& td = sc.traversal_data_gpu();
        d->accel_layout = IMBA_ACCEL_BVH2; d->root = td.root; d->nodes = td.nodes.data();
        d->node_count = count_nodes<traversal_gpu::Node>(td.nodes.data(), td.root);
        d->tris = (const imba_vec4*)td.tris.data(); d->tri_vec4_count = td.tris.size();
        d->inst_nodes = (const imba_instance_node*)td.instances.data(); d->inst_node_count = (int)td.instances.size();
        d->mask_count = (int)td.masks.size(); d->masks = (const imba_mask*)td.masks.data();
        d->mask_buffer = (const uint8_t*)td.mask_buffer.data(); d->mask_buffer_size = td.mask_buffer.size();
        d->mask_indices = td.indices.data(); d->mask_texcoords = (const imba_vec2*)td.texcoords.data(); d->mask_texcoord_count = td.texcoords.size();
    } else {
        auto& td = sc.traversal_data_cpu();
        d->accel_layout = IMBA_ACCEL_MBVH4; d->root = td.root; d->nodes = td.nodes.data();
        d->node_count = count_nodes<traversal_cpu::Node>(td.nodes.data(), td.root);
        d->tris = (const imba_vec4*)td.tris.data(); d->tri_vec4_count = td.tris.size();
        d->inst_nodes = (const imba_instance_node*)td.instances.data(); d->inst_node_count = (int)td.instances.size();
        d->mask_count = (int)td.masks.size(); d->masks = (const imba_mask*)td.masks.data();
        d->mask_buffer = (const uint8_t*)td.mask_buffer.data(); d->mask_buffer_size = td.mask_buffer.size();
        d->mask_indices = td.indices.data(); d->mask_texcoords = (const imba_vec2*)td.texcoords.data(); d->mask_texcoord_count = td.texcoords.size();
    }
    return 0;
}

/// Full render with the reference's own scheduler and integrators (frontend/main.cpp:87-163 minus SDL).
/// out_rgb: w*h*3 floats, the SUM over iterations (not normalised).  stats: [0] primary rays, [1] shadow rays.
/// seconds: wall clock of the render() calls only (render_window.cpp:60-66).
int ref_render(void* h, int algo, int w, int hgt, float fov, int spp, int iterations, int threads, int tile,
               int max_path_len, int light_paths, int connections, int knn, float radius_factor,
               float* out_rgb, uint64_t* stats, double* seconds, float* pixel_size_out) {
    Quiet q;
    auto rs = (RefScene*)h;
    Scene& scene = *rs->scene;
    UserSettings s;
    s.width = w; s.height = hgt; s.fov = fov; s.concurrent_spp = spp; s.thread_count = threads; s.tile_size = tile;
    s.max_path_len = max_path_len; s.light_path_count = light_paths > 0 ? light_paths : w * hgt / 2;
    s.num_connections = connections; s.num_knn = knn; s.radius_factor = radius_factor;
    s.traversal_platform = rs->gpu_layout ? UserSettings::gpu : UserSettings::cpu;
    s.algorithm = (UserSettings::Algorithm)algo;
    const bool gpu = rs->gpu_layout;

    PerspectiveCamera cam(s.width, s.height, s.fov);
    setup_camera(cam, rs->cam_pos, rs->cam_dir, rs->cam_up);
    AtomicImage img(w, hgt); img.clear();
    double secs = 0.0; uint64_t prim = 0, shad = 0; float px = 0.0f;
    auto loop = [&](Integrator& integ) {
        integ.preprocess(); px = integ.pixel_size();
        auto t0 = std::chrono::high_resolution_clock::now();
        for (int i = 0; i < iterations; ++i) integ.render(img);
        secs = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    };
    if (s.algorithm == UserSettings::PT) {
        DefaultTileGen<PTState> ray_gen(s.width, s.height, s.concurrent_spp, s.tile_size);
        TileScheduler<PTState, ShadowState> scheduler(ray_gen, scene, 1, s.thread_count, s.tile_size * s.tile_size * s.concurrent_spp, gpu);
        PathTracer integ(scene, cam, scheduler, s.max_path_len);
        loop(integ);
        prim = scheduler.total_prim_rays_; shad = scheduler.total_shadow_rays_;
    } else {
        DefaultTileGen<VCMState> ray_gen(s.width, s.height, s.concurrent_spp, s.tile_size);
        TileScheduler<VCMState, VCMShadowState> scheduler(ray_gen, scene, s.num_connections + 1, s.thread_count, s.tile_size * s.tile_size * s.concurrent_spp, gpu);
        std::unique_ptr<Integrator> integ;
        uint64_t lp = 0, ls = 0;
        switch (s.algorithm) {
            case UserSettings::BPT:    integ.reset(new BPT(scene, cam, scheduler, s)); break;
            case UserSettings::PPM:    integ.reset(new PPM(scene, cam, scheduler, s)); break;
            case UserSettings::LT:     integ.reset(new LT(scene, cam, scheduler, s)); break;
            case UserSettings::VCM_PT: integ.reset(new VCM_PT(scene, cam, scheduler, s)); break;
            default:                   integ.reset(new VCM(scene, cam, scheduler, s)); break;
        }
        loop(*integ);
        switch (s.algorithm) {
#define LS(T) { auto p = static_cast<T*>(integ.get()); lp = p->light_scheduler_.total_prim_rays_; ls = p->light_scheduler_.total_shadow_rays_; }
            case UserSettings::BPT: LS(BPT) break; case UserSettings::PPM: LS(PPM) break; case UserSettings::LT: LS(LT) break;
            case UserSettings::VCM_PT: LS(VCM_PT) break; default: LS(VCM) break;
#undef LS
        }
        prim = scheduler.total_prim_rays_ + lp; shad = scheduler.total_shadow_rays_ + ls;
        integ.reset();
    }
    for (int i = 0; i < w * hgt; ++i) { out_rgb[3 * i] = img.pixels()[i][0]; out_rgb[3 * i + 1] = img.pixels()[i][1]; out_rgb[3 * i + 2] = img.pixels()[i][2]; }
    if (stats) { stats[0] = prim; stats[1] = shad; }
    if (seconds) *seconds = secs;
    if (pixel_size_out) *pixel_size_out = px;
    return 0;
}

#ifdef IMBA_REF_ON_B200
/// The product's plug-ins at the reference's own interfaces (include/imba_b200_ref.hpp), driven like frontend/main.cpp:117-163 +
/// render_window.cpp:93-95 drive the reference's classes.
///   mode 0: the reference's PathTracer / VCMIntegrator<algo> on imba::B200WavefrontScheduler (a RayScheduler subclass);
///   mode 1: imba::B200Integrator (an Integrator subclass, everything on the device) -- needs the host ("cpu") layout for the export.
/// out_rgb: SUM over iterations; stats: primary / shadow rays; pixel_sizes[0] = Integrator::pixel_size() (the reference's probe through
/// the linked traversal symbols), pixel_sizes[1] = the device integrator's own probe (mode 1).
int ref_render_plugin(void* h, int mode, int algo, int w, int hgt, float fov, int spp, int iterations, int threads, int tile, int max_path_len,
                      int light_paths, int connections, int knn, float radius_factor, unsigned seed, float* out_rgb, uint64_t* stats, float* pixel_sizes) {
    Quiet q;
    auto rs = (RefScene*)h;
    Scene& scene = *rs->scene;
    UserSettings s;
    s.width = w; s.height = hgt; s.fov = fov; s.concurrent_spp = spp; s.thread_count = threads; s.tile_size = tile;
    s.max_path_len = max_path_len; s.light_path_count = light_paths > 0 ? light_paths : w * hgt / 2;
    s.num_connections = connections; s.num_knn = knn; s.radius_factor = radius_factor;
    s.traversal_platform = rs->gpu_layout ? UserSettings::gpu : UserSettings::cpu;
    s.algorithm = (UserSettings::Algorithm)algo;
    const bool gpu = rs->gpu_layout;
    PerspectiveCamera cam(s.width, s.height, s.fov);
    setup_camera(cam, rs->cam_pos, rs->cam_dir, rs->cam_up);
    AtomicImage img(w, hgt); img.clear();
    pixel_sizes[0] = pixel_sizes[1] = 0.0f; stats[0] = stats[1] = 0;
    try {
        if (mode == 1) {
            imba_scene_desc d;
            const int rc = ref_scene_export(h, &d);
            if (rc != 0) return rc;
            imba_scene* dev = imba_scene_create(&d);
            if (!dev) { fprintf(stderr, "ref_render_plugin: %s\n", imba_last_error()); return -5; }
            {
                B200Integrator integ(scene, cam, dev, b200_render_settings(s, rs->cam_pos, rs->cam_dir, rs->cam_up, seed));
                Integrator& base = integ;            // from here on only the reference's interface is used
                base.preprocess();
                pixel_sizes[0] = base.pixel_size(); pixel_sizes[1] = integ.device_pixel_size();
                for (int i = 0; i < iterations; ++i) base.render(img);
                base.reset();                        // camera unchanged: restarts the schedule, must not disturb the image
                integ.ray_counts(stats);
            }
            imba_scene_destroy(dev);
        } else if (s.algorithm == UserSettings::PT) {
            PixelRayGen<PTState> ray_gen(s.width, s.height, s.concurrent_spp);
            B200WavefrontScheduler<PTState, ShadowState> scheduler(ray_gen, scene, 1, w * hgt * spp, gpu);
            PathTracer integ(scene, cam, scheduler, s.max_path_len);
            integ.preprocess(); pixel_sizes[0] = integ.pixel_size();
            for (int i = 0; i < iterations; ++i) integ.render(img);
            stats[0] = scheduler.total_prim_rays(); stats[1] = scheduler.total_shadow_rays();
        } else {
            PixelRayGen<VCMState> ray_gen(s.width, s.height, s.concurrent_spp);
            B200WavefrontScheduler<VCMState, VCMShadowState> scheduler(ray_gen, scene, s.num_connections + 1, w * hgt * spp, gpu);
            std::unique_ptr<Integrator> integ;
            switch (s.algorithm) {
                case UserSettings::BPT:    integ.reset(new BPT(scene, cam, scheduler, s)); break;
                case UserSettings::PPM:    integ.reset(new PPM(scene, cam, scheduler, s)); break;
                case UserSettings::LT:     integ.reset(new LT(scene, cam, scheduler, s)); break;
                case UserSettings::VCM_PT: integ.reset(new VCM_PT(scene, cam, scheduler, s)); break;
                default:                   integ.reset(new VCM(scene, cam, scheduler, s)); break;
            }
            integ->preprocess(); pixel_sizes[0] = integ->pixel_size();
            for (int i = 0; i < iterations; ++i) integ->render(img);
            stats[0] = scheduler.total_prim_rays(); stats[1] = scheduler.total_shadow_rays();   // camera pass only (the light pass has its own TileScheduler)
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "ref_render_plugin: %s\n", e.what());
        return -6;
    }
    for (int i = 0; i < w * hgt; ++i) { out_rgb[3 * i] = img.pixels()[i][0]; out_rgb[3 * i + 1] = img.pixels()[i][1]; out_rgb[3 * i + 2] = img.pixels()[i][2]; }
    return 0;
}
#endif

/// Camera rays exactly as PathTracer::render's sample function makes them (pt.cpp:168-177) for pixel
/// centres + given jitter; rays[i] for pixel (x = i % w, y = i / w).
int ref_camera_rays(void* h, int w, int hgt, float fov, const float* jitter_xy, imba_ray* rays) {
    auto rs = (RefScene*)h;
    PerspectiveCamera cam(w, hgt, fov);
    setup_camera(cam, rs->cam_pos, rs->cam_dir, rs->cam_up);
    for (int y = 0; y < hgt; ++y) for (int x = 0; x < w; ++x) {
        const int i = y * w + x;
        const float sx = (float)x + (jitter_xy ? jitter_xy[2 * i] : 0.5f), sy = (float)y + (jitter_xy ? jitter_xy[2 * i + 1] : 0.5f);
        ::Ray r = cam.generate_ray(sx, sy);
        memcpy(&rays[i], &r, sizeof(r));
    }
    return 0;
}

/// PathTracer::process_primary_rays (pt.cpp:82-160) on an injected batch.
/// states: PTState[count] (40 B). Outputs: the compacted continuation queue and the shadow queue.
/// fb_rgb (w*h*3, may be null) receives add_contribution()s.
int ref_pt_shade_batch(void* h, int w, int hgt, float fov, int max_path_len, int count, const imba_ray* rays,
                       const imba_hit* hits, const void* states, imba_ray* rays_out, void* states_out, int* out_count,
                       imba_ray* shadow_rays, void* shadow_states, int* shadow_count, float* fb_rgb) {
    Quiet q;
    static_assert(sizeof(PTState) == 40 && sizeof(ShadowState) == 16, "state sizes");
    auto rs = (RefScene*)h;
    Scene& scene = *rs->scene;
    PerspectiveCamera cam(w, hgt, fov);
    setup_camera(cam, rs->cam_pos, rs->cam_dir, rs->cam_up);
    BatchScheduler<PTState, ShadowState> sched(scene);
    RayQueue<PTState> prim(count, false); RayQueue<ShadowState> shadow(count, false);
    sched.prim = &prim; sched.shadow = &shadow;
    const PTState* st = (const PTState*)states;
    for (int i = 0; i < count; ++i) { ::Ray r; memcpy(&r, &rays[i], 32); prim.push(r, st[i]); }
    memcpy(prim.hits(), hits, sizeof(Hit) * count);
    PathTracer pt(scene, cam, sched, max_path_len);
    AtomicImage img(w, hgt); img.clear();
    pt.render(img);
    *out_count = prim.size(); *shadow_count = shadow.size();
    memcpy(rays_out, prim.rays(), 32 * prim.size()); memcpy(states_out, prim.states(), sizeof(PTState) * prim.size());
    memcpy(shadow_rays, shadow.rays(), 32 * shadow.size()); memcpy(shadow_states, shadow.states(), sizeof(ShadowState) * shadow.size());
    if (fb_rgb) for (int i = 0; i < w * hgt; ++i) { fb_rgb[3 * i] = img.pixels()[i][0]; fb_rgb[3 * i + 1] = img.pixels()[i][1]; fb_rgb[3 * i + 2] = img.pixels()[i][2]; }
    return 0;
}

/// VCMIntegrator<algo>::process_light_rays (stage 0, vcm.cpp:193-258) or process_camera_rays (stage 1, vcm.cpp:323-439) of
/// the unmodified reference on an injected batch, with the iteration constants render() would have set.
int ref_vcm_stage(void* h, int algo, int w, int hgt, float fov, int max_path_len, int light_path_count, int num_connections, int num_knn, float base_radius,
                  int iteration, int stage, int count, const imba_ray* rays, const imba_hit* hits, const void* states, const void* cache_in, int cache_in_count,
                  imba_ray* rays_out, void* states_out, int* out_count, imba_ray* shadow_rays, void* shadow_states, int* shadow_count, float* fb_rgb,
                  void* cache_out, int* cache_out_count) {
    Quiet q;
    auto rs = (RefScene*)h; Scene& scene = *rs->scene;
    UserSettings s;
    s.width = w; s.height = hgt; s.fov = fov; s.concurrent_spp = 1; s.thread_count = 1; s.tile_size = 64; s.max_path_len = max_path_len;
    s.light_path_count = light_path_count; s.num_connections = num_connections; s.num_knn = num_knn; s.radius_factor = 2.0f;
    s.traversal_platform = UserSettings::cpu; s.algorithm = (UserSettings::Algorithm)algo;
    PerspectiveCamera cam(w, hgt, fov);
    setup_camera(cam, rs->cam_pos, rs->cam_dir, rs->cam_up);
#define STAGE(T) return run_vcm_stage<T>(scene, cam, s, base_radius, iteration, stage, count, rays, hits, states, (const ExportVertex*)cache_in, cache_in_count, rays_out, \
                                         states_out, out_count, shadow_rays, shadow_states, shadow_count, fb_rgb, (ExportVertex*)cache_out, cache_out_count)
    switch (s.algorithm) {
        case UserSettings::BPT: STAGE(BPT); case UserSettings::PPM: STAGE(PPM); case UserSettings::LT: STAGE(LT);
        case UserSettings::VCM_PT: STAGE(VCM_PT); default: STAGE(VCM);
    }
#undef STAGE
}

/// calculate_intersection (integrator.h:71-123) on a batch; out: 20 floats per hit
/// (pos3, out_dir3, normal3, uv2, geom_normal3, u_tangent3, v_tangent3) + material index.
int ref_calc_intersections(void* h, int count, const imba_ray* rays, const imba_hit* hits, float* out20, int* mat_out) {
    auto rs = (RefScene*)h; Scene& scene = *rs->scene;
    for (int i = 0; i < count; ++i) {
        Hit hit; memcpy(&hit, &hits[i], 16); ::Ray r; memcpy(&r, &rays[i], 32);
        if (hit.tri_id < 0) { mat_out[i] = -1; continue; }
        Intersection is = calculate_intersection(scene, hit, r);
        float* o = out20 + 20 * i;
        o[0] = is.pos.x; o[1] = is.pos.y; o[2] = is.pos.z; o[3] = is.out_dir.x; o[4] = is.out_dir.y; o[5] = is.out_dir.z;
        o[6] = is.normal.x; o[7] = is.normal.y; o[8] = is.normal.z; o[9] = is.uv.x; o[10] = is.uv.y;
        o[11] = is.geom_normal.x; o[12] = is.geom_normal.y; o[13] = is.geom_normal.z;
        o[14] = is.u_tangent.x; o[15] = is.u_tangent.y; o[16] = is.u_tangent.z; o[17] = is.v_tangent.x; o[18] = is.v_tangent.y; o[19] = is.v_tangent.z;
        int m = -1; for (size_t k = 0; k < scene.material_count(); ++k) if (scene.material(k).get() == is.mat) { m = (int)k; break; }
        mat_out[i] = m;
    }
    return 0;
}

/// TEST HOOK: gives the (texture-less) Cornell scenes procedural textures, bump maps and texture coordinates so that the
/// reference's TextureSampler::sample (texture_sampler.h:18-46), Material::bump (materials.h:34-55) and the textured
/// branches of DiffuseMaterial / GlossyMaterial::get_bsdf run in the parity harness.  Every non-emissive diffuse material
/// gets a colour texture, every second one a bump map, glossy materials a diffuse texture; vertices get seeded texcoords.
int ref_scene_add_test_textures(void* h, uint64_t seed) {
    auto rs = (RefScene*)h; Scene& sc = *rs->scene;
    RNG rng(seed);
    auto make_tex = [&](int w, int hgt, bool smooth) -> TextureSampler* {
        Image img(w, hgt);
        for (int i = 0; i < w * hgt; ++i) {
            const float a = rng.random_float(), b = rng.random_float(), c = rng.random_float();
            img.pixels()[i] = smooth ? rgba(0.5f + 0.5f * sinf(0.7f * i), a * 0.2f, b, 1.0f) : rgba(0.2f + 0.8f * a, 0.2f + 0.8f * b, 0.2f + 0.8f * c, 1.0f);
        }
        sc.textures().emplace_back(new TextureSampler(std::move(img)));
        return sc.textures().back().get();
    };
    int changed = 0;
    for (size_t m = 0; m < sc.material_count(); ++m) {
        Material* mat = sc.material(m).get();
        if (mat->emitter()) continue;
        if (dynamic_cast<DiffuseMaterial*>(mat)) {
            TextureSampler* col = make_tex(7 + (int)m, 5 + (int)m % 3, false);
            TextureSampler* bump = (m % 2 == 0) ? make_tex(16, 9, true) : nullptr;
            sc.materials()[m].reset(new DiffuseMaterial(col, bump)); ++changed;
        } else if (auto g = dynamic_cast<GlossyMaterial*>(mat)) {
            TextureSampler* col = make_tex(11, 13, false);
            sc.materials()[m].reset(new GlossyMaterial(g->exponent_, g->specular_color_, col, make_tex(8, 8, true))); ++changed;
        }
    }
    for (size_t k = 0; k < sc.mesh_count(); ++k) {
        Mesh& mesh = sc.meshes()[k];
        auto tc = mesh.attribute<float2>(MeshAttributes::TEXCOORDS);
        for (size_t v = 0; v < mesh.vertex_count(); ++v) { tc[v].x = 4.0f * rng.random_float() - 1.5f; tc[v].y = 4.0f * rng.random_float() - 1.5f; }
    }
    return changed;
}

/// Replaces the non-emissive diffuse / glossy materials alternately by a PhongTestMaterial (coefficient from its colour, exponent
/// 4 .. 67) and an OrenNayarTestMaterial (reflectance from its colour, roughness 5 .. 44 degrees).
int ref_scene_add_test_phong(void* h, uint64_t seed) {
    auto rs = (RefScene*)h; Scene& sc = *rs->scene;
    RNG rng(seed);
    int changed = 0;
    for (size_t m = 0; m < sc.material_count(); ++m) {
        Material* mat = sc.material(m).get();
        if (mat->emitter()) continue;
        rgb c(0.0f);
        if (auto d = dynamic_cast<DiffuseMaterial*>(mat)) c = d->sampler_ ? rgb(0.6f, 0.5f, 0.4f) : d->color_;
        else if (auto g = dynamic_cast<GlossyMaterial*>(mat)) c = g->specular_color_;
        else continue;
        const float r = rng.random_float();
        if (m % 2 == 0) sc.materials()[m].reset(new PhongTestMaterial(0.25f * c + rgb(0.05f), 4.0f + floorf(64.0f * r)));
        else            sc.materials()[m].reset(new OrenNayarTestMaterial(0.8f * c + rgb(0.05f), 5.0f + floorf(40.0f * r)));
        ++changed;
    }
    return changed;
}

/// TEST HOOK: adds a procedural environment map + EnvLight (what build_scene.cpp:505-527 does for an `env` line; the only
/// scene of the reference that has one, Car, lacks its HDR file here).
int ref_scene_add_test_envmap(void* h, uint64_t seed, float intensity) {
    auto rs = (RefScene*)h; Scene& sc = *rs->scene;
    if (sc.env_map()) return 0;
    RNG rng(seed);
    const int w = 24, hgt = 12;
    Image img(w, hgt);
    for (int y = 0; y < hgt; ++y) for (int x = 0; x < w; ++x)
        img(x, y) = rgba(0.3f + 0.7f * rng.random_float(), 0.4f + 0.5f * (float)y / hgt, 0.2f + 0.8f * (float)x / w, 1.0f);
    sc.set_env_map(new EnvMap(img, intensity, sc.bounding_sphere()));
    sc.lights().emplace_back(new EnvLight(sc.env_map(), sc.bounding_sphere()));
    return 1;
}


/// Start of the light paths: the sample function VCMIntegrator::trace_light_paths hands to its ray generator (vcm.cpp:64-96),
/// i.e. Light::sample_emit of the reference's lights (light.h) + the initial partial MIS weights, for given (light, RNG state)
/// pairs -- the seeding itself (ray_gen.h:122-152) is covered by ref_seed_states.  The few assignments of that lambda are
/// restated here because it cannot be called outside run_iteration; all arithmetic is the reference's sample_emit.
int ref_vcm_light_emit(void* h, int count, const int* light_ids, const uint64_t* rng_in, float mis_eta_vc, imba_ray* rays_out, void* states_out) {
    auto rs = (RefScene*)h; Scene& scene = *rs->scene;
    VCMState* out = (VCMState*)states_out;
    const float pdf_lightpick = 1.0f / scene.light_count();
    for (int i = 0; i < count; ++i) {
        if (light_ids[i] < 0 || light_ids[i] >= (int)scene.light_count()) return -1;
        VCMState st; memset(&st, 0, sizeof(st));
        st.ray_id = i; st.light_id = light_ids[i]; st.rng.state_ = rng_in[i];
        auto& l = scene.light(light_ids[i]);
        Light::EmitSample sample = l->sample_emit(st.rng);
        ::Ray r;
        r.org.x = sample.pos.x; r.org.y = sample.pos.y; r.org.z = sample.pos.z; r.org.w = 1e-3f;
        r.dir.x = sample.dir.x; r.dir.y = sample.dir.y; r.dir.z = sample.dir.z; r.dir.w = FLT_MAX;
        st.throughput = sample.radiance / pdf_lightpick;
        st.path_length = 1;
        st.dVCM = sample.pdf_direct_a / sample.pdf_emit_w;
        st.dVC = l->is_delta() ? 0.0f : sample.cos_out / (sample.pdf_emit_w * pdf_lightpick);
        st.dVM = st.dVC * mis_eta_vc;
        st.finite_light = l->is_finite();
        memcpy(&rays_out[i], &r, 32); out[i] = st;
    }
    return 0;
}

/// Start of the camera paths: the sample function of VCMIntegrator::trace_camera_paths (vcm.cpp:109-129) for given pixels and
/// RNG states -- the reference's camera (camera.h:57-65) and the conversion of the image-plane pdf that initialises dVCM.
int ref_vcm_camera_start(void* h, int w, int hgt, float fov, int light_path_count, int count, const int* pixel_ids, const uint64_t* rng_in,
                         imba_ray* rays_out, void* states_out) {
    auto rs = (RefScene*)h;
    PerspectiveCamera cam(w, hgt, fov);
    setup_camera(cam, rs->cam_pos, rs->cam_dir, rs->cam_up);
    VCMState* out = (VCMState*)states_out;
    for (int i = 0; i < count; ++i) {
        VCMState st; memset(&st, 0, sizeof(st));
        st.pixel_id = pixel_ids[i]; st.rng.state_ = rng_in[i];
        const int x = pixel_ids[i] % w, y = pixel_ids[i] / w;
        const float sample_x = static_cast<float>(x) + st.rng.random_float();
        const float sample_y = static_cast<float>(y) + st.rng.random_float();
        ::Ray r = cam.generate_ray(sample_x, sample_y);
        st.throughput = rgb(1.0f);
        st.path_length = 1;
        const float3 dir(r.dir.x, r.dir.y, r.dir.z);
        const float cos_theta_o = dot(dir, cam.dir());
        const float pdf_cam_w = sqr(cam.image_plane_dist() / cos_theta_o) / cos_theta_o;
        st.dVC = 0.0f; st.dVM = 0.0f; st.dVCM = light_path_count / pdf_cam_w;
        memcpy(&rays_out[i], &r, 32); out[i] = st;
    }
    return 0;
}

/// RNG known-answer helper (render/random.h:15-47).
void ref_rng_floats(uint64_t seed, int discard, int n, float* out, uint64_t* state_out) {
    RNG rng(seed); rng.discard(discard);
    for (int i = 0; i < n; ++i) out[i] = rng.random_float();
    if (state_out) *state_out = rng.state_;
}

/// Seeding of primary-ray states (ray_gen.h:57-72) for a given seed_base: returns RNG state after discard.
void ref_seed_states(uint64_t seed_base, int first, int count, int spp, uint64_t* states_out) {
    for (int i = first; i < first + count; ++i) {
        int pixel_idx = i / spp;
        int seed = seed_base;
        seed = 33 * seed ^ i; seed = 33 * seed ^ i; seed = 33 * seed ^ i; seed = 33 * seed ^ i;
        RNG rng(seed);
        rng.discard((seed % 5) + 16 + pixel_idx % 5);
        states_out[i - first] = rng.state_;
    }
}

int ref_struct_sizes(int* out) {
    out[0] = sizeof(::Ray); out[1] = sizeof(Hit); out[2] = sizeof(RayState); out[3] = sizeof(ShadowState); out[4] = sizeof(PTState);
    out[5] = sizeof(VCMState); out[6] = sizeof(VCMShadowState); out[7] = sizeof(Intersection); out[8] = sizeof(LightPathVertex);
    out[9] = sizeof(VCMPhoton); out[10] = sizeof(traversal_cpu::Node); out[11] = sizeof(traversal_gpu::Node); out[12] = sizeof(InstanceNode);
    return 13;
}

} // extern "C"
