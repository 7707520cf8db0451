/* imba_b200_ref.hpp -- the reference's OWN plug-in points implemented over libimba_b200.so.
 *
 * Unlike imba_b200.hpp (a stand-alone mirror in its own namespace) this header is compiled INSIDE the reference's tree: it
 * includes the reference's headers and derives from its classes, so that frontend/main.cpp:117-163 can construct these
 * objects and hand them to RenderWindow (frontend/render_window.cpp:50-95) unchanged:
 *
 *   imba::B200Integrator            : public imba::Integrator (render/integrators/integrator.h:18-39)
 *       render(AtomicImage&) runs one progressive iteration of PT / BPT / VCM / PPM / LT / VCM_PT entirely on the device
 *       (imba_integrator_render) and ADDS its un-normalised radiance into `out`, as every reference integrator does and
 *       render_window.cpp:105,178 expects; reset() follows the camera; preprocess() runs the reference's own pixel-size probe
 *       (Integrator::preprocess -> estimate_pixel_size, integrator.cpp:14-71, through the traversal symbols this library
 *       exports) and the device integrator's.
 *   imba::B200WavefrontScheduler    : public imba::RayScheduler<State, ShadowState> (render/scheduling/ray_scheduler.h:15-38)
 *       run_iteration() drives ONE wavefront over all samples of the frame instead of 256^2 tiles on CPU threads
 *       (tile_scheduler.h:100-144): fill -> closest-hit traversal on the B200 -> the integrator's process_primary_rays ->
 *       occlusion traversal on the B200 -> process_shadow_rays.  The shading callbacks stay the reference's CPU code -- this
 *       is the plug-in point for a host that wants its own integrator but the device's traversal.
 *
 * The flattened scene (imba_scene_desc) comes from the host program: INTEGRATION.md shows the export against imba::Scene,
 * oracle/ref_driver.cpp (ref_scene_export) is a compiled example.
 */
#ifndef IMBA_B200_REF_HPP
#define IMBA_B200_REF_HPP
#include <atomic>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "imbatracer/frontend/cmd_line.h"
#include "imbatracer/render/integrators/integrator.h"
#include "imbatracer/render/scheduling/ray_scheduler.h"

#ifndef IMBA_B200_NO_L1_DECLS
#define IMBA_B200_NO_L1_DECLS   /* the reference's traversal_interface.h declares the four entry points inside its namespaces */
#endif
#include "imba_b200.h"

namespace imba {

/// UserSettings (frontend/cmd_line.h:13-80) + the camera of the .scene file -> imba_render_settings.
inline imba_render_settings b200_render_settings(const UserSettings& s, const float3& cam_pos, const float3& cam_dir, const float3& cam_up, unsigned seed = 1) {
    imba_render_settings r; std::memset(&r, 0, sizeof(r));
    switch (s.algorithm) {
        case UserSettings::PT:     r.algorithm = IMBA_ALGO_PT; break;
        case UserSettings::BPT:    r.algorithm = IMBA_ALGO_BPT; break;
        case UserSettings::PPM:    r.algorithm = IMBA_ALGO_PPM; break;
        case UserSettings::LT:     r.algorithm = IMBA_ALGO_LT; break;
        case UserSettings::VCM_PT: r.algorithm = IMBA_ALGO_VCM_PT; break;
        default:                   r.algorithm = IMBA_ALGO_VCM; break;
    }
    r.width = (int)s.width; r.height = (int)s.height; r.fov = s.fov;
    r.cam_pos[0] = cam_pos.x; r.cam_pos[1] = cam_pos.y; r.cam_pos[2] = cam_pos.z;
    r.cam_dir[0] = cam_dir.x; r.cam_dir[1] = cam_dir.y; r.cam_dir[2] = cam_dir.z;
    r.cam_up[0] = cam_up.x; r.cam_up[1] = cam_up.y; r.cam_up[2] = cam_up.z;
    r.spp = (int)s.concurrent_spp; r.max_path_len = (int)s.max_path_len;
    r.light_path_count = (int)s.light_path_count; r.num_connections = (int)s.num_connections; r.num_knn = (int)s.num_knn;
    r.radius_factor = s.radius_factor;
    r.seed = seed; r.rank = 0; r.rank_count = 1; r.iterations_per_pass = 1;
    return r;
}

/// Any of the reference's integrators, on the device, behind the reference's Integrator interface.
class B200Integrator : public Integrator {
public:
    /// `device_scene` = imba_scene_create() of the flattened scene; it must outlive the integrator.
    B200Integrator(Scene& scene, PerspectiveCamera& cam, imba_scene* device_scene, const imba_render_settings& settings)
        : Integrator(scene, cam), settings_(settings), h_(imba_integrator_create(device_scene, &settings)),
          staging_((size_t)settings.width * settings.height * 3)
    {
        if (!h_) throw std::runtime_error(std::string("imba_integrator_create: ") + imba_last_error());
    }
    ~B200Integrator() override { imba_integrator_destroy(h_); }
    B200Integrator(const B200Integrator&) = delete;
    B200Integrator& operator=(const B200Integrator&) = delete;

    /// One progressive iteration; its radiance is ADDED to `out` (un-normalised: the caller divides by frames * spp).
    void render(AtomicImage& out) override {
        check(imba_integrator_clear(h_), "clear");
        check(imba_integrator_render(h_, 1), "render");
        check(imba_integrator_download(h_, staging_.data()), "download");
        uint64_t rays[2];                            // (imba_integrator_clear also restarts the device's ray counters)
        check(imba_integrator_ray_counts(h_, rays), "ray_counts");
        total_rays_[0] += rays[0]; total_rays_[1] += rays[1];
        const int n = settings_.width * settings_.height;
        for (int i = 0; i < n; ++i)
            out.pixels()[i].apply<std::plus<float> >(rgb(staging_[3 * i], staging_[3 * i + 1], staging_[3 * i + 2]));
    }

    /// Camera moved (render_window.cpp:97-101): follow cam_ and restart the progressive radius schedule (vcm.h:69-72).
    void reset() override {
        const float3 p = cam_.pos(), d = cam_.dir();
        const float pos[3] = { p.x, p.y, p.z }, dir[3] = { d.x, d.y, d.z };
        check(imba_integrator_set_camera(h_, pos, dir, settings_.cam_up), "set_camera");
        for (int k = 0; k < 3; ++k) { settings_.cam_pos[k] = pos[k]; settings_.cam_dir[k] = dir[k]; }
    }

    /// The reference's probe (sets pixel_size()) and the device integrator's (pixel size + light-vertex cache sizing).
    void preprocess() override {
        Integrator::preprocess();
        check(imba_integrator_preprocess(h_), "preprocess");
    }

    float device_pixel_size() const { return imba_integrator_pixel_size(h_); }
    /// Primary and shadow rays of all render() calls so far, like TileScheduler's counters (tile_scheduler.h:59-60,117-127).
    void ray_counts(uint64_t out[2]) const { out[0] = total_rays_[0]; out[1] = total_rays_[1]; }
    imba_integrator* handle() const { return h_; }

private:
    static void check(int rc, const char* what) {
        if (rc != IMBA_OK) throw std::runtime_error(std::string("B200Integrator::") + what + ": " + imba_last_error());
    }
    imba_render_settings settings_;
    imba_integrator* h_;
    std::vector<float> staging_;
    uint64_t total_rays_[2] = { 0, 0 };
};

/// One wavefront per frame: a single primary queue over every sample RayGen produces, traversal on the device through the
/// reference's own RayQueue::traverse_* (i.e. the four entry points libimba_b200.so exports), shading by the callbacks.
template <typename StateType, typename ShadowStateType>
class B200WavefrontScheduler : public RayScheduler<StateType, ShadowStateType> {
    using BaseType = RayScheduler<StateType, ShadowStateType>;
    using SamplePixelFn = typename RayGen<StateType>::SamplePixelFn;
    using ProcessPrimaryFn = typename BaseType::ProcessPrimaryFn;
    using ProcessShadowFn = typename BaseType::ProcessShadowFn;
protected:
    using BaseType::scene_;
public:
    using BaseType::gpu_traversal;

    /// `ray_gen` enumerates the whole frame (e.g. PixelRayGen(width, height, spp)); `q_size` >= the rays it produces per frame.
    B200WavefrontScheduler(RayGen<StateType>& ray_gen, Scene& scene, int max_shadow_rays_per_hit, int q_size, bool gpu_traversal)
        : BaseType(scene, gpu_traversal), ray_gen_(ray_gen)
        , prim_q_(q_size, gpu_traversal), shadow_q_(q_size * max_shadow_rays_per_hit, gpu_traversal)
    {
        total_prim_rays_ = 0; total_shadow_rays_ = 0;
    }

    void run_iteration(AtomicImage& out, ProcessShadowFn process_shadow_rays, ProcessPrimaryFn process_primary_rays,
                       SamplePixelFn sample_fn) override {
        ray_gen_.start_frame();
        while (!ray_gen_.is_empty() || prim_q_.size() > 0) {
            ray_gen_.fill_queue(prim_q_, sample_fn);
            total_prim_rays_ += prim_q_.size();
            if (gpu_traversal) prim_q_.traverse_gpu(scene_.traversal_data_gpu());
            else               prim_q_.traverse_cpu(scene_.traversal_data_cpu());
            process_primary_rays(prim_q_, shadow_q_, out);
            if (shadow_q_.size() > 0) {
                total_shadow_rays_ += shadow_q_.size();
                if (gpu_traversal) shadow_q_.traverse_occluded_gpu(scene_.traversal_data_gpu());
                else               shadow_q_.traverse_occluded_cpu(scene_.traversal_data_cpu());
                process_shadow_rays(shadow_q_, out);
            }
            shadow_q_.clear();
        }
    }

    uint64_t total_prim_rays() const { return total_prim_rays_; }
    uint64_t total_shadow_rays() const { return total_shadow_rays_; }

private:
    RayGen<StateType>& ray_gen_;
    RayQueue<StateType> prim_q_;
    RayQueue<ShadowStateType> shadow_q_;
    std::atomic<uint64_t> total_prim_rays_, total_shadow_rays_;
};

} // namespace imba
#endif
