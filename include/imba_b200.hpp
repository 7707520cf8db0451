/* imba_b200.hpp -- C++ host-side mirror of the reference's plug-in points over the C ABI (imba_b200.h).
 *
 * The reference is a C++ program; these header-only classes carry the names, argument meaning and error behaviour of
 * its own interfaces for the hot path so that frontend code written against them keeps its shape:
 *
 *   imba_b200::RayQueue<State>   render/scheduling/ray_queue.h:58-360   push / clear / shrink / size / rays / hits /
 *                                states / traverse_gpu / traverse_occluded_gpu; the traversal runs on the B200
 *   imba_b200::Integrator        render/integrators/integrator.h:19-39  render / reset / preprocess / pixel_size
 *   imba_b200::PathTracer, VCM, BPT, PPM, LT, VCM_PT   pt.h:18-39, vcm.h:144-148
 *
 * Differences that are part of the design: the integrators keep their queues and the accumulation buffer
 * in device memory -- render() ADDS one iteration of un-normalised radiance (as the reference does,
 * render_window.cpp:105,178) into the device buffer, download() copies it out; there is no RayScheduler argument because
 * one wavefront covers the whole image.  Errors of the C ABI are thrown as std::runtime_error.
 */
#ifndef IMBA_B200_HPP
#define IMBA_B200_HPP
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "imba_b200.h"

namespace imba_b200 {

inline void check(int rc, const char* what) {
    if (rc != IMBA_OK) throw std::runtime_error(std::string(what) + ": " + imba_last_error());
}

/// The flattened scene on the current CUDA device (replaces Scene::upload_*, render/scene.cpp:132-188).
class Scene {
public:
    explicit Scene(const imba_scene_desc& desc) : h_(imba_scene_create(&desc)) {
        if (!h_) throw std::runtime_error(std::string("imba_scene_create: ") + imba_last_error());
    }
    ~Scene() { imba_scene_destroy(h_); }
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;
    imba_scene* handle() const { return h_; }
private:
    imba_scene* h_;
};

/// Host-side ray queue with the reference's interface; traversal is delegated to the device.
template <typename StateType>
class RayQueue {
public:
    explicit RayQueue(int capacity) : last_(-1) {
        // traversal is called with the count rounded up to 64, padding rays are zero (ray_queue.h:59-61,74,273)
        const size_t n = (size_t)((capacity + 63) / 64) * 64;
        rays_.assign(n, imba_ray()); hits_.resize(n); states_.resize(n);
        std::memset(rays_.data(), 0, n * sizeof(imba_ray));
        pinned_ = imba_host_register(rays_.data(), n * sizeof(imba_ray)) == IMBA_OK &&
                  imba_host_register(hits_.data(), n * sizeof(imba_hit)) == IMBA_OK;
    }
    ~RayQueue() { if (pinned_) { imba_host_unregister(rays_.data()); imba_host_unregister(hits_.data()); } }
    RayQueue(const RayQueue&) = delete;
    RayQueue& operator=(const RayQueue&) = delete;

    int size() const { return last_ + 1; }
    int capacity() const { return (int)states_.size(); }
    void shrink(int size) { last_ = size - 1; }
    void clear() { last_ = -1; }
    imba_ray* rays() { return rays_.data(); }
    imba_hit* hits() { return hits_.data(); }
    StateType* states() { return states_.data(); }
    void push(const imba_ray& ray, const StateType& state) {
        const int id = ++last_;
        assert(id < (int)rays_.size() && "ray queue full");
        rays_[id] = ray; states_[id] = state;
    }
    /// RayQueue::traverse_gpu (ray_queue.h:269-293): closest hit for every ray of the queue.
    void traverse_gpu(const Scene& scene) { check(imba_trace_host(scene.handle(), rays_.data(), hits_.data(), size(), 0), "traverse_gpu"); }
    /// RayQueue::traverse_occluded_gpu (ray_queue.h:319-343): consumers read hits()[i].tri_id < 0 only.
    void traverse_occluded_gpu(const Scene& scene) { check(imba_trace_host(scene.handle(), rays_.data(), hits_.data(), size(), 1), "traverse_occluded_gpu"); }
private:
    std::vector<imba_ray> rays_; std::vector<imba_hit> hits_; std::vector<StateType> states_;
    int last_; bool pinned_;
};

/// Integrator::{render, reset, preprocess, pixel_size} (integrator.h:19-39) on the device.
class Integrator {
public:
    Integrator(Scene& scene, const imba_render_settings& settings) : settings_(settings), h_(imba_integrator_create(scene.handle(), &settings)) {
        if (!h_) throw std::runtime_error(std::string("imba_integrator_create: ") + imba_last_error());
    }
    virtual ~Integrator() { imba_integrator_destroy(h_); }
    Integrator(const Integrator&) = delete;
    Integrator& operator=(const Integrator&) = delete;

    /// Adds `iterations` progressive iterations to the device accumulation buffer (asynchronous).
    virtual void render(int iterations = 1) { check(imba_integrator_render(h_, iterations), "render"); }
    virtual void reset() { check(imba_integrator_reset(h_), "reset"); }
    virtual void preprocess() { check(imba_integrator_preprocess(h_), "preprocess"); }
    float pixel_size() const { return imba_integrator_pixel_size(h_); }
    /// Copies the accumulation buffer (width * height * 3 floats, a SUM over iterations) to the host.
    void download(float* rgb) { check(imba_integrator_download(h_, rgb), "download"); }
    void clear() { check(imba_integrator_clear(h_), "clear"); }
    float* device_framebuffer() { return imba_integrator_framebuffer(h_); }
    imba_integrator* handle() const { return h_; }
    const imba_render_settings& settings() const { return settings_; }
protected:
    imba_render_settings settings_;
    imba_integrator* h_;
};

template <int ALGO>
class AlgoIntegrator : public Integrator {
public:
    AlgoIntegrator(Scene& scene, imba_render_settings settings) : Integrator(scene, with_algo(settings)) {}
private:
    static imba_render_settings with_algo(imba_render_settings s) { s.algorithm = ALGO; return s; }
};
typedef AlgoIntegrator<IMBA_ALGO_PT>     PathTracer;
typedef AlgoIntegrator<IMBA_ALGO_BPT>    BPT;
typedef AlgoIntegrator<IMBA_ALGO_VCM>    VCM;
typedef AlgoIntegrator<IMBA_ALGO_PPM>    PPM;
typedef AlgoIntegrator<IMBA_ALGO_VCM_PT> VCM_PT;
typedef AlgoIntegrator<IMBA_ALGO_LT>     LT;

/// UserSettings defaults (frontend/cmd_line.h:61-79).
inline imba_render_settings default_settings(int width = 512, int height = 512) {
    imba_render_settings s; std::memset(&s, 0, sizeof(s));
    s.algorithm = IMBA_ALGO_PT; s.width = width; s.height = height; s.fov = 60.0f;
    s.cam_dir[2] = 1.0f; s.cam_up[1] = 1.0f;
    s.spp = 1; s.max_path_len = 10; s.light_path_count = width * height / 2; s.num_connections = 1; s.num_knn = 10;
    s.radius_factor = 2.0f; s.seed = 1; s.rank = 0; s.rank_count = 1;
    return s;
}

} // namespace imba_b200
#endif
