/* imba_b200.hpp -- C++ host-side mirror of the reference's plug-in points over the C ABI (imba_b200.h).
 *
 * The reference is a C++ program; these header-only classes carry the names, argument meaning and error behaviour of
 * its own interfaces for the hot path so that frontend code written against them keeps its shape:
 *
 *   imba_b200::RayQueue<State>   render/scheduling/ray_queue.h:58-360   atomic push / range push / append / clear / shrink / size /
 *                                rays / hits / states / ray(i) / hit(i) / state(i) / compact_hits / compact_rays /
 *                                sort_by_material / traverse_{gpu,cpu} / traverse_occluded_{gpu,cpu}; compaction, sorting and
 *                                traversal run on the B200
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
#include <algorithm>
#include <atomic>
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

/// Host-side ray queue with the reference's interface (ray_queue.h:58-360): storage in host memory, thread-safe push / append,
/// the sorted_indices_ indirection, and compact_hits / compact_rays / sort_by_material / traverse_* -- which the reference runs
/// as serial or TBB loops on the CPU -- delegated to the device through the C ABI.  StateType starts with `int pixel_id`
/// (RayState, ray_queue.h:23-31) and its size is a multiple of 8.
template <typename StateType>
class RayQueue {
    static int align(int v) { return (v + 63) / 64 * 64; }   // ray_queue.h:59-61: max(align_cpu 8, align_gpu 64)
public:
    RayQueue() : last_(-1), pinned_(false) {}
    /// `gpu_buffers` is accepted for source compatibility (ray_queue.h:66); the staging buffers belong to the library.
    explicit RayQueue(int capacity, bool gpu_buffers = true) : last_(-1), pinned_(false) {
        (void)gpu_buffers;
        // traversal is called with the count rounded up to 64, padding rays are zero (ray_queue.h:59-61,74,273)
        const size_t n = (size_t)align(capacity);
        rays_.assign(n, imba_ray()); hits_.resize(n); states_.resize(n); sorted_indices_.resize(n);
        std::memset(rays_.data(), 0, n * sizeof(imba_ray));
        if (imba_host_register(rays_.data(), n * sizeof(imba_ray)) == IMBA_OK) {
            if (imba_host_register(hits_.data(), n * sizeof(imba_hit)) == IMBA_OK) pinned_ = true;
            else imba_host_unregister(rays_.data());
        }
    }
    ~RayQueue() { if (pinned_) { imba_host_unregister(rays_.data()); imba_host_unregister(hits_.data()); } }
    RayQueue(const RayQueue&) = delete;
    RayQueue& operator=(const RayQueue&) = delete;

    int size() const { return last_.load() + 1; }
    int capacity() const { return (int)states_.size(); }
    void shrink(int size) { last_ = size - 1; }
    void clear() { last_ = -1; }
    imba_ray* rays() { return rays_.data(); }
    imba_hit* hits() { return hits_.data(); }
    StateType* states() { return states_.data(); }
    /// Access through the order sort_by_material / compact_hits established (ray_queue.h:114-116).
    imba_ray& ray(int idx) { return rays_[sorted_indices_[idx]]; }
    imba_hit& hit(int idx) { return hits_[sorted_indices_[idx]]; }
    StateType& state(int idx) { return states_[sorted_indices_[idx]]; }

    /// Adds a single secondary or shadow ray.  Thread-safe (ray_queue.h:123-131).
    void push(const imba_ray& ray, const StateType& state) {
        const int id = ++last_;                      // atomic increment
        assert(id < (int)rays_.size() && "ray queue full");
        rays_[id] = ray; states_[id] = state;
    }
    /// Adds a set of camera rays.  Thread-safe (ray_queue.h:134-145).
    template <typename RayIter, typename StateIter>
    void push(RayIter rays_begin, RayIter rays_end, StateIter states_begin, StateIter states_end) {
        const int count = (int)(rays_end - rays_begin);
        const int end_idx = (last_ += count);        // atomic add
        const int start_idx = end_idx - (count - 1);
        assert(end_idx < (int)rays_.size() && "ray queue full");
        std::copy(rays_begin, rays_end, rays_.begin() + start_idx);
        std::copy(states_begin, states_end, states_.begin() + start_idx);
    }
    /// Appends the rays and states of another queue; hits are not copied (ray_queue.h:148-158; the reference copies the other
    /// queue's whole capacity, this copies its size() live entries).
    void append(const RayQueue<StateType>& other) {
        const int count = other.size();
        const int end_idx = (last_ += count);
        const int start_idx = end_idx - (count - 1);
        assert(end_idx < (int)rays_.size() && "ray queue full");
        std::copy(other.rays_.begin(), other.rays_.begin() + count, rays_.begin() + start_idx);
        std::copy(other.states_.begin(), other.states_.begin() + count, states_.begin() + start_idx);
    }

    /// Moves all rays that hit something (with their states and hits) to the front, resets the index order and returns the
    /// number of hits (ray_queue.h:161-185).
    int compact_hits() {
        int hit_count = 0;
        check(imba_queue_compact_hits_host(rays_.data(), hits_.data(), states_.data(), (int)sizeof(StateType), size(), &hit_count), "compact_hits");
        for (int i = 0, n = size(); i < n; ++i) sorted_indices_[i] = i;
        return hit_count;
    }
    /// Moves all continued rays (pixel_id >= 0) to the front and shrinks the queue; hits are not moved (ray_queue.h:188-207).
    void compact_rays() {
        int kept = 0;
        check(imba_queue_compact_rays_host(rays_.data(), states_.data(), (int)sizeof(StateType), size(), &kept), "compact_rays");
        shrink(kept);
    }
    /// Groups the first `count` entries by material: afterwards ray(i) / hit(i) / state(i) enumerate them material by material
    /// (ray_queue.h:211-244).  Like the reference it leaves the material id in rays()[i].dir.w.
    template <typename GetMatIDFn>
    void sort_by_material(GetMatIDFn get_mat_id, int num_mats, int count) {
        keys_.resize((size_t)count);
        for (int i = 0; i < count; ++i) {
            const int m = get_mat_id(hits_[i]);
            keys_[i] = m;
            std::memcpy(&rays_[i].dir.w, &m, sizeof(int));
        }
        check(imba_queue_sort_keys_host(keys_.data(), count, num_mats, sorted_indices_.data()), "sort_by_material");
    }

    /// RayQueue::traverse_gpu / traverse_cpu (ray_queue.h:247-293): closest hit for every ray of the queue.
    void traverse_gpu(const Scene& scene) { check(imba_trace_host(scene.handle(), rays_.data(), hits_.data(), size(), 0), "traverse_gpu"); }
    void traverse_cpu(const Scene& scene) { traverse_gpu(scene); }   // there is one traversal path, and it is the device's
    /// RayQueue::traverse_occluded_gpu / _cpu (ray_queue.h:296-343): consumers read hits()[i].tri_id < 0 only.
    void traverse_occluded_gpu(const Scene& scene) { check(imba_trace_host(scene.handle(), rays_.data(), hits_.data(), size(), 1), "traverse_occluded_gpu"); }
    void traverse_occluded_cpu(const Scene& scene) { traverse_occluded_gpu(scene); }
private:
    std::vector<imba_ray> rays_; std::vector<imba_hit> hits_; std::vector<StateType> states_;
    std::vector<int> sorted_indices_, keys_;
    std::atomic<int> last_; bool pinned_;
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
