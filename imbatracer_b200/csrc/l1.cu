/* l1.cu -- symbol-for-symbol replacements of the four entry points the reference binds from the external
 * AnyDSL/traversal library (core/traversal_interface.h:4-10; call sites render/scheduling/ray_queue.h:254-265,
 * 279-290, 304-315, 329-340).  An unmodified imbatracer build links against libimba_b200.so instead of
 * TRAVERSAL_LIBRARY_{CPU,GPU} (src/imbatracer/CMakeLists.txt:76-87) and its traversal runs on the B200.
 *
 * The reference passes raw arrays without sizes.  The compressed 8-wide structure is built on the first call for a
 * given (nodes, tris, instances, root) and cached:
 *   - gpu variant: the arrays are device allocations; their extents come from cuMemGetAddressRange and the reference
 *     layout is copied back once for conversion;
 *   - cpu variant: the arrays are host memory; the converter follows the tree, array extents for the mask data are
 *     derived from the largest triangle / material / texcoord index.
 * No error channel exists in these signatures (ray_queue.h:248,270-271 assert): failures print and abort().
 */
#include "api_common.h"
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {

struct Entry {
    const void* nodes; const void* tris; const void* insts; int root; int layout; uint64_t fingerprint;
    imba_scene* scene;
};
std::mutex g_mutex;
std::vector<Entry> g_cache;

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "libimba_b200: %s: %s\n", what, imba_last_error());
    abort();
}

uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
GetRangeFn range_fn() {
    static GetRangeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &st) != cudaSuccess || !p) {
            imba_set_error("cuMemGetAddressRange unavailable"); die("driver entry point");
        }
        fn = (GetRangeFn)p;
    }
    return fn;
}

// Copies a whole device allocation, starting at `ptr`, to the host.
std::vector<unsigned char> fetch_allocation(const void* ptr) {
    unsigned long long base = 0; size_t size = 0;
    if (range_fn()(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0) { imba_set_error("pointer %p is not a device allocation", ptr); die("cuMemGetAddressRange"); }
    const size_t bytes = (size_t)(base + size - (unsigned long long)(uintptr_t)ptr);
    std::vector<unsigned char> h(bytes);
    if (cudaMemcpy(h.data(), ptr, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { imba_set_error("%s", cudaGetErrorString(cudaGetLastError())); die("device-to-host copy of the reference structure"); }
    return h;
}

imba_scene* new_scene() {
    imba_scene* s = new imba_scene;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { imba_set_error("no CUDA device (this library has no CPU path)"); die("init"); }
    cudaGetDevice(&s->device);
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    return s;
}

bool any_real_mask(const imba_mask* m, size_t count) {
    for (size_t i = 0; i < count; ++i) if (!(m[i].width == 1 && m[i].height == 1 && m[i].offset == 0)) return true;
    return false;
}

imba_scene* lookup(const void* nodes, const void* tris, const void* insts, int root, int layout, uint64_t fp) {
    for (const Entry& e : g_cache)
        if (e.nodes == nodes && e.tris == tris && e.insts == insts && e.root == root && e.layout == layout && e.fingerprint == fp) return e.scene;
    return nullptr;
}

imba_scene* scene_for_gpu(int root, imba_node2* nodes, imba_instance_node* instances, imba_vec4* tris, int* indices,
                          imba_vec2* texcoords, imba_mask* masks, char* mask_buffer) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (imba_scene* s = lookup(nodes, tris, instances, root, IMBA_ACCEL_BVH2, 0)) return s;
    const std::vector<unsigned char> hn = fetch_allocation(nodes), ht = fetch_allocation(tris), hi = fetch_allocation(instances);
    Bvh8 b; std::string err;
    if (!imba_build_bvh8(IMBA_ACCEL_BVH2, root, hn.data(), (const imba_vec4*)ht.data(), (const imba_instance_node*)hi.data(), b, err)) {
        imba_set_error("%s", err.c_str()); die("BVH conversion");
    }
    imba_scene* s = new_scene();
    if (imba_scene_upload_accel(s, b) != IMBA_OK) die("upload");
    if (masks && indices && texcoords && mask_buffer) {
        const std::vector<unsigned char> hm = fetch_allocation(masks);
        // the allocation may be larger than material_count descriptors; trailing bytes of a cudaMalloc are not
        // descriptors, so only look at as many as there are materials referenced by the index buffer
        const std::vector<unsigned char> hidx = fetch_allocation(indices);
        const size_t ntri = (size_t)(b.tris.size() - 1);
        int max_mat = -1;
        for (size_t t = 0; t < ntri && (t * 4 + 3) * 4 < hidx.size(); ++t) max_mat = std::max(max_mat, ((const int*)hidx.data())[t * 4 + 3]);
        const size_t nm = std::min((size_t)(max_mat + 1), hm.size() / sizeof(imba_mask));
        s->has_masks = any_real_mask((const imba_mask*)hm.data(), nm);
        s->owns_masks = false; // device arrays belong to the caller; passed per call
        s->d_mask_indices = indices; s->d_mask_texcoords = (float2*)texcoords; s->d_masks = (int*)masks; s->d_mask_buffer = (uint8_t*)mask_buffer;
    }
    g_cache.push_back(Entry{ nodes, tris, instances, root, IMBA_ACCEL_BVH2, 0, s });
    return s;
}

template <typename T> int upload_array(T** dst, const void* src, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (cudaMalloc((void**)dst, bytes) != cudaSuccess) return -1;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
}

imba_scene* scene_for_cpu(int root, imba_node4* nodes, imba_instance_node* instances, imba_vec4* tris, int* indices,
                          imba_vec2* texcoords, imba_mask* masks, char* mask_buffer) {
    // cheap staleness check: the root node and the first instance record change when the host rebuilds in place
    const uint64_t fp = fnv(instances, sizeof(imba_instance_node), fnv(nodes + root, sizeof(imba_node4)));
    std::lock_guard<std::mutex> lock(g_mutex);
    if (imba_scene* s = lookup(nodes, tris, instances, root, IMBA_ACCEL_MBVH4, fp)) return s;
    Bvh8 b; std::string err;
    if (!imba_build_bvh8(IMBA_ACCEL_MBVH4, root, nodes, tris, instances, b, err)) { imba_set_error("%s", err.c_str()); die("BVH conversion"); }
    imba_scene* s = new_scene();
    if (imba_scene_upload_accel(s, b) != IMBA_OK) die("upload");
    if (masks && indices && texcoords && mask_buffer) {
        int max_tri = -1;
        for (size_t i = 0; i + 1 < b.tris.size(); ++i) max_tri = std::max(max_tri, b.tris[i].id);
        const size_t ntri = (size_t)(max_tri + 1);
        int max_mat = -1, max_tc = -1;
        for (size_t t = 0; t < ntri; ++t) {
            max_mat = std::max(max_mat, indices[t * 4 + 3]);
            for (int k = 0; k < 3; ++k) max_tc = std::max(max_tc, indices[t * 4 + k]);
        }
        const size_t nm = (size_t)(max_mat + 1);
        if (any_real_mask(masks, nm)) {
            size_t mb = 4;
            for (size_t i = 0; i < nm; ++i) mb = std::max(mb, (size_t)masks[i].offset + (size_t)masks[i].width * masks[i].height);
            if (upload_array(&s->d_mask_indices, indices, ntri * 16) || upload_array(&s->d_mask_texcoords, texcoords, (size_t)(max_tc + 1) * 8) ||
                upload_array(&s->d_masks, masks, nm * 12) || upload_array(&s->d_mask_buffer, mask_buffer, mb)) {
                imba_set_error("%s", cudaGetErrorString(cudaGetLastError())); die("mask upload");
            }
            s->has_masks = true;
        }
    }
    g_cache.push_back(Entry{ nodes, tris, instances, root, IMBA_ACCEL_MBVH4, fp, s });
    return s;
}

void run_gpu(imba_scene* s, imba_ray* rays, imba_hit* hits, int count, int any_hit) {
    if (imba_scene_trace(s, rays, hits, count, any_hit, nullptr, cudaStreamPerThread) != IMBA_OK) die("traversal launch");
    if (cudaStreamSynchronize(cudaStreamPerThread) != cudaSuccess) { imba_set_error("%s", cudaGetErrorString(cudaGetLastError())); die("traversal"); }
}

} // namespace

extern "C" {

void intersect_gpu_masked_instanced(int root, imba_node2* nodes, imba_instance_node* instances, imba_vec4* tris, imba_ray* rays,
                                    imba_hit* hits, int* indices, imba_vec2* texcoords, imba_mask* masks, char* mask_buffer, int count) {
    run_gpu(scene_for_gpu(root, nodes, instances, tris, indices, texcoords, masks, mask_buffer), rays, hits, count, 0);
}
void occluded_gpu_masked_instanced(int root, imba_node2* nodes, imba_instance_node* instances, imba_vec4* tris, imba_ray* rays,
                                   imba_hit* hits, int* indices, imba_vec2* texcoords, imba_mask* masks, char* mask_buffer, int count) {
    run_gpu(scene_for_gpu(root, nodes, instances, tris, indices, texcoords, masks, mask_buffer), rays, hits, count, 1);
}
void intersect_cpu_masked_instanced(int root, imba_node4* nodes, imba_instance_node* instances, imba_vec4* tris, imba_ray* rays,
                                    imba_hit* hits, int* indices, imba_vec2* texcoords, imba_mask* masks, char* mask_buffer, int count) {
    if (imba_trace_host(scene_for_cpu(root, nodes, instances, tris, indices, texcoords, masks, mask_buffer), rays, hits, count, 0) != IMBA_OK) die("traversal");
}
void occluded_cpu_masked_instanced(int root, imba_node4* nodes, imba_instance_node* instances, imba_vec4* tris, imba_ray* rays,
                                   imba_hit* hits, int* indices, imba_vec2* texcoords, imba_mask* masks, char* mask_buffer, int count) {
    if (imba_trace_host(scene_for_cpu(root, nodes, instances, tris, indices, texcoords, masks, mask_buffer), rays, hits, count, 1) != IMBA_OK) die("traversal");
}

void imba_l1_invalidate(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    for (Entry& e : g_cache) imba_scene_destroy(e.scene);
    g_cache.clear();
}

} // extern "C"
