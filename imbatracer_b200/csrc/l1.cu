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
 * A cached structure is recognised by the array addresses, the root, the extents of the allocations (gpu variant) and a
 * CONTENT FINGERPRINT, so that a host which frees a scene and uploads another one at the same addresses -- cudaMalloc and
 * malloc both hand freed blocks out again -- does not traverse the stale tree: the gpu variant hashes the root node and 256
 * evenly spaced 16-byte words of each of the three arrays in a one-block kernel launched in front of the traversal (checked
 * after the call's own synchronisation, no extra round trip; on a mismatch the structure is rebuilt and the queue traversed
 * again), the cpu variant hashes the top 64 nodes of the tree, the first instance record and the first triangle packet on
 * the host.  Edits elsewhere inside an unchanged allocation are not seen: call imba_l1_invalidate() after such a rebuild.
 * No error channel exists in these signatures (ray_queue.h:248,270-271 assert): failures print and abort().
 */
#include "api_common.h"
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {

struct Entry {
    const void* nodes; const void* tris; const void* insts; int root; int layout; uint64_t fingerprint;
    size_t bytes[3];
    imba_scene* scene;
};
std::mutex g_mutex;
std::vector<Entry> g_cache;
std::vector<imba_scene*> g_retired;   // structures found stale: other threads may still traverse them, freed by imba_l1_invalidate

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

__device__ __host__ inline uint64_t mix64(uint64_t h, uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xBF58476D1CE4E5B9ull; return h ^ (h >> 29); }

/* One block: every thread hashes one 16-byte word of each array (evenly spaced over the allocation), thread 0 the root node
 * as well; the per-thread hashes are position-salted and XOR-reduced, so the result does not depend on scheduling. */
__global__ void __launch_bounds__(256) fingerprint_kernel(const uint4* nodes, size_t node_words, const uint4* tris, size_t tri_words, const uint4* insts, size_t inst_words,
                                                          size_t root_word, unsigned long long* out) {
    __shared__ unsigned long long acc;
    if (threadIdx.x == 0) acc = 0ull;
    __syncthreads();
    const uint4* arr[3] = { nodes, tris, insts }; const size_t words[3] = { node_words, tri_words, inst_words };
    uint64_t h = 0x243F6A8885A308D3ull + threadIdx.x;
    for (int a = 0; a < 3; ++a) {
        if (words[a] == 0) continue;
        const size_t w = (size_t)((unsigned __int128)threadIdx.x * words[a] / 256);
        const uint4 v = arr[a][w];
        h = mix64(h, ((uint64_t)v.y << 32) | v.x); h = mix64(h, ((uint64_t)v.w << 32) | v.z); h = mix64(h, (uint64_t)a * 1315423911ull + w);
    }
    if (threadIdx.x < 3 && root_word + threadIdx.x < node_words) {   // the 56-byte root node: 3.5 words
        const uint4 v = nodes[root_word + threadIdx.x];
        h = mix64(h, ((uint64_t)v.y << 32) | v.x); h = mix64(h, ((uint64_t)v.w << 32) | v.z);
    }
    atomicXor(&acc, (unsigned long long)h);
    __syncthreads();
    if (threadIdx.x == 0) *out = acc;
}

// a pinned result slot per calling thread (TileScheduler threads call concurrently, tile_scheduler.h:69-78)
unsigned long long* fingerprint_slot() {
    static thread_local unsigned long long* slot = nullptr;
    if (!slot && cudaHostAlloc((void**)&slot, sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess) { imba_set_error("%s", cudaGetErrorString(cudaGetLastError())); die("pinned fingerprint slot"); }
    return slot;
}

size_t bytes_from(const void* ptr) {
    unsigned long long base = 0; size_t size = 0;
    if (range_fn()(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0) { imba_set_error("pointer %p is not a device allocation", ptr); die("cuMemGetAddressRange"); }
    return (size_t)(base + size - (unsigned long long)(uintptr_t)ptr);
}

void launch_fingerprint(const void* nodes, const void* tris, const void* insts, int root, const size_t bytes[3], unsigned long long* slot) {
    fingerprint_kernel<<<1, 256, 0, cudaStreamPerThread>>>((const uint4*)nodes, bytes[0] / 16, (const uint4*)tris, bytes[1] / 16, (const uint4*)insts, bytes[2] / 16,
                                                         (size_t)root * sizeof(imba_node2) / 16, slot);
    if (cudaGetLastError() != cudaSuccess) { imba_set_error("%s", cudaGetErrorString(cudaGetLastError())); die("fingerprint launch"); }
    g_imba_launches.fetch_add(1);
}

Entry* lookup(const void* nodes, const void* tris, const void* insts, int root, int layout, const size_t* bytes) {
    for (Entry& e : g_cache)
        if (e.nodes == nodes && e.tris == tris && e.insts == insts && e.root == root && e.layout == layout &&
            (!bytes || (e.bytes[0] == bytes[0] && e.bytes[1] == bytes[1] && e.bytes[2] == bytes[2]))) return &e;
    return nullptr;
}

void retire(const void* nodes, const void* tris, const void* insts, int root, int layout) {
    for (size_t i = 0; i < g_cache.size();) {
        Entry& e = g_cache[i];
        if (e.nodes == nodes && e.tris == tris && e.insts == insts && e.root == root && e.layout == layout) { g_retired.push_back(e.scene); g_cache.erase(g_cache.begin() + i); }
        else ++i;
    }
}

imba_scene* scene_for_gpu(int root, imba_node2* nodes, imba_instance_node* instances, imba_vec4* tris, int* indices,
                          imba_vec2* texcoords, imba_mask* masks, char* mask_buffer, const size_t bytes[3], uint64_t fp, bool have_fp) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (Entry* e = lookup(nodes, tris, instances, root, IMBA_ACCEL_BVH2, bytes)) {
        if (!have_fp || e->fingerprint == fp) return e->scene;
    }
    if (!have_fp) return nullptr;   // unknown structure: the caller waits for the fingerprint first
    retire(nodes, tris, instances, root, IMBA_ACCEL_BVH2);
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
    g_cache.push_back(Entry{ nodes, tris, instances, root, IMBA_ACCEL_BVH2, fp, { bytes[0], bytes[1], bytes[2] }, s });
    return s;
}

template <typename T> int upload_array(T** dst, const void* src, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (cudaMalloc((void**)dst, bytes) != cudaSuccess) return -1;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
}

imba_scene* scene_for_cpu(int root, imba_node4* nodes, imba_instance_node* instances, imba_vec4* tris, int* indices,
                          imba_vec2* texcoords, imba_mask* masks, char* mask_buffer) {
    // staleness check on the host: the top of the tree (breadth first from the root, 64 nodes), the instance records and triangles
    // its first leaves point at -- what changes when the host rebuilds in place or loads another scene at the same addresses
    uint64_t fp = 0;
    {
        int queue[64]; int qh = 0, qt = 0; queue[qt++] = root;
        while (qh < qt) {
            const imba_node4& n = nodes[queue[qh++]];
            fp = fnv(&n, sizeof(imba_node4), fp ? fp : 1469598103934665603ull);
            for (int k = 0; k < 4; ++k) {
                const int c = n.children[k];
                if (c > 0 && qt < 64) queue[qt++] = c;
            }
        }
        if (instances) fp = fnv(instances, sizeof(imba_instance_node), fp);
        fp = fnv(tris, 13 * sizeof(imba_vec4), fp);
    }
    std::lock_guard<std::mutex> lock(g_mutex);
    if (Entry* e = lookup(nodes, tris, instances, root, IMBA_ACCEL_MBVH4, nullptr)) {
        if (e->fingerprint == fp) return e->scene;
        retire(nodes, tris, instances, root, IMBA_ACCEL_MBVH4);
    }
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
    g_cache.push_back(Entry{ nodes, tris, instances, root, IMBA_ACCEL_MBVH4, fp, { 0, 0, 0 }, s });
    return s;
}

void trace_sync(imba_scene* s, imba_ray* rays, imba_hit* hits, int count, int any_hit) {
    if (imba_scene_trace(s, rays, hits, count, any_hit, nullptr, cudaStreamPerThread) != IMBA_OK) die("traversal launch");
    if (cudaStreamSynchronize(cudaStreamPerThread) != cudaSuccess) { imba_set_error("%s", cudaGetErrorString(cudaGetLastError())); die("traversal"); }
}

/* gpu variant of a call: fingerprint kernel, traversal on the cached structure, ONE synchronisation, fingerprint check. */
void run_gpu(int root, imba_node2* nodes, imba_instance_node* instances, imba_vec4* tris, imba_ray* rays, imba_hit* hits, int* indices,
             imba_vec2* texcoords, imba_mask* masks, char* mask_buffer, int count, int any_hit) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { imba_set_error("no CUDA device (this library has no CPU path)"); die("init"); }
    const size_t bytes[3] = { bytes_from(nodes), bytes_from(tris), bytes_from(instances) };
    unsigned long long* slot = fingerprint_slot();
    launch_fingerprint(nodes, tris, instances, root, bytes, slot);
    imba_scene* s = scene_for_gpu(root, nodes, instances, tris, indices, texcoords, masks, mask_buffer, bytes, 0, false);
    if (s) {
        trace_sync(s, rays, hits, count, any_hit);
        imba_scene* s2 = scene_for_gpu(root, nodes, instances, tris, indices, texcoords, masks, mask_buffer, bytes, *slot, true);
        if (s2 == s) return;
        s = s2;   // the cached structure was stale (rebuilt just now): traverse again
    } else {
        if (cudaStreamSynchronize(cudaStreamPerThread) != cudaSuccess) { imba_set_error("%s", cudaGetErrorString(cudaGetLastError())); die("fingerprint"); }
        s = scene_for_gpu(root, nodes, instances, tris, indices, texcoords, masks, mask_buffer, bytes, *slot, true);
    }
    trace_sync(s, rays, hits, count, any_hit);
}

} // namespace

extern "C" {

void intersect_gpu_masked_instanced(int root, imba_node2* nodes, imba_instance_node* instances, imba_vec4* tris, imba_ray* rays,
                                    imba_hit* hits, int* indices, imba_vec2* texcoords, imba_mask* masks, char* mask_buffer, int count) {
    run_gpu(root, nodes, instances, tris, rays, hits, indices, texcoords, masks, mask_buffer, count, 0);
}
void occluded_gpu_masked_instanced(int root, imba_node2* nodes, imba_instance_node* instances, imba_vec4* tris, imba_ray* rays,
                                   imba_hit* hits, int* indices, imba_vec2* texcoords, imba_mask* masks, char* mask_buffer, int count) {
    run_gpu(root, nodes, instances, tris, rays, hits, indices, texcoords, masks, mask_buffer, count, 1);
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
    for (imba_scene* s : g_retired) imba_scene_destroy(s);
    g_cache.clear(); g_retired.clear();
}

} // extern "C"
