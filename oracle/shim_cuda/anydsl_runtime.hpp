/* ORACLE / TEST INFRASTRUCTURE.  The AnyDSL runtime's Array<T> / copy API (as used by render/scene.cpp:14-188 and
 * render/scheduling/ray_queue.h:42-54,78-79,277,292) over the CUDA runtime: Platform::Cuda arrays are cudaMalloc'ed on the
 * given device, copies go through cudaMemcpy.  With this shim (oracle/Makefile: ref_b200) the UNMODIFIED reference uploads its
 * scene to the GPU, copies every ray queue up and every hit queue down itself, and calls the four traversal entry points with
 * device pointers -- exactly what it does with the real AnyDSL runtime -- so libimba_b200.so is exercised as the drop-in it is. */
#ifndef ORACLE_ANYDSL_RUNTIME_CUDA_HPP
#define ORACLE_ANYDSL_RUNTIME_CUDA_HPP
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
/* The five CUDA runtime entry points used, declared by hand: <cuda_runtime_api.h> puts float3 / float4 / ... into the global
 * namespace, where they collide with the reference's imba::float3 under `using namespace imba`. */
extern "C" {
int cudaMalloc(void** ptr, size_t bytes);
int cudaFree(void* ptr);
int cudaMemcpy(void* dst, const void* src, size_t bytes, int kind);
int cudaSetDevice(int device);
const char* cudaGetErrorString(int error);
}
namespace anydsl {
enum { kCudaMemcpyDefault = 4 };
enum class Platform : int32_t { Host = 0, Cuda = 1, OpenCL = 2, HSA = 3 };
struct Device { int32_t id; explicit Device(int32_t i = 0) : id(i) {} };
inline void cuda_check(int e, const char* what) {
    if (e != 0) { fprintf(stderr, "anydsl shim: %s: %s\n", what, cudaGetErrorString(e)); abort(); }
}
template <typename T>
class Array {
public:
    Array() : data_(nullptr), size_(0), dev_(0) {}
    Array(int64_t size) : Array(Platform::Host, Device(0), size) {}
    Array(Platform p, Device d, int64_t size) : size_(size), dev_((int32_t)p | (d.id << 4)) {
        void* ptr = nullptr;
        const size_t bytes = (size_t)(size > 0 ? size : 1) * sizeof(T);
        if (p == Platform::Cuda) { cuda_check(cudaSetDevice(d.id), "cudaSetDevice"); cuda_check(cudaMalloc(&ptr, bytes), "cudaMalloc"); }
        else if (posix_memalign(&ptr, 64, bytes) != 0) abort();
        data_ = (T*)ptr;
    }
    Array(Array&& o) : data_(o.data_), size_(o.size_), dev_(o.dev_) { o.data_ = nullptr; o.size_ = 0; }
    Array& operator=(Array&& o) { if (this != &o) { release(); data_ = o.data_; size_ = o.size_; dev_ = o.dev_; o.data_ = nullptr; o.size_ = 0; } return *this; }
    Array(const Array&) = delete;
    Array& operator=(const Array&) = delete;
    ~Array() { release(); }
    T* begin() { return data_; }  const T* begin() const { return data_; }
    T* end() { return data_ + size_; }  const T* end() const { return data_ + size_; }
    T* data() { return data_; }  const T* data() const { return data_; }
    int64_t size() const { return size_; }
    int32_t device() const { return dev_; }
    T& operator[](int64_t i) { return data_[i]; }
    const T& operator[](int64_t i) const { return data_[i]; }
private:
    void release() { if (!data_) return; if ((dev_ & 0xf) == (int32_t)Platform::Cuda) cudaFree(data_); else free(data_); data_ = nullptr; }
    T* data_; int64_t size_; int32_t dev_;
};
template <typename T> void copy(const Array<T>& a, Array<T>& b, int64_t n) {
    if (((a.device() | b.device()) & 0xf) == 0) memcpy(b.data(), a.data(), sizeof(T) * n);
    else cuda_check(cudaMemcpy(b.data(), a.data(), sizeof(T) * n, kCudaMemcpyDefault), "cudaMemcpy");
}
template <typename T> void copy(const Array<T>& a, Array<T>& b) { copy(a, b, a.size()); }
}
inline void anydsl_copy(int32_t dev_src, const void* src, int64_t off_src, int32_t dev_dst, void* dst, int64_t off_dst, int64_t size) {
    if (((dev_src | dev_dst) & 0xf) == 0) memcpy((char*)dst + off_dst, (const char*)src + off_src, (size_t)size);
    else anydsl::cuda_check(cudaMemcpy((char*)dst + off_dst, (const char*)src + off_src, (size_t)size, anydsl::kCudaMemcpyDefault), "anydsl_copy");
}
#endif
