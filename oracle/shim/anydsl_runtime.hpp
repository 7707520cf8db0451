/* ORACLE / TEST INFRASTRUCTURE. Host-only stand-in for the AnyDSL runtime's
 * Array<T>/copy API as used by render/scene.cpp:14-188 and
 * render/scheduling/ray_queue.h:42-54,78-79,277,292. "Cuda" arrays are plain
 * host memory here: the oracle never touches a device. */
#ifndef ORACLE_ANYDSL_RUNTIME_HPP
#define ORACLE_ANYDSL_RUNTIME_HPP
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <utility>
namespace anydsl {
enum class Platform : int32_t { Host = 0, Cuda = 1, OpenCL = 2, HSA = 3 };
struct Device { int32_t id; explicit Device(int32_t i = 0) : id(i) {} };
template <typename T>
class Array {
public:
    Array() : data_(nullptr), size_(0), dev_(0) {}
    Array(int64_t size) : Array(Platform::Host, Device(0), size) {}
    Array(Platform p, Device d, int64_t size) : size_(size), dev_((int32_t)p | (d.id << 4)) {
        void* ptr = nullptr;
        if (posix_memalign(&ptr, 64, (size_t)(size > 0 ? size : 1) * sizeof(T)) != 0) abort();
        data_ = (T*)ptr;
    }
    Array(Array&& o) : data_(o.data_), size_(o.size_), dev_(o.dev_) { o.data_ = nullptr; o.size_ = 0; }
    Array& operator=(Array&& o) { if (this != &o) { free(data_); data_ = o.data_; size_ = o.size_; dev_ = o.dev_; o.data_ = nullptr; o.size_ = 0; } return *this; }
    Array(const Array&) = delete;
    Array& operator=(const Array&) = delete;
    ~Array() { free(data_); }
    T* begin() { return data_; }  const T* begin() const { return data_; }
    T* end() { return data_ + size_; }  const T* end() const { return data_ + size_; }
    T* data() { return data_; }  const T* data() const { return data_; }
    int64_t size() const { return size_; }
    int32_t device() const { return dev_; }
    T& operator[](int64_t i) { return data_[i]; }
    const T& operator[](int64_t i) const { return data_[i]; }
private:
    T* data_; int64_t size_; int32_t dev_;
};
template <typename T> void copy(const Array<T>& a, Array<T>& b, int64_t n) { memcpy(b.data(), a.data(), sizeof(T) * n); }
template <typename T> void copy(const Array<T>& a, Array<T>& b) { copy(a, b, a.size()); }
}
inline void anydsl_copy(int32_t, const void* src, int64_t off_src, int32_t, void* dst, int64_t off_dst, int64_t size) {
    memcpy((char*)dst + off_dst, (const char*)src + off_src, (size_t)size);
}
#endif
