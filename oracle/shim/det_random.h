/* ORACLE / TEST INFRASTRUCTURE, force-included (-include) in front of every reference translation unit.
 * The reference seeds every queue refill from a `static std::random_device rd` (render/ray_gen/ray_gen.h:49-50,128-129,
 * render/integrators/light_vertices.cpp:34-35), so two runs never trace the same rays.  To compare the SAME render on two
 * traversal back ends ray for ray, std::random_device is replaced here by a class that behaves exactly like it until
 * imba_det::seed(s != 0) is called (oracle/ref_driver.cpp: ref_set_seed) and then returns the splitmix64 sequence of s.
 * No reference source line changes; with one scheduler thread the n-th refill always gets the n-th value. */
#ifndef ORACLE_DET_RANDOM_H
#define ORACLE_DET_RANDOM_H
#include <atomic>
#include <cstdint>
#include <random>
namespace imba_det {
inline std::atomic<uint64_t>& state() { static std::atomic<uint64_t> s{ 0 }; return s; }
inline std::atomic<bool>& enabled() { static std::atomic<bool> e{ false }; return e; }
inline void seed(uint64_t s) { enabled() = s != 0; state() = s; }
}
namespace std {
class imba_random_device {
public:
    typedef unsigned int result_type;
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xffffffffu; }
    result_type operator()() {
        if (!imba_det::enabled()) return real_();
        uint64_t z = imba_det::state().fetch_add(0x9E3779B97F4A7C15ull) + 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        return (result_type)(z >> 32);
    }
private:
    random_device real_;
};
}
#define random_device imba_random_device
#endif
