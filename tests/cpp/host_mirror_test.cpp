// Compiles the C++ host-side mirror (include/imba_b200.hpp) and links it against libimba_b200.so.
// Without a CUDA device the library must refuse to work (no CPU fallback): exit code 0 = refused as expected,
// 2 = a scene was created (a GPU is present; the caller then expects traversal results on stdout).
#include <cstdio>
#include "imba_b200.hpp"

struct MyState { int pixel_id; int sample_id; unsigned long long rng; };

int main() {
    imba_b200::RayQueue<MyState> q(100);
    if (q.capacity() != 128 || q.size() != 0) return 10;
    imba_ray r; std::memset(&r, 0, sizeof(r)); r.dir.z = 1.0f; r.dir.w = 1e30f;
    MyState s = { 7, 0, 42ull };
    q.push(r, s);
    if (q.size() != 1 || q.states()[0].pixel_id != 7) return 11;
    imba_render_settings rs = imba_b200::default_settings(64, 48);
    rs.iterations_per_pass = 4;                       // PT / LT: four progressive iterations per wavefront pass
    if (rs.spp != 1 || rs.max_path_len != 10 || rs.light_path_count != 64 * 48 / 2 || sizeof(rs) % 8 != 0) return 13;
    imba_scene_desc d; std::memset(&d, 0, sizeof(d));
    d.accel_layout = IMBA_ACCEL_DEVICE_BUILD;         // the enumerator must exist for C++ hosts too
    try {
        imba_b200::Scene scene(d);
        std::printf("scene created\n");
        return 2;
    } catch (const std::runtime_error& e) {
        std::printf("refused: %s\n", e.what());
        return std::string(e.what()).find("no CUDA device") != std::string::npos || std::string(e.what()).find("BVH conversion") != std::string::npos ? 0 : 12;
    }
}
