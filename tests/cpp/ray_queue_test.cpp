// GPU test of the C++ RayQueue mirror (include/imba_b200.hpp) against serial restatements of the reference's own loops
// (render/scheduling/ray_queue.h:123-244): concurrent push from 8 threads, range push, append, compact_hits, compact_rays,
// sort_by_material with the sorted_indices_ indirection.  Prints "ok" and exits 0; any mismatch exits with its check number.
#include <cstdio>
#include <cstdint>
#include <random>
#include <thread>
#include <vector>
#include "imba_b200.hpp"

struct State { int pixel_id; int sample_id; unsigned long long rng; float throughput[4]; };   // 32 bytes, pixel_id first (RayState)
static_assert(sizeof(State) % 8 == 0, "state size");

#define CHECK(n, cond) do { if (!(cond)) { std::printf("check %d failed: %s\n", n, #cond); return n; } } while (0)

int main() {
    const int N = 100003;
    imba_b200::RayQueue<State> q(N);
    CHECK(1, q.capacity() == (N + 63) / 64 * 64 && q.size() == 0);

    // ---- thread-safe push: 8 threads, every (thread, k) pair must land in exactly one slot (ray_queue.h:123-131)
    const int T = 8, PER = 5000;
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([&q, t] {
        for (int k = 0; k < PER; ++k) {
            imba_ray r; std::memset(&r, 0, sizeof(r)); r.org.x = (float)t; r.org.y = (float)k;
            State s; std::memset(&s, 0, sizeof(s)); s.pixel_id = t * PER + k; s.rng = 0x9E3779B97F4A7C15ull * (unsigned)(t * PER + k);
            q.push(r, s);
        }
    });
    for (auto& x : th) x.join();
    CHECK(2, q.size() == T * PER);
    std::vector<int> seen(T * PER, 0);
    for (int i = 0; i < q.size(); ++i) {
        const State& s = q.states()[i]; const imba_ray& r = q.rays()[i];
        CHECK(3, s.pixel_id >= 0 && s.pixel_id < T * PER && (int)r.org.x * PER + (int)r.org.y == s.pixel_id);
        seen[s.pixel_id]++;
    }
    for (int v : seen) CHECK(4, v == 1);

    // ---- range push (ray_queue.h:134-145) and append (:148-158)
    std::vector<imba_ray> rr(1000); std::vector<State> ss(1000);
    for (int i = 0; i < 1000; ++i) { std::memset(&rr[i], 0, sizeof(imba_ray)); rr[i].org.z = (float)i; std::memset(&ss[i], 0, sizeof(State)); ss[i].pixel_id = 100000 + i; }
    const int before = q.size();
    q.push(rr.begin(), rr.end(), ss.begin(), ss.end());
    CHECK(5, q.size() == before + 1000 && q.states()[before + 999].pixel_id == 100999 && q.rays()[before + 7].org.z == 7.0f);
    imba_b200::RayQueue<State> other(64);
    for (int i = 0; i < 10; ++i) other.push(rr[i], ss[i]);
    q.append(other);
    CHECK(6, q.size() == before + 1010 && q.states()[before + 1009].pixel_id == 100009);

    // ---- compact_hits / sort_by_material / compact_rays on random contents
    std::mt19937 gen(12345);
    q.clear();
    const int NUM_MATS = 13;
    std::vector<int> tri_mat(5000);
    for (auto& m : tri_mat) m = (int)(gen() % NUM_MATS);
    for (int i = 0; i < N; ++i) {
        imba_ray r; std::memset(&r, 0, sizeof(r)); r.org.x = (float)i; r.dir.w = 1e30f;
        State s; std::memset(&s, 0, sizeof(s)); s.pixel_id = (gen() % 10 < 7) ? i : -1; s.rng = gen();
        q.push(r, s);
        q.hits()[i].tri_id = (gen() % 10 < 6) ? (int)(gen() % tri_mat.size()) : -1; q.hits()[i].inst_id = 0; q.hits()[i].tmax = (float)i; q.hits()[i].u = 0.25f;
    }
    std::vector<imba_hit> h0(q.hits(), q.hits() + N); std::vector<State> s0(q.states(), q.states() + N); std::vector<imba_ray> r0(q.rays(), q.rays() + N);
    const int hit_count = q.compact_hits();
    int want_hits = 0; for (int i = 0; i < N; ++i) want_hits += h0[i].tri_id >= 0;
    CHECK(7, hit_count == want_hits && q.size() == N);
    // hits first, in their original order (what the reference's swap loop guarantees for the hits), each with its own ray and state
    int j = 0;
    for (int i = 0; i < N; ++i) if (h0[i].tri_id >= 0) {
        CHECK(8, q.hits()[j].tri_id == h0[i].tri_id && q.hits()[j].tmax == h0[i].tmax && q.rays()[j].org.x == r0[i].org.x && q.states()[j].rng == s0[i].rng && q.states()[j].pixel_id == s0[i].pixel_id);
        ++j;
    }
    for (int i = hit_count; i < N; ++i) CHECK(9, q.hits()[i].tri_id < 0);
    for (int i = 0; i < N; ++i) CHECK(10, &q.ray(i) == &q.rays()[i]);   // identity order after compact_hits (ray_queue.h:178-179)

    q.sort_by_material([&](const imba_hit& h) { return tri_mat[h.tri_id]; }, NUM_MATS, hit_count);
    std::vector<int> used(hit_count, 0); int last_mat = -1;
    for (int i = 0; i < hit_count; ++i) {
        const imba_hit& h = q.hit(i);
        const int m = tri_mat[h.tri_id];
        CHECK(11, m >= last_mat); last_mat = m;                       // grouped by material
        const int idx = (int)(&h - q.hits());
        CHECK(12, idx >= 0 && idx < hit_count && !used[idx]); used[idx] = 1;          // a permutation of the hit entries
        int stash; std::memcpy(&stash, &q.rays()[idx].dir.w, 4);
        CHECK(13, stash == m && &q.state(i) == &q.states()[idx] && &q.ray(i) == &q.rays()[idx]);   // material id left in dir.w (ray_queue.h:222)
    }

    q.shrink(hit_count);
    std::vector<State> s1(q.states(), q.states() + hit_count); std::vector<imba_ray> r1(q.rays(), q.rays() + hit_count);
    q.compact_rays();
    int k = 0;
    for (int i = 0; i < hit_count; ++i) if (s1[i].pixel_id >= 0) {
        CHECK(14, q.states()[k].pixel_id == s1[i].pixel_id && q.states()[k].rng == s1[i].rng && q.rays()[k].org.x == r1[i].org.x);
        ++k;
    }
    CHECK(15, q.size() == k);

    // empty queue: all three are no-ops
    q.clear();
    CHECK(16, q.compact_hits() == 0); q.compact_rays(); CHECK(17, q.size() == 0);
    std::printf("ok\n");
    return 0;
}
