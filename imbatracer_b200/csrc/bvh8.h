/* bvh8.h -- device acceleration-structure layout of the B200 traversal kernels.
 *
 * The reference hands its traversal library either MBVH4 nodes + 4-triangle SoA packets
 * (core/cpu_adapter.cpp:56-208) or BVH2 nodes + 3-Vec4 triangles (core/gpu_adapter.cpp:62-143),
 * plus a top-level tree whose leaves are InstanceNode lists (cpu_adapter.cpp:268-367).  Both are
 * converted once (bvh8_build.cpp) into ONE array of compressed 8-wide nodes shared by the top level
 * and every mesh, one array of 48-byte triangle blocks and one array of 64-byte instance records.
 *
 * Node8 (80 bytes = five 128-bit loads), after Ylitie/Karras/Laine 2017:
 *   n0: origin p (3 x f32) | ex, ey, ez (biased power-of-two exponents of the quantisation grid), imask
 *   n1: child_base (first inner child, absolute node index) | prim_base (first primitive of the first
 *       leaf child: triangle index in a mesh tree, instance-record index in the top-level tree) | meta[8]
 *   n2: qlo_x[8] qlo_y[8]     n3: qlo_z[8] qhi_x[8]     n4: qhi_y[8] qhi_z[8]
 *   child box = p + q * 2^e, conservative with at least 1/64 cell of slack (the builder keeps 2^e >= ulp of the
 *   node's coordinates so the grid is never finer than float resolution)
 *   meta[slot]: 0 = empty; inner child: 0b001_11sss (sss = slot); leaf child: unary primitive count
 *   (1 -> 001, 2 -> 011, 3 -> 111) in the high 3 bits, offset from prim_base (< 24) in the low 5.
 *   imask bit s = slot s holds an inner child; inner children are stored contiguously in slot order.
 *
 * Tri48: v0 (w = global triangle id), e1 = v0 - v1, e2 = v2 - v0  -- the quantities the reference's CPU layout
 *   stores (cpu_adapter.cpp:146-163); n = e1 x e2 is recomputed unfused, bit-identical to the stored one.
 * Inst64: the reference's InstanceNode (inverse 3x4 transform, instance id) with `next` rewritten to the
 *   Node8 index of the mesh root.
 */
#ifndef IMBA_BVH8_H
#define IMBA_BVH8_H
#include <stdint.h>

struct Node8 {
    float    p[3];
    uint8_t  e[3];
    uint8_t  imask;
    uint32_t child_base;
    uint32_t prim_base;
    uint8_t  meta[8];
    uint8_t  qlo_x[8], qlo_y[8], qlo_z[8], qhi_x[8], qhi_y[8], qhi_z[8];
};
/* nx, ny = the first two components of n = e1 x e2, computed unfused at build time exactly as the kernel computes the third
 * (cpu_adapter.cpp:149 stores the whole n; two floats fit the padding of the 48-byte record): 6 instructions less per test. */
struct Tri48  { float v0[3]; int32_t id; float e1[3]; float nx; float e2[3]; float ny; };
#define IMBA_TRI48_FINISH(t) do { (t).nx = (t).e1[1] * (t).e2[2] - (t).e1[2] * (t).e2[1]; (t).ny = (t).e1[2] * (t).e2[0] - (t).e1[0] * (t).e2[2]; } while (0)
struct Inst64 { float m[12]; int32_t inst_id; int32_t root; int32_t pad[2]; };

#if defined(__cplusplus)
static_assert(sizeof(Node8) == 80 && sizeof(Tri48) == 48 && sizeof(Inst64) == 64, "device layout sizes");
#endif

#define IMBA_ROBUST_TFAR 1.00000048f /* 1 + 2^-21 (Ize 2013), same constant as the oracle */
#define IMBA_T_PREFILTER  1.000001f  /* slack of the T <= |det|*tbest pre-filter; the decision is taken on the rounded t */
#define IMBA_NO_HIT       0x7fffffff /* initial (tri, inst) of a ray: loses every tie */
#define IMBA_RCP_EPS     8.2718061e-25f /* 2^-80 */

#endif
