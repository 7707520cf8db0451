/* vecmath.cuh -- float3 helpers for the shading kernels.  Operation order follows the reference's core/float3.h,
 * core/common.h (dot = x*x + y*y + z*z left to right, normalize = v * (1/sqrt(dot)), lerp = a*(1-u-v) + b*u + c*v) so
 * that, compiled with -fmad=false, algebraic results are bit-identical to the host code and only the transcendental
 * functions (sinf, cosf, powf, ...) differ by their library rounding. */
#ifndef IMBA_VECMATH_CUH
#define IMBA_VECMATH_CUH
#include <cuda_runtime.h>

struct V3 { float x, y, z; };
__host__ __device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ V3 v3(float a) { return v3(a, a, a); }
__host__ __device__ __forceinline__ V3 v3(const float4& f) { return v3(f.x, f.y, f.z); }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float b) { return v3(a.x * b, a.y * b, a.z * b); }
__host__ __device__ __forceinline__ V3 operator*(float a, V3 b) { return v3(a * b.x, a * b.y, a * b.z); }
__host__ __device__ __forceinline__ V3 operator/(V3 a, float b) { return v3(a.x / b, a.y / b, a.z / b); }
__host__ __device__ __forceinline__ V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
__host__ __device__ __forceinline__ V3& operator*=(V3& a, float b) { a = a * b; return a; }
__host__ __device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__host__ __device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__host__ __device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / length(a)); }
__host__ __device__ __forceinline__ float sqr(float x) { return x * x; }
__host__ __device__ __forceinline__ float clampf(float a, float lo, float hi) { return a < lo ? lo : (a > hi ? hi : a); }
__host__ __device__ __forceinline__ float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__host__ __device__ __forceinline__ bool is_black(V3 a) { return a.x <= 0.0f && a.y <= 0.0f && a.z <= 0.0f; }
__host__ __device__ __forceinline__ float luminance(V3 c) { return dot(c, v3(0.2126f, 0.7152f, 0.0722f)); }

#define IMBA_PI 3.14159265359f

/* render/random.h:57-72 */
__host__ __device__ __forceinline__ void local_coordinates(V3 n, V3& tangent, V3& binormal) {
    const bool xy = fabsf(n.x) > fabsf(n.y);
    const float n0 = xy ? n.x : n.y;
    const float sig = xy ? -1.f : 1.f;
    const float inv_len = 1.f / sqrtf(sqr(n0) + sqr(n.z));
    const float t0 = n.z * sig * inv_len;
    const float tz = n0 * -1.f * sig * inv_len;
    tangent = xy ? v3(t0, 0.f, tz) : v3(0.f, t0, tz);
    binormal = cross(n, tangent);
    tangent = normalize(tangent);
    binormal = normalize(binormal);
}

#endif
