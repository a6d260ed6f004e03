// Lane types of the quadrotor kernels: the same float32 arithmetic written once and instantiated for
//   T = float : one env per thread (scalar FFMA / FMUL / FADD), and
//   T = f2    : two envs per thread in packed registers (sm_100 FFMA2 / FADD2: one instruction = the same IEEE
//               round-to-nearest operation on both halves, so every env computes bit for bit what the scalar code does
//               while the kernel issues half as many floating-point instructions).
//
// Contraction hazard (measured with CUDA 12.9's ptxas): `mul.rn.f32x2` followed by `add.rn.f32x2` IS fused into one
// FFMA2 even under --fmad=false (the scalar `mul.rn.f32` + `add.rn.f32` pair is not).  The reference rounds every
// product before adding, so the packed product is issued as fma(a, b, -0) with the -0 read from constant memory at run
// time: x*y + (-0) == x*y for every x*y (including +-0, inf, NaN), it is still one FFMA2, and ptxas has no FMUL2 to
// contract and no way to know the addend is -0.  tests/test_quadrotor_gpu.py::test_packed_kernel_equals_scalar_kernel
// is the guard (bit equality of whole trajectories against the scalar kernel).
#pragma once
#include <cuda_runtime.h>

struct f2 {
    float2 v;
};

__constant__ float2 mgb_neg_zero2 = {-0.0f, -0.0f};

template <class T> struct Lanes;
template <> struct Lanes<float> { static constexpr int N = 1; };
template <> struct Lanes<f2> { static constexpr int N = 2; };

// ---- lane access
__device__ __forceinline__ float lane(float x, int) { return x; }
__device__ __forceinline__ float lane(f2 x, int h) { return h ? x.v.y : x.v.x; }
__device__ __forceinline__ void set_lane(float &x, int, float s) { x = s; }
__device__ __forceinline__ void set_lane(f2 &x, int h, float s)
{
    if (h) x.v.y = s;
    else x.v.x = s;
}
template <class T> __device__ __forceinline__ T bc(float s);
template <> __device__ __forceinline__ float bc<float>(float s) { return s; }
template <> __device__ __forceinline__ f2 bc<f2>(float s) { return f2{make_float2(s, s)}; }
__device__ __forceinline__ f2 pack2(float a, float b) { return f2{make_float2(a, b)}; }

// ---- IEEE float32 building blocks (round to nearest even, no contraction across calls)
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vmul(float a, float b) { return a * b; }       // file is compiled with -fmad=false
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float vabs(float a) { return fabsf(a); }

__device__ __forceinline__ f2 vfma(f2 a, f2 b, f2 c) { return f2{__ffma2_rn(a.v, b.v, c.v)}; }
__device__ __forceinline__ f2 vmul(f2 a, f2 b) { return f2{__ffma2_rn(a.v, b.v, mgb_neg_zero2)}; }
__device__ __forceinline__ f2 vadd(f2 a, f2 b) { return f2{__fadd2_rn(a.v, b.v)}; }
__device__ __forceinline__ f2 vneg(f2 a) { return f2{make_float2(-a.v.x, -a.v.y)}; }
__device__ __forceinline__ f2 vsub(f2 a, f2 b) { return f2{__fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y))}; }
__device__ __forceinline__ f2 vabs(f2 a) { return f2{make_float2(fabsf(a.v.x), fabsf(a.v.y))}; }

// mixed forms: uniform (per-launch constant) operands are broadcast; ptxas encodes them as `UR.F32` / `R.F32` operands
template <class T> __device__ __forceinline__ T vfma(float a, T b, T c) { return vfma(bc<T>(a), b, c); }
template <class T> __device__ __forceinline__ T vfma(T a, float b, T c) { return vfma(a, bc<T>(b), c); }
template <class T> __device__ __forceinline__ T vmul(float a, T b) { return vmul(bc<T>(a), b); }
template <class T> __device__ __forceinline__ T vmul(T a, float b) { return vmul(a, bc<T>(b)); }
template <class T> __device__ __forceinline__ T vadd(T a, float b) { return vadd(a, bc<T>(b)); }
template <class T> __device__ __forceinline__ T vsub(T a, float b) { return vsub(a, bc<T>(b)); }
template <class T> __device__ __forceinline__ T vsub(float a, T b) { return vsub(bc<T>(a), b); }

// ---- per-lane (scalar pipe) helpers: min / max / select / special functions
__device__ __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float vmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ f2 vmax(f2 a, f2 b) { return f2{make_float2(fmaxf(a.v.x, b.v.x), fmaxf(a.v.y, b.v.y))}; }
__device__ __forceinline__ f2 vmin(f2 a, f2 b) { return f2{make_float2(fminf(a.v.x, b.v.x), fminf(a.v.y, b.v.y))}; }

// sqrt.approx: MUFU.SQRT, <= 1 ulp-ish, sqrt(0) = 0
__device__ __forceinline__ float vsqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ f2 vsqrt(f2 x) { return f2{make_float2(vsqrt(x.v.x), vsqrt(x.v.y))}; }
__device__ __forceinline__ float vrcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ f2 vrcp_approx(f2 x) { return f2{make_float2(vrcp_approx(x.v.x), vrcp_approx(x.v.y))}; }
// MUFU.RCP + one Newton step (~0.5 ulp)
template <class T> __device__ __forceinline__ T vrcp(T x)
{
    const T r = vrcp_approx(x);
    return vfma(r, vfma(vneg(x), r, bc<T>(1.0f)), r);
}
