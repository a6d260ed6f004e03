// Throughput of the float64 instructions the MetaMaze raycaster is made of, per SM (B200, through gpurun):
//   DFMA / DMUL / DADD / DADD.RZ, F2I.F64.TRUNC, I2F.F64, FRND.F64.FLOOR, the float64 division sequence, and FFMA as the yardstick.
// Every thread runs 8 independent chains, 16 warps per SM sub-partition-quad (512 threads), one CTA per SM, so the numbers are
// issue/pipe throughput, not latency.  Printed unit: lane-operations per clock per SM.
// Second table: latency of a dependent chain (one warp, one chain per thread), cycles per chain step.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate.bin fp64_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int CH = 8;
// conversions through inline PTX so that the compiler cannot fold (int)(double)(int)x chains
__device__ __forceinline__ int f2i(double x) { int r; asm volatile("cvt.rzi.s32.f64 %0, %1;" : "=r"(r) : "d"(x)); return r; }
__device__ __forceinline__ double i2f(int x) { double r; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(r) : "r"(x)); return r; }
__device__ __forceinline__ double frnd(double x) { double r; asm volatile("cvt.rmi.f64.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
enum { M_DFMA, M_DMUL, M_DADD, M_DADD_RZ, M_F2I, M_I2F, M_FRND, M_DDIV, M_FFMA, M_MAGIC_TRUNC, M_COUNT };
static const char *kNames[M_COUNT] = {"DFMA", "DMUL", "DADD", "DADD.RZ", "F2I.F64.TRUNC (+I2F.F64 back)", "I2F.F64 (+F2I back)",
                                     "FRND.F64.FLOOR", "float64 division", "FFMA (float32)", "trunc via DADD.RZ + 2^52"};
// how many counted operations one chain step performs
static const double kOps[M_COUNT] = {1, 1, 1, 1, 2, 2, 1, 1, 1, 1};

template <int MODE> __global__ void __launch_bounds__(512, 1) rate_kernel(double *out, long long *cycles, int iters, double a, double b)
{
    double acc[CH];
    float facc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) { acc[k] = threadIdx.x * 0.37 + k + 1.5; facc[k] = (float)acc[k]; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                if (MODE == M_DFMA) acc[k] = fma(acc[k], a, b);
                else if (MODE == M_DMUL) acc[k] = __dmul_rn(acc[k], a);
                else if (MODE == M_DADD) acc[k] = __dadd_rn(acc[k], b);
                else if (MODE == M_DADD_RZ) acc[k] = __dadd_rz(acc[k], b);
                else if (MODE == M_F2I) acc[k] = i2f(f2i(acc[k]));
                else if (MODE == M_I2F) acc[k] = i2f(f2i(acc[k]) ^ 3);
                else if (MODE == M_FRND) acc[k] = frnd(acc[k]);
                else if (MODE == M_DDIV) acc[k] = acc[k] / a;
                else if (MODE == M_FFMA) facc[k] = fmaf(facc[k], (float)a, (float)b);
                else if (MODE == M_MAGIC_TRUNC)
                    acc[k] = __hiloint2double(0x40200000, __double2loint(__dadd_rz(acc[k], 4503599627370496.0)));
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += acc[k] + facc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> __global__ void latency_kernel(double *out, long long *cycles, int iters, double a, double b)
{
    double acc = threadIdx.x * 0.37 + 1.5;
    float facc = (float)acc;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == M_DFMA) acc = fma(acc, a, b);
            else if (MODE == M_DMUL) acc = __dmul_rn(acc, a);
            else if (MODE == M_DADD) acc = __dadd_rn(acc, b);
            else if (MODE == M_DADD_RZ) acc = __dadd_rz(acc, b);
            else if (MODE == M_F2I) acc = i2f(f2i(acc));
            else if (MODE == M_I2F) acc = i2f(f2i(acc) ^ 3);
            else if (MODE == M_FRND) acc = frnd(acc);
            else if (MODE == M_DDIV) acc = acc / a;
            else if (MODE == M_FFMA) facc = fmaf(facc, (float)a, (float)b);
            else if (MODE == M_MAGIC_TRUNC) acc = __hiloint2double(0x40200000, __double2loint(__dadd_rz(acc, 4503599627370496.0)));
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc + facc;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int MODE> static void run_latency(double *out, long long *cyc)
{
    const int iters = 500;
    const double a = (MODE == M_DDIV) ? 1.0000001 : 0.999999, b = 1.0e-3;
    latency_kernel<MODE><<<1, 32>>>(out, cyc, iters, a, b);
    CK(cudaDeviceSynchronize());
    latency_kernel<MODE><<<1, 32>>>(out, cyc, iters, a, b);
    CK(cudaDeviceSynchronize());
    long long h = 0;
    CK(cudaMemcpy(&h, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    printf("%-34s %8.1f cycles per dependent step (%s)\n", kNames[MODE], (double)h / (iters * 16.0),
           kOps[MODE] == 2 ? "two conversions (+ xor)" : MODE == M_FRND ? "one instruction" : MODE == M_MAGIC_TRUNC ? "DADD.RZ + mov" : "one instruction");
}

template <int MODE> static void run(double *out, long long *cyc, int sms)
{
    const int iters = 2000;
    const double a = (MODE == M_DDIV) ? 1.0000001 : 0.999999, b = 1.0e-3;
    rate_kernel<MODE><<<sms, 512>>>(out, cyc, iters, a, b);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    rate_kernel<MODE><<<sms, 512>>>(out, cyc, iters, a, b);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[1024];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    mean /= sms;
    const double ops = (double)iters * 4 * CH * 512 * kOps[MODE];
    printf("%-34s %8.2f lane-ops/clk/SM   (%.0f cycles, %.3f ms, %.2f Tops/s over %d SMs)\n", kNames[MODE], ops / mean, mean, ms,
           ops * sms / (ms * 1e-3) * 1e-12, sms);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    double *out; long long *cyc;
    CK(cudaMalloc(&out, sizeof(double) * sms * 512));
    CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
    run<M_FFMA>(out, cyc, sms);
    run<M_DFMA>(out, cyc, sms);
    run<M_DMUL>(out, cyc, sms);
    run<M_DADD>(out, cyc, sms);
    run<M_DADD_RZ>(out, cyc, sms);
    run<M_F2I>(out, cyc, sms);
    run<M_I2F>(out, cyc, sms);
    run<M_FRND>(out, cyc, sms);
    run<M_DDIV>(out, cyc, sms);
    run<M_MAGIC_TRUNC>(out, cyc, sms);
    printf("latency, one warp, one chain:\n");
    run_latency<M_FFMA>(out, cyc);
    run_latency<M_DFMA>(out, cyc);
    run_latency<M_DMUL>(out, cyc);
    run_latency<M_DADD>(out, cyc);
    run_latency<M_F2I>(out, cyc);
    run_latency<M_I2F>(out, cyc);
    run_latency<M_FRND>(out, cyc);
    run_latency<M_DDIV>(out, cyc);
    run_latency<M_MAGIC_TRUNC>(out, cyc);
    return 0;
}
