// Micro-measurements (B200, through gpurun):
//   A. latency / throughput of dependent FFMA vs FFMA2 chains: 1 or 2 warps per scheduler, 1..8 independent chains per thread
//   B. programmatic dependent launch: does ONE thread executing griddepcontrol.launch_dependents release the dependent grid,
//      and when do the dependent grid's CTAs actually start relative to the primary's?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_latency.bin fma_latency.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int MODE, int CH> __global__ void chain_kernel(float *out, long long *cycles, int iters, float a, float b)
{
    float2 acc[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[k] = make_float2(threadIdx.x * 0.001f + k, k * 0.5f);
    const float2 av = make_float2(a, a), bv = make_float2(b, b);
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                if (MODE == 0) acc[k].x = fmaf(acc[k].x, a, b);
                else acc[k] = __ffma2_rn(acc[k], av, bv);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// primary: records its start, triggers (all threads / thread 0 only / never), then spins ~spin_ns, records its end
template <int TRIG> __global__ void primary_kernel(unsigned long long *ts, unsigned long long spin_ns)
{
    const unsigned long long t0 = gtime();
    if (TRIG == 1) asm volatile("griddepcontrol.launch_dependents;");
    if (TRIG == 2 && threadIdx.x == 0) asm volatile("griddepcontrol.launch_dependents;");
    while (gtime() - t0 < spin_ns) { }
    if (threadIdx.x == 0) { ts[2 * blockIdx.x] = t0; ts[2 * blockIdx.x + 1] = gtime(); }
}
// secondary: never waits on the grid dependency; records when each CTA started
__global__ void secondary_kernel(unsigned long long *ts)
{
    if (threadIdx.x == 0) ts[blockIdx.x] = gtime();
}

template <typename K, typename... A> static void launch_ex(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, A... args)
{
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, args...));
}

template <int MODE, int CH> static void run_chain(int warps_per_sm, float *out, long long *cyc)
{
    const int iters = 4000;
    const int threads = warps_per_sm * 32;
    chain_kernel<MODE, CH><<<148, threads>>>(out, cyc, iters, 1.0001f, 0.5f);
    CK(cudaDeviceSynchronize());
    chain_kernel<MODE, CH><<<148, threads>>>(out, cyc, iters, 1.0001f, 0.5f);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    const double per_instr = (double)c / (iters * 8.0 * CH);
    printf("  %s chains/thread %d, warps/SM %2d: %.2f cycles per warp-instruction (%.1f lane-FMA/clk/SM)\n", MODE ? "FFMA2" : "FFMA ", CH, warps_per_sm,
           per_instr, warps_per_sm * 32.0 * (MODE ? 2 : 1) / per_instr);
}

int main()
{
    float *out; long long *cyc;
    CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 8));
    printf("== A. dependent chains (cycles per warp-instruction of ONE warp's stream; 4 warps/SM = 1 per scheduler)\n");
    run_chain<0, 1>(4, out, cyc); run_chain<1, 1>(4, out, cyc);
    run_chain<0, 2>(4, out, cyc); run_chain<1, 2>(4, out, cyc);
    run_chain<0, 4>(4, out, cyc); run_chain<1, 4>(4, out, cyc);
    run_chain<0, 8>(4, out, cyc); run_chain<1, 8>(4, out, cyc);
    run_chain<0, 1>(8, out, cyc); run_chain<1, 1>(8, out, cyc);
    run_chain<0, 2>(8, out, cyc); run_chain<1, 2>(8, out, cyc);
    run_chain<0, 4>(8, out, cyc); run_chain<1, 4>(8, out, cyc);
    run_chain<0, 4>(16, out, cyc); run_chain<1, 4>(16, out, cyc);

    printf("== B. programmatic dependent launch: primary 148 CTAs x 64 threads spinning 20 us; secondary 148 CTAs, no grid wait\n");
    cudaStream_t st; CK(cudaStreamCreate(&st));
    unsigned long long *tp, *ts2; CK(cudaMalloc(&tp, 148 * 16)); CK(cudaMalloc(&ts2, 148 * 8));
    unsigned long long hp[296], hs[148];
    for (int trig = 0; trig < 3; ++trig) {
        for (int rep = 0; rep < 2; ++rep) {
            if (trig == 0) launch_ex(primary_kernel<0>, dim3(148), dim3(64), 0, st, false, tp, 20000ull);
            if (trig == 1) launch_ex(primary_kernel<1>, dim3(148), dim3(64), 0, st, false, tp, 20000ull);
            if (trig == 2) launch_ex(primary_kernel<2>, dim3(148), dim3(64), 0, st, false, tp, 20000ull);
            launch_ex(secondary_kernel, dim3(148), dim3(64), 0, st, true, ts2);
            CK(cudaStreamSynchronize(st));
        }
        CK(cudaMemcpy(hp, tp, sizeof(hp), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hs, ts2, sizeof(hs), cudaMemcpyDeviceToHost));
        unsigned long long p0 = ~0ull, p1 = 0, s0 = ~0ull, s1 = 0;
        for (int i = 0; i < 148; ++i) { if (hp[2 * i] < p0) p0 = hp[2 * i]; if (hp[2 * i + 1] > p1) p1 = hp[2 * i + 1]; if (hs[i] < s0) s0 = hs[i]; if (hs[i] > s1) s1 = hs[i]; }
        printf("  trigger %s: primary ran %.2f us; first secondary CTA started %.2f us after the primary's start, last %.2f us (primary end = %.2f)\n",
               trig == 0 ? "never       " : (trig == 1 ? "all threads " : "thread 0 only"), (p1 - p0) * 1e-3, ((double)s0 - (double)p0) * 1e-3, ((double)s1 - (double)p0) * 1e-3, (p1 - p0) * 1e-3);
    }
    return 0;
}
