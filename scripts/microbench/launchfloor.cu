// Micro-measurements behind the round-2 design of the quadrotor step kernel (run on a B200 through gpurun):
//   1. per-launch floor of a dependent chain of kernels inside a CUDA graph: plain / programmatic dependent launch (PDL),
//      one CTA per SM vs a grid whose CTAs of launch k+1 can be co-resident with those of launch k;
//   2. memory-only step (loads 112 B + stores 169 B per env, no arithmetic) at 65 536 envs;
//   3. FFMA vs FFMA2 issue/throughput (lane-FMAs per clock per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launchfloor.bin launchfloor.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <bool PDL> __global__ void empty_kernel(float *p)
{
    if (PDL) {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f;
}

// memory-only step: thread = 2 envs; 12 float4 state loads, 2 action float4, 12 state stores, obs rows through smem + bulk
// store is approximated by plain float4 stores of 2 x 76 B -> 10 float4 (160 B), reward 8 B, done 2 B.
template <bool PDL> __global__ void __launch_bounds__(256) memstep_kernel(const float4 *__restrict__ st, float4 *__restrict__ st_out,
                                                                       const float4 *__restrict__ act, float4 *__restrict__ obs,
                                                                       float2 *__restrict__ rew, int n_pairs)
{
    if (PDL) {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const int per = (n_pairs + gridDim.x - 1) / gridDim.x;
    const int j = blockIdx.x * per + threadIdx.x;
    if (threadIdx.x >= per || j >= n_pairs) return;
    const int tile = j / 64, lane = j % 64;
    const float4 *b = st + (size_t)tile * 12 * 64 + lane;
    float4 q[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) q[k] = b[k * 64];
    const float4 a0 = act[2 * j], a1 = act[2 * j + 1];
    float s = a0.x + a1.y;
#pragma unroll
    for (int k = 0; k < 12; ++k) { q[k].x += s; q[k].y += a0.z; q[k].z += a1.w; q[k].w += 1.f; }
    float4 *o = st_out + (size_t)tile * 12 * 64 + lane;
#pragma unroll
    for (int k = 0; k < 12; ++k) o[k * 64] = q[k];
#pragma unroll
    for (int k = 0; k < 10; ++k) obs[(size_t)j * 10 + k] = q[k];
    rew[j] = make_float2(q[0].x, q[1].y);
}

template <int MODE> __global__ void __launch_bounds__(256) fma_kernel(float *out, int iters, float a, float b)
{
    float2 acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float2(threadIdx.x * 0.001f + k, k * 0.5f);
    const float2 av = make_float2(a, a), bv = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) { acc[k].x = fmaf(acc[k].x, a, b); acc[k].y = fmaf(acc[k].y, a, b); }
            else if (MODE == 1) acc[k] = __ffma2_rn(acc[k], av, bv);
            else { acc[k] = __ffma2_rn(acc[k], av, acc[(k + 1) & 7]); }      // 3 register operands
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> static float time_graph(cudaStream_t st, int nodes, int replays, F launch)
{
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int i = 0; i < nodes; ++i) launch(i);
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < replays; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    return ms * 1e3f / (nodes * replays);
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

int main()
{
    cudaStream_t st; CK(cudaStreamCreate(&st));
    float *p; CK(cudaMalloc(&p, 1 << 20));
    CK(cudaMemset(p, 0, 1 << 20));
    CK(cudaFuncSetAttribute(empty_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(empty_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    printf("== 1. launch floor, graph of 256 dependent launches, us per launch\n");
    struct Cfg { int grid, block; size_t smem; const char *what; } cfgs[] = {
        {148, 448, 0, "148 x 448, no smem (co-residency possible)"},
        {148, 448, 160 * 1024, "148 x 448, 160 KB smem (1 CTA/SM: no co-residency)"},
        {148, 224, 0, "148 x 224"},
        {296, 224, 0, "296 x 224"},
        {1024, 64, 0, "1024 x 64"},
        {148, 1024, 0, "148 x 1024"},
    };
    for (auto &c : cfgs) {
        float a = time_graph(st, 256, 40, [&](int) { launch_ex(empty_kernel<false>, dim3(c.grid), dim3(c.block), c.smem, st, false, p); });
        float b = time_graph(st, 256, 40, [&](int) { launch_ex(empty_kernel<true>, dim3(c.grid), dim3(c.block), c.smem, st, true, p); });
        printf("  %-55s plain %.3f  pdl %.3f\n", c.what, a, b);
    }
    printf("== 2. memory-only step, 65536 envs (32768 pairs), us per launch (algorithmic 281 B/env = 18.4 MB)\n");
    const int n_pairs = 32768;
    float4 *sa, *sb, *act, *obs; float2 *rew;
    CK(cudaMalloc(&sa, (size_t)n_pairs * 192)); CK(cudaMalloc(&sb, (size_t)n_pairs * 192));
    CK(cudaMemset(sa, 0, (size_t)n_pairs * 192));
    const int slots = 32;
    CK(cudaMalloc(&act, (size_t)slots * n_pairs * 32)); CK(cudaMalloc(&obs, (size_t)slots * n_pairs * 160));
    CK(cudaMemset(act, 0, (size_t)slots * n_pairs * 32));
    CK(cudaMalloc(&rew, (size_t)slots * n_pairs * 8));
    for (int grid : {148, 296, 592}) {
        const int per = (n_pairs + grid - 1) / grid;
        const int block = (per + 31) / 32 * 32;
        for (int pdl = 0; pdl < 2; ++pdl) {
            float t = time_graph(st, 256, 40, [&](int i) {
                const float4 *src = (i & 1) ? sb : sa; float4 *dst = (i & 1) ? sa : sb;
                if (pdl) launch_ex(memstep_kernel<true>, dim3(grid), dim3(block), 0, st, true, src, dst, (const float4 *)(act + (size_t)(i % slots) * n_pairs * 2), obs + (size_t)(i % slots) * n_pairs * 10, rew + (size_t)(i % slots) * n_pairs, n_pairs);
                else launch_ex(memstep_kernel<false>, dim3(grid), dim3(block), 0, st, false, src, dst, (const float4 *)(act + (size_t)(i % slots) * n_pairs * 2), obs + (size_t)(i % slots) * n_pairs * 10, rew + (size_t)(i % slots) * n_pairs, n_pairs);
            });
            printf("  grid %4d x %4d threads  pdl %d : %.3f us\n", grid, block, pdl, t);
        }
    }
    printf("== 3. FFMA vs FFMA2: 148 x 4 CTAs x 256 threads, 8 independent float2 accumulators, lane-FMAs/clk/SM at 1.965 GHz\n");
    float *out; CK(cudaMalloc(&out, 148 * 8 * 256 * 4));
    const int iters = 20000;
    for (int warps_mode = 0; warps_mode < 2; ++warps_mode) {
        const int ctas = warps_mode ? 148 : 148 * 4;    // 8 or 2 warps per scheduler
        for (int mode = 0; mode < 3; ++mode) {
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0, st));
                if (mode == 0) fma_kernel<0><<<ctas, 256, 0, st>>>(out, iters, 1.0001f, 0.5f);
                else if (mode == 1) fma_kernel<1><<<ctas, 256, 0, st>>>(out, iters, 1.0001f, 0.5f);
                else fma_kernel<2><<<ctas, 256, 0, st>>>(out, iters, 1.0001f, 0.5f);
                CK(cudaEventRecord(e1, st));
                CK(cudaStreamSynchronize(st));
            }
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double fmas = (double)ctas * 256 * iters * 16;
            printf("  %s, %d warps/SM: %.3f ms  %.1f lane-FMA/clk/SM (at 1.965 GHz)\n",
                   mode == 0 ? "FFMA (scalar x2)" : (mode == 1 ? "FFMA2 (imm/uniform operands)" : "FFMA2 (3 register operands)"),
                   ctas * 8 / 148, ms, fmas / (ms * 1e-3) / 1.965e9 / 148);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
