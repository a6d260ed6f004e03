// How fast can ONE single-wave launch move the quadrotor step's bytes at 65 536 envs (state 96 B in + 96 B out, action 16 B
// in, obs 76 B + reward 4 B + done 1 B out = 289 B/env with the int ct/episode words; 18.9 MB per launch), with no
// arithmetic?  Variants: per-thread LDG/STG.128 (what the step kernel did in round 1) vs bulk copies by the TMA engine
// (cp.async.bulk global<->shared, one 12 KB request per 128-env state tile), loads only / stores only / both.
// Launch = 148 CTAs (one per SM), graph of 256 launches, programmatic dependent launch on/off.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o memfloor.bin memfloor.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_load(void *s, const void *g, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(s)), "l"(g), "r"(bytes), "r"(smem_addr(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase)
{
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_addr(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_store(void *g, const void *s, uint32_t bytes) { asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_addr(s)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Args {
    const float4 *st_in; float4 *st_out; const float4 *act; float *obs; float *rew; uint8_t *done;
    int n_tiles;     // 128-env tiles (12 KB of state each)
    int do_load, do_store;
};

// LDG/STG variant: thread = env pair (12 float4 of state), CTA = contiguous range of tiles, 64 threads per tile
template <bool PDL> __global__ void __launch_bounds__(256) ldg_kernel(const __grid_constant__ Args a)
{
    extern __shared__ __align__(128) float tile[];           // obs rows [envs][19]
    if (PDL) { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
    const int tiles_per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
    const int t0 = blockIdx.x * tiles_per, t1 = min(a.n_tiles, t0 + tiles_per);
    const int npairs = (t1 - t0) * 64;
    const int j = threadIdx.x;
    float4 q[12];
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
    if (j < npairs) {
        const size_t P = (size_t)t0 * 64 + j;
        const float4 *b = a.st_in + (P / 64) * 768 + (P % 64);
        if (a.do_load) {
#pragma unroll
            for (int k = 0; k < 12; ++k) q[k] = __ldcg(b + k * 64);
            a0 = __ldg(a.act + 2 * P); a1 = __ldg(a.act + 2 * P + 1);
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k) q[k] = make_float4(j, k, 1.f, 2.f);
        }
        const float s = a0.x + a1.y;
#pragma unroll
        for (int k = 0; k < 12; ++k) { q[k].x += s; q[k].w += 1.f; }
        if (a.do_store) {
            float4 *o = a.st_out + (P / 64) * 768 + (P % 64);
#pragma unroll
            for (int k = 0; k < 12; ++k) o[k * 64] = q[k];
            float *r0 = tile + (2 * j) * 19, *r1 = r0 + 19;
#pragma unroll
            for (int k = 0; k < 19; ++k) { r0[k] = q[k % 12].x; r1[k] = q[k % 12].y; }
            *reinterpret_cast<float2 *>(a.rew + 2 * P) = make_float2(q[0].x, q[1].y);
            *reinterpret_cast<uchar2 *>(a.done + 2 * P) = make_uchar2(1, 0);
        } else if (q[3].x == 123.456f) a.rew[0] = q[5].y;
    }
    if (a.do_store) {
        fence_async();
        __syncthreads();
        if (threadIdx.x == 0 && npairs > 0) { bulk_store(a.obs + (size_t)t0 * 128 * 19, tile, (uint32_t)npairs * 2 * 76); bulk_commit(); bulk_wait_read0(); }
    }
}

// TMA variant: the CTA's tiles are fetched with ONE bulk copy each (12 KB + 2 KB of actions) into shared memory, threads
// read their pair from smem, write the new state back into the same smem tile, and one bulk store per tile writes it out.
template <bool PDL> __global__ void __launch_bounds__(256) tma_kernel(const __grid_constant__ Args a)
{
    extern __shared__ __align__(128) float4 smem4[];          // [tiles_per][768 + 128] float4, then obs rows
    __shared__ __align__(8) uint64_t bar;
    if (PDL) { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
    const int tiles_per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
    const int t0 = blockIdx.x * tiles_per, t1 = min(a.n_tiles, t0 + tiles_per);
    const int nt = t1 - t0, npairs = nt * 64;
    float *tile = reinterpret_cast<float *>(smem4 + (size_t)tiles_per * 896);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (a.do_load && nt > 0) {
            mbar_expect(&bar, (uint32_t)nt * (12288 + 2048));
            for (int t = 0; t < nt; ++t) {
                bulk_load(smem4 + (size_t)t * 896, a.st_in + (size_t)(t0 + t) * 768, 12288, &bar);
                bulk_load(smem4 + (size_t)t * 896 + 768, a.act + (size_t)(t0 + t) * 128, 2048, &bar);
            }
        }
    }
    __syncthreads();
    if (a.do_load && nt > 0) mbar_wait(&bar, 0);
    const int j = threadIdx.x;
    if (j < npairs) {
        float4 *b = smem4 + (size_t)(j / 64) * 896 + (j % 64);
        float4 q[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) q[k] = a.do_load ? b[k * 64] : make_float4(j, k, 1.f, 2.f);
        const float4 a0 = a.do_load ? b[768 - (j % 64) + 2 * (j % 64)] : make_float4(0, 0, 0, 0);
        const float s = a0.x + a0.y;
#pragma unroll
        for (int k = 0; k < 12; ++k) { q[k].x += s; q[k].w += 1.f; }
        if (a.do_store) {
#pragma unroll
            for (int k = 0; k < 12; ++k) b[k * 64] = q[k];
            float *r0 = tile + (2 * j) * 19, *r1 = r0 + 19;
#pragma unroll
            for (int k = 0; k < 19; ++k) { r0[k] = q[k % 12].x; r1[k] = q[k % 12].y; }
            const size_t P = (size_t)t0 * 64 + j;
            *reinterpret_cast<float2 *>(a.rew + 2 * P) = make_float2(q[0].x, q[1].y);
            *reinterpret_cast<uchar2 *>(a.done + 2 * P) = make_uchar2(1, 0);
        } else if (q[3].x == 123.456f) a.rew[0] = q[5].y;
    }
    if (a.do_store) {
        fence_async();
        __syncthreads();
        if (threadIdx.x == 0 && nt > 0) {
            for (int t = 0; t < nt; ++t) bulk_store(a.st_out + (size_t)(t0 + t) * 768, smem4 + (size_t)t * 896, 12288);
            bulk_store(a.obs + (size_t)t0 * 128 * 19, tile, (uint32_t)npairs * 2 * 76);
            bulk_commit();
            bulk_wait_read0();
        }
    }
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
template <typename K> static void launch_ex(K kern, int grid, int block, size_t smem, cudaStream_t st, bool pdl, const Args &a)
{
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, a));
}

int main()
{
    cudaStream_t st; CK(cudaStreamCreate(&st));
    const int n_tiles = 512, n = n_tiles * 128, slots = 32;
    float4 *sa, *sb, *act; float *obs, *rew; uint8_t *done;
    CK(cudaMalloc(&sa, (size_t)n_tiles * 12288)); CK(cudaMalloc(&sb, (size_t)n_tiles * 12288));
    CK(cudaMemset(sa, 0, (size_t)n_tiles * 12288)); CK(cudaMemset(sb, 0, (size_t)n_tiles * 12288));
    CK(cudaMalloc(&act, (size_t)slots * n * 16)); CK(cudaMemset(act, 0, (size_t)slots * n * 16));
    CK(cudaMalloc(&obs, (size_t)slots * n * 76)); CK(cudaMalloc(&rew, (size_t)slots * n * 4)); CK(cudaMalloc(&done, (size_t)slots * n));
    const int grid = 148, tiles_per = (n_tiles + grid - 1) / grid;      // 4 tiles = 512 envs = 256 threads
    const size_t smem_ldg = (size_t)tiles_per * 128 * 76, smem_tma = (size_t)tiles_per * 896 * 16 + smem_ldg;
    CK(cudaFuncSetAttribute(ldg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ldg));
    CK(cudaFuncSetAttribute(ldg_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ldg));
    CK(cudaFuncSetAttribute(tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
    CK(cudaFuncSetAttribute(tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
    printf("65536 envs, 148 CTAs x 256 threads (grid of whole 128-env tiles: 4 or 3 tiles per CTA); us per launch\n");
    const char *names[3] = {"loads only (7.3 MB)", "stores only (11.6 MB)", "loads + stores (18.9 MB)"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int pdl = 0; pdl < 2; ++pdl) {
            float tl = time_graph(st, 256, 40, [&](int i) {
                Args a = {(i & 1) ? sb : sa, (i & 1) ? sa : sb, act + (size_t)(i % slots) * n, obs + (size_t)(i % slots) * n * 19,
                          rew + (size_t)(i % slots) * n, done + (size_t)(i % slots) * n, n_tiles, mode != 1, mode != 0};
                if (pdl) launch_ex(ldg_kernel<true>, grid, 256, smem_ldg, st, true, a); else launch_ex(ldg_kernel<false>, grid, 256, smem_ldg, st, false, a);
            });
            float tt = time_graph(st, 256, 40, [&](int i) {
                Args a = {(i & 1) ? sb : sa, (i & 1) ? sa : sb, act + (size_t)(i % slots) * n, obs + (size_t)(i % slots) * n * 19,
                          rew + (size_t)(i % slots) * n, done + (size_t)(i % slots) * n, n_tiles, mode != 1, mode != 0};
                if (pdl) launch_ex(tma_kernel<true>, grid, 256, smem_tma, st, true, a); else launch_ex(tma_kernel<false>, grid, 256, smem_tma, st, false, a);
            });
            printf("  %-26s pdl %d :  LDG/STG %.3f us   TMA bulk %.3f us\n", names[mode], pdl, tl, tt);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
