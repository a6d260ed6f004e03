// What does moving 1024 cached 48 KB frames (scattered in a multi-GB pool, so they come from DRAM) into a contiguous 48 MB
// observation buffer cost on a B200, by itself?  This is the data movement of maze3d_step_kernel without its step logic.
//   A  ring of bulk (TMA) copies through shared memory, one issuing thread per CTA, 2 CTAs per SM (the step kernel's scheme)
//   B  plain LDG.128 / STG.128 copy, UNROLL independent 16-byte loads per thread, many CTAs per SM
//   C  like A with a dedicated load-issuing warp (producer / consumer mbarriers)
// Every launch uses a fresh random selection of frames (the pool is much larger than L2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o framecopy.bin framecopy.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr uint32_t kFrame = 128 * 128 * 3;       // 49152 B
__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph)
{
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(saddr(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_load(void *s, const void *g, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(s)), "l"(g), "r"(bytes), "r"(saddr(b)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *g, const void *s, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(saddr(s)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ---- A / C: ring through shared memory
template <int SLOTS, int CHUNK, int SLACK, bool SPLIT>
__global__ void __launch_bounds__(128) ring_kernel(const uint8_t *pool, const int *sel, uint8_t *out, int n)
{
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ __align__(8) uint64_t full[SLOTS], empty[SLOTS];
    const int tid = threadIdx.x;
    constexpr int NCH = kFrame / CHUNK;
    if (tid == 0) {
        for (int k = 0; k < SLOTS; ++k) { mbar_init(&full[k], 1); mbar_init(&empty[k], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int frames = (n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int M = frames * NCH;
    auto src = [&](int u) { return pool + (size_t)sel[blockIdx.x + (u / NCH) * gridDim.x] * kFrame + (size_t)(u % NCH) * CHUNK; };
    auto dst = [&](int u) { return out + (size_t)(blockIdx.x + (u / NCH) * gridDim.x) * kFrame + (size_t)(u % NCH) * CHUNK; };
    if (!SPLIT) {
        if (tid == 0) {
            for (int u = 0; u < SLOTS - SLACK && u < M; ++u) { mbar_expect(&full[u % SLOTS], CHUNK); bulk_load(ring + (size_t)(u % SLOTS) * CHUNK, src(u), CHUNK, &full[u % SLOTS]); }
            for (int u = 0; u < M; ++u) {
                const int slot = u % SLOTS;
                mbar_wait(&full[slot], (u / SLOTS) & 1);
                bulk_store(dst(u), ring + (size_t)slot * CHUNK, CHUNK);
                bulk_commit();
                const int v = u + SLOTS - SLACK;
                if (v < M) {
                    bulk_wait_read<SLACK>();
                    mbar_expect(&full[v % SLOTS], CHUNK);
                    bulk_load(ring + (size_t)(v % SLOTS) * CHUNK, src(v), CHUNK, &full[v % SLOTS]);
                }
            }
            bulk_wait_read<0>();
        }
    } else {
        if (tid == 32) {
            for (int u = 0; u < M; ++u) {
                const int slot = u % SLOTS, use = u / SLOTS;
                if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
                mbar_expect(&full[slot], CHUNK);
                bulk_load(ring + (size_t)slot * CHUNK, src(u), CHUNK, &full[slot]);
            }
        } else if (tid == 0) {
            for (int u = 0; u < M; ++u) {
                const int slot = u % SLOTS;
                mbar_wait(&full[slot], (u / SLOTS) & 1);
                bulk_store(dst(u), ring + (size_t)slot * CHUNK, CHUNK);
                bulk_commit();
                if (u >= SLACK) { bulk_wait_read<SLACK>(); mbar_arrive(&empty[(u - SLACK) % SLOTS]); }
            }
            bulk_wait_read<0>();
        }
    }
}

// ---- B: plain vector copy; one CTA moves PART bytes of one frame
template <int UNROLL>
__global__ void __launch_bounds__(256) plain_kernel(const uint8_t *pool, const int *sel, uint8_t *out, int parts)
{
    const int f = blockIdx.x / parts, p = blockIdx.x % parts;
    const uint32_t part = kFrame / parts;
    const uint4 *s = reinterpret_cast<const uint4 *>(pool + (size_t)sel[f] * kFrame + (size_t)p * part);
    uint4 *d = reinterpret_cast<uint4 *>(out + (size_t)f * kFrame + (size_t)p * part);
    const int n16 = part / 16;
    for (int i = threadIdx.x; i < n16; i += 256 * UNROLL) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) if (i + u * 256 < n16) v[u] = __ldg(s + i + u * 256);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) if (i + u * 256 < n16) d[i + u * 256] = v[u];
    }
}

static int *d_sel[64];
template <typename F> static float time_it(F launch, const char *name, int n)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 8; ++i) launch(d_sel[i]);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 48;
    for (int i = 0; i < reps; ++i) launch(d_sel[8 + i]);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps;
    printf("%-58s %7.2f us per %d frames   %6.0f GB/s read + %6.0f GB/s written\n", name, us, n, n * (double)kFrame / us * 1e-3, n * (double)kFrame / us * 1e-3);
    return (float)us;
}

int main()
{
    const int n = 1024;
    const size_t pool_frames = 40000;                       // 1.97 GB
    uint8_t *pool, *out;
    CK(cudaMalloc(&pool, pool_frames * kFrame));
    CK(cudaMalloc(&out, (size_t)n * kFrame));
    CK(cudaMemset(pool, 7, pool_frames * kFrame));
    srand(1);
    for (int k = 0; k < 64; ++k) {
        std::vector<int> h(n);
        for (int i = 0; i < n; ++i) h[i] = (int)(((size_t)rand() * 7919u + (size_t)rand()) % pool_frames);
        CK(cudaMalloc(&d_sel[k], n * sizeof(int)));
        CK(cudaMemcpy(d_sel[k], h.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs; %d frames of %u B per launch, back-to-back launches (no graph)\n", prop.name, sms, n, kFrame);
#define RING(SLOTS, CHUNK, SLACK, SPLIT, CPS, NAME)                                                                       \
    {                                                                                                                       \
        auto k = ring_kernel<SLOTS, CHUNK, SLACK, SPLIT>;                                                                   \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SLOTS * CHUNK));                           \
        time_it([&](int *sel) { k<<<sms * CPS, 128, SLOTS * CHUNK>>>(pool, sel, out, n); }, NAME, n);                      \
    }
    RING(8, 12288, 3, false, 2, "A ring 8 x 12 KB, slack 3, 2 CTAs/SM (step kernel)");
    RING(8, 12288, 3, true, 2, "C same, separate load-issuing warp");
    RING(4, 24576, 1, true, 2, "C ring 4 x 24 KB, slack 1, 2 CTAs/SM");
    RING(8, 6144, 3, true, 4, "C ring 8 x 6 KB, slack 3, 4 CTAs/SM");
    RING(4, 12288, 1, true, 4, "C ring 4 x 12 KB, slack 1, 4 CTAs/SM");
    RING(6, 8192, 2, true, 4, "C ring 6 x 8 KB, slack 2, 4 CTAs/SM");
    RING(16, 6144, 6, true, 2, "C ring 16 x 6 KB, slack 6, 2 CTAs/SM");
    RING(3, 16384, 1, true, 4, "C ring 3 x 16 KB, slack 1, 4 CTAs/SM");
    time_it([&](int *sel) { plain_kernel<4><<<n * 1, 256>>>(pool, sel, out, 1); }, "B LDG/STG.128 x4, one CTA per frame", n);
    time_it([&](int *sel) { plain_kernel<4><<<n * 4, 256>>>(pool, sel, out, 4); }, "B LDG/STG.128 x4, four CTAs per frame", n);
    time_it([&](int *sel) { plain_kernel<8><<<n * 2, 256>>>(pool, sel, out, 2); }, "B LDG/STG.128 x8, two CTAs per frame", n);
    time_it([&](int *sel) { plain_kernel<2><<<n * 8, 256>>>(pool, sel, out, 8); }, "B LDG/STG.128 x2, eight CTAs per frame", n);
    time_it([&](int *sel) { CK(cudaMemcpyAsync(out, pool + (size_t)(rand() % 30000) * kFrame, (size_t)n * kFrame, cudaMemcpyDeviceToDevice)); },
            "cudaMemcpy D2D of one contiguous 48 MB block", n);
    return 0;
}
