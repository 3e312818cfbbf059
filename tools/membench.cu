// Micro-benchmark: achievable HBM bandwidth for warp-wide row reads of 256 B .. 2 KB at scattered
// addresses (the access pattern of the tiled SoA layout) versus sequential streaming.
// Index generation is a 32-bit LCG + mask (a few instructions) so that the loads dominate.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/membench tools/membench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int VEC, int UNROLL>
__global__ void scatter_read(const double *buf, unsigned row_mask, int iters, double *sink, int sequential)
{
    const int lane = threadIdx.x & 31;
    const unsigned warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    unsigned s = warp * 2654435761u + 12345u;
    double acc = 0;
    unsigned seq = warp * 7919u;
    for (int it = 0; it < iters; it++)
    {
        unsigned r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            s = s * 1664525u + 1013904223u;
            r[u] = sequential ? ((seq++) & row_mask) : ((s >> 7) & row_mask);
        }
        double v[UNROLL][VEC];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int k = 0; k < VEC; k++)
                v[u][k] = buf[((size_t)r[u] * VEC + k) * 32 + lane]; // VEC consecutive 256-B rows = one VEC*256 B burst
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int k = 0; k < VEC; k++)
                acc += v[u][k];
    }
    if (acc == 1.2345)
        sink[0] = acc;
}

// four warps of a CTA read the four quarters of the same random 1-KB row
template <int UNROLL>
__global__ void coop_read(const double *buf, unsigned row_mask, int iters, double *sink)
{
    const int lane = threadIdx.x & 31, wk = threadIdx.x >> 5, nwk = blockDim.x >> 5;
    unsigned s = blockIdx.x * 2654435761u + 12345u;
    double acc = 0;
    for (int it = 0; it < iters; it++)
    {
        unsigned r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            s = s * 1664525u + 1013904223u;
            r[u] = (s >> 7) & row_mask;
        }
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
            v[u] = buf[((size_t)r[u] * nwk + wk) * 32 + lane];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
            acc += v[u];
    }
    if (acc == 1.2345)
        sink[0] = acc;
}

static float timed(void (*launch)(int), int iters)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch(2);
    cudaEventRecord(a);
    launch(iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

static const double *g_buf;
static double *g_sink;
static int g_blocks, g_threads;
static unsigned g_mask;
template <int VEC, int UNROLL, int SEQ>
static void launch_scatter(int iters) { scatter_read<VEC, UNROLL><<<g_blocks, g_threads>>>(g_buf, g_mask, iters, g_sink, SEQ); }
template <int UNROLL>
static void launch_coop(int iters) { coop_read<UNROLL><<<g_blocks, g_threads>>>(g_buf, g_mask, iters, g_sink); }

template <int VEC, int UNROLL, int SEQ>
static double bw(size_t nd, int iters)
{
    size_t rows = nd / (32 * VEC), p2 = 1;
    while (p2 * 2 <= rows) p2 *= 2;
    g_mask = (unsigned)(p2 - 1);
    const float ms = timed(launch_scatter<VEC, UNROLL, SEQ>, iters);
    return (double)g_blocks * (g_threads / 32) * iters * UNROLL * VEC * 256.0 / (ms * 1e-3) / 1e9;
}
template <int UNROLL>
static double bwc(size_t nd, int iters)
{
    size_t rows = nd / g_threads, p2 = 1;
    while (p2 * 2 <= rows) p2 *= 2;
    g_mask = (unsigned)(p2 - 1);
    const float ms = timed(launch_coop<UNROLL>, iters);
    return (double)g_blocks * (g_threads / 32) * iters * UNROLL * 256.0 / (ms * 1e-3) / 1e9;
}

int main()
{
    const size_t bytes = 64ULL << 30;
    double *buf;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&g_sink, 8);
    cudaMemset(buf, 0, bytes);
    g_buf = buf;
    const size_t nd = bytes / 8;
    for (int warps_per_sm : {4, 8, 16, 32, 64})
    {
        g_threads = 128;
        g_blocks = 148 * warps_per_sm * 32 / g_threads;
        printf("warps/SM=%2d | rand256 u1 %5.0f u4 %5.0f u8 %5.0f | rand512 u1 %5.0f u4 %5.0f | rand1K u1 %5.0f u4 %5.0f | rand2K u1 %5.0f | seq u1 %5.0f u8 %5.0f | coop4x256 u1 %5.0f u4 %5.0f u8 %5.0f GB/s\n",
               warps_per_sm,
               bw<1, 1, 0>(nd, 4000), bw<1, 4, 0>(nd, 1000), bw<1, 8, 0>(nd, 500),
               bw<2, 1, 0>(nd, 4000), bw<2, 4, 0>(nd, 1000),
               bw<4, 1, 0>(nd, 2000), bw<4, 4, 0>(nd, 500),
               bw<8, 1, 0>(nd, 1000),
               bw<1, 1, 1>(nd, 4000), bw<1, 8, 1>(nd, 500),
               bwc<1>(nd, 4000), bwc<4>(nd, 1000), bwc<8>(nd, 500));
    }
    return 0;
}
