// Thin device-runtime shim.  The product build (nvcc) maps it onto the CUDA runtime; the
// tests/emu build (-DEICOS_EMU, plain g++) maps it onto the heap so that the host-side
// orchestration in engine.cu can be exercised without a GPU.  The emulator is test
// infrastructure: it is not compiled into libeicos_b200.so and nothing in the product loads it.
#pragma once

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#ifndef EICOS_EMU
#include <cuda_runtime.h>

#define EI_CUDA(call)                                                                                      \
    do                                                                                                     \
    {                                                                                                      \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call); \
    } while (0)

namespace eicos
{
namespace be
{
typedef cudaStream_t stream_t;
inline void set_device(int d) { EI_CUDA(cudaSetDevice(d)); }
inline void *alloc(size_t bytes)
{
    void *p = nullptr;
    EI_CUDA(cudaMalloc(&p, bytes ? bytes : 8));
    return p;
}
inline void dfree(void *p) { cudaFree(p); }
inline void h2d(void *d, const void *h, size_t n, stream_t s) { if (n) EI_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
inline void d2h(void *h, const void *d, size_t n, stream_t s) { if (n) EI_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline void zero(void *d, size_t n, stream_t s) { if (n) EI_CUDA(cudaMemsetAsync(d, 0, n, s)); }
inline void d2h_2d(void *h, size_t hpitch, const void *d, size_t dpitch, size_t width, size_t rows, stream_t s)
{
    if (width && rows)
        EI_CUDA(cudaMemcpy2DAsync(h, hpitch, d, dpitch, width, rows, cudaMemcpyDeviceToHost, s));
}
inline void sync(stream_t s) { EI_CUDA(cudaStreamSynchronize(s)); }
inline stream_t make_stream()
{
    cudaStream_t s;
    EI_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return s;
}
inline void drop_stream(stream_t s) { cudaStreamDestroy(s); }
inline void *pinned(size_t n)
{
    void *p = nullptr;
    EI_CUDA(cudaMallocHost(&p, n));
    return p;
}
inline void unpin(void *p) { cudaFreeHost(p); }
} // namespace be
} // namespace eicos

#else // ------------------------------------------------------------ emulator

namespace eicos
{
namespace be
{
typedef int stream_t;
inline void set_device(int) {}
inline void *alloc(size_t bytes) { return std::calloc(bytes ? bytes : 8, 1); }
inline void dfree(void *p) { std::free(p); }
inline void h2d(void *d, const void *h, size_t n, stream_t) { if (n) std::memcpy(d, h, n); }
inline void d2h(void *h, const void *d, size_t n, stream_t) { if (n) std::memcpy(h, d, n); }
inline void zero(void *d, size_t n, stream_t) { if (n) std::memset(d, 0, n); }
inline void d2h_2d(void *h, size_t hpitch, const void *d, size_t dpitch, size_t width, size_t rows, stream_t)
{
    for (size_t r = 0; r < rows; r++)
        std::memcpy((char *)h + r * hpitch, (const char *)d + r * dpitch, width);
}
inline void sync(stream_t) {}
inline stream_t make_stream() { return 0; }
inline void drop_stream(stream_t) {}
inline void *pinned(size_t n) { return std::calloc(n, 1); }
inline void unpin(void *p) { std::free(p); }
} // namespace be
} // namespace eicos
#endif
