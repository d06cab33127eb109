// Runtime glue shared by every translation unit of libps3d_cuda.
//
// Product build (nvcc, sm_100a): thin checked wrappers over the CUDA runtime.
// Test-only build (-DPS3D_EMU, plain g++): the same kernel sources run on a
// fibre-based CPU model of a thread block (emu.h) so that index arithmetic can
// be unit-tested without a GPU.  The emulated library is built into
// tests/_emu/ and is never loaded by the ps3d_b200 package: the product has no
// CPU fallback.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>

#ifdef PS3D_EMU
#include "emu.h"
#else
#include <cuda_runtime.h>
#endif

namespace ps3d {

extern std::string g_last_error;
void set_error(const char* fmt, ...);

#ifndef PS3D_EMU
#define PS_CUDA_TRY(expr)                                                          \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            ::ps3d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                 \
            throw ::ps3d::DeviceError();                                           \
        }                                                                          \
    } while (0)
#endif

struct DeviceError {};

// 8-byte asynchronous copy global -> shared (LDGSTS) and the wait for this thread's outstanding copies; the
// test-only emulator copies at once
#ifdef PS3D_EMU
inline void ps_cp_async8(double* smem_dst, const double* gsrc) { *smem_dst = *gsrc; }
inline void ps_cp_async_wait() {}
#else
__device__ __forceinline__ void ps_cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ps_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

#ifdef PS3D_EMU
typedef int ps_stream_t;
#define PS_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ::emu::launch((grid), (block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define PS_SMEM(type, name) type* name = reinterpret_cast<type*>(::emu::smem_base())
inline void* ps_malloc(size_t bytes) { void* p = calloc(bytes ? bytes : 1, 1); if (!p) { set_error("calloc(%zu) failed", bytes); throw DeviceError(); } return p; }
inline void ps_free(void* p) { free(p); }
inline void ps_h2d(void* d, const void* h, size_t n, ps_stream_t) { memcpy(d, h, n); }
inline void ps_d2h(void* h, const void* d, size_t n, ps_stream_t) { memcpy(h, d, n); }
inline void ps_d2d(void* d, const void* s, size_t n, ps_stream_t) { memmove(d, s, n); }
inline void ps_memset(void* d, int v, size_t n, ps_stream_t) { memset(d, v, n); }
inline void ps_sync(ps_stream_t) {}
inline void ps_check_launch() {}
#else
typedef cudaStream_t ps_stream_t;
#define PS_UNPAREN(...) __VA_ARGS__
// PS3D_TRACE=1: CUDA events around every launch, aggregated by kernel name at ps3d_cuda_finalise (in-situ times
// of a whole run, launch gaps included; development aid, off by default)
void ps_trace_begin(const char* name, cudaStream_t s);
void ps_trace_end(cudaStream_t s);
extern int g_trace;
#define PS_LAUNCH(kernel, grid, block, smem, stream, ...)                    \
    do {                                                                     \
        if (::ps3d::g_trace) ::ps3d::ps_trace_begin(#kernel, (stream));      \
        PS_UNPAREN kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
        if (::ps3d::g_trace) ::ps3d::ps_trace_end((stream));                 \
        ::ps3d::ps_check_launch();                                           \
    } while (0)
#define PS_SMEM(type, name) extern __shared__ __align__(16) unsigned char _ps_smem_raw[]; \
    type* name = reinterpret_cast<type*>(_ps_smem_raw)
inline void* ps_malloc(size_t bytes) {
    void* p = nullptr;
    PS_CUDA_TRY(cudaMalloc(&p, bytes ? bytes : 1));
    PS_CUDA_TRY(cudaMemset(p, 0, bytes ? bytes : 1));
    // the memset runs on the legacy default stream, which the library's non-blocking stream does not
    // synchronise with: finish it before the buffer can be used
    PS_CUDA_TRY(cudaDeviceSynchronize());
    return p;
}
inline void ps_free(void* p) { if (p) cudaFree(p); }
inline void ps_h2d(void* d, const void* h, size_t n, ps_stream_t s) { PS_CUDA_TRY(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
inline void ps_d2h(void* h, const void* d, size_t n, ps_stream_t s) { PS_CUDA_TRY(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline void ps_d2d(void* d, const void* s_, size_t n, ps_stream_t s) { PS_CUDA_TRY(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); }
inline void ps_memset(void* d, int v, size_t n, ps_stream_t s) { PS_CUDA_TRY(cudaMemsetAsync(d, v, n, s)); }
inline void ps_sync(ps_stream_t s) { PS_CUDA_TRY(cudaStreamSynchronize(s)); }
inline void ps_check_launch() { PS_CUDA_TRY(cudaGetLastError()); }
#endif

}  // namespace ps3d
