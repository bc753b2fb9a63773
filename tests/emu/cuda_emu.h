// TEST INFRASTRUCTURE ONLY -- a tiny SIMT emulator that runs the repo's CUDA-core kernels FROM THEIR OWN SOURCE on the host CPU
// (tests/test_kernel_emulation_cpu.py), so that index arithmetic, shared-memory carve-ups, warp reductions and barrier
// placement of a kernel can be checked in a container without a GPU.  One OS thread per CUDA thread of a block, blocks one
// after another; __syncthreads / __syncwarp / __shfl_*_sync are real barriers, so a missing or misplaced synchronisation
// shows up as a wrong result or a data race here too (not guaranteed, but the schedules differ enough to be a useful probe).
// Not emulated: tcgen05 / TMA / mbarrier / inline PTX (those kernels are validated on the GPU only), textures.
// The product never includes this file.
#pragma once
#include <cuda_runtime.h>   // vector types and make_* only: no CUDA runtime call is made (nothing links against libcudart)
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <thread>
#include <vector>

#undef __launch_bounds__
#define __launch_bounds__(...)

// ---- execution context ---------------------------------------------------------------------------------------------
struct EmuWarp {
  pthread_barrier_t bar;
  uint32_t slot[32];
};
struct EmuBlock {
  pthread_barrier_t bar;
  std::vector<EmuWarp> warps;
};
inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;
inline thread_local EmuBlock* emu_block = nullptr;
inline thread_local EmuWarp* emu_warp = nullptr;
inline thread_local int emu_lane = 0;
constexpr int warpSize = 32;

inline void __syncthreads() { pthread_barrier_wait(&emu_block->bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu_warp->bar); }

template <class T>
inline T emu_exchange(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  memcpy(&emu_warp->slot[emu_lane], &v, 4);
  pthread_barrier_wait(&emu_warp->bar);
  T r;
  memcpy(&r, &emu_warp->slot[src_lane & 31], 4);
  pthread_barrier_wait(&emu_warp->bar);
  return r;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int mask) { return emu_exchange(v, emu_lane ^ mask); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) { return emu_exchange(v, emu_lane + d < 32 ? emu_lane + d : emu_lane); }

// ---- intrinsics the CUDA-core kernels use ----------------------------------------------------------------------------
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float rsqrtf(float v) { return 1.0f / sqrtf(v); }
inline float __saturatef(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float v) { return expf(v); }
inline float __fmaf_rz(float a, float b, float c) {           // round-toward-zero fma via long double (exact enough for the tests)
  const long double e = (long double)a * b + c;
  float r = (float)e;
  if ((long double)r != e && ((r > 0) == (e > 0)) && fabsl((long double)r) > fabsl(e)) r = nextafterf(r, 0.f);
  return r;
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }      // one rounding per operation, never contracted
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicMax(int* p, int v) { int o = *p; while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline float atomicAdd(float* p, float v) {                    // CAS loop: order-dependent rounding, like the hardware
  uint32_t* u = reinterpret_cast<uint32_t*>(p);
  uint32_t o = __atomic_load_n(u, __ATOMIC_RELAXED), nw;
  float f;
  do { memcpy(&f, &o, 4); f += v; memcpy(&nw, &f, 4); } while (!__atomic_compare_exchange_n(u, &o, nw, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  memcpy(&f, &o, 4);
  return f;
}
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline double min(double a, double b) { return fmin(a, b); }
inline double max(double a, double b) { return fmax(a, b); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

// dynamic shared memory: `extern __shared__ T name[];` is rewritten into a pointer to this pool (blocks run one at a time)
alignas(128) inline unsigned char emu_smem_pool[256 * 1024];

// ---- launches ----------------------------------------------------------------------------------------------------------
// `kernel<<<grid, block, smem, stream>>>(args...)` is rewritten by the test into emu_launch(grid, block, smem, [=] { kernel(args...); })
inline dim3 emu_dim(dim3 d) { return d; }
inline dim3 emu_dim(unsigned x) { return dim3(x, 1, 1); }
inline dim3 emu_dim(int x) { return dim3((unsigned)x, 1, 1); }
inline dim3 emu_dim(size_t x) { return dim3((unsigned)x, 1, 1); }

template <class G, class B>
inline void emu_launch(G grid_, B block_, size_t smem_bytes, const std::function<void()>& body) {
  const dim3 grid = emu_dim(grid_), block = emu_dim(block_);
  if (smem_bytes > sizeof(emu_smem_pool)) { fprintf(stderr, "emu: %zu bytes of dynamic shared memory requested\n", smem_bytes); abort(); }
  const unsigned nthreads = block.x * block.y * block.z, nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        EmuBlock blk;
        blk.warps.resize(nwarps);
        pthread_barrier_init(&blk.bar, nullptr, nthreads);
        for (unsigned w = 0; w < nwarps; ++w) {
          const unsigned lanes = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
          pthread_barrier_init(&blk.warps[w].bar, nullptr, lanes);
        }
        std::vector<std::thread> threads;
        threads.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
          threads.emplace_back([&, t] {
            threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            blockIdx = uint3{bx, by, bz};
            blockDim = block;
            gridDim = grid;
            emu_block = &blk;
            emu_warp = &blk.warps[t / 32];
            emu_lane = (int)(t % 32);
            body();
          });
        for (auto& th : threads) th.join();
        for (auto& w : blk.warps) pthread_barrier_destroy(&w.bar);
        pthread_barrier_destroy(&blk.bar);
      }
}

// ---- the few CUDA runtime calls the entry points make, on host memory (nothing links against libcudart) ----------------------
extern "C" {
inline cudaError_t cudaGetLastError(void) { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulator"; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 2; return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
}

template <class T> inline cudaError_t cudaFuncSetAttribute(T*, cudaFuncAttribute, int) { return cudaSuccess; }   // kernel pointers

// ---- host-side plumbing of common.cuh that needs a device ------------------------------------------------------------------
#ifndef EMU_API_TU            // api.cu defines these two itself
inline int b200r_num_sms() { return 1; }     // grid caps are multiples of the SM count: keep the emulated grids small
inline thread_local char emu_error[512];
inline void b200r_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(emu_error, sizeof(emu_error), fmt, ap);
  va_end(ap);
}
#endif
