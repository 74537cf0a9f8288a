// cuda_emu.h -- TEST HARNESS ONLY.  A minimal CPU stand-in for the CUDA execution model so that the
// index arithmetic of the kernels in this directory can be exercised by `pytest -m "not gpu"` in a
// container that has no GPU.  It is compiled ONLY into tests/_emu/libainmf_emu.so (see
// audio-inpainting_b200/build.py: build_emulator), never into libainmf.so, and the product loader
// (audio-inpainting_b200/_lib.py) cannot load it.  It is not a fallback: it is ~1000x slower than
// one CPU core running numpy and exists to catch indexing bugs before GPU time is spent.
//
// Model: blocks run one after another; the threads of a block are real OS threads;
// __syncthreads() is a barrier; warp shuffles exchange through a per-warp mailbox.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

namespace emu {
struct Ctx {
    dim3 grid, block;
    unsigned nthreads = 0;
    std::unique_ptr<std::barrier<>> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<uint64_t> mailbox;     // [warp][32]
    std::vector<unsigned char> dyn_smem;
};
inline Ctx*& ctx() { static Ctx* c = nullptr; return c; }
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline thread_local unsigned t_linear = 0;
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::ctx()->block)
#define gridDim (emu::ctx()->grid)

static inline void __syncthreads() { emu::ctx()->block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::ctx()->warp_bar[emu::t_linear / 32]->arrive_and_wait(); }
static inline long long clock64() { return 0; }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

namespace emu {
template <class T>
inline T exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Ctx* c = ctx();
    unsigned w = t_linear / 32, l = t_linear % 32;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    c->mailbox[w * 32 + l] = raw;
    c->warp_bar[w]->arrive_and_wait();
    uint64_t got = c->mailbox[w * 32 + (src_lane & 31)];
    c->warp_bar[w]->arrive_and_wait();
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}
}  // namespace emu
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    int l = emu::t_linear % 32; (void)width; return emu::exchange(v, l ^ m);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int l = emu::t_linear % 32; int base = l & ~(width - 1); return emu::exchange(v, base + (src & (width - 1)));
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d, int width = 32) {
    int l = emu::t_linear % 32; int s = l + d; if ((s & ~(width - 1)) != (l & ~(width - 1))) s = l; return emu::exchange(v, s);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d, int width = 32) {
    int l = emu::t_linear % 32; int s = l - d; if (s < (l & ~(width - 1))) s = l; return emu::exchange(v, s);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned mine = pred ? 1u : 0u, out = 0;
    // gather every lane's bit
    emu::Ctx* c = emu::ctx();
    unsigned w = emu::t_linear / 32, l = emu::t_linear % 32;
    c->mailbox[w * 32 + l] = mine;
    c->warp_bar[w]->arrive_and_wait();
    for (int i = 0; i < 32; ++i) out |= (unsigned)(c->mailbox[w * 32 + i] & 1u) << i;
    c->warp_bar[w]->arrive_and_wait();
    return out;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline void sincospif(float x, float* s, float* c) { *s = sinf((float)M_PI * x); *c = cosf((float)M_PI * x); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }

static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v) {
    unsigned* up = reinterpret_cast<unsigned*>(p);
    unsigned old = __atomic_load_n(up, __ATOMIC_SEQ_CST), nw;
    float f;
    do { std::memcpy(&f, &old, 4); f += v; std::memcpy(&nw, &f, 4); }
    while (!__atomic_compare_exchange_n(up, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    std::memcpy(&f, &old, 4);
    return f;
}
static inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}

// ---- runtime API subset ---------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0) {
    for (size_t i = 0; i < h; ++i) std::memmove((char*)d + i * dp, (const char*)s + i * sp, w);
    return 0;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
typedef void* cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)1; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)1; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
struct cudaDeviceProp { int multiProcessorCount; int major, minor; size_t sharedMemPerBlockOptin; int l2CacheSize; char name[64]; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    p->multiProcessorCount = 4; p->major = 10; p->minor = 0; p->sharedMemPerBlockOptin = 227 * 1024; p->l2CacheSize = 1 << 20;
    std::snprintf(p->name, sizeof p->name, "cpu-emulator");
    return 0;
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

namespace emu {
template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 block, size_t smem, A... args) {
    Ctx c;
    c.grid = grid; c.block = block;
    c.nthreads = block.x * block.y * block.z;
    if (c.nthreads == 0 || c.nthreads % 32 != 0 || c.nthreads > 1024) { std::fprintf(stderr, "emu: bad block size %u\n", c.nthreads); std::abort(); }
    c.block_bar.reset(new std::barrier<>(c.nthreads));
    for (unsigned w = 0; w < c.nthreads / 32; ++w) c.warp_bar.emplace_back(new std::barrier<>(32));
    c.mailbox.assign(c.nthreads, 0);
    c.dyn_smem.assign(smem + 1024, 0);
    ctx() = &c;
    auto body = [&](unsigned lin) {
        t_linear = lin;
        t_threadIdx = {lin % block.x, (lin / block.x) % block.y, lin / (block.x * block.y)};
        for (unsigned bz = 0; bz < grid.z; ++bz)
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    t_blockIdx = {bx, by, bz};
                    kernel(args...);
                    c.block_bar->arrive_and_wait();   // statics ("shared") are reused by the next block
                }
    };
    std::vector<std::thread> th;
    for (unsigned i = 0; i < c.nthreads; ++i) th.emplace_back(body, i);
    for (auto& t : th) t.join();
    ctx() = nullptr;
}
inline unsigned char* dyn_smem_ptr() {
    uintptr_t p = (uintptr_t)ctx()->dyn_smem.data();
    return (unsigned char*)((p + 1023) & ~(uintptr_t)1023);
}
}  // namespace emu

namespace ainmf { extern unsigned long long g_launch_count; }
#define AINMF_LAUNCH(kernel, grid, block, smem, stream, ...) (++ainmf::g_launch_count, emu::launch(kernel, grid, block, smem, __VA_ARGS__))
#define AINMF_DYN_SMEM(name) unsigned char* name = emu::dyn_smem_ptr()
