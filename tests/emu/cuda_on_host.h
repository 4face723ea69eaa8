// cuda_on_host.h — TEST INFRASTRUCTURE: a small SIMT emulator that lets the product's CUDA sources (tbraymarcherplugin_b200/csrc/*.cu)
// be compiled by g++ and run on the CPU, so that a machine without a GPU can execute the kernels' own source — indexing, bounds, barriers,
// launch geometry, host drivers — against the oracle (tests/test_kernels_emulated_cpu.py). It is not a fallback of the product: nothing in
// tbraymarcherplugin_b200/ knows about it, libtbrm.so is never built from it, and only tests/ load the library it produces
// (tests/emu/_build/libtbrm_emu.so, built by tests/emu/build_emu.py).
//
// Execution model: a launch runs its blocks one after the other on the calling OS thread; the threads of a block are ucontext fibers that are
// resumed round-robin. __syncthreads() and the full-mask warp shuffles are barriers between fibers (exited threads count as arrived, as on the
// device). A cooperative launch makes the fibers of ALL blocks co-resident, and every relaxed / volatile global load of a spin loop
// (tbrm_emu::spin_hint()) yields, so inter-block flag protocols make progress. Everything is synchronous: streams and events are no-ops.
//
// Arithmetic: the sources are compiled with -ffp-contract=off, __fmaf_rn is fmaf, sqrt and division are IEEE — the same fp32 contract as the
// device build (--fmad=false), so results are expected to equal the oracle bit for bit wherever the GPU's do. Transcendentals come from libm
// here and from CUDA's math library on the device (the Mandelbulb's trigonometric formulation): same tolerance as the GPU tests.
//
// Device memory: cudaMalloc places every allocation so that it ends at a PROT_NONE guard page (16-byte slack at most), and starts after one:
// a kernel that reads or writes past an allocation faults instead of passing silently.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <vector>

#define TBRM_HOST_EMULATION 1

// ---- qualifiers --------------------------------------------------------------------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

// ---- vector types ------------------------------------------------------------------------------------------------------------------------
struct dim3 {
    unsigned int x = 1, y = 1, z = 1;
    dim3() = default;
    dim3(unsigned int x_, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
#define TBRM_EMU_VEC(T, name)                                                             \
    struct name##1 { T x; };                                                              \
    struct alignas(2 * sizeof(T)) name##2 { T x, y; };                                    \
    struct name##3 { T x, y, z; };                                                        \
    struct alignas(4 * sizeof(T) > 16 ? 16 : 4 * sizeof(T)) name##4 { T x, y, z, w; };    \
    static inline name##2 make_##name##2(T x, T y) { return name##2{x, y}; }              \
    static inline name##3 make_##name##3(T x, T y, T z) { return name##3{x, y, z}; }      \
    static inline name##4 make_##name##4(T x, T y, T z, T w) { return name##4{x, y, z, w}; }
TBRM_EMU_VEC(float, float)
TBRM_EMU_VEC(int, int)
TBRM_EMU_VEC(unsigned int, uint)
TBRM_EMU_VEC(unsigned char, uchar)
TBRM_EMU_VEC(unsigned short, ushort)
TBRM_EMU_VEC(short, short)
TBRM_EMU_VEC(char, char)
struct alignas(16) ulonglong2 {
    unsigned long long x, y;
};
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

// ---- runtime API (synchronous, one device) -------------------------------------------------------------------------------------------------
enum cudaError_t {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorNotSupported = 801,
    cudaErrorUnknown = 999
};
typedef struct tbrm_emu_stream* cudaStream_t;
typedef struct tbrm_emu_event* cudaEvent_t;
#define cudaStreamPerThread ((cudaStream_t) 2)
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrCooperativeLaunch = 95 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1, cudaEnableDefault = 0 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
struct cudaIpcMemHandle_t {
    char reserved[64];
};

namespace tbrm_emu {
void* device_alloc(size_t bytes);
void device_free(void* p);
struct Event {
    std::chrono::steady_clock::time_point t;
};
// 1: cooperative launches are offered (inter-block protocols run with all blocks' fibers co-resident)
int cooperative_supported();
}  // namespace tbrm_emu

static inline cudaError_t cudaMalloc(void** p, size_t bytes) {
    *p = tbrm_emu::device_alloc(bytes);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    return cudaMalloc((void**) p, bytes);
}
static inline cudaError_t cudaFree(void* p) {
    tbrm_emu::device_free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) { return cudaMemcpy(d, s, n, k); }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) {
    if (n) memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { return cudaMemset(d, v, n); }
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaGetDevice(int* d) {
    *d = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceCount(int* n) {
    *n = 1;
    return cudaSuccess;
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorNotSupported ? "operation not supported" : "emulated CUDA error"); }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) {
    *s = (cudaStream_t) malloc(8);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
    free(s);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = (cudaEvent_t) new tbrm_emu::Event();
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete (tbrm_emu::Event*) e;
    return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    ((tbrm_emu::Event*) e)->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(((tbrm_emu::Event*) b)->t - ((tbrm_emu::Event*) a)->t).count();
    return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    *v = a == cudaDevAttrMultiProcessorCount ? 148 : (a == cudaDevAttrCooperativeLaunch ? tbrm_emu::cooperative_supported() : 0);
    return cudaSuccess;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) {
    *n = 4;
    return cudaSuccess;
}
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long flags, cudaDriverEntryPointQueryResult* q = nullptr);

// ---- driver types of the tensor-map API (cuda.h) -----------------------------------------------------------------------------------------------
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
enum CUresult { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
struct alignas(64) CUtensorMap {
    cuuint64_t opaque[16];
};
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT8 = 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
namespace tbrm_emu {
constexpr uint64_t kTensorMapMagic = 0x74626d7470616d31ull;
struct TensorMap {  // what the emulated cuTensorMapEncodeTiled keeps in the 128 opaque bytes (rank 3, tiled, no interleave / swizzle)
    uint64_t magic;
    void* base;
    uint64_t dims[3];
    uint64_t strides[2];  // bytes, of dimensions 1 and 2
    uint32_t box[3];
    uint32_t elem;
};
static_assert(sizeof(TensorMap) <= sizeof(CUtensorMap), "the emulated tensor map must fit the opaque one");
// the driver's argument checks that matter to the sources: 16-byte aligned base and strides, box sides 1..256, inner box extent a multiple of 16 bytes
CUresult encode_tiled(CUtensorMap* out, CUtensorMapDataType type, cuuint32_t rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t* elem_strides, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                      CUtensorMapFloatOOBfill);
}  // namespace tbrm_emu

// ---- SIMT engine -------------------------------------------------------------------------------------------------------------------------
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

namespace tbrm_emu {
void run_grid(dim3 grid, dim3 block, size_t smem_bytes, bool cooperative, const std::function<void()>& thread_body);
void sync_block();
void sync_warp();                    // barrier of the alive lanes of the calling thread's warp
unsigned long long* warp_slot(int lane);  // 8-byte exchange slot of a lane of the calling thread's warp
int lane_id();
bool lane_alive(int lane);
void spin_hint();                    // inside a spin loop: let the other fibers run
unsigned char* dynamic_smem();       // the calling block's dynamic shared memory (zero-length launches get a valid pointer too)
[[noreturn]] void fail(const char* what);

unsigned long long globaltimer_ns();

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, F&& body) {
    run_grid(grid, block, smem, false, std::function<void()>(body));
}
// cudaLaunchCooperativeKernel with the kernel's function-pointer type kept: all blocks co-resident
template <class... A, size_t... I>
inline void call_with(void (*kernel)(A...), void** args, std::index_sequence<I...>) {
    kernel(*(typename std::remove_cv<typename std::remove_reference<A>::type>::type*) args[I]...);
}
template <class... A>
inline cudaError_t launch_cooperative(void (*kernel)(A...), dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t) {
    run_grid(grid, block, smem, true, std::function<void()>([=] { call_with(kernel, args, std::index_sequence_for<A...>()); }));
    return cudaSuccess;
}
// the inline-PTX global accesses of the flag / ring protocols (ld.acquire / ld.relaxed, st.release / st.relaxed at gpu or sys scope):
// acquire / release atomics; a load also yields, because every spin loop of the sources polls through one
template <class T>
inline T ld_global(const T* p) {
    const T v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
    spin_hint();
    return v;
}
template <class T, class V>
inline void st_global(T* p, V v) {
    __atomic_store_n(p, (T) v, __ATOMIC_RELEASE);
}
}  // namespace tbrm_emu

static inline void __syncthreads() { tbrm_emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { tbrm_emu::sync_warp(); }
static inline int __syncthreads_and(int pred) {  // three barriers: vote, read, reset (fibers of a block share the accumulator)
    static thread_local int acc = 1;
    if (!pred) acc = 0;
    tbrm_emu::sync_block();
    const int r = acc;
    tbrm_emu::sync_block();
    acc = 1;
    tbrm_emu::sync_block();
    return r;
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8 && std::is_trivially_copyable<T>::value, "shuffle of a type wider than 64 bits");
    if (mask != 0xffffffffu) tbrm_emu::fail("__shfl_xor_sync: only full-mask shuffles are emulated");
    const int lane = tbrm_emu::lane_id(), src = lane ^ lane_mask;
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    *tbrm_emu::warp_slot(lane) = bits;
    tbrm_emu::sync_warp();
    T out = v;  // a source lane outside the warp returns the caller's own value
    if (src < 32) {
        if (!tbrm_emu::lane_alive(src)) tbrm_emu::fail("__shfl_xor_sync: the source lane has exited (undefined on the device)");
        const unsigned long long got = *tbrm_emu::warp_slot(src);
        memcpy(&out, &got, sizeof(T));
    }
    tbrm_emu::sync_warp();
    return out;
}

// ---- intrinsics --------------------------------------------------------------------------------------------------------------------------
template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __ldcg(const T* p) { return *p; }
template <class T>
static inline T __ldcs(const T* p) { return *p; }
template <class T>
static inline void __stcg(T* p, T v) { *p = v; }
template <class T>
static inline void __stcs(T* p, T v) { *p = v; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline int __float_as_int(float f) {
    int i;
    memcpy(&i, &f, 4);
    return i;
}
static inline unsigned int __float_as_uint(float f) {
    unsigned int i;
    memcpy(&i, &f, 4);
    return i;
}
static inline float __int_as_float(int i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}
static inline float __uint_as_float(unsigned int i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}
static inline unsigned int __funnelshift_r(unsigned int lo, unsigned int hi, unsigned int shift) {
    return (unsigned int) ((((unsigned long long) hi << 32) | lo) >> (shift & 31u));
}
static inline unsigned int __byte_perm(unsigned int a, unsigned int b, unsigned int sel) {  // PRMT, default mode (no sign replication)
    const unsigned long long v = ((unsigned long long) b << 32) | a;
    unsigned int r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned int) ((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffull) << (8 * i);
    return r;
}
static inline unsigned int __vmaxu4(unsigned int a, unsigned int b) {  // per-byte unsigned maximum
    unsigned int r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned int x = (a >> (8 * i)) & 0xffu, y = (b >> (8 * i)) & 0xffu;
        r |= (x > y ? x : y) << (8 * i);
    }
    return r;
}
// binary16 (round to nearest even), through the compiler's _Float16
struct __half {
    _Float16 v;
};
typedef __half half;
static inline __half __float2half_rn(float f) { return __half{(_Float16) f}; }
static inline __half __float2half(float f) { return __half{(_Float16) f}; }
static inline float __half2float(__half h) { return (float) h.v; }

// CUDA's overloaded integer / mixed min and max
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }

// atomics on global / shared memory: blocks run on one OS thread, fibers switch only at barriers and spin hints
template <class T>
static inline T atomicAdd(T* p, T v) {
    const T old = *p;
    *p = old + v;
    return old;
}
template <class T>
static inline T atomicMax(T* p, T v) {
    const T old = *p;
    if (v > old) *p = v;
    return old;
}
template <class T>
static inline T atomicMin(T* p, T v) {
    const T old = *p;
    if (v < old) *p = v;
    return old;
}
template <class T>
static inline T atomicExch(T* p, T v) {
    const T old = *p;
    *p = v;
    return old;
}
template <class T>
static inline T atomicOr(T* p, T v) {
    const T old = *p;
    *p = old | v;
    return old;
}
template <class T>
static inline T atomicCAS(T* p, T cmp, T v) {
    const T old = *p;
    if (old == cmp) *p = v;
    return old;
}
