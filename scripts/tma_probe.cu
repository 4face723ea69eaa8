// Probe which TMA box shapes / coordinates are legal on this device (debug aid, not part of the library).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, int c1, int c2, uint32_t bytes, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    uint32_t b = (uint32_t) __cvta_generic_to_shared(&bar), d = (uint32_t) __cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d), "l"(&m), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
        uint32_t done = 0; int it = 0;
        while (!done && it++ < 1000000) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(b) : "memory");
        out[0] = done; out[1] = smem[0] | (smem[1] << 8) | (smem[2] << 16) | (smem[3] << 24); out[2] = d;
    }
}
int main() {
    void* fn; cudaDriverEntryPointQueryResult q; cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q); PFN_enc enc = (PFN_enc) fn;
    void* buf; cudaMalloc(&buf, 1 << 22); cudaMemset(buf, 0x11, 1 << 22);
    uint32_t* out; cudaMalloc(&out, 64);
    struct T { const char* name; CUtensorMapDataType t; int es; int d[3]; int b[3]; int c[3]; } tests[] = {
        {"f32 box(64,8,4)", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, {32,32,32}, {64,8,4}, {0,0,0}},
        {"f32 box(4,64,8)", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, {32,32,32}, {4,64,8}, {0,0,16}},
        {"f32 box(8,64,8)", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, {32,32,32}, {8,64,8}, {0,0,16}},
        {"u8 box(80,10,6) c(-1,15,-1)", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, {32,32,32}, {80,10,6}, {-1,15,-1}},
        {"u8 box(80,10,6) c(-1,-1,-1)", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, {32,32,32}, {80,10,6}, {-1,-1,-1}},
        {"u8 box(80,9,5) c(0,16,-1)", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, {32,32,32}, {80,9,5}, {0,16,-1}},
        {"u8 box(80,6,10) c(-1,-1,15)", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, {32,32,32}, {80,6,10}, {-1,-1,15}},
        {"u8 box(80,10,6) c(-1,24,-1)", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, {32,32,32}, {80,10,6}, {-1,24,-1}},
        {"u8 box(80,10,6) c(-1,31,-1)", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, {32,32,32}, {80,10,6}, {-1,31,-1}},
    };
    for (auto& t : tests) {
        CUtensorMap m; cuuint64_t gd[3] = {(cuuint64_t)t.d[0], (cuuint64_t)t.d[1], (cuuint64_t)t.d[2]}; cuuint64_t gs[2] = {(cuuint64_t)t.d[0]*t.es, (cuuint64_t)t.d[0]*t.d[1]*t.es};
        cuuint32_t bd[3] = {(cuuint32_t)t.b[0], (cuuint32_t)t.b[1], (cuuint32_t)t.b[2]}, es[3] = {1,1,1};
        CUresult r = enc(&m, t.t, 3, buf, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        uint32_t bytes = t.b[0]*t.b[1]*t.b[2]*t.es;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        k<<<1, 32, bytes + 128>>>(m, t.c[0], t.c[1], t.c[2], bytes, out);
        cudaError_t e = cudaDeviceSynchronize(); uint32_t h[3] = {0,0,0}; if (e == cudaSuccess) cudaMemcpy(h, out, 12, cudaMemcpyDeviceToHost);
        printf("%-32s encode=%d run=%s done=%u first=%08x smem=%u\n", t.name, (int) r, cudaGetErrorString(e), h[0], h[1], h[2]);
        if (e != cudaSuccess) { printf("(context dead, stopping)\n"); break; }
    }
    return 0;
}
