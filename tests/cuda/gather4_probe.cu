// gather4_probe.cu -- micro-probe of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a:
// which box shape the tensor map needs, where the 4 gathered rows land in shared memory (with and without
// SWIZZLE_128B) and what an out-of-range row index yields.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, float *out, int *flag)
{
    __shared__ __align__(1024) float tile[4 * 32 * 2];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 256; ++i) tile[i] = -1.f;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(512) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            ::"r"(smem_u32(tile)), "l"(&map), "r"(smem_u32(&bar)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
        int ok = 0;
        for (uint32_t it = 0; it < (1u << 22); ++it) {
            uint32_t done;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
            if (done) { ok = 1; break; }
        }
        *flag = ok;
        for (int i = 0; i < 256; ++i) out[i] = tile[i];
    }
}

int main()
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)p;
    const int R = 64, C = 32;
    float h[R * C];
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = r * 100.f + c;
    float *d, *out; int *flag;
    cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 256 * 4); cudaMalloc(&flag, 4);
    for (int variant = 0; variant < 4; variant += 2) {
        cuuint32_t boxrows = (variant & 1) ? 4 : 1;
        CUtensorMapSwizzle sw = (variant & 2) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
        CUtensorMap m;
        cuuint64_t dims[2] = {C, R}; cuuint64_t str[1] = {C * 4}; cuuint32_t box[2] = {C, boxrows}; cuuint32_t es[2] = {1, 1};
        CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("variant %d: boxrows=%u swizzle=%s encode rc=%d\n", variant, boxrows, (variant & 2) ? "128B" : "none", (int)rc);
        if (rc != CUDA_SUCCESS) continue;
        cudaMemset(out, 0, 1024); cudaMemset(flag, 0, 4);
        probe<<<1, 32>>>(m, 5, 17, 60, 200, out, flag);
        cudaError_t e = cudaDeviceSynchronize();
        float o[256]; int f = 0;
        if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return 1; }
        cudaMemcpy(o, out, 1024, cudaMemcpyDeviceToHost); cudaMemcpy(&f, flag, 4, cudaMemcpyDeviceToHost);
        printf("  barrier completed=%d\n", f);
        for (int r = 0; r < 5; ++r) {
            printf("  smem row %d:", r);
            for (int c = 0; c < 32; c += 4) printf(" %g", o[r * 32 + c]);
            printf("\n");
        }
    }
    return 0;
}
