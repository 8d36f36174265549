// mma_probe.cu -- micro-benchmark of tcgen05.mma issue cost on sm_100a (no data movement: operands are whatever
// sits in shared memory).  One warp issues `reps` rounds of a pattern of MMAs, commits, waits, and reports
// cycles per MMA.  Answers: what does a kind::tf32 MMA cost as a function of N, of operand sharing between
// consecutive MMAs, and of the A collector hints -- the numbers the kernels' tile shapes are chosen from
// (DESIGN.md section 6).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o mma_probe mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a)
{
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
__host__ __device__ constexpr uint32_t idesc_bf16(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

#define MMA(KIND, COLL, d, a, b, i)                                                                         \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"                                           \
                 "tcgen05.mma.cta_group::1.kind::" KIND COLL " [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), \
                 "l"(b), "r"(i) : "memory")

struct Result { long long cycles; int mmas; };

// pattern ids
//  0: tf32 N=n, same A, same B          1: tf32 N=n, A and B rotate over 3 distinct tiles each
//  2: tf32 3xTF32 triple (a_lo,b_hi)(a_hi,b_lo)(a_hi,b_hi) at N=n
//  3: tf32 concat pair (a_hi,[b_hi;b_lo] N=2n) (a_lo,b_hi N=n)
//  4: tf32 N=n pair sharing A with collector fill/lastuse     5: same pair without hints
//  6: bf16 N=n same A,B                                        7: tf32 N=n rotating over 2 accumulators
__global__ void __launch_bounds__(128, 1) probe(int pattern, int n, int reps, Result *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = 1e-3f * (i & 255);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);
    if (warp == 1) {
        const uint32_t s0 = smem_u32(smem);
        // A tiles: 16 KB each at 0, 16K, 32K;  B tiles: 32 KB each at 48K, 80K, 112K (N up to 256)
        const uint64_t a0 = desc_sw128(s0), a1 = desc_sw128(s0 + 16384), a2 = desc_sw128(s0 + 32768);
        const uint64_t b0 = desc_sw128(s0 + 49152), b1 = desc_sw128(s0 + 81920), b2 = desc_sw128(s0 + 114688);
        const uint32_t id = idesc_tf32(n), id2 = idesc_tf32(2 * n), idb = idesc_bf16(n);
        long long t0 = 0, t1 = 0;
        int count = 0;
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
        if (pred) {
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                const uint64_t ko = (uint64_t)((r & 3) * 2);          // K step inside the 128-byte row
                switch (pattern) {
                    case 0: MMA("tf32", "", tmem, a0 + ko, b0 + ko, id); count += 1; break;
                    case 1: {
                        const uint64_t a = (r % 3 == 0) ? a0 : (r % 3 == 1) ? a1 : a2;
                        const uint64_t b = (r % 3 == 0) ? b0 : (r % 3 == 1) ? b1 : b2;
                        MMA("tf32", "", tmem, a + ko, b + ko, id); count += 1; break;
                    }
                    case 2:
                        MMA("tf32", "", tmem, a1 + ko, b0 + ko, id);
                        MMA("tf32", "", tmem, a0 + ko, b1 + ko, id);
                        MMA("tf32", "", tmem, a0 + ko, b0 + ko, id); count += 3; break;
                    case 3:
                        MMA("tf32", "", tmem, a0 + ko, b0 + ko, id2);
                        MMA("tf32", "", tmem, a1 + ko, b0 + ko, id); count += 2; break;
                    case 4:
                        MMA("tf32", ".collector::a::fill", tmem, a0 + ko, b0 + ko, id);
                        MMA("tf32", ".collector::a::lastuse", tmem, a0 + ko, b1 + ko, id); count += 2; break;
                    case 5:
                        MMA("tf32", "", tmem, a0 + ko, b0 + ko, id);
                        MMA("tf32", "", tmem, a0 + ko, b1 + ko, id); count += 2; break;
                    case 6: MMA("f16", "", tmem, a0 + ko, b0 + ko, idb); count += 1; break;
                    case 7: MMA("tf32", "", tmem + (uint32_t)((r & 1) * 256), a0 + ko, b0 + ko, id); count += 1; break;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            for (uint32_t it = 0; it < (1u << 26); ++it) {
                uint32_t done;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
                if (done) break;
            }
            t1 = clock64();
            if (blockIdx.x == 0) { out->cycles = t1 - t0; out->mmas = count; }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

int main()
{
    Result *d;
    cudaMalloc(&d, sizeof(Result));
    const size_t smem = 161 * 1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char *names[] = {"tf32 same A,B", "tf32 rotating A,B (3 tiles)", "tf32 3xTF32 triple", "tf32 concat pair (2N + N)",
                           "tf32 pair shared A, collector fill/lastuse", "tf32 pair shared A, no hints", "bf16 same A,B",
                           "tf32 same A,B, 2 accumulators"};
    const int reps = 4000;
    for (int grid = 1; grid <= 148; grid += 147) {
        printf("== grid %d CTA(s), %d rounds per pattern\n", grid, reps);
        for (int pat = 0; pat < 8; ++pat)
            for (int n = 32; n <= 256; n *= 2) {
                if (pat == 3 && n > 128) continue;
                if (pat == 7 && n > 256) continue;
                Result h = {0, 0};
                cudaMemcpy(d, &h, sizeof(h), cudaMemcpyHostToDevice);
                probe<<<grid, 128, smem>>>(pat, n, reps, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("pattern %d n %d: %s\n", pat, n, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
                printf("%-44s N=%3d  %7.1f cycles/MMA  (%d MMAs)\n", names[pat], n, (double)h.cycles / h.mmas, h.mmas);
            }
    }
    return 0;
}
