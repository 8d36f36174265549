// mma_probe2.cu -- tcgen05.mma throughput with the operand ADDRESS PATTERNS of the real kernels and optional
// interference (shared-memory writers, TMEM readers).  Fully unrolled issue (no per-MMA control flow).
//   S3: sparse_conv_tc as first written: per K step (a_lo,b_hi)(a_hi,b_lo)(a_hi,b_hi), N = 64, 4 stages of 48 KB
//   S2: B-concatenation variant:          per K step (a_hi,[b_hi;b_lo] N=128)(a_lo,b_hi N=64)
//   C2: conv_tc2: N = 256, activation stage 72 KB x2, weight slots 16 KB x5, three vertical views
//   R1: one N-wide MMA repeated on the same operands (pipe rate), N = 64 / 128 / 256
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe2 mma_probe2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t dsc(uint32_t a)
{
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idt(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t i)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(i) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t *bar, uint32_t par)
{
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(par) : "memory");
        if (done) return;
    }
}

#ifndef RANDOM_DATA
#define RANDOM_DATA 1
#endif
struct Result { long long cycles; int mmas; };

// interference bit 1: warps 4-7 write shared memory continuously (st.shared.v4, 512 B per warp instruction)
// interference bit 2: warps 8-11 read TMEM continuously (tcgen05.ld 32x32b.x16)
template <int PAT, int N>
__global__ void __launch_bounds__(384, 1) probe(int reps, int interference, Result *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    __shared__ volatile int s_done;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 208 * 1024 / 4; i += blockDim.x) {
        // random sign/mantissa, magnitudes in [0.5, 2): realistic toggle rates (tensor-pipe power is data dependent)
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t *>(smem)[i] = RANDOM_DATA ? ((h & 0x807FE000u) | 0x3F000000u | ((h >> 3) & 0x00800000u))
                                                            : __float_as_uint(1e-3f * (i & 255));
    }
    if (threadIdx.x == 0) {
        s_done = 0;
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
        long long t0 = 0, t1 = 0;
        int count = 0;
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
        if (pred) {
            t0 = clock64();
            int st = 0, xs = 0, ws = 0;
            for (int r = 0; r < reps; ++r) {
                if (PAT == 0) {            // S3
                    const uint32_t sa = s0 + st * 49152;
                    const uint64_t ah = dsc(sa), al = dsc(sa + 16384), bh = dsc(sa + 32768), bl = dsc(sa + 32768 + N * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        mma(tmem, al + 2 * k, bh + 2 * k, idt(N));
                        mma(tmem, ah + 2 * k, bl + 2 * k, idt(N));
                        mma(tmem, ah + 2 * k, bh + 2 * k, idt(N));
                    }
                    count += 12;
                    st = (st + 1) & 3;
                } else if (PAT == 1) {     // S2
                    const uint32_t sa = s0 + st * 49152;
                    const uint64_t ah = dsc(sa), al = dsc(sa + 16384), bh = dsc(sa + 32768);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        mma(tmem, ah + 2 * k, bh + 2 * k, idt(2 * N));
                        mma(tmem, al + 2 * k, bh + 2 * k, idt(N));
                    }
                    count += 8;
                    st = (st + 1) & 3;
                } else if (PAT == 2) {     // C2 (N = 256): one (dx, chunk) group = 3 dy x 12 MMAs
                    const uint32_t sx = s0 + xs * 73728;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint64_t xh = dsc(sx + dy * 2048), xl = dsc(sx + 36864 + dy * 2048);
                        const uint64_t wh = dsc(s0 + 147456 + ws * 16384);
                        ws = (ws + 1) % 3;
                        const uint64_t wl = dsc(s0 + 147456 + ws * 16384);
                        ws = (ws + 1) % 3;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            mma(tmem, wh + 2 * k, xl + 2 * k, idt(256));
                            mma(tmem, wh + 2 * k, xh + 2 * k, idt(256));
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma(tmem, wl + 2 * k, xh + 2 * k, idt(256));
                    }
                    count += 36;
                    xs ^= 1;
                } else {                   // R1
#pragma unroll
                    for (int k = 0; k < 8; ++k) mma(tmem, dsc(s0) + 2 * (k & 3), dsc(s0 + 65536) + 2 * (k & 3), idt(N));
                    count += 8;
                }
            }
            commit(&bar);
            wait(&bar, 0);
            t1 = clock64();
            s_done = 1;
            if (blockIdx.x == 0) { out->cycles = t1 - t0; out->mmas = count; }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 8 && (interference & 1)) {
        // shared-memory writers into the last 8 KB of the operand area (not read by the MMAs of S3/S2/R1 at N<=64;
        // overlapping for C2 -- values are irrelevant)
        const uint32_t dst = smem_u32(smem) + 200 * 1024 + (warp - 4) * 2048 + lane * 16;
        while (!s_done) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(dst + i * 512), "f"(0.f) : "memory");
        }
    } else if (warp >= 8 && (interference & 2)) {
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        float acc = 0.f;
        while (!s_done) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(r[0]);
        }
        if (acc == 123.456f) out->mmas = -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

template <int PAT, int N>
void run(const char *name, int grid, Result *d)
{
    const size_t smem = 209 * 1024 + 1024;
    cudaFuncSetAttribute(probe<PAT, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int inter = 0; inter < 4; inter += 3) {
        Result h = {0, 0};
        cudaMemcpy(d, &h, sizeof(h), cudaMemcpyHostToDevice);
        probe<PAT, N><<<grid, 384, smem>>>(20000, inter, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-34s N=%3d grid=%3d interference=%d (1=smem writers 2=tmem readers): %6.1f cycles/MMA\n", name, N, grid,
               inter, (double)h.cycles / h.mmas);
    }
}

int main()
{
    Result *d;
    cudaMalloc(&d, sizeof(Result));
    for (int grid = 1; grid <= 148; grid += 147) {
        run<3, 64>("R1 repeated MMA", grid, d);
        run<3, 256>("R1 repeated MMA", grid, d);
        run<0, 64>("S3 sparse 3xTF32 triple", grid, d);
        run<1, 64>("S2 sparse concat pair", grid, d);
        run<2, 256>("C2 conv_tc2 group", grid, d);
    }
    return 0;
}
