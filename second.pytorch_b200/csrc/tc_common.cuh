// tc_common.cuh -- tcgen05 / TMEM / TMA / mbarrier helpers shared by the tensor-core kernels (sm_100a).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace b2s_tc {

// Operand format ("3xF16"): every fp32 operand x is carried as two fp16 planes, hi = fp16_rn(x) and
// lo = fp16_rn(x - hi): 22 significand bits while lo is a normal fp16, 2^-25 absolute otherwise (fp16 subnormals are
// exact operands for the tensor core).  Each K step issues A_hi*B_hi + A_hi*B_lo + A_lo*B_hi with fp32 accumulation in
// TMEM (the dropped lo*lo term is ~2^-22 relative) -- the same three products as the 3xTF32 split of round 1, but
// kind::f16 runs at twice the tf32 rate and the planes are half the bytes.  Weights are pre-scaled by a power of two
// per layer (folded into the epilogue scale) so that their lo plane stays in the normal range
// (b2second/tc.py: split_f16); activations are stored unscaled (|x| <= 65504, saturating conversion).
constexpr int BLOCK_M = 128;   // accumulator rows (TMEM lanes)
constexpr int BLOCK_K = 64;    // 64 fp16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;     // f16: 32 bytes per MMA K step
constexpr int ELEM_BYTES = 2;
constexpr uint32_t A_TILE_BYTES = BLOCK_M * BLOCK_K * ELEM_BYTES;   // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Warp-uniform role dispatch.  tcgen05.mma / tcgen05.commit take their operands from UNIFORM registers.  If the
// issuing code sits in a divergent region (`if (lane == 0) { ... }`), the compiler cannot prove the descriptors
// warp-uniform and wraps every UTCHMMA in an ELECT / R2UR / BRA.U.ANY "waterfall" loop: ~190 issue cycles per MMA
// measured with ncu source counters (round 1) -- more than the 64..128 tensor cycles the MMA itself takes, i.e. the
// issuing thread, not the tensor pipe or shared memory, was the bottleneck of all three tcgen05 kernels.
// So: the whole warp runs the role's loops with provably uniform values (warp index through a shuffle, the way
// CUTLASS does it) and one lane is elected only around the asynchronous instructions themselves.
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: ~seconds of spinning, then trap (a protocol bug must not hang the device)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t addr = smem_u32(bar);
    // not unrolled: the compiler otherwise replicates the try_wait ~60x at every call site, and these warp-specialised
    // kernels (several roles = several instruction streams per SM) are sensitive to instruction-cache footprint
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 28); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
// the same wait with a suspend-time hint (the form CUTLASS's ClusterBarrier::wait uses): the thread sleeps in the
// barrier unit (SASS: TRYWAIT + NANOSLEEP.SYNCS) instead of re-polling.  Measured: fewer shared-memory operations from
// the many waiting roles of the sparse kernel, but a later wake-up -- the dense conv kernels were 4 % slower with it.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity)
{
    uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 28); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity), "r"(0x989680u) : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address, 16-byte units
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major): 1
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // layout type: SWIZZLE_128B
    return d;
}

// instruction descriptor: D=f32 (bits 4-5 = 1), A=B=f16 (format code 0 in bits 7-9 / 10-12), both K-major, M=128, N
__host__ __device__ constexpr uint32_t make_idesc_f16(int n)
{
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Two consecutive MMAs that share the A operand: the first keeps A in the tensor core's collector buffer
// (SASS UTCHMMA ...A_KEEP), the second reads it from there (A_REUSE) instead of fetching it from shared memory again.
__device__ __forceinline__ void umma_f16_afill(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_alast(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 -> (hi, lo) fp16 pair packed in one register: hi in bits 0-15, lo in bits 16-31.  Saturating conversions
// (a value beyond +-65504 clamps instead of turning into inf and poisoning the accumulators).
__device__ __forceinline__ uint32_t split_f16(float v)
{
    uint16_t h, l;
    float hf;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"(v - hf));
    return (uint32_t)h | ((uint32_t)l << 16);
}
__device__ __forceinline__ float f16_bits_to_float(uint16_t h)
{
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
    return f;
}


// ---- host side: per-device caches (a process may drive several GPUs: SM count and the opt-in dynamic shared memory
// attribute belong to the CURRENT device, so they are cached per device index, not in a process-wide static) ----
constexpr int kMaxDevices = 64;
inline int current_device()
{
    int d = 0;
    cudaGetDevice(&d);
    return d & (kMaxDevices - 1);
}
inline int num_sms_current()
{
    static int sms[kMaxDevices] = {0};
    const int d = current_device();
    if (!sms[d]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148;
        sms[d] = v;
    }
    return sms[d];
}
// opt a kernel into `bytes` of dynamic shared memory on the current device (once per device and kernel instance)
#define B2S_SMEM_OPT_IN(kernel, bytes)                                                                           \
    do {                                                                                                         \
        static bool _done[b2s_tc::kMaxDevices] = {false};                                                        \
        const int _d = b2s_tc::current_device();                                                                 \
        if (!_done[_d]) {                                                                                        \
            B2S_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            _done[_d] = true;                                                                                    \
        }                                                                                                        \
    } while (0)

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency) ----
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (PFN_encodeTiled)p;
    return fn;
}

// elem_strides (optional): traversal stride per dimension -- {1, s, s, 1} fetches every s-th pixel (strided conv)
inline int make_map(CUtensorMap *m, const void *base, int rank, const cuuint64_t *dims,
                    const cuuint64_t *strides_bytes, const cuuint32_t *box, const cuuint32_t *elem_strides = nullptr)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) { b2s_set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (elem_strides)
        for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, (void *)base, dims, strides_bytes, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b2s_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
    return 0;
}

}  // namespace b2s_tc
