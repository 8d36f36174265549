// conv3x3_tc.cu -- the 3x3 RPN convolution with HALO-TILE operand reuse (b2s_conv2d_tc, taps = 9, Cout = 128).
//
// conv_tc.cu fetches a fresh [128 pixels x 32 channels] A tile for every filter tap: 9 x (A 32 KB + W 32 KB)
// = 576 KB of L2->SM traffic per tile and 32-channel chunk, which is what bounded it at ~60 % tensor-pipe
// activity (ncu, round 1).  Here the A operand of a chunk is ONE halo tile
//     [18 rows x 10 cols x 32 channels]  (hi + lo = 45 KB)
// loaded once by TMA, and the nine taps are nine *views* of it: the output tile is 16 rows x 8 cols, so
// each group of 8 accumulator rows is one image row of the tile, consecutive groups are one halo row (10 pixels
// = 1280 B) apart, and tap (dy,dx) merely offsets the descriptor start by (dy*10 + dx)*128 B.  Only the weight
// tiles (32 KB per tap) stream.  Traffic per tile and chunk: 45 + 9*32 = 333 KB (-42 %).
//
//   UMMA descriptor of a view: K-major SWIZZLE_128B, stride-byte-offset 1280, start = halo + (dy*10+dx)*128,
//   base_offset = (start >> 7) & 7 (the start is 128-B but not 1024-B aligned).  TMA wrote the halo with the
//   address-based 128B swizzle, so any 128-B aligned row window is a valid operand.
//
// Accumulation chains are kept short (tensor-core accumulate is not round-to-nearest, see conv_tc.cu): one chain =
// (32-channel chunk, dy) = 3 taps x 12 MMAs = 36; the epilogue warps drain 12 partial sums per tile from a ring of
// 4 TMEM accumulators with round-to-nearest adds.
//
// Warps: 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 4-7 epilogue.  Persistent over tiles.
#include "tc_common.cuh"

namespace {

using namespace b2s_tc;
constexpr int kThreads = 256;
constexpr int TH = 16, TW = 8;                 // output tile: 16 rows x 8 cols = 128 pixels
constexpr int HH = TH + 2, HW = TW + 2;        // halo tile
constexpr int N = 128;                         // Cout
constexpr int ACC_SLOTS = 4;
constexpr int B_STAGES = 3;
constexpr uint32_t HALO_BYTES = HH * HW * BLOCK_K * 4;                 // 23040
constexpr uint32_t HALO_PLANE = (HALO_BYTES + 1023) / 1024 * 1024;     // 23552: keep every plane 1024-B aligned
constexpr uint32_t A_BUF_BYTES = 2 * HALO_PLANE;                       // hi + lo
constexpr uint32_t B_TILE = N * BLOCK_K * 4;                           // 16 KB
constexpr uint32_t B_STAGE_BYTES = 2 * B_TILE;                         // hi + lo
constexpr uint32_t SMEM_BYTES = 2 * A_BUF_BYTES + B_STAGES * B_STAGE_BYTES;

struct Conv3Params {
    int B, H, W, Cin, Cout, relu;
    int tiles_h, tiles_w, num_tiles;
    int out_stride;
    const float *scale, *shift;
    float *out_hi, *out_lo;
};

// view descriptor: rows 128 B apart inside a group of 8, groups `sbo` bytes apart, arbitrary 128-B aligned start
__device__ __forceinline__ uint64_t make_desc_view(uint32_t smem_addr, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)((smem_addr >> 7) & 7) << 49;       // base offset: swizzle phase of the start address
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(kThreads, 1)
k_conv3x3_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
             const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
             const Conv3Params p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem;                                   // [2 buffers][hi plane | lo plane]
    uint8_t *smem_b = smem + 2 * A_BUF_BYTES;                 // [B_STAGES][hi | lo]
    __shared__ __align__(8) uint64_t bar_afull[2], bar_aempty[2], bar_bfull[B_STAGES], bar_bempty[B_STAGES];
    __shared__ __align__(8) uint64_t bar_tfull[ACC_SLOTS], bar_tempty[ACC_SLOTS];
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_scale[N], s_shift[N];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = p.Cin / BLOCK_K;

    if (threadIdx.x < N) {
        int c = threadIdx.x;
        s_scale[c] = (p.scale && c < p.Cout) ? p.scale[c] : 1.f;
        s_shift[c] = (p.shift && c < p.Cout) ? p.shift[c] : 0.f;
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_afull[i], 1); mbar_init(&bar_aempty[i], 1); }
        for (int i = 0; i < B_STAGES; ++i) { mbar_init(&bar_bfull[i], 1); mbar_init(&bar_bempty[i], 1); }
        for (int i = 0; i < ACC_SLOTS; ++i) { mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int abuf = 0, bst = 0;
            uint32_t aphase = 0, bphase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int tw = tile % p.tiles_w;
                int th = (tile / p.tiles_w) % p.tiles_h;
                int b = tile / (p.tiles_w * p.tiles_h);
                for (int chunk = 0; chunk < kchunks; ++chunk) {
                    // halo tile of this chunk (padded coordinates: output (h,w) + tap (dy,dx) = input (h+dy, w+dx))
                    mbar_wait(&bar_aempty[abuf], aphase ^ 1);
                    uint8_t *ab = smem_a + (size_t)abuf * A_BUF_BYTES;
                    mbar_arrive_expect_tx(&bar_afull[abuf], 2 * HALO_BYTES);
                    tma_load_4d(ab, &map_a_hi, &bar_afull[abuf], chunk * BLOCK_K, tw * TW, th * TH, b);
                    tma_load_4d(ab + HALO_PLANE, &map_a_lo, &bar_afull[abuf], chunk * BLOCK_K, tw * TW, th * TH, b);
                    if (++abuf == 2) { abuf = 0; aphase ^= 1; }
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&bar_bempty[bst], bphase ^ 1);
                        uint8_t *bb = smem_b + (size_t)bst * B_STAGE_BYTES;
                        mbar_arrive_expect_tx(&bar_bfull[bst], B_STAGE_BYTES);
                        tma_load_3d(bb, &map_b_hi, &bar_bfull[bst], chunk * BLOCK_K, 0, tap);
                        tma_load_3d(bb + B_TILE, &map_b_lo, &bar_bfull[bst], chunk * BLOCK_K, 0, tap);
                        if (++bst == B_STAGES) { bst = 0; bphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(N);
            constexpr uint32_t SBO = HW * 128;        // next image row of the tile = next halo row
            int abuf = 0, bst = 0, acc = 0;
            uint32_t aphase = 0, bphase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int chunk = 0; chunk < kchunks; ++chunk) {
                    mbar_wait(&bar_afull[abuf], aphase);
                    tc_fence_after();
                    const uint32_t a_hi0 = smem_u32(smem_a + (size_t)abuf * A_BUF_BYTES);
                    const uint32_t a_lo0 = a_hi0 + HALO_PLANE;
                    for (int dy = 0; dy < 3; ++dy) {
                        mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                        tc_fence_after();
                        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * N);
                        for (int dx = 0; dx < 3; ++dx) {
                            mbar_wait(&bar_bfull[bst], bphase);
                            tc_fence_after();
                            const uint32_t voff = (uint32_t)(dy * HW + dx) * 128u;
                            const uint32_t sb = smem_u32(smem_b + (size_t)bst * B_STAGE_BYTES);
                            const uint64_t b_hi = make_desc_sw128(sb), b_lo = make_desc_sw128(sb + B_TILE);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                const uint32_t kb = (uint32_t)(k * UMMA_K * 4);
                                const uint64_t a_hi = make_desc_view(a_hi0 + voff + kb, SBO);
                                const uint64_t a_lo = make_desc_view(a_lo0 + voff + kb, SBO);
                                const uint64_t koff = (uint64_t)(kb >> 4);
                                umma_tf32(tmem_d, a_lo, b_hi + koff, idesc, (dx | k) != 0);
                                umma_tf32(tmem_d, a_hi, b_lo + koff, idesc, 1);
                                umma_tf32(tmem_d, a_hi, b_hi + koff, idesc, 1);
                            }
                            umma_commit(&bar_bempty[bst]);
                            if (++bst == B_STAGES) { bst = 0; bphase ^= 1; }
                        }
                        umma_commit(&bar_tfull[acc]);           // (chunk, dy) partial sum complete
                        if (++acc == ACC_SLOTS) { acc = 0; acc_phase ^= 1; }
                    }
                    umma_commit(&bar_aempty[abuf]);             // all nine views of this halo have been consumed
                    if (++abuf == 2) { abuf = 0; aphase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int groups = kchunks * 3;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            int tw = tile % p.tiles_w;
            int th = (tile / p.tiles_w) % p.tiles_h;
            int b = tile / (p.tiles_w * p.tiles_h);
            const int m = ew * 32 + lane;                      // accumulator row: image row m/8, column m%8 of the tile
            const int h = th * TH + m / TW, w = tw * TW + m % TW;
            const bool valid = (h < p.H) && (w < p.W);
            const size_t pix = ((size_t)b * (p.H + 2) + (h + 1)) * (p.W + 2) + (w + 1);
            float *oh = p.out_hi + pix * p.out_stride;
            float *ol = p.out_lo + pix * p.out_stride;
            float sum[N];
#pragma unroll
            for (int j = 0; j < N; ++j) sum[j] = 0.f;
            for (int g = 0; g < groups; ++g) {
                mbar_wait(&bar_tfull[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * N);
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[c0 + j] = __fadd_rn(sum[c0 + j], __uint_as_float(r[j]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tempty[acc]);
                if (++acc == ACC_SLOTS) { acc = 0; acc_phase ^= 1; }
            }
            if (valid) {
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += 4) {
                    if (c0 < p.Cout) {
                        float v[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x = fmaf(sum[c0 + j], s_scale[c0 + j], s_shift[c0 + j]);
                            if (p.relu) x = fmaxf(x, 0.f);
                            float hi = to_tf32_rn(x);
                            lo[j] = to_tf32_rn(x - hi);
                            v[j] = hi;
                        }
                        *reinterpret_cast<float4 *>(oh + c0) = make_float4(v[0], v[1], v[2], v[3]);
                        *reinterpret_cast<float4 *>(ol + c0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace

// called by b2s_conv2d_tc for taps == 9, n_pad == 128, padded hi/lo output
int b2s_conv3x3_tc_halo(const float *in_hi, const float *in_lo, int B, int H, int W, int Cin, const float *w_hi,
                        const float *w_lo, int Cout, const float *scale, const float *shift, int relu, float *out_hi,
                        float *out_lo, int out_stride, int num_sms, cudaStream_t stream)
{
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)(W + 2), (cuuint64_t)(H + 2), (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)(W + 2) * Cin * 4, (cuuint64_t)(H + 2) * (W + 2) * Cin * 4};
        cuuint32_t box[4] = {BLOCK_K, HW, HH, 1};
        if (make_map(&a_hi, in_hi, 4, dims, str, box) || make_map(&a_lo, in_lo, 4, dims, str, box)) return -1;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)N, 9};
        cuuint64_t str[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)N * Cin * 4};
        cuuint32_t box[3] = {BLOCK_K, N, 1};
        if (make_map(&b_hi, w_hi, 3, dims, str, box) || make_map(&b_lo, w_lo, 3, dims, str, box)) return -1;
    }
    Conv3Params p;
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.relu = relu;
    p.tiles_h = (H + TH - 1) / TH;
    p.tiles_w = (W + TW - 1) / TW;
    p.num_tiles = B * p.tiles_h * p.tiles_w;
    p.out_stride = out_stride;
    p.scale = scale; p.shift = shift; p.out_hi = out_hi; p.out_lo = out_lo;
    static bool attr = false;
    const size_t smem = SMEM_BYTES + 1024;
    if (!attr) {
        B2S_CUDA_OK(cudaFuncSetAttribute(k_conv3x3_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    k_conv3x3_tc<<<grid, kThreads, smem, stream>>>(a_hi, a_lo, b_hi, b_lo, p);
    B2S_LAUNCH_OK();
    return 0;
}
