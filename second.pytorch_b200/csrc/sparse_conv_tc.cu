// sparse_conv_tc.cu -- sparse convolution inner product on the tensor pipe (b2s_sparse_conv_tc).
//
// Same output-stationary formulation as sparse_conv.cu (one CTA owns 128 output rows and walks the K kernel
// offsets; nbr[o][k] names the input row feeding output row o through offset k), but the per-offset product
//     D[128 rows, Cout] += A[128 gathered rows, 32 channels] * W[k][Cout, 32 channels]^T
// is a tcgen05.mma (kind::tf32) on the 3xTF32 hi/lo split (see conv_tc.cu), accumulating in TMEM.
//
// The gather is done by the TMA unit: cp.async.bulk.tensor.2d ... tile::gather4 fetches FOUR arbitrary rows of
// the [rows, Cin] feature matrix (32 channels = 128 B each) straight into the K-major SWIZZLE_128B A tile; a
// missing neighbour is an out-of-range row index, which TMA zero-fills without touching memory.  (Probe:
// tests/cuda/gather4_probe.cu -- box {32,1}, rows land 128 B apart with the standard address swizzle.)
// One warp issues, per K block, 32 hi + 32 lo gather4 ops (lane l = tile rows 4l..4l+3) plus the two weight
// tiles, all completing on the stage's mbarrier -- no LSU instructions, no shared-memory bank conflicts
// (the first version gathered with per-thread 16-byte cp.async: 25-30 % tensor-pipe activity).
//
// Warp roles (8 warps):
//   warp 0      TMA producer (gather4 of A hi/lo + bulk tensor load of W[k] hi/lo)
//   warp 1      MMA issuer (one elected lane)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: drain per-group partial sums from TMEM (round-to-nearest adds in registers; the
//               tensor core's own accumulate is not RN -- see conv_tc.cu), BN scale/shift + ReLU, hi/lo split,
//               one contiguous row store per thread
// K block = (kernel offset, 32-channel chunk); accumulation group = GROUP offsets (short chains).
#include "tc_common.cuh"

namespace {

using namespace b2s_tc;
constexpr int kThreads = 256;
constexpr int GROUP = 3;        // kernel offsets per accumulation chain
constexpr int ACC_SLOTS = 4;
constexpr int kOobRow = 0x3FFFFFFF;   // row coordinate beyond any tensor: TMA zero-fills

__device__ __forceinline__ void tma_gather4(uint32_t smem_dst, const CUtensorMap *map, uint64_t *bar, int col, int r0,
                                            int r1, int r2, int r3)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6, %7}], [%2];"
        ::"r"(smem_dst), "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

struct SpParams {
    const int *nbr;
    const int *n_out_dev;
    int cap_out, K, relu;
    const float *scale, *shift;
    float *out_hi, *out_lo;
};

// CIN in {32, 64}; COUT (= UMMA N) in {32, 64}
template <int CIN, int COUT, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
k_sparse_conv_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                 const SpParams p)
{
    constexpr int N = COUT;
    constexpr int KCH = CIN / BLOCK_K;                        // 32-channel chunks per offset
    constexpr uint32_t B_TILE_BYTES = N * BLOCK_K * 4;
    constexpr uint32_t STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    constexpr uint32_t TMEM_COLS = (ACC_SLOTS * N <= 128) ? 128 : (ACC_SLOTS * N <= 256) ? 256 : 512;
    static_assert(N % 16 == 0 && ACC_SLOTS * N <= 512, "TMEM capacity");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_tfull[ACC_SLOTS], bar_tempty[ACC_SLOTS];
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_scale[N], s_shift[N];
    __shared__ int s_nbr[BLOCK_M * 27];                      // the tile's neighbour table (K <= 27)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_out = min(*p.n_out_dev, p.cap_out);
    const int num_tiles = (n_out + BLOCK_M - 1) / BLOCK_M;
    const int K = p.K;
    const int num_groups = (K + GROUP - 1) / GROUP;

    if (threadIdx.x < N) {
        s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
        s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
        for (int i = 0; i < ACC_SLOTS; ++i) { mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        // ===================== TMA producer: gather4 of A (hi, lo) + weight tiles =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            // stage the tile's neighbour table (128 x K ints, contiguous in global memory)
            __syncwarp();
            {
                const int row0 = tile * BLOCK_M;
                const int valid = min(BLOCK_M, n_out - row0) * K;
                const int *src = p.nbr + (size_t)row0 * K;
                for (int i = lane; i < BLOCK_M * K; i += 32) {
                    int v = i < valid ? __ldg(&src[i]) : -1;
                    s_nbr[i] = v >= 0 ? v : kOobRow;
                }
            }
            __syncwarp();
            for (int k = 0; k < K; ++k) {
                const int r0 = s_nbr[(lane * 4 + 0) * K + k], r1 = s_nbr[(lane * 4 + 1) * K + k];
                const int r2 = s_nbr[(lane * 4 + 2) * K + k], r3 = s_nbr[(lane * 4 + 3) * K + k];
                for (int ch = 0; ch < KCH; ++ch) {
                    if (lane == 0) {
                        mbar_wait(&bar_empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&bar_full[stage], STAGE_BYTES);
                    }
                    __syncwarp();
                    uint8_t *st = smem + (size_t)stage * STAGE_BYTES;
                    const uint32_t dst = smem_u32(st) + (uint32_t)lane * 512u;     // 4 rows x 128 B per lane
                    tma_gather4(dst, &map_a_hi, &bar_full[stage], ch * BLOCK_K, r0, r1, r2, r3);
                    tma_gather4(dst + A_TILE_BYTES, &map_a_lo, &bar_full[stage], ch * BLOCK_K, r0, r1, r2, r3);
                    if (lane == 0) {
                        tma_load_3d(st + 2 * A_TILE_BYTES, &map_w_hi, &bar_full[stage], ch * BLOCK_K, 0, k);
                        tma_load_3d(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &map_w_lo, &bar_full[stage], ch * BLOCK_K, 0, k);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int g = 0; g < num_groups; ++g) {
                    mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * N);
                    const int k_end = min(K, (g + 1) * GROUP);
                    bool first = true;
                    for (int k = g * GROUP; k < k_end; ++k)
                        for (int ch = 0; ch < KCH; ++ch) {
                            mbar_wait(&bar_full[stage], phase);
                            tc_fence_after();
                            const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                            const uint64_t a_hi = make_desc_sw128(sa), a_lo = make_desc_sw128(sa + A_TILE_BYTES);
                            const uint64_t b_hi = make_desc_sw128(sa + 2 * A_TILE_BYTES);
                            const uint64_t b_lo = make_desc_sw128(sa + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
                            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                                const uint64_t koff = (uint64_t)((kk * UMMA_K * 4) >> 4);
                                umma_tf32(tmem_d, a_lo + koff, b_hi + koff, idesc, first ? 0u : 1u);
                                first = false;
                                umma_tf32(tmem_d, a_hi + koff, b_lo + koff, idesc, 1);
                                umma_tf32(tmem_d, a_hi + koff, b_hi + koff, idesc, 1);
                            }
                            umma_commit(&bar_empty[stage]);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    umma_commit(&bar_tfull[acc]);
                    if (++acc == ACC_SLOTS) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;                 // == warp % 4: TMEM lane quarter
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int row = tile * BLOCK_M + ew * 32 + lane;
            float sum[N];
#pragma unroll
            for (int j = 0; j < N; ++j) sum[j] = 0.f;
            for (int g = 0; g < num_groups; ++g) {
                mbar_wait(&bar_tfull[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * N);
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += 16) {
                    uint32_t rr[16];
                    tmem_ld16(taddr + c0, rr);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[c0 + j] = __fadd_rn(sum[c0 + j], __uint_as_float(rr[j]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tempty[acc]);
                if (++acc == ACC_SLOTS) { acc = 0; acc_phase ^= 1; }
            }
            if (row < n_out) {
                float *oh = p.out_hi + (size_t)row * N;
                float *ol = p.out_lo ? p.out_lo + (size_t)row * N : nullptr;
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += 4) {
                    float v[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float x = fmaf(sum[c0 + j], s_scale[c0 + j], s_shift[c0 + j]);
                        if (p.relu) x = fmaxf(x, 0.f);
                        if (ol) { float hi = to_tf32_rn(x); lo[j] = to_tf32_rn(x - hi); x = hi; }
                        v[j] = x;
                    }
                    *reinterpret_cast<float4 *>(oh + c0) = make_float4(v[0], v[1], v[2], v[3]);
                    if (ol) *reinterpret_cast<float4 *>(ol + c0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

template <int CIN, int COUT, int STAGES>
int launch(const CUtensorMap &a_hi, const CUtensorMap &a_lo, const CUtensorMap &w_hi, const CUtensorMap &w_lo,
           const SpParams &p, int num_sms, cudaStream_t stream)
{
    constexpr size_t stage = 2 * A_TILE_BYTES + 2 * (size_t)COUT * BLOCK_K * 4;
    size_t smem = stage * STAGES + 1024;
    static bool attr = false;
    if (!attr) {
        B2S_CUDA_OK(cudaFuncSetAttribute(k_sparse_conv_tc<CIN, COUT, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    int tiles_cap = (p.cap_out + BLOCK_M - 1) / BLOCK_M;
    int grid = tiles_cap < num_sms ? tiles_cap : num_sms;
    k_sparse_conv_tc<CIN, COUT, STAGES><<<grid, kThreads, smem, stream>>>(a_hi, a_lo, w_hi, w_lo, p);
    B2S_LAUNCH_OK();
    return 0;
}

int make_map_sw(CUtensorMap *m, const float *base, int rank, const cuuint64_t *dims, const cuuint64_t *str,
                const cuuint32_t *box)
{
    return make_map(m, base, rank, dims, str, box);
}

}  // namespace

extern "C" int b2s_sparse_conv_tc(const float *feat_hi, const float *feat_lo, int rows_in, int cin, const float *w_hi,
                                  const float *w_lo, const int *nbr, int K, const int *num_out_dev, int cap_out,
                                  const float *scale, const float *shift, int relu, float *out_hi, float *out_lo,
                                  int cout, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2S_REQUIRE((cin == 32 || cin == 64) && (cout == 32 || cout == 64),
                "b2s_sparse_conv_tc: built for Cin, Cout in {32, 64} (thin layers use b2s_sparse_conv)");
    B2S_REQUIRE(K >= 1 && K <= 27 && cap_out >= 0 && rows_in >= 1, "b2s_sparse_conv_tc: K must be 1..27, rows_in >= 1");
    if (cap_out == 0) return 0;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        B2S_CUDA_OK(cudaGetDevice(&dev));
        B2S_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // A: feature planes [rows_in, Cin], gathered 4 rows x 32 channels at a time (box {32, 1})
    CUtensorMap a_hi, a_lo, m_hi, m_lo;
    {
        cuuint64_t dims[2] = {(cuuint64_t)cin, (cuuint64_t)rows_in};
        cuuint64_t str[1] = {(cuuint64_t)cin * 4};
        cuuint32_t box[2] = {BLOCK_K, 1};
        if (make_map_sw(&a_hi, feat_hi, 2, dims, str, box) || make_map_sw(&a_lo, feat_lo, 2, dims, str, box)) return -1;
    }
    {
        // weights [K][Cout][Cin] (K-major B operand), hi and lo planes
        cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)K};
        cuuint64_t str[2] = {(cuuint64_t)cin * 4, (cuuint64_t)cout * cin * 4};
        cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)cout, 1};
        if (make_map_sw(&m_hi, w_hi, 3, dims, str, box) || make_map_sw(&m_lo, w_lo, 3, dims, str, box)) return -1;
    }
    SpParams p;
    p.nbr = nbr; p.n_out_dev = num_out_dev; p.cap_out = cap_out; p.K = K;
    p.relu = relu; p.scale = scale; p.shift = shift; p.out_hi = out_hi; p.out_lo = out_lo;
    if (cin == 64 && cout == 64) return launch<64, 64, 4>(a_hi, a_lo, m_hi, m_lo, p, num_sms, stream);
    if (cin == 32 && cout == 64) return launch<32, 64, 4>(a_hi, a_lo, m_hi, m_lo, p, num_sms, stream);
    if (cin == 32 && cout == 32) return launch<32, 32, 4>(a_hi, a_lo, m_hi, m_lo, p, num_sms, stream);
    return launch<64, 32, 4>(a_hi, a_lo, m_hi, m_lo, p, num_sms, stream);
}
