// sparse_conv_tc.cu -- sparse convolution inner product on the tensor pipe (b2s_sparse_conv_tc).
//
// Same output-stationary formulation as sparse_conv.cu (one CTA owns 128 output rows and walks the K kernel
// offsets; nbr[o][k] names the input row feeding output row o through offset k), but the per-offset product
//     D[128 rows, Cout] += A[128 gathered rows, 64 channels] * W[k][Cout, 64 channels]^T
// is a tcgen05.mma (kind::f16) on the 3xF16 hi/lo split (tc_common.cuh), accumulating in TMEM.
//
// K block = one 128-byte-wide slice of the reduction (64 fp16 channels):
//   Cin >= 64 : (kernel offset k, 64-channel chunk ch)                      -- K * Cin/64 K blocks per tile
//   Cin <  64 : PACK = 64/Cin consecutive kernel offsets side by side        -- ceil(K / PACK) K blocks per tile
//               (Cin = 32: 2 offsets, 16: 4 offsets, 8: 8 offsets per K block; the weights arrive pre-packed as
//               [K block][Cout][64], see b2second/tc.py: pack_sparse_weights.  A 3- or 4-feature input layer is
//               zero-padded to 8 channels = one 16-byte cp.async per row and plane.)
//
// Warp roles (16 warps):
//   warp 0      TMA producer for the weight tile of each K block (bulk tensor load, SWIZZLE_128B)
//   warp 1      MMA issuer (one elected lane, software-pipelined)
//   warp 2      TMEM allocator
//   warps 4-11  gather producers: 16-byte cp.async copies of the neighbour rows (hi and lo planes) into the
//               K-major SWIZZLE_128B A tile, zero-fill for missing neighbours, completion signalled on the
//               stage's mbarrier with cp.async.mbarrier.arrive.noinc.  (A TMA tile::gather4 producer was tried
//               -- tests/cuda/gather4_probe.cu proves the instruction works -- but 64 four-row TMA ops per K
//               block ran 3x SLOWER than cp.async: ~60+ cycles per gather4 issue, measured round 1.)
//   warps 12-15 epilogue: drain per-chain partial sums from TMEM (round-to-nearest adds in registers, the
//               tensor core's own accumulate is not RN -- see conv_tc.cu), BN scale/shift + ReLU, hi/lo split,
//               coalesced row stores through a small staging tile
//
// What bounds it (clock64 instrumentation, B2S_SP_ZSKIP bit 16, round 1): the gather producers' own instruction
// stream.  With 4 gather warps and address arithmetic inside the K-block loop the MMA issuer waited ~940 of
// ~1340 cycles per K block for them although the copies themselves cost 4 %; hoisting everything that is constant
// per kernel offset, unrolling the channel chunks and doubling the gather warps brought the K block to ~760
// cycles (tensor pipe 448 of them, tests/cuda/mma_probe2.cu).
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace {

using namespace b2s_tc;
constexpr int GW = 8;                     // gather warps (warps 4 .. 4+GW-1); epilogue = the 4 warps after them
constexpr int RI = 128 / GW / 4;          // 4-row copy iterations per gather warp and K block
constexpr int kThreads = 32 * (4 + GW + 4);
constexpr int GROUP = 3;                  // kernel offsets per accumulation chain (Cin >= 32)
constexpr int ACC_SLOTS = 4;

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gsrc, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct SpParams {
    const __half *in_hi, *in_lo;
    int in_stride, out_stride;   // halves between consecutive rows of the input / output planes
    int *status;
    const int *nbr;
    const int *n_out_dev;
    int cap_out, K, relu;
    int flags;               // B2S_SP_ZSKIP: bit 0 zero-slot skip (default on); diagnostics (wrong results): 2 no gather
                             // copies, 4 no weight loads, 8 no neighbour-table staging; 16 print the issuer's wait times
    const float *scale, *shift;
    void *out_hi;                // fp16 hi plane, or (out_lo == NULL) fp32 rows [cap_out, COUT]
    __half *out_lo;
};

// CIN in {8, 16, 32, 64}; COUT (= UMMA N) in {16, 32, 64}
template <int CIN, int COUT, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
k_sparse_conv_tc(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                 const SpParams p)
{
    constexpr int N = COUT;
    constexpr bool PACKED = CIN < BLOCK_K;
    constexpr int PACK = PACKED ? BLOCK_K / CIN : 1;          // kernel offsets per K block
    constexpr int KCH = PACKED ? 1 : CIN / BLOCK_K;           // 64-channel chunks per offset
    constexpr int CPO = 8 / PACK;                             // 16-byte chunks per offset inside a 128-byte row
    constexpr int CHAIN_KB = PACKED ? 2 : GROUP * KCH;        // K blocks per accumulation chain (short chains)
    constexpr uint32_t B_TILE_BYTES = N * BLOCK_K * ELEM_BYTES;
    constexpr uint32_t STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    // B-operand concatenation: the stage holds W_hi (N rows) directly followed by W_lo (N rows), so ONE MMA with
    // UMMA N = 2N computes A_hi*W_hi (accumulator columns [0,N)) and A_hi*W_lo (columns [N,2N)); a second MMA
    // (UMMA N = N) adds A_lo*W_hi into columns [0,N).  8 instead of 12 MMAs per K block for the same products
    // (tests/cuda/mma_probe2.cu: 448 instead of 576 tensor cycles at N = 64).  The epilogue adds the two halves.
    constexpr int ACC_W = 2 * N;                              // accumulator slot width in TMEM columns
    constexpr uint32_t TMEM_COLS = (ACC_SLOTS * ACC_W <= 128) ? 128 : (ACC_SLOTS * ACC_W <= 256) ? 256 : 512;
    static_assert(N % 16 == 0 && ACC_SLOTS * ACC_W <= 512, "TMEM capacity");
    static_assert(STAGES * 8 <= 32, "zero-slot bookkeeping uses 8 bits per stage");

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_tfull[ACC_SLOTS], bar_tempty[ACC_SLOTS];
    __shared__ uint32_t s_tmem_base;
    __shared__ float s_scale[N], s_shift[N];
    __shared__ int s_nbr[BLOCK_M * 27];                      // the tile's neighbour table (K <= 27)
    __shared__ __align__(16) uint32_t s_stage[4][32 * 36];   // per epilogue warp: 32 rows x 32 words transpose tile

    const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
    const int n_out = min(*p.n_out_dev, p.cap_out);
    const int num_tiles = (n_out + BLOCK_M - 1) / BLOCK_M;
    const int K = p.K;
    const int num_kb = PACKED ? (K + PACK - 1) / PACK : K * KCH;
    const int num_chains = (num_kb + CHAIN_KB - 1) / CHAIN_KB;

    if (threadIdx.x < N) {
        s_scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.f;
        s_shift[threadIdx.x] = p.shift ? p.shift[threadIdx.x] : 0.f;
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&bar_full[i], 32 * GW + 1); mbar_init(&bar_empty[i], 1); }
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
        // ===================== weight TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    uint8_t *st = smem + (size_t)stage * STAGE_BYTES;
                    if (p.flags & 4) {                     // diagnostic: no weight loads (results wrong)
                        mbar_arrive(&bar_full[stage]);
                    } else {
                        const int c0 = PACKED ? 0 : (kb % KCH) * BLOCK_K, c2 = PACKED ? kb : kb / KCH;
                        mbar_arrive_expect_tx(&bar_full[stage], 2 * B_TILE_BYTES);
                        tma_load_3d(st + 2 * A_TILE_BYTES, &map_w_hi, &bar_full[stage], c0, 0, c2);
                        tma_load_3d(st + 2 * A_TILE_BYTES + B_TILE_BYTES, &map_w_lo, &bar_full[stage], c0, 0, c2);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // Software-pipelined single-lane issue: the barriers of the NEXT K block (and, at a chain end, of the next
        // accumulator) are waited for before the current K block's last two MMAs are issued, so the tensor queue
        // does not drain between the short 8-MMA bursts.
        constexpr uint32_t idesc = make_idesc_f16(N), idesc2 = make_idesc_f16(2 * N);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t smem0 = smem_u32(smem);
        const int num_tiles_u = __shfl_sync(0xffffffffu, num_tiles, 0);
        if (elect_one_sync() && blockIdx.x < (unsigned)num_tiles_u) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const bool timing = (p.flags & 16) != 0;
            long long t_full = 0, t_tempty = 0, t_fence = 0, t_begin = clock64();
            int n_kb = 0;
            mbar_wait(&bar_tempty[0], 1);
            mbar_wait(&bar_full[0], 0);
            tc_fence_after();
            for (int tile = blockIdx.x; tile < num_tiles_u; tile += gridDim.x) {
                const bool last_tile = tile + (int)gridDim.x >= num_tiles_u;
                for (int g = 0; g < num_chains; ++g) {
                    const uint32_t tmem_d = tmem_u + (uint32_t)(acc * ACC_W);
                    const int kb_end = min(num_kb, (g + 1) * CHAIN_KB) - g * CHAIN_KB;   // K blocks of this chain
                    int accn = acc + 1;
                    uint32_t acc_phase_n = acc_phase;
                    if (accn == ACC_SLOTS) { accn = 0; acc_phase_n ^= 1; }
                    for (int kb = 0; kb < kb_end; ++kb) {
                        const uint32_t sa = smem0 + (uint32_t)stage * STAGE_BYTES;
                        const uint64_t a_hi = make_desc_sw128(sa), a_lo = make_desc_sw128(sa + A_TILE_BYTES);
                        const uint64_t b_hl = make_desc_sw128(sa + 2 * A_TILE_BYTES);   // [W_hi; W_lo], 2N rows
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) {
                            const uint64_t koff = (uint64_t)((kk * UMMA_K * ELEM_BYTES) >> 4);
                            umma_f16(tmem_d, a_hi + koff, b_hl + koff, idesc2, (kb | kk) != 0);   // cols [0,2N)
                            umma_f16(tmem_d, a_lo + koff, b_hl + koff, idesc, 1);                 // cols [0,N)
                        }
                        // look ahead
                        int stn = stage + 1;
                        uint32_t phn = phase;
                        if (stn == STAGES) { stn = 0; phn ^= 1; }
                        const bool chain_end = kb == kb_end - 1;
                        const bool last = last_tile && chain_end && g == num_chains - 1;
                        if (!last) {
                            const long long w0 = timing ? clock64() : 0;
                            mbar_wait(&bar_full[stn], phn);
                            const long long w1 = timing ? clock64() : 0;
                            if (chain_end) mbar_wait(&bar_tempty[accn], acc_phase_n ^ 1);
                            const long long w2 = timing ? clock64() : 0;
                            tc_fence_after();
                            if (timing) { t_full += w1 - w0; t_tempty += w2 - w1; t_fence += clock64() - w2; ++n_kb; }
                        }
                        {
                            const uint64_t koff = (uint64_t)((3 * UMMA_K * ELEM_BYTES) >> 4);
                            umma_f16(tmem_d, a_hi + koff, b_hl + koff, idesc2, 1);
                            umma_f16(tmem_d, a_lo + koff, b_hl + koff, idesc, 1);
                        }
                        umma_commit(&bar_empty[stage]);
                        if (chain_end) umma_commit(&bar_tfull[acc]);
                        stage = stn; phase = phn;
                    }
                    acc = accn; acc_phase = acc_phase_n;
                }
            }
            if (timing && blockIdx.x == 0)
                printf("[sparse_tc<%d,%d>] issuer: %d K blocks, total %lld cyc (%.0f/kb), wait full %lld (%.0f/kb), "
                       "wait tempty %lld (%.0f/kb), fence %lld (%.0f/kb)\n", CIN, COUT, n_kb, clock64() - t_begin,
                       (double)(clock64() - t_begin) / (n_kb + 1), t_full, (double)t_full / (n_kb + 1), t_tempty,
                       (double)t_tempty / (n_kb + 1), t_fence, (double)t_fence / (n_kb + 1));
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + GW) {
        // ===================== gather producers =====================
        // Lane mapping: one warp instruction covers 4 rows x 8 sixteen-byte chunks, so the 32 lanes write 4 whole
        // 128-byte smem rows (bank-conflict free under the 128B swizzle) and read 4 x 128 contiguous global bytes.
        // (One lane per row -- the first version -- was a 4-way bank conflict on every cp.async.)
        const int gw = warp - 4;                               // rows 4*RI*gw .. of the tile
        const int sub = lane >> 3;                             // row within a group of 4
        const uint32_t chunk = (uint32_t)(lane & 7);           // 16-byte chunk of the 128-byte row
        int stage = 0;
        uint32_t phase = 0;
        // ~70 % of the neighbour slots are empty.  Bit (stage*8 + i) of `zeroed` remembers that this lane's 16-byte
        // chunk of row slot i in that stage already holds zeros (from an earlier empty neighbour), so an empty
        // neighbour needs no shared-memory write at all (predication, uniform issue).
        uint32_t zeroed = 0;
        uint32_t dst_off[RI];                                   // swizzled byte offset of (row slot i, chunk) in a stage
#pragma unroll
        for (int i = 0; i < RI; ++i) {
            const uint32_t rl = (uint32_t)(gw * (4 * RI) + i * 4 + sub);
            dst_off[i] = rl * 128u + ((chunk ^ (rl & 7u)) << 4);
        }
        const ptrdiff_t lo_delta = reinterpret_cast<const char *>(p.in_lo) - reinterpret_cast<const char *>(p.in_hi);
        const bool zskip_on = (p.flags & 1) != 0;
        const uint32_t copy_mask = (p.flags & 2) ? 0u : 0xFFu;  // diagnostic bit 2: no copies at all
        const uint32_t smem0 = smem_u32(smem);

        const bool timing_g = (p.flags & 16) != 0 && blockIdx.x == 0 && gw == 0 && lane == 0;
        long long tg_empty = 0, tg_stage = 0, tg_begin = clock64();
        int tg_kb = 0;
        // one K block: wait for the stage, issue this lane's (predicated) copies, arrive
        auto copy_block = [&](uint32_t valid, const char *const *g_hi, const char *const *g_lo, int byte_off) {
            const long long te0 = timing_g ? clock64() : 0;
            mbar_wait(&bar_empty[stage], phase ^ 1);
            if (timing_g) { tg_empty += clock64() - te0; ++tg_kb; }
            const uint32_t sa = smem0 + (uint32_t)stage * STAGE_BYTES;
            const uint32_t zst = (zeroed >> (stage * 8)) & 0xFFu;          // slots of this stage holding zeros
            const uint32_t need = (valid | ~zst | (zskip_on ? 0u : 0xFFu)) & copy_mask;
#pragma unroll
            for (int i = 0; i < RI; ++i) {
                if (need & (1u << i)) {
                    const uint32_t nbytes = (valid >> i) & 1u ? 16u : 0u;     // src-size 0 -> 16 bytes of zeros
                    cp_async16(sa + dst_off[i], g_hi[i] + byte_off, nbytes);
                    cp_async16(sa + dst_off[i] + A_TILE_BYTES, g_lo[i] + byte_off, nbytes);
                }
            }
            zeroed = (zeroed & ~(0xFFu << (stage * 8))) | ((~valid & 0xFFu) << (stage * 8));
            cp_async_mbar_arrive_noinc(&bar_full[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };

        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            // stage the tile's neighbour table in shared memory (coalesced), shared by the gather warps
            const long long ts0 = timing_g ? clock64() : 0;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * GW) : "memory");     // previous tile's readers are done
            {
                const int row0 = tile * BLOCK_M;
                const int valid_n = min(BLOCK_M, n_out - row0) * K;
                const int *src = p.nbr + (size_t)row0 * K;
                if (!(p.flags & 8)) {                  // (diagnostic bit 8: skip the table staging)
                    // fixed trip count (K <= 27) so the loads are all in flight together instead of one L2 round
                    // trip per iteration
                    constexpr int NLD = (BLOCK_M * 27 + 32 * GW - 1) / (32 * GW);
                    int v[NLD];
#pragma unroll
                    for (int j = 0; j < NLD; ++j) {
                        const int i = gw * 32 + lane + j * 32 * GW;
                        v[j] = i < valid_n ? __ldg(&src[i]) : -1;
                    }
#pragma unroll
                    for (int j = 0; j < NLD; ++j) {
                        const int i = gw * 32 + lane + j * 32 * GW;
                        if (i < BLOCK_M * K) s_nbr[i] = v[j];
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * GW) : "memory");
            if (timing_g) tg_stage += clock64() - ts0;
            if constexpr (!PACKED) {
                // addresses and validity are constant per kernel offset; the channel chunks are unrolled so that
                // ch * 128 folds into the copy instructions' address immediates
                for (int k = 0; k < K; ++k) {
                    const char *g_hi[RI], *g_lo[RI];
                    uint32_t valid = 0;
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        const int src = s_nbr[(gw * (4 * RI) + i * 4 + sub) * K + k];
                        valid |= (src >= 0 ? 1u : 0u) << i;
                        g_hi[i] = reinterpret_cast<const char *>(p.in_hi + (size_t)(src >= 0 ? src : 0) * p.in_stride + chunk * 8);
                        g_lo[i] = g_hi[i] + lo_delta;
                    }
#pragma unroll
                    for (int ch = 0; ch < KCH; ++ch) copy_block(valid, g_hi, g_lo, ch * (BLOCK_K * ELEM_BYTES));
                }
            } else {
                // PACK offsets side by side: this lane's chunk belongs to offset kb*PACK + chunk/CPO and carries
                // channels (chunk % CPO)*8 .. +7 of that neighbour's row
                const int ko = (int)chunk / CPO;
                const int cofs = ((int)chunk % CPO) * 8;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int k = kb * PACK + ko;
                    const char *g_hi[RI], *g_lo[RI];
                    uint32_t valid = 0;
#pragma unroll
                    for (int i = 0; i < RI; ++i) {
                        const int src = k < K ? s_nbr[(gw * (4 * RI) + i * 4 + sub) * K + k] : -1;
                        valid |= (src >= 0 ? 1u : 0u) << i;
                        g_hi[i] = reinterpret_cast<const char *>(p.in_hi + (size_t)(src >= 0 ? src : 0) * p.in_stride + cofs);
                        g_lo[i] = g_hi[i] + lo_delta;
                    }
                    copy_block(valid, g_hi, g_lo, 0);
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (timing_g)
            printf("[sparse_tc<%d,%d>] gather warp 0: %d K blocks, total %lld cyc (%.0f/kb), wait empty %lld (%.0f/kb), "
                   "nbr-table staging %lld (%.0f/kb)\n", CIN, COUT, tg_kb, clock64() - tg_begin,
                   (double)(clock64() - tg_begin) / (tg_kb + 1), tg_empty, (double)tg_empty / (tg_kb + 1), tg_stage,
                   (double)tg_stage / (tg_kb + 1));
    } else if (warp >= 4 + GW) {
        // ===================== epilogue =====================
        const int ew = warp - (4 + GW);                // == warp % 4: TMEM lane quarter
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool timing_e = (p.flags & 16) != 0 && blockIdx.x == 0 && ew == 0 && lane == 0;
        long long te_wait = 0, te_drain = 0, te_store = 0, te_begin = clock64();
        int te_chains = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            float sum[N];
#pragma unroll
            for (int j = 0; j < N; ++j) sum[j] = 0.f;
            for (int g = 0; g < num_chains; ++g) {
                const long long e0 = timing_e ? clock64() : 0;
                mbar_wait(&bar_tfull[acc], acc_phase);
                tc_fence_after();
                const long long e1 = timing_e ? clock64() : 0;
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * ACC_W);
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += 16) {
                    uint32_t rr[16];
                    tmem_ld16(taddr + c0, rr);            // A_hi*W_hi + A_lo*W_hi
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[c0 + j] = __fadd_rn(sum[c0 + j], __uint_as_float(rr[j]));
                    tmem_ld16(taddr + N + c0, rr);        // A_hi*W_lo
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[c0 + j] = __fadd_rn(sum[c0 + j], __uint_as_float(rr[j]));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tempty[acc]);
                if (++acc == ACC_SLOTS) { acc = 0; acc_phase ^= 1; }
                if (timing_e) { te_wait += e1 - e0; te_drain += clock64() - e1; ++te_chains; }
            }
            const long long es0 = timing_e ? clock64() : 0;
            // coalesced stores: transpose 32 rows x CW channels through a padded shared tile so every store
            // instruction writes whole contiguous row segments (a lane-per-row store is 16 B at a row stride).
            // fp16 planes: a staged row = WP words of hi pairs, then (at word 16) WP words of lo pairs; the first CP
            // lanes of a row group write the hi segment, the next CP lanes the lo segment.
            constexpr int CW = N < 32 ? N : 32;                // channels per pass
            uint32_t *stg = s_stage[ew];
            const int row_w0 = tile * BLOCK_M + ew * 32;       // first row of this warp
            bool range_bad = false;
            if (p.out_lo) {
                constexpr int WP = CW / 2;                     // words per plane and row
                constexpr int CP = WP / 4;                     // 16-byte chunks per plane and row
                constexpr int LPR = 2 * CP;                    // lanes per row
                constexpr int RPI = 32 / LPR;                  // rows per store instruction
                const int sp = lane / LPR, sq = lane % LPR;
                __half *outp = (sq < CP) ? reinterpret_cast<__half *>(p.out_hi) : p.out_lo;
                const int coff = (sq % CP) * 8;                // first of this lane's 8 channels inside the pass
#pragma unroll
                for (int cc = 0; cc < N; cc += CW) {
                    __syncwarp();
#pragma unroll
                    for (int c0 = 0; c0 < CW; c0 += 8) {
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x0 = fmaf(sum[cc + c0 + 2 * j], s_scale[cc + c0 + 2 * j], s_shift[cc + c0 + 2 * j]);
                            float x1 = fmaf(sum[cc + c0 + 2 * j + 1], s_scale[cc + c0 + 2 * j + 1], s_shift[cc + c0 + 2 * j + 1]);
                            if (p.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                            range_bad |= (fabsf(x0) > 65504.f) | (fabsf(x1) > 65504.f);
                            const uint32_t p0 = split_f16(x0), p1 = split_f16(x1);
                            hw[j] = __byte_perm(p0, p1, 0x5410);
                            lw[j] = __byte_perm(p0, p1, 0x7632);
                        }
                        *reinterpret_cast<uint4 *>(stg + lane * 36 + c0 / 2) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4 *>(stg + lane * 36 + 16 + c0 / 2) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 32 / RPI; ++it) {
                        const int rr = row_w0 + it * RPI + sp;
                        if (rr < n_out)
                            *reinterpret_cast<uint4 *>(outp + (size_t)rr * p.out_stride + cc + coff) =
                                *reinterpret_cast<const uint4 *>(stg + (it * RPI + sp) * 36 + (sq < CP ? 0 : 16) + (sq % CP) * 4);
                    }
                }
            } else {
                constexpr int CH4 = CW / 4;                        // 16-byte chunks per row segment
                constexpr int RPI = 32 / CH4;                      // rows per store instruction
                const int sp = lane / CH4, sq = lane % CH4;
                float *outp = reinterpret_cast<float *>(p.out_hi);
#pragma unroll
                for (int cc = 0; cc < N; cc += CW) {
                    __syncwarp();
#pragma unroll
                    for (int c0 = 0; c0 < CW; c0 += 4) {
                        uint32_t v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x = fmaf(sum[cc + c0 + j], s_scale[cc + c0 + j], s_shift[cc + c0 + j]);
                            if (p.relu) x = fmaxf(x, 0.f);
                            v[j] = __float_as_uint(x);
                        }
                        *reinterpret_cast<uint4 *>(stg + lane * 36 + c0) = make_uint4(v[0], v[1], v[2], v[3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 32 / RPI; ++it) {
                        const int rr = row_w0 + it * RPI + sp;
                        if (rr < n_out)
                            *reinterpret_cast<uint4 *>(outp + (size_t)rr * N + cc + sq * 4) =
                                *reinterpret_cast<const uint4 *>(stg + (it * RPI + sp) * 36 + sq * 4);
                    }
                }
            }
            if (__any_sync(0xffffffffu, range_bad) && lane == 0 && p.status) atomicOr(p.status, B2S_STATUS_F16_RANGE);
            if (timing_e) te_store += clock64() - es0;
        }
        if (timing_e)
            printf("[sparse_tc<%d,%d>] epilogue warp 0: %d chains, total %lld cyc, wait tfull %lld (%.0f/chain), drain %lld "
                   "(%.0f/chain), output stage %lld\n", CIN, COUT, te_chains, clock64() - te_begin, te_wait,
                   (double)te_wait / (te_chains + 1), te_drain, (double)te_drain / (te_chains + 1), te_store);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

template <int CIN, int COUT, int STAGES>
int launch(const CUtensorMap &w_hi, const CUtensorMap &w_lo, const SpParams &p, int num_sms, cudaStream_t stream)
{
    constexpr size_t stage = 2 * A_TILE_BYTES + 2 * (size_t)COUT * BLOCK_K * ELEM_BYTES;
    size_t smem = stage * STAGES + 1024;
    B2S_SMEM_OPT_IN((k_sparse_conv_tc<CIN, COUT, STAGES>), smem);
    int tiles_cap = (p.cap_out + BLOCK_M - 1) / BLOCK_M;
    int grid = tiles_cap < num_sms ? tiles_cap : num_sms;
    k_sparse_conv_tc<CIN, COUT, STAGES><<<grid, kThreads, smem, stream>>>(w_hi, w_lo, p);
    B2S_LAUNCH_OK();
    return 0;
}

}  // namespace

extern "C" int b2s_sparse_conv_tc_supported(int cin, int cout)
{
    return (cin == 8 || cin == 16 || cin == 32 || cin == 64) && (cout == 16 || cout == 32 || cout == 64);
}

// Weight layout: Cin = 64: [K][Cout][64]; Cin < 64 (8, 16, 32): packed [ceil(K / (64/Cin))][Cout][64] with column
// (offset-in-pack * Cin + cin) and zero columns for offsets >= K (b2second/tc.py: pack_sparse_weights).
extern "C" int b2s_sparse_conv_tc(const b2s_half *feat_hi, const b2s_half *feat_lo, int in_stride, int rows_in, int cin,
                                  const b2s_half *w_hi, const b2s_half *w_lo, const int *nbr, int K,
                                  const int *num_out_dev, int cap_out, const float *scale, const float *shift, int relu,
                                  void *out_hi, b2s_half *out_lo, int out_stride, int cout, unsigned *status_dev,
                                  void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2S_REQUIRE(b2s_sparse_conv_tc_supported(cin, cout),
                "b2s_sparse_conv_tc: built for Cin in {8, 16, 32, 64}, Cout in {16, 32, 64} (others: b2s_sparse_conv)");
    B2S_REQUIRE(K >= 1 && K <= 27 && cap_out >= 0 && rows_in >= 0, "b2s_sparse_conv_tc: K must be 1..27");
    B2S_REQUIRE(in_stride >= cin && in_stride % 8 == 0 && ((uintptr_t)feat_hi & 15) == 0 && ((uintptr_t)feat_lo & 15) == 0,
                "b2s_sparse_conv_tc: input rows must be 16-byte aligned (in_stride multiple of 8 halves)");
    B2S_REQUIRE(out_lo == nullptr || (out_stride >= cout && out_stride % 8 == 0 && ((uintptr_t)out_hi & 15) == 0 &&
                                      ((uintptr_t)out_lo & 15) == 0),
                "b2s_sparse_conv_tc: output rows must be 16-byte aligned (out_stride multiple of 8 halves)");
    if (cap_out == 0) return 0;
    const int num_sms = num_sms_current();
    CUtensorMap m_hi, m_lo;
    {
        const bool packed = cin < BLOCK_K;
        const int pack = packed ? BLOCK_K / cin : 1;
        cuuint64_t row = packed ? BLOCK_K : (cuuint64_t)cin;     // halves per weight row
        cuuint64_t nkb = packed ? (cuuint64_t)((K + pack - 1) / pack) : (cuuint64_t)K;
        cuuint64_t dims[3] = {row, (cuuint64_t)cout, nkb};
        cuuint64_t str[2] = {row * ELEM_BYTES, (cuuint64_t)cout * row * ELEM_BYTES};
        cuuint32_t box[3] = {BLOCK_K, (cuuint32_t)cout, 1};
        if (make_map(&m_hi, w_hi, 3, dims, str, box) || make_map(&m_lo, w_lo, 3, dims, str, box)) return -1;
    }
    SpParams p;
    p.in_hi = reinterpret_cast<const __half *>(feat_hi); p.in_lo = reinterpret_cast<const __half *>(feat_lo);
    p.in_stride = in_stride; p.out_stride = out_stride; p.status = (int *)status_dev;
    p.nbr = nbr; p.n_out_dev = num_out_dev; p.cap_out = cap_out; p.K = K;
    {
        static int fl = -1;
        if (fl < 0) { const char *e = getenv("B2S_SP_ZSKIP"); fl = e ? atoi(e) : 1; }
#ifdef B2S_DIAG
        p.flags = fl;           // bits 2/4/8 corrupt results: only in `make DIAG=1` builds
#else
        p.flags = fl & (1 | 16);
#endif
    }
    p.relu = relu; p.scale = scale; p.shift = shift; p.out_hi = out_hi; p.out_lo = reinterpret_cast<__half *>(out_lo);
#define B2S_TC_CASE(CI, CO) if (cin == CI && cout == CO) return launch<CI, CO, 4>(m_hi, m_lo, p, num_sms, stream);
    B2S_TC_CASE(64, 64) B2S_TC_CASE(64, 32) B2S_TC_CASE(64, 16) B2S_TC_CASE(32, 64) B2S_TC_CASE(32, 32) B2S_TC_CASE(32, 16)
    B2S_TC_CASE(16, 64) B2S_TC_CASE(16, 32) B2S_TC_CASE(16, 16) B2S_TC_CASE(8, 64) B2S_TC_CASE(8, 32) B2S_TC_CASE(8, 16)
#undef B2S_TC_CASE
    b2s_set_error("b2s_sparse_conv_tc: Cin=%d Cout=%d not built", cin, cout);
    return -2;
}
