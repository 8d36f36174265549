// sparse_conv.cu -- output-stationary sparse convolution (b2s_sparse_conv, include/b2second.h).
//
//   out[o,:] = act( (sum_k W[k]^T in[nbr[o,k],:]) * scale + shift )
//
// One CTA owns a tile of TM output rows and walks the K kernel offsets: for each offset it gathers
// the tile's neighbour rows (zero rows where the neighbour is inactive) into shared memory, stages
// W[k] in shared memory and accumulates a TM x COUT register tile.  Every output row is written
// exactly once (no atomics, no scatter-add), with the BatchNorm1d(eval)+ReLU that always follows a
// sparse conv in second.pytorch (middle.py:146-191) folded into the epilogue.
//
// fp32 FMA register tiles (exact fp32 parity with the oracle): the fallback for channel widths the tensor-pipe
// kernel (sparse_conv_tc.cu: Cin in {4,16,32,64}) does not cover, and the `sparse_impl="fma"` cross-check.
//
// Algorithmic traffic per layer: 4*(N_in*Cin + N_out*Cout) + 4*K*N_out (nbr table)
// + 4*K*Cin*Cout (weights) bytes; flops = 2*pairs*Cin*Cout.
#include "common.cuh"

namespace {

constexpr int TM = 64;        // output rows per CTA
constexpr int kThreads = 256; // 16 row groups (4 rows) x 16 column groups

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads)
k_sparse_conv(const float *__restrict__ feat_in, const float *__restrict__ weight,
              const int *__restrict__ nbr, int K, const int *__restrict__ n_out_dev, int cap_out,
              const float *__restrict__ scale, const float *__restrict__ shift, int relu,
              float *__restrict__ feat_out)
{
    constexpr int NC = COUT / 16;          // output columns per thread
    constexpr int CINP = (CIN + 3) / 4 * 4; // padded K-dim of the smem tiles
    constexpr int LDA = TM + 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *s_nbr = reinterpret_cast<int *>(smem_raw);                    // [TM][K]
    float *s_w = reinterpret_cast<float *>(s_nbr + TM * 27);           // [CINP][COUT]
    float *s_at = s_w + CINP * COUT;                                   // [CINP][LDA]  (A transposed)
    __shared__ int s_any[27];

    const int n_out = min(*n_out_dev, cap_out);
    const int row0 = blockIdx.x * TM;
    if (row0 >= n_out) return;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int rows_here = min(TM, n_out - row0);

    if (tid < 27) s_any[tid] = 0;
    __syncthreads();
    for (int i = tid; i < TM * K; i += kThreads) {
        int r = i / K;
        int v = (r < rows_here) ? __ldg(&nbr[(size_t)row0 * K + i]) : -1;
        s_nbr[i] = v;
        if (v >= 0) s_any[i - r * K] = 1;  // benign race: all writers store 1
    }
    __syncthreads();

    float acc[4][NC];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[i][j] = 0.f;

    for (int k = 0; k < K; ++k) {
        if (!s_any[k]) continue;  // block-uniform
        // stage W[k]
        const float *wk = weight + (size_t)k * CIN * COUT;
        for (int i = tid; i < CIN * COUT; i += kThreads) s_w[i] = __ldg(&wk[i]);
        // gather A (transposed): s_at[c][r]
        if constexpr (CIN % 4 == 0) {
            constexpr int V = CIN / 4;
            for (int i = tid; i < TM * V; i += kThreads) {
                int r = i / V, v = i - r * V;
                int src = s_nbr[r * K + k];
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src >= 0) val = __ldg(reinterpret_cast<const float4 *>(feat_in + (size_t)src * CIN) + v);
                s_at[(v * 4 + 0) * LDA + r] = val.x;
                s_at[(v * 4 + 1) * LDA + r] = val.y;
                s_at[(v * 4 + 2) * LDA + r] = val.z;
                s_at[(v * 4 + 3) * LDA + r] = val.w;
            }
        } else {
            for (int i = tid; i < TM * CIN; i += kThreads) {
                int r = i / CIN, c = i - r * CIN;
                int src = s_nbr[r * K + k];
                s_at[c * LDA + r] = (src >= 0) ? __ldg(&feat_in[(size_t)src * CIN + c]) : 0.f;
            }
        }
        __syncthreads();
        // warp-uniform skip: a warp covers rows [8*w, 8*w+8)
        bool any = false;
        {
            int wbase = (tid >> 5) * 8;
#pragma unroll
            for (int r = 0; r < 8; ++r) any |= (s_nbr[(wbase + r) * K + k] >= 0);
        }
        if (any) {
#pragma unroll 4
            for (int c = 0; c < CIN; ++c) {
                float4 a = *reinterpret_cast<const float4 *>(&s_at[c * LDA + ty * 4]);
                float w[NC];
                if constexpr (NC == 4) {
                    float4 t = *reinterpret_cast<const float4 *>(&s_w[c * COUT + tx * 4]);
                    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
                } else if constexpr (NC == 2) {
                    float2 t = *reinterpret_cast<const float2 *>(&s_w[c * COUT + tx * 2]);
                    w[0] = t.x; w[1] = t.y;
                } else {
                    w[0] = s_w[c * COUT + tx];
                }
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    acc[0][j] = fmaf(a.x, w[j], acc[0][j]);
                    acc[1][j] = fmaf(a.y, w[j], acc[1][j]);
                    acc[2][j] = fmaf(a.z, w[j], acc[2][j]);
                    acc[3][j] = fmaf(a.w, w[j], acc[3][j]);
                }
            }
        }
        __syncthreads();
    }

    // epilogue: BN(eval) scale/shift + ReLU, one coalesced write per row
    float sc[NC], sh[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        int col = tx * NC + j;
        sc[j] = scale ? __ldg(&scale[col]) : 1.f;
        sh[j] = shift ? __ldg(&shift[col]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int r = ty * 4 + i;
        if (r >= rows_here) continue;
        float o[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            float v = scale ? fmaf(acc[i][j], sc[j], sh[j]) : (acc[i][j] + sh[j]);
            if (relu) v = fmaxf(v, 0.f);
            o[j] = v;
        }
        float *dst = feat_out + (size_t)(row0 + r) * COUT + tx * NC;
        if constexpr (NC == 4) *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        else if constexpr (NC == 2) *reinterpret_cast<float2 *>(dst) = make_float2(o[0], o[1]);
        else dst[0] = o[0];
    }
}

// any channel counts / K: one thread per output element (API completeness, not a hot path)
__global__ void k_sparse_conv_generic(const float *__restrict__ feat_in, int cin, const float *__restrict__ weight,
                                      const int *__restrict__ nbr, int K, const int *__restrict__ n_out_dev,
                                      int cap_out, const float *__restrict__ scale,
                                      const float *__restrict__ shift, int relu, float *__restrict__ feat_out,
                                      int cout)
{
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int n_out = min(*n_out_dev, cap_out);
    if (gid >= (long long)n_out * cout) return;
    int row = (int)(gid / cout), col = (int)(gid % cout);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        int src = __ldg(&nbr[(size_t)row * K + k]);
        if (src < 0) continue;
        const float *a = feat_in + (size_t)src * cin;
        const float *w = weight + (size_t)k * cin * cout + col;
        for (int c = 0; c < cin; ++c) acc = fmaf(__ldg(&a[c]), __ldg(&w[(size_t)c * cout]), acc);
    }
    float v = scale ? fmaf(acc, scale[col], shift ? shift[col] : 0.f) : (acc + (shift ? shift[col] : 0.f));
    if (relu) v = fmaxf(v, 0.f);
    feat_out[gid] = v;
}

template <int CIN, int COUT>
int launch(const float *feat_in, const float *weight, const int *nbr, int K, const int *n_out_dev, int cap_out,
           const float *scale, const float *shift, int relu, float *feat_out, cudaStream_t stream)
{
    constexpr int CINP = (CIN + 3) / 4 * 4;
    size_t smem = sizeof(int) * TM * 27 + sizeof(float) * (CINP * COUT + CINP * (TM + 4));
    k_sparse_conv<CIN, COUT><<<b2s_cdiv(cap_out, TM), kThreads, smem, stream>>>(
        feat_in, weight, nbr, K, n_out_dev, cap_out, scale, shift, relu, feat_out);
    B2S_LAUNCH_OK();
    return 0;
}

}  // namespace

extern "C" int b2s_sparse_conv(const float *feat_in, int cin, const float *weight, const int *nbr, int K,
                               const int *num_out_dev, int cap_out, const float *scale, const float *shift,
                               int relu, float *feat_out, int cout, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2S_REQUIRE(cin >= 1 && cout >= 1 && K >= 1 && cap_out >= 0, "b2s_sparse_conv: bad sizes");
    if (cap_out == 0) return 0;
#define B2S_CASE(CI, CO)                                                                                  \
    if (cin == CI && cout == CO && K <= 27)                                                               \
        return launch<CI, CO>(feat_in, weight, nbr, K, num_out_dev, cap_out, scale, shift, relu, feat_out, \
                              stream);
    B2S_CASE(3, 16) B2S_CASE(4, 16) B2S_CASE(16, 16) B2S_CASE(16, 32) B2S_CASE(32, 32) B2S_CASE(32, 64)
    B2S_CASE(64, 64)
#undef B2S_CASE
    k_sparse_conv_generic<<<b2s_cdiv((long long)cap_out * cout, 256), 256, 0, stream>>>(
        feat_in, cin, weight, nbr, K, num_out_dev, cap_out, scale, shift, relu, feat_out, cout);
    B2S_LAUNCH_OK();
    return 0;
}
