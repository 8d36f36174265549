// tc_util.cu -- hi/lo (3xTF32) plane conversions between the fp32 kernels and the tensor-core kernels.
#include "common.cuh"

namespace {

__device__ __forceinline__ float rn_tf32(float v)
{
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

__global__ void k_split_tf32(const float4 *__restrict__ x, float4 *__restrict__ hi, float4 *__restrict__ lo,
                             const int *__restrict__ n_rows_dev, int row_floats, long long cap4)
{
    long long n4 = n_rows_dev ? min(cap4, (long long)(*n_rows_dev) * row_floats / 4) : cap4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(&x[i]);
        float4 h = make_float4(rn_tf32(v.x), rn_tf32(v.y), rn_tf32(v.z), rn_tf32(v.w));
        hi[i] = h;
        lo[i] = make_float4(rn_tf32(v.x - h.x), rn_tf32(v.y - h.y), rn_tf32(v.z - h.z), rn_tf32(v.w - h.w));
    }
}

__global__ void k_merge_hilo(const float4 *__restrict__ hi, const float4 *__restrict__ lo, float4 *__restrict__ x,
                             const int *__restrict__ n_rows_dev, int row_floats, long long cap4)
{
    long long n4 = n_rows_dev ? min(cap4, (long long)(*n_rows_dev) * row_floats / 4) : cap4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 h = __ldg(&hi[i]), l = __ldg(&lo[i]);
        x[i] = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
    }
}

}  // namespace

// x [rows, row_floats] fp32 -> hi = tf32-rounded, lo = tf32-rounded remainder.  rows = *num_rows_dev (or cap_rows
// when num_rows_dev is NULL); row_floats must be a multiple of 4.
extern "C" int b2s_split_tf32(const float *x, float *hi, float *lo, const int *num_rows_dev, int cap_rows,
                              int row_floats, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2S_REQUIRE(row_floats % 4 == 0 && cap_rows >= 0, "b2s_split_tf32: row_floats must be a multiple of 4");
    long long cap4 = (long long)cap_rows * row_floats / 4;
    if (cap4 == 0) return 0;
    k_split_tf32<<<(int)(b2s_cdiv(cap4, 256) < 1184 ? b2s_cdiv(cap4, 256) : 1184), 256, 0, stream>>>((const float4 *)x, (float4 *)hi, (float4 *)lo, num_rows_dev,
                                                          row_floats, cap4);
    B2S_LAUNCH_OK();
    return 0;
}

extern "C" int b2s_merge_hilo(const float *hi, const float *lo, float *x, const int *num_rows_dev, int cap_rows,
                              int row_floats, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    B2S_REQUIRE(row_floats % 4 == 0 && cap_rows >= 0, "b2s_merge_hilo: row_floats must be a multiple of 4");
    long long cap4 = (long long)cap_rows * row_floats / 4;
    if (cap4 == 0) return 0;
    k_merge_hilo<<<(int)(b2s_cdiv(cap4, 256) < 1184 ? b2s_cdiv(cap4, 256) : 1184), 256, 0, stream>>>((const float4 *)hi, (const float4 *)lo, (float4 *)x,
                                                          num_rows_dev, row_floats, cap4);
    B2S_LAUNCH_OK();
    return 0;
}
