// icn.cu -- the pieces of the ICN generator G_Resnet (warp_learn/models.py:15-208 of the reference) that sit
// between its convolutions: reflection-padded layout conversion, InstanceNorm / LayerNorm statistics, and the fused
// normalise + activation + residual + nearest-upsample + reflection-pad pass that writes the next convolution's
// (bordered) input.  The convolutions run on fusg_conv2d with pad_mode = 1 (conv.cu).  All HBM-bound, NHWC.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/fusg.h"
#include "fusg_common.h"

namespace fusg_icn {

__device__ __forceinline__ int reflect(int i, int n) {      // nn.ReflectionPad2d index map, |overhang| < n
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

template <typename T> struct Vec;
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static void load(const __nv_bfloat16 *p, float *v) {
        const uint4 q = *reinterpret_cast<const uint4 *>(p);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
    __device__ static void store(__nv_bfloat16 *p, const float *v) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct Vec<__half> {
    static constexpr int N = 8;
    __device__ static void load(const __half *p, float *v) {
        const uint4 q = *reinterpret_cast<const uint4 *>(p);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    __device__ static void store(__half *p, const float *v) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct Vec<float> {
    static constexpr int N = 4;
    __device__ static void load(const float *p, float *v) {
        const float4 q = *reinterpret_cast<const float4 *>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    __device__ static void store(float *p, const float *v) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

// ---- NCHW fp32 -> bordered NHWC, reflection padded; one thread per output pixel --------------------------------
template <typename T>
__global__ void k_nchw_to_nhwc_reflect(const float *__restrict__ in, T *__restrict__ out, int C, int H, int W, int cpad, int border, size_t npix) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int Hp = H + 2 * border, Wp = W + 2 * border;
    const size_t b = i / ((size_t)Hp * Wp);
    const int r = (int)(i - b * (size_t)Hp * Wp), Y = r / Wp, X = r - Y * Wp;
    const int sy = reflect(Y - border, H), sx = reflect(X - border, W);
    const float *src = in + b * (size_t)C * H * W + (size_t)sy * W + sx;
    T *dst = out + i * cpad;
    constexpr int N = Vec<T>::N;
    for (int c0 = 0; c0 < cpad; c0 += N) {
        float v[N];
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = c0 + k < C ? __ldg(src + (size_t)(c0 + k) * H * W) : 0.f;
        Vec<T>::store(dst + c0, v);
    }
}

// ---- partial sums: grid (nsplit, B), block 256 = (C/N channel groups) x (256/(C/N) pixel lanes) ---------------------
constexpr int ST_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(ST_THREADS) k_norm_stats(const T *__restrict__ x, float *__restrict__ partial, int HW, int C, int nsplit) {
    constexpr int N = Vec<T>::N;
    __shared__ float s_sum[ST_THREADS * 8], s_sq[ST_THREADS * 8];
    const int groups = C / N, lanes = ST_THREADS / groups;
    const int g = threadIdx.x % groups, l = threadIdx.x / groups;
    const int split = blockIdx.x, b = blockIdx.y;
    const int per = (HW + nsplit - 1) / nsplit, p0 = split * per, p1 = min(HW, p0 + per);
    // Shifted sums: the statistics are taken of (x - k) with k = the channel's value at pixel 0 of the sample, so that
    // E[(x-k)^2] - E[x-k]^2 does not cancel when |mean| >> std (a one-pass sum of x and x^2 in fp32 does; the reference's
    // InstanceNorm uses Welford).  Every split of a sample uses the same k; the finalize kernel adds it back.
    float sum[N], sq[N], shift[N];
#pragma unroll
    for (int k = 0; k < N; ++k) sum[k] = sq[k] = 0.f;
    if (l < lanes) {
        const T *base = x + ((size_t)b * HW) * C + g * N;
        Vec<T>::load(base, shift);
        for (int p = p0 + l; p < p1; p += lanes) {
            float v[N];
            Vec<T>::load(base + (size_t)p * C, v);
#pragma unroll
            for (int k = 0; k < N; ++k) { const float dv = v[k] - shift[k]; sum[k] += dv; sq[k] += dv * dv; }
        }
        if (split == 0 && l == 0) {
#pragma unroll
            for (int k = 0; k < N; ++k) partial[(size_t)gridDim.y * nsplit * C * 2 + (size_t)b * C + g * N + k] = shift[k];
        }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) { s_sum[threadIdx.x * N + k] = sum[k]; s_sq[threadIdx.x * N + k] = sq[k]; }
    __syncthreads();
    // fixed-order reduction over the pixel lanes: thread c < C sums lane 0, 1, ... for channel c
    for (int c = threadIdx.x; c < C; c += ST_THREADS) {
        const int gg = c / N, k = c - gg * N;
        float a = 0.f, q = 0.f;
        for (int ll = 0; ll < lanes; ++ll) { a += s_sum[(ll * groups + gg) * N + k]; q += s_sq[(ll * groups + gg) * N + k]; }
        float *o = partial + (((size_t)b * nsplit + split) * C + c) * 2;
        o[0] = a; o[1] = q;
    }
}

// ---- finalize: one block per sample -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_norm_finalize(const float *__restrict__ partial, const float *__restrict__ gamma,
                                                       const float *__restrict__ beta, float *__restrict__ ss, int HW, int C, int nsplit,
                                                       int kind, float eps) {
    __shared__ double s_a[256], s_q[256];
    const int b = blockIdx.x;
    const float *shifts = partial + (size_t)gridDim.x * nsplit * C * 2 + (size_t)b * C;      // k of the shifted sums
    double tot_a = 0, tot_q = 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double a = 0, q = 0;
        for (int s = 0; s < nsplit; ++s) {
            const float *pp = partial + (((size_t)b * nsplit + s) * C + c) * 2;
            a += (double)pp[0]; q += (double)pp[1];
        }
        const double k = (double)shifts[c];
        if (kind == 0) {
            const double dm = a / HW, mean = k + dm, var = fmax(q / HW - dm * dm, 0.0);
            const double sc = 1.0 / sqrt(var + (double)eps);
            ss[((size_t)b * C + c) * 2] = (float)sc;
            ss[((size_t)b * C + c) * 2 + 1] = (float)(-mean * sc);
        }
        // un-shift for the per-sample statistics: sum x = a + HW k, sum x^2 = q + 2 k a + HW k^2 (fp64)
        tot_a += a + (double)HW * k;
        tot_q += q + 2.0 * k * a + (double)HW * k * k;
    }
    if (kind != 1) return;
    s_a[threadIdx.x] = tot_a; s_q[threadIdx.x] = tot_q;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) { s_a[threadIdx.x] += s_a[threadIdx.x + off]; s_q[threadIdx.x] += s_q[threadIdx.x + off]; }
        __syncthreads();
    }
    const double n = (double)HW * C, mean = s_a[0] / n;
    const double var = fmax((s_q[0] - n * mean * mean) / (n - 1.0), 0.0);      // torch.std: unbiased
    const double inv = 1.0 / (sqrt(var) + (double)eps);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double gm = gamma ? (double)gamma[c] : 1.0, bt = beta ? (double)beta[c] : 0.0;
        ss[((size_t)b * C + c) * 2] = (float)(gm * inv);
        ss[((size_t)b * C + c) * 2 + 1] = (float)(bt - mean * inv * gm);
    }
}

// ---- apply: one CTA per output row (b, Y); threads stride over (X, N-channel group) with 32-bit index math ---------
template <typename T>
__global__ void __launch_bounds__(256) k_norm_apply(const T *__restrict__ x, const float *__restrict__ ss, const T *__restrict__ residual, int rb,
                                                    T *__restrict__ out, int H, int W, int C, int relu, int up, int border) {
    constexpr int N = Vec<T>::N;
    extern __shared__ float2 s_ss[];                      // this sample's C scale/shift pairs
    const int Y = blockIdx.x, b = blockIdx.y;
    const int groups = C / N;
    const int Hu = H * up, Wu = W * up, Wp = Wu + 2 * border, Hp = Hu + 2 * border;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s_ss[c] = reinterpret_cast<const float2 *>(ss)[(size_t)b * C + c];
    __syncthreads();
    const int sy = reflect(Y - border, Hu) / up;
    const T *xrow = x + ((size_t)b * H + sy) * (size_t)W * C;
    const T *rrow = residual ? residual + (((size_t)b * (H + 2 * rb) + sy + rb) * (size_t)(W + 2 * rb) + rb) * C : nullptr;
    T *orow = out + ((size_t)b * Hp + Y) * (size_t)Wp * C;
    const int total = Wp * groups;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int X = i / groups, g = i - X * groups;
        const int sx = reflect(X - border, Wu) / up;
        float v[N];
        Vec<T>::load(xrow + (size_t)sx * C + g * N, v);
#pragma unroll
        for (int k = 0; k < N; ++k) { const float2 t = s_ss[g * N + k]; v[k] = v[k] * t.x + t.y; }
        if (rrow) {
            float rv[N];
            Vec<T>::load(rrow + (size_t)sx * C + g * N, rv);
#pragma unroll
            for (int k = 0; k < N; ++k) v[k] += rv[k];
        }
        if (relu) {
#pragma unroll
            for (int k = 0; k < N; ++k) v[k] = fmaxf(v[k], 0.f);
        }
        Vec<T>::store(orow + (size_t)i * N, v);
    }
}

}  // namespace fusg_icn

using namespace fusg_icn;

extern "C" int fusg_nchw_to_nhwc_reflect(const float *in, void *out, int B, int C, int H, int W, int cpad, int border, int dtype, void *stream) {
    if (!in || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || cpad < C || border < 0) return FUSG_ERR_ARG;
    if (cpad % 8 != 0 || border >= H || border >= W) return FUSG_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npix = (size_t)B * (H + 2 * border) * (W + 2 * border);
    const unsigned grid = (unsigned)((npix + 255) / 256);
    if (dtype == FUSG_DTYPE_BF16) k_nchw_to_nhwc_reflect<__nv_bfloat16><<<grid, 256, 0, st>>>(in, (__nv_bfloat16 *)out, C, H, W, cpad, border, npix);
    else if (dtype == FUSG_DTYPE_F16) k_nchw_to_nhwc_reflect<__half><<<grid, 256, 0, st>>>(in, (__half *)out, C, H, W, cpad, border, npix);
    else k_nchw_to_nhwc_reflect<float><<<grid, 256, 0, st>>>(in, (float *)out, C, H, W, cpad, border, npix);
    fusg_count_launch(1);
    return fusg_check_launch();
}

extern "C" int fusg_norm_stats(const void *x, float *partial, int B, int HW, int C, int nsplit, int dtype, void *stream) {
    if (!x || !partial || B <= 0 || HW <= 0 || C <= 0 || nsplit <= 0) return FUSG_ERR_ARG;
    const bool half16 = dtype == FUSG_DTYPE_BF16 || dtype == FUSG_DTYPE_F16;
    if (C % 8 != 0 || C > 256 || B > 65535 || (256 % (C / 8)) != 0 || (!half16 && (256 % (C / 4)) != 0)) return FUSG_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)nsplit, (unsigned)B);
    if (dtype == FUSG_DTYPE_BF16) k_norm_stats<__nv_bfloat16><<<grid, ST_THREADS, 0, st>>>((const __nv_bfloat16 *)x, partial, HW, C, nsplit);
    else if (dtype == FUSG_DTYPE_F16) k_norm_stats<__half><<<grid, ST_THREADS, 0, st>>>((const __half *)x, partial, HW, C, nsplit);
    else k_norm_stats<float><<<grid, ST_THREADS, 0, st>>>((const float *)x, partial, HW, C, nsplit);
    fusg_count_launch(1);
    return fusg_check_launch();
}

extern "C" int fusg_norm_finalize(const float *partial, const float *gamma, const float *beta, float *ss, int B, int HW, int C, int nsplit,
                                  int kind, float eps, void *stream) {
    if (!partial || !ss || B <= 0 || HW <= 0 || C <= 0 || nsplit <= 0 || (kind != 0 && kind != 1)) return FUSG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    k_norm_finalize<<<B, 256, 0, st>>>(partial, gamma, beta, ss, HW, C, nsplit, kind, eps);
    fusg_count_launch(1);
    return fusg_check_launch();
}

extern "C" int fusg_norm_apply(const void *x, const float *ss, const void *residual, int rb, void *out, int B, int H, int W, int C, int relu,
                               int up, int border, int dtype, void *stream) {
    if (!x || !ss || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || border < 0 || rb < 0) return FUSG_ERR_ARG;
    if (C % 8 != 0 || (up != 1 && up != 2) || border >= H * up || border >= W * up) return FUSG_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (B > 65535) return FUSG_ERR_UNSUPPORTED;
    dim3 grid((unsigned)(H * up + 2 * border), (unsigned)B);
    const size_t smem = (size_t)C * sizeof(float2);
    if (dtype == FUSG_DTYPE_BF16)
        k_norm_apply<__nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16 *)x, ss, (const __nv_bfloat16 *)residual, rb, (__nv_bfloat16 *)out, H, W, C, relu, up, border);
    else if (dtype == FUSG_DTYPE_F16)
        k_norm_apply<__half><<<grid, 256, smem, st>>>((const __half *)x, ss, (const __half *)residual, rb, (__half *)out, H, W, C, relu, up, border);
    else
        k_norm_apply<float><<<grid, 256, smem, st>>>((const float *)x, ss, (const float *)residual, rb, (float *)out, H, W, C, relu, up, border);
    fusg_count_launch(1);
    return fusg_check_launch();
}
