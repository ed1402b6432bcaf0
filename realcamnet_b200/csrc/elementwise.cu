// Bandwidth-bound helper kernels of the RAW->bitstream path (NHWC fp32): layout changes,
// pooling / normalisation statistics, gating, resampling, Haar DWT, depthwise convolutions.
#include <cuda_bf16.h>

#include "common.cuh"

namespace rcn {
namespace {

// ---------------------------------------------------------------- layout
// x: [N][C][HW]  ->  y: [N][HW][ld]   (32x32 smem tile transpose)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, long long HW, float* __restrict__ y, int ldy) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float* xn = x + (long long)n * C * HW;
    float* yn = y + (long long)n * HW * ldy;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? xn[(long long)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long p = p0 + i;
        const int c = c0 + threadIdx.x;
        if (c < C && p < HW) yn[p * ldy + c] = tile[threadIdx.x][i];
    }
}

// few-channel inputs (packed Bayer: 4, coordinates: 2): one thread per pixel, plane reads coalesced across threads, one contiguous
// C-float store per pixel (the 32 x 32 transposing tile above wastes 28 / 32 of its channel rows on them)
template <int C>
__global__ void nchw_to_nhwc_small_kernel(const float* __restrict__ x, long long HW, long long total, float* __restrict__ y, int ldy) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / HW, p = i - n * HW;
        const float* xn = x + n * C * HW + p;
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = xn[(long long)c * HW];
        float* yo = y + i * ldy;
        if constexpr (C == 4) {
            if ((ldy & 3) == 0 && ((uintptr_t)y & 15) == 0) { *reinterpret_cast<float4*>(yo) = make_float4(v[0], v[1], v[2], v[3]); continue; }
        }
        if constexpr (C == 2) {
            if ((ldy & 1) == 0 && ((uintptr_t)y & 7) == 0) { *reinterpret_cast<float2*>(yo) = make_float2(v[0], v[1]); continue; }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) yo[c] = v[c];
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int ldx, int C, long long HW, float* __restrict__ y) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float* xn = x + (long long)n * HW * ldx;
    float* yn = y + (long long)n * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long p = p0 + i;
        const int c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? xn[p * ldx + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long p = p0 + threadIdx.x;
        if (c < C && p < HW) yn[(long long)c * HW + p] = tile[threadIdx.x][i];
    }
}

// ---------------------------------------------------------------- channel mean (AdaptiveAvgPool2d(1))
// grid (chunks, N); block (CX, RY); deterministic two-stage reduction
__global__ void channel_sum_partial_kernel(const float* __restrict__ x, long long HW, int C, int ldx,
                                           float* __restrict__ partial, int chunks) {
    extern __shared__ float red[];  // [RY][CX]
    const int n = blockIdx.y, chunk = blockIdx.x;
    const long long per = (HW + chunks - 1) / chunks;
    const long long pbeg = chunk * per, pend = (pbeg + per < HW) ? pbeg + per : HW;
    const float* xn = x + (long long)n * HW * ldx;
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        float s = 0.f;
        if (c < C)
            for (long long p = pbeg + threadIdx.y; p < pend; p += blockDim.y) s += xn[p * ldx + c];
        red[threadIdx.y * blockDim.x + threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.y == 0 && c < C) {
            float t = 0.f;
            for (int r = 0; r < blockDim.y; ++r) t += red[r * blockDim.x + threadIdx.x];
            partial[((long long)n * chunks + chunk) * C + c] = t;
        }
        __syncthreads();
    }
}

__global__ void channel_mean_final_kernel(const float* __restrict__ partial, int chunks, int C, long long HW,
                                          float* __restrict__ mean) {
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int k = 0; k < chunks; ++k) t += partial[((long long)n * chunks + k) * C + c];
        mean[n * C + c] = t / (float)HW;
    }
}

// instance-norm statistics of small maps: one block per n, two passes (mean, centred variance)
__global__ void channel_meanvar_kernel(const float* __restrict__ x, long long HW, int C, int ldx,
                                       float* __restrict__ mean, float* __restrict__ var) {
    extern __shared__ float red[];
    const int n = blockIdx.x;
    const float* xn = x + (long long)n * HW * ldx;
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        float s = 0.f;
        if (c < C)
            for (long long p = threadIdx.y; p < HW; p += blockDim.y) s += xn[p * ldx + c];
        red[threadIdx.y * blockDim.x + threadIdx.x] = s;
        __syncthreads();
        float m = 0.f;
        for (int r = 0; r < blockDim.y; ++r) m += red[r * blockDim.x + threadIdx.x];
        m /= (float)HW;
        __syncthreads();
        float q = 0.f;
        if (c < C)
            for (long long p = threadIdx.y; p < HW; p += blockDim.y) { const float d = xn[p * ldx + c] - m; q += d * d; }
        red[threadIdx.y * blockDim.x + threadIdx.x] = q;
        __syncthreads();
        if (threadIdx.y == 0 && c < C) {
            float t = 0.f;
            for (int r = 0; r < blockDim.y; ++r) t += red[r * blockDim.x + threadIdx.x];
            mean[n * C + c] = m;
            var[n * C + c] = t / (float)HW;
        }
        __syncthreads();
    }
}

// y = (x - mean[n,c]) * rsqrt(var[n,c] + eps) * gamma[c] + beta[c]
__global__ void norm_apply_kernel(const float* __restrict__ x, int ldx, long long HW, int C, long long total,
                                  const float* __restrict__ mean, const float* __restrict__ var,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                  float* __restrict__ y, int ldy) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const int n = (int)(pix / HW);
        const float v = (x[pix * ldx + c] - mean[n * C + c]) * rsqrtf(var[n * C + c] + eps);
        y[pix * ldy + c] = v * (gamma ? gamma[c] : 1.f) + (beta ? beta[c] : 0.f);
    }
}

// y = act(x * g[n,c] + b[n,c]) (+ r)   -- CALayer scale + skip; BatchNorm(eval) affine when g,b are per-channel
__global__ void scale_add_kernel(const float* __restrict__ x, int ldx, long long HW, int C, long long total,
                                 const float* __restrict__ g, const float* __restrict__ b, int per_n,
                                 const float* __restrict__ r, int ldr, float* __restrict__ y, int ldy, int act) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const int gi = per_n ? (int)(pix / HW) * C + c : c;
        float v = x[pix * ldx + c] * g[gi];
        if (b) v += b[gi];
        v = act_apply(v, act, 0.f);
        if (r) v += r[pix * ldr + c];
        y[pix * ldy + c] = v;
    }
}

__global__ void scale_add_vec4_kernel(const float* __restrict__ x, int ldx, long long HW, int C4, long long total,
                                      const float* __restrict__ g, const float* __restrict__ b, int per_n,
                                      const float* __restrict__ r, int ldr, float* __restrict__ y, int ldy, int act) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const long long pix = i / C4;
        const int gi = per_n ? (int)(pix / HW) * C4 * 4 + c : c;
        const float4 xv = *reinterpret_cast<const float4*>(x + pix * ldx + c), gv = *reinterpret_cast<const float4*>(g + gi);
        float v[4] = {xv.x * gv.x, xv.y * gv.y, xv.z * gv.z, xv.w * gv.w};
        if (b) {
            const float4 bv = *reinterpret_cast<const float4*>(b + gi);
            v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = act_apply(v[e], act, 0.f);
        if (r) {
            const float4 rv = *reinterpret_cast<const float4*>(r + pix * ldr + c);
            v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
        }
        *reinterpret_cast<float4*>(y + pix * ldy + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// ---------------------------------------------------------------- resampling
// AvgPool2d(3, stride 2, pad 1, count_include_pad=True) followed by LeakyReLU(slope)
__global__ void avgpool3s2_kernel(const float* __restrict__ x, int H, int W, int C, int ldx, int Ho, int Wo,
                                  long long total, float slope, float* __restrict__ y, int ldy) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float s = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int hi = 2 * ho - 1 + dy;
            if ((unsigned)hi >= (unsigned)H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int wi = 2 * wo - 1 + dx;
                if ((unsigned)wi >= (unsigned)W) continue;
                s += x[((long long)(n * H + hi) * W + wi) * ldx + c];
            }
        }
        s *= (1.f / 9.f);
        y[((long long)(n * Ho + ho) * Wo + wo) * ldy + c] = s > 0.f ? s : s * slope;
    }
}

// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
__global__ void upsample2x_kernel(const float* __restrict__ x, int H, int W, int C, int ldx, long long total,
                                  float* __restrict__ y, int ldy) {
    const int Ho = 2 * H, Wo = 2 * W;
    const float sh = (Ho > 1) ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
    const float sw = (Wo > 1) ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const float hr = sh * ho, wr = sw * wo;
        const int h1 = (int)hr, w1 = (int)wr;
        const int hp = (h1 < H - 1) ? 1 : 0, wp = (w1 < W - 1) ? 1 : 0;
        const float hl1 = hr - h1, hl0 = 1.f - hl1, wl1 = wr - w1, wl0 = 1.f - wl1;
        const float* b = x + ((long long)(n * H + h1) * W + w1) * ldx + c;
        const float v = hl0 * (wl0 * b[0] + wl1 * b[(long long)wp * ldx]) +
                        hl1 * (wl0 * b[(long long)hp * W * ldx] + wl1 * b[((long long)hp * W + wp) * ldx]);
        y[((long long)(n * Ho + ho) * Wo + wo) * ldy + c] = v;
    }
}

// same, four channels per thread (16-byte accesses): the scalar kernel ran the 32-channel 1024^2 -> 2048^2 map of the condition UNet at
// 0.86 TB/s
__global__ void upsample2x_vec4_kernel(const float* __restrict__ x, int H, int W, int C4, int ldx, long long total,
                                       float* __restrict__ y, int ldy) {
    const int Ho = 2 * H, Wo = 2 * W;
    const float sh = (Ho > 1) ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
    const float sw = (Wo > 1) ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        long long t = i / C4;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const float hr = sh * ho, wr = sw * wo;
        const int h1 = (int)hr, w1 = (int)wr;
        const int hp = (h1 < H - 1) ? 1 : 0, wp = (w1 < W - 1) ? 1 : 0;
        const float hl1 = hr - h1, hl0 = 1.f - hl1, wl1 = wr - w1, wl0 = 1.f - wl1;
        const float* b = x + ((long long)(n * H + h1) * W + w1) * ldx + c;
        const float4 a00 = *reinterpret_cast<const float4*>(b), a01 = *reinterpret_cast<const float4*>(b + (long long)wp * ldx);
        const float4 a10 = *reinterpret_cast<const float4*>(b + (long long)hp * W * ldx),
                     a11 = *reinterpret_cast<const float4*>(b + ((long long)hp * W + wp) * ldx);
        float4 v;   // same expression (and rounding) as the scalar kernel
        v.x = hl0 * (wl0 * a00.x + wl1 * a01.x) + hl1 * (wl0 * a10.x + wl1 * a11.x);
        v.y = hl0 * (wl0 * a00.y + wl1 * a01.y) + hl1 * (wl0 * a10.y + wl1 * a11.y);
        v.z = hl0 * (wl0 * a00.z + wl1 * a01.z) + hl1 * (wl0 * a10.z + wl1 * a11.z);
        v.w = hl0 * (wl0 * a00.w + wl1 * a01.w) + hl1 * (wl0 * a10.w + wl1 * a11.w);
        *reinterpret_cast<float4*>(y + ((long long)(n * Ho + ho) * Wo + wo) * ldy + c) = v;
    }
}

// same interpolation, result written as the bf16 hi / lo operand planes of the conv that reads it (pixel stride ldp): the fp32 map of
// the condition UNet's up path (32 channels at 2048^2) and its rcn_split_bf16 pass disappear
__global__ void upsample2x_planes_kernel(const float* __restrict__ x, int H, int W, int C4, int ldx, long long total,
                                         __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, int ldp) {
    const int Ho = 2 * H, Wo = 2 * W;
    const float sh = (Ho > 1) ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
    const float sw = (Wo > 1) ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        long long t = i / C4;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const float hr = sh * ho, wr = sw * wo;
        const int h1 = (int)hr, w1 = (int)wr;
        const int hp = (h1 < H - 1) ? 1 : 0, wp = (w1 < W - 1) ? 1 : 0;
        const float hl1 = hr - h1, hl0 = 1.f - hl1, wl1 = wr - w1, wl0 = 1.f - wl1;
        const float* b = x + ((long long)(n * H + h1) * W + w1) * ldx + c;
        const float4 a00 = *reinterpret_cast<const float4*>(b), a01 = *reinterpret_cast<const float4*>(b + (long long)wp * ldx);
        const float4 a10 = *reinterpret_cast<const float4*>(b + (long long)hp * W * ldx),
                     a11 = *reinterpret_cast<const float4*>(b + ((long long)hp * W + wp) * ldx);
        float v[4];   // same expression (and rounding) as the fp32 kernels
        v[0] = hl0 * (wl0 * a00.x + wl1 * a01.x) + hl1 * (wl0 * a10.x + wl1 * a11.x);
        v[1] = hl0 * (wl0 * a00.y + wl1 * a01.y) + hl1 * (wl0 * a10.y + wl1 * a11.y);
        v[2] = hl0 * (wl0 * a00.z + wl1 * a01.z) + hl1 * (wl0 * a10.z + wl1 * a11.z);
        v[3] = hl0 * (wl0 * a00.w + wl1 * a01.w) + hl1 * (wl0 * a10.w + wl1 * a11.w);
        __nv_bfloat16 hb[4], lb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {     // rounding of rcn_split_bf16: hi = rn(v), lo = rn(v - hi)
            hb[e] = __float2bfloat16_rn(v[e]);
            lb[e] = __float2bfloat16_rn(v[e] - __bfloat162float(hb[e]));
        }
        const long long po = ((long long)(n * Ho + ho) * Wo + wo) * ldp + c;
        *reinterpret_cast<uint2*>(y_hi + po) = *reinterpret_cast<uint2*>(hb);
        *reinterpret_cast<uint2*>(y_lo + po) = *reinterpret_cast<uint2*>(lb);
    }
}

// Haar analysis: (N,H,W,C) -> (N,H/2,W/2,4C), channel c*4 + {LL,LH,HL,HH}
__global__ void dwt_forward_kernel(const float* __restrict__ x, int H, int W, int C, int ldx, long long total,
                                   float* __restrict__ y, int ldy) {
    const int Ho = H / 2, Wo = W / 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const float* b = x + ((long long)(n * H + 2 * ho) * W + 2 * wo) * ldx + c;
        const float a = b[0], bb = b[ldx], cc = b[(long long)W * ldx], d = b[(long long)(W + 1) * ldx];
        float* o = y + ((long long)(n * Ho + ho) * Wo + wo) * ldy + 4 * c;
        o[0] = 0.5f * a + 0.5f * bb + 0.5f * cc + 0.5f * d;
        o[1] = 0.5f * a + 0.5f * bb - 0.5f * cc - 0.5f * d;
        o[2] = 0.5f * a - 0.5f * bb + 0.5f * cc - 0.5f * d;
        o[3] = 0.5f * a - 0.5f * bb - 0.5f * cc + 0.5f * d;
    }
}

// space-to-depth by 2: (N,H,W,C) -> (N,H/2,W/2,4C), channel (i*2+j)*C + c <- pixel (2y+i, 2x+j) -- turns a learned 2x2 stride-2
// conv (the `down` layers of ISPUNet_GFM_LSC / ResUNet, models/LiteISP.py:1253,2056) into a 1x1 contraction
__global__ void space_to_depth2_kernel(const float* __restrict__ x, int H, int W, int C, int ldx, long long total,
                                       float* __restrict__ y, int ldy) {
    const int Ho = H / 2, Wo = W / 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int q = (int)(t & 3); t >>= 2;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        y[((long long)(n * Ho + ho) * Wo + wo) * ldy + q * C + c] =
            x[((long long)(n * H + 2 * ho + (q >> 1)) * W + 2 * wo + (q & 1)) * ldx + c];
    }
}

// Haar synthesis: (N,H,W,4C) -> (N,2H,2W,C)
__global__ void dwt_inverse_kernel(const float* __restrict__ x, int H, int W, int C, int ldx, long long total,
                                   float* __restrict__ y, int ldy) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        const float* b = x + ((long long)(n * H + h) * W + w) * ldx + 4 * c;
        const float ll = b[0], lh = b[1], hl = b[2], hh = b[3];
        float* o = y + ((long long)(n * 2 * H + 2 * h) * (2 * W) + 2 * w) * ldy + c;
        o[0] = 0.5f * ll + 0.5f * lh + 0.5f * hl + 0.5f * hh;
        o[ldy] = 0.5f * ll + 0.5f * lh - 0.5f * hl - 0.5f * hh;
        o[(long long)2 * W * ldy] = 0.5f * ll - 0.5f * lh + 0.5f * hl - 0.5f * hh;
        o[(long long)(2 * W + 1) * ldy] = 0.5f * ll - 0.5f * lh - 0.5f * hl + 0.5f * hh;
    }
}

// depthwise k x k conv, pad k/2, optional bias, optional "+ x" (ConvPosEnc), optional gate multiply (crpe: q * conv(v))
__global__ void depthwise_kernel(const float* __restrict__ x, int H, int W, int C, int ldx, long long total,
                                 const float* __restrict__ w /*[k*k][C]*/, const float* __restrict__ bias, int k,
                                 int add_input, const float* __restrict__ mul, int ldm, float* __restrict__ y, int ldy) {
    const int pad = k / 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int wo = (int)(t % W); t /= W;
        const int ho = (int)(t % H);
        const int n = (int)(t / H);
        float s = bias ? bias[c] : 0.f;
        for (int ky = 0; ky < k; ++ky) {
            const int hi = ho + ky - pad;
            if ((unsigned)hi >= (unsigned)H) continue;
            for (int kx = 0; kx < k; ++kx) {
                const int wi = wo + kx - pad;
                if ((unsigned)wi >= (unsigned)W) continue;
                s = fmaf(x[((long long)(n * H + hi) * W + wi) * ldx + c], w[(ky * k + kx) * C + c], s);
            }
        }
        const long long pix = (long long)(n * H + ho) * W + wo;
        if (add_input) s += x[pix * ldx + c];
        if (mul) s *= mul[pix * ldm + c];
        y[pix * ldy + c] = s;
    }
}

// Vectorised depthwise conv: one thread = 4 channels x a vertical strip of R output pixels.  Every input row of the strip is
// loaded once (K float4 per row) and feeds the up-to-K outputs it is a tap of, so a 7x7 filter needs 12 loads per output instead
// of 49; the filter taps sit in shared memory (a warp reads at most C/4 distinct float4 of them per instruction).  The scalar
// kernel above -- one thread per element, 2 scalar loads per FMA -- ran the GroupMix aggregator at 3 % of the HBM roofline.
template <int K>
__global__ void __launch_bounds__(256)
depthwise_strip_kernel(const float* __restrict__ x, int N, int H, int W, int C, int ldx, const float* __restrict__ w /*[K*K][C]*/,
                       const float* __restrict__ bias, int add_input, const float* __restrict__ mul, int ldm,
                       float* __restrict__ y, int ldy) {
    constexpr int R = 8, PAD = K / 2;
    extern __shared__ float sw[];
    for (int i = threadIdx.x; i < K * K * C; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int C4 = C >> 2, strips = (H + R - 1) / R;
    const long long total = (long long)N * strips * W * C4;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C4) * 4;
        long long t = idx / C4;
        const int wo = (int)(t % W); t /= W;
        const int st = (int)(t % strips);
        const int n = (int)(t / strips);
        const int ho0 = st * R;
        float4 acc[R];
        const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = b4;
#pragma unroll
        for (int ir = 0; ir < R + K - 1; ++ir) {
            const int hi = ho0 - PAD + ir;
            if ((unsigned)hi >= (unsigned)H) continue;
            float4 xr[K];
            const float* row = x + ((long long)(n * H + hi) * W) * ldx + c;
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int wi = wo + kx - PAD;
                xr[kx] = ((unsigned)wi < (unsigned)W) ? *reinterpret_cast<const float4*>(row + (long long)wi * ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int ky = ir - r;                 // compile-time after unrolling
                if (ky < 0 || ky >= K) continue;
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float4 wv = *reinterpret_cast<const float4*>(sw + (ky * K + kx) * C + c);
                    acc[r].x = fmaf(xr[kx].x, wv.x, acc[r].x);
                    acc[r].y = fmaf(xr[kx].y, wv.y, acc[r].y);
                    acc[r].z = fmaf(xr[kx].z, wv.z, acc[r].z);
                    acc[r].w = fmaf(xr[kx].w, wv.w, acc[r].w);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ho = ho0 + r;
            if (ho >= H) continue;
            const long long pix = (long long)(n * H + ho) * W + wo;
            float4 v = acc[r];
            if (add_input) {
                const float4 xc = *reinterpret_cast<const float4*>(x + pix * ldx + c);
                v.x += xc.x; v.y += xc.y; v.z += xc.z; v.w += xc.w;
            }
            if (mul) {
                const float4 mm = *reinterpret_cast<const float4*>(mul + pix * ldm + c);
                v.x *= mm.x; v.y *= mm.y; v.z *= mm.z; v.w *= mm.w;
            }
            *reinterpret_cast<float4*>(y + pix * ldy + c) = v;
        }
    }
}

template <int K>
void launch_depthwise_strip(const float* x, int N, int H, int W, int C, int ldx, const float* w, const float* bias, int add_input,
                            const float* mul, int ldm, float* y, int ldy, cudaStream_t s) {
    const long long total = (long long)N * ((H + 7) / 8) * W * (C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    const size_t smem = (size_t)K * K * C * sizeof(float);
    depthwise_strip_kernel<K><<<(int)blocks, 256, smem, s>>>(x, N, H, W, C, ldx, w, bias, add_input, mul, ldm, y, ldy);
}

__global__ void copy_channels_kernel(const float* __restrict__ x, int ldx, int C, long long total, float* __restrict__ y, int ldy) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        y[pix * ldy + c] = x[pix * ldx + c];
    }
}

inline int ew_blocks(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = 148LL * 32;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace
}  // namespace rcn

using namespace rcn;

extern "C" int rcn_nchw_to_nhwc(const float* x, int N, int C, int H, int W, float* y, int ldy, void* stream) {
    RCN_CHECK_ARG(x && y && N > 0 && C > 0 && ldy >= C, "rcn_nchw_to_nhwc: bad arguments");
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
    if (C == 4 || C == 2 || C == 3) {
        const long long total = (long long)N * HW;
        if (C == 4) nchw_to_nhwc_small_kernel<4><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, HW, total, y, ldy);
        else if (C == 3) nchw_to_nhwc_small_kernel<3><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, HW, total, y, ldy);
        else nchw_to_nhwc_small_kernel<2><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, HW, total, y, ldy);
    } else
        nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, C, HW, y, ldy);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_nchw_to_nhwc");
    return RCN_OK;
}

extern "C" int rcn_nhwc_to_nchw(const float* x, int ldx, int N, int C, int H, int W, float* y, void* stream) {
    RCN_CHECK_ARG(x && y && N > 0 && C > 0 && ldx >= C, "rcn_nhwc_to_nchw: bad arguments");
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
    nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, ldx, C, HW, y);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_nhwc_to_nchw");
    return RCN_OK;
}

extern "C" int rcn_channel_mean(const float* x, int N, long long HW, int C, int ldx, float* mean, float* workspace,
                                long long workspace_floats, void* stream) {
    RCN_CHECK_ARG(x && mean && workspace && N > 0 && HW > 0 && C > 0, "rcn_channel_mean: bad arguments");
    int cx = ((C + 31) / 32) * 32;
    if (cx > 256) cx = 256;
    const int ry = 256 / cx > 0 ? 256 / cx : 1;
    long long chunks = HW / (ry * 64);
    if (chunks < 1) chunks = 1;
    if (chunks > 1024) chunks = 1024;
    while (chunks > 1 && (long long)N * chunks * C > workspace_floats) chunks /= 2;
    RCN_CHECK_ARG((long long)N * chunks * C <= workspace_floats, "rcn_channel_mean: workspace too small");
    dim3 grid((unsigned)chunks, N), block(cx, ry);
    channel_sum_partial_kernel<<<grid, block, cx * ry * sizeof(float), (cudaStream_t)stream>>>(x, HW, C, ldx, workspace, (int)chunks);
    channel_mean_final_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(workspace, (int)chunks, C, HW, mean);
    count_launch(2);
    RCN_CHECK_LAUNCH("rcn_channel_mean");
    return RCN_OK;
}

extern "C" int rcn_channel_meanvar(const float* x, int N, long long HW, int C, int ldx, float* mean, float* var, void* stream) {
    RCN_CHECK_ARG(x && mean && var && N > 0 && HW > 0 && C > 0, "rcn_channel_meanvar: bad arguments");
    int cx = ((C + 31) / 32) * 32;
    if (cx > 128) cx = 128;
    const int ry = 1024 / cx;
    dim3 block(cx, ry);
    channel_meanvar_kernel<<<N, block, cx * ry * sizeof(float), (cudaStream_t)stream>>>(x, HW, C, ldx, mean, var);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_channel_meanvar");
    return RCN_OK;
}

extern "C" int rcn_norm_apply(const float* x, int ldx, int N, long long HW, int C, const float* mean, const float* var,
                              const float* gamma, const float* beta, float eps, float* y, int ldy, void* stream) {
    RCN_CHECK_ARG(x && y && mean && var, "rcn_norm_apply: null pointer");
    const long long total = (long long)N * HW * C;
    norm_apply_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, HW, C, total, mean, var, gamma, beta, eps, y, ldy);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_norm_apply");
    return RCN_OK;
}

extern "C" int rcn_scale_add(const float* x, int ldx, int N, long long HW, int C, const float* g, const float* b, int per_n,
                             const float* r, int ldr, float* y, int ldy, int act, void* stream) {
    RCN_CHECK_ARG(x && y && g, "rcn_scale_add: null pointer");
    const long long total = (long long)N * HW * C;
    auto a16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    if (C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && (!r || ldr % 4 == 0) && a16(x) && a16(y) && a16(g) && (!b || a16(b)) && (!r || a16(r)))
        scale_add_vec4_kernel<<<ew_blocks(total / 4), 256, 0, (cudaStream_t)stream>>>(x, ldx, HW, C / 4, total / 4, g, b, per_n, r, ldr, y, ldy, act);
    else
        scale_add_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, HW, C, total, g, b, per_n, r, ldr, y, ldy, act);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_scale_add");
    return RCN_OK;
}

extern "C" int rcn_avgpool3s2_lrelu(const float* x, int N, int H, int W, int C, int ldx, float slope, float* y, int ldy, void* stream) {
    RCN_CHECK_ARG(x && y, "rcn_avgpool3s2_lrelu: null pointer");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * Ho * Wo * C;
    avgpool3s2_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, H, W, C, ldx, Ho, Wo, total, slope, y, ldy);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_avgpool3s2_lrelu");
    return RCN_OK;
}

extern "C" int rcn_upsample_bilinear2x(const float* x, int N, int H, int W, int C, int ldx, float* y, int ldy, void* stream) {
    RCN_CHECK_ARG(x && y, "rcn_upsample_bilinear2x: null pointer");
    const long long total = (long long)N * 4 * H * W * C;
    if (C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0)
        upsample2x_vec4_kernel<<<ew_blocks(total / 4), 256, 0, (cudaStream_t)stream>>>(x, H, W, C / 4, ldx, total / 4, y, ldy);
    else
        upsample2x_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, H, W, C, ldx, total, y, ldy);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_upsample_bilinear2x");
    return RCN_OK;
}

extern "C" int rcn_upsample_bilinear2x_planes(const float* x, int N, int H, int W, int C, int ldx, void* y_hi, void* y_lo, int ldp,
                                              void* stream) {
    RCN_CHECK_ARG(x && y_hi && y_lo, "rcn_upsample_bilinear2x_planes: null pointer");
    RCN_CHECK_ARG(C % 4 == 0 && ldx % 4 == 0 && ldp >= C && ldp % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)y_hi % 8 == 0 &&
                      (uintptr_t)y_lo % 8 == 0,
                  "rcn_upsample_bilinear2x_planes: needs C %% 4 == 0, 16-byte aligned rows and 8-byte aligned planes");
    const long long total = (long long)N * 4 * H * W * (C / 4);
    upsample2x_planes_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, H, W, C / 4, ldx, total, (__nv_bfloat16*)y_hi,
                                                                                 (__nv_bfloat16*)y_lo, ldp);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_upsample_bilinear2x_planes");
    return RCN_OK;
}

extern "C" int rcn_dwt_forward(const float* x, int N, int H, int W, int C, int ldx, float* y, int ldy, void* stream) {
    RCN_CHECK_ARG(x && y && H % 2 == 0 && W % 2 == 0, "rcn_dwt_forward: bad arguments");
    const long long total = (long long)N * (H / 2) * (W / 2) * C;
    dwt_forward_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, H, W, C, ldx, total, y, ldy);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_dwt_forward");
    return RCN_OK;
}

extern "C" int rcn_space_to_depth2(const float* x, int N, int H, int W, int C, int ldx, float* y, int ldy, void* stream) {
    RCN_CHECK_ARG(x && y && H % 2 == 0 && W % 2 == 0 && ldy >= 4 * C, "rcn_space_to_depth2: bad arguments");
    const long long total = (long long)N * (H / 2) * (W / 2) * 4 * C;
    space_to_depth2_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, H, W, C, ldx, total, y, ldy);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_space_to_depth2");
    return RCN_OK;
}

extern "C" int rcn_dwt_inverse(const float* x, int N, int H, int W, int C4, int ldx, float* y, int ldy, void* stream) {
    RCN_CHECK_ARG(x && y && C4 % 4 == 0, "rcn_dwt_inverse: bad arguments");
    const int C = C4 / 4;
    const long long total = (long long)N * H * W * C;
    dwt_inverse_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, H, W, C, ldx, total, y, ldy);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_dwt_inverse");
    return RCN_OK;
}

extern "C" int rcn_depthwise_conv(const float* x, int N, int H, int W, int C, int ldx, const float* w, const float* bias, int k,
                                  int add_input, const float* mul, int ldm, float* y, int ldy, void* stream) {
    RCN_CHECK_ARG(x && y && w && (k & 1), "rcn_depthwise_conv: bad arguments");
    const long long total = (long long)N * H * W * C;
    const bool vec = (C % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && (!mul || ldm % 4 == 0) && ((uintptr_t)x % 16 == 0) &&
                     ((uintptr_t)y % 16 == 0) && ((uintptr_t)w % 16 == 0) && (!bias || (uintptr_t)bias % 16 == 0) &&
                     (!mul || (uintptr_t)mul % 16 == 0) && (k * k * C * 4 <= 40 * 1024);
    if (vec && k == 3) launch_depthwise_strip<3>(x, N, H, W, C, ldx, w, bias, add_input, mul, ldm, y, ldy, (cudaStream_t)stream);
    else if (vec && k == 5) launch_depthwise_strip<5>(x, N, H, W, C, ldx, w, bias, add_input, mul, ldm, y, ldy, (cudaStream_t)stream);
    else if (vec && k == 7) launch_depthwise_strip<7>(x, N, H, W, C, ldx, w, bias, add_input, mul, ldm, y, ldy, (cudaStream_t)stream);
    else
        depthwise_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, H, W, C, ldx, total, w, bias, k, add_input, mul, ldm, y, ldy);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_depthwise_conv");
    return RCN_OK;
}

extern "C" int rcn_copy_channels(const float* x, int ldx, long long npix, int C, float* y, int ldy, void* stream) {
    RCN_CHECK_ARG(x && y, "rcn_copy_channels: null pointer");
    const long long total = npix * C;
    copy_channels_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, C, total, y, ldy);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_copy_channels");
    return RCN_OK;
}
