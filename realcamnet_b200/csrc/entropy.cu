// Entropy-model kernels: factorized-prior bottleneck on z and the Gaussian conditional on y slices.
// Pure element-wise work (HBM-bound): ~20 B/symbol for the Gaussian kernel.
#include "common.cuh"

namespace rcn {
namespace {

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// per-channel packed parameters (58 floats): softplus(matrix) / bias / tanh(factor) for the
// 1-3-3-3-3-1 cumulative MLP, in layer order: [m0(3) b0(3) f0(3)] [m(9) b(3) f(3)]x3 [m4(3) b4(1)]
__device__ __forceinline__ float eb_logits(const float* __restrict__ P, float v) {
    float h[3], g[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float t = P[i] * v + P[3 + i];
        h[i] = t + P[6 + i] * tanhf(t);
    }
    const float* q = P + 9;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float t = q[i * 3 + 0] * h[0];
            t += q[i * 3 + 1] * h[1];
            t += q[i * 3 + 2] * h[2];
            t += q[9 + i];
            g[i] = t + q[12 + i] * tanhf(t);
        }
        h[0] = g[0]; h[1] = g[1]; h[2] = g[2];
        q += 15;
    }
    float t = q[0] * h[0];
    t += q[1] * h[1];
    t += q[2] * h[2];
    return t + q[3];
}

// z: NHWC (npix, C) ; outputs optional
__global__ void eb_forward_kernel(const float* __restrict__ z, int ldz, long long npix, int C, long long HW,
                                  const float* __restrict__ params, const float* __restrict__ medians,
                                  float* __restrict__ z_hat, int ldzh, float* __restrict__ lik, int ldl,
                                  int* __restrict__ symbols /* NCHW order */, float lik_bound) {
    const long long total = npix * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const float med = medians[c];
        const float q = rintf(z[pix * ldz + c] - med);
        const float zh = q + med;
        if (z_hat) z_hat[pix * ldzh + c] = zh;
        if (symbols) {
            const long long n = pix / HW, p = pix - n * HW;
            symbols[(n * C + c) * HW + p] = (int)q;
        }
        if (lik) {
            const float* P = params + (size_t)c * 58;
            const float lower = eb_logits(P, zh - 0.5f);
            const float upper = eb_logits(P, zh + 0.5f);
            const float sum = lower + upper;
            const float sign = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
            float l = fabsf(sigmoidf_(sign * upper) - sigmoidf_(sign * lower));
            lik[pix * ldl + c] = fmaxf(l, lik_bound);
        }
    }
}

// dequantise decoded symbols: z_hat = symbol + median (EntropyBottleneck.decompress)
__global__ void eb_dequant_kernel(const int* __restrict__ symbols /*NCHW*/, long long npix, int C, long long HW,
                                  const float* __restrict__ medians, float* __restrict__ z_hat, int ldzh) {
    const long long total = npix * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const long long n = pix / HW, p = pix - n * HW;
        z_hat[pix * ldzh + c] = (float)symbols[(n * C + c) * HW + p] + medians[c];
    }
}

struct RansPrep {
    const int* cdf; int cdf_stride; const int* cdf_len; const int* cdf_off;
    unsigned* packed; unsigned* raw; unsigned char* flags;   // per symbol, coding order; raw valid where flags != 0
};

// Gaussian conditional on one slice: y, mu, scale are NHWC channel slices (npix, C).
__global__ void gaussian_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ mu, int ldm,
                                const float* __restrict__ scale, int lds, long long npix, int C, long long HW,
                                const float* __restrict__ table, int ntable, float scale_bound, float lik_bound,
                                float* __restrict__ y_hat, int ldyh, float* __restrict__ lik, int ldl,
                                int* __restrict__ symbols, int* __restrict__ indexes, RansPrep rp) {
    __shared__ float tab[64];
    for (int i = threadIdx.x; i < ntable && i < 64; i += blockDim.x) tab[i] = table[i];
    __syncthreads();
    const long long total = npix * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const float m = mu[pix * ldm + c];
        const float q = rintf(y[pix * ldy + c] - m);  // torch.round = round-half-to-even
        const float yh = q + m;
        if (y_hat) y_hat[pix * ldyh + c] = yh;
        const float s = fmaxf(scale[pix * lds + c], scale_bound);
        if (lik) {
            // likelihood of the dequantised value: values = |y_hat - mu| (GaussianConditional._likelihood)
            const float v = fabsf(yh - m);
            const float k = -0.70710678118654752440f;  // -(2 ** -0.5)
            const float upper = 0.5f * erfcf(k * ((0.5f - v) / s));
            const float lower = 0.5f * erfcf(k * ((-0.5f - v) / s));
            lik[pix * ldl + c] = fmaxf(upper - lower, lik_bound);
        }
        if (symbols) {
            const long long n = pix / HW, p = pix - n * HW;
            const long long o = (n * C + c) * HW + p;
            symbols[o] = (int)q;
            int idx = ntable - 1;
            for (int t = 0; t < ntable - 1; ++t) idx -= (s <= tab[t]) ? 1 : 0;
            indexes[o] = idx;
            if (rp.packed) {
                // range-coder front end (the per-symbol CDF lookup of BufferedRansEncoder.encode_with_indexes,
                // raw2bit.py:1956): (start << 16) | (freq - 1); out-of-range values take the sentinel bin, set their
                // flag byte and leave the bypass payload in raw[] for the host state chain.
                const int sentinel = rp.cdf_len[idx] - 2;
                int v = (int)q - rp.cdf_off[idx];
                unsigned raw = 0u;
                bool escaped = false;
                if (v < 0) { raw = (unsigned)(-2 * (long long)v - 1); v = sentinel; escaped = true; }
                else if (v >= sentinel) { raw = (unsigned)(2 * ((long long)v - sentinel)); v = sentinel; escaped = true; }
                const int* row = rp.cdf + (long long)idx * rp.cdf_stride;
                const unsigned st = (unsigned)row[v], fr = (unsigned)(row[v + 1] - row[v]);
                rp.packed[o] = (st << 16) | ((fr - 1u) & 0xFFFFu);
                rp.flags[o] = escaped ? 1 : 0;
                if (escaped) rp.raw[o] = raw;
            }
        }
    }
}

// y_hat = symbol + mu  (GaussianConditional.dequantize in decompress(), raw2bit.py:2014-2015)
__global__ void gaussian_dequant_kernel(const int* __restrict__ symbols, const float* __restrict__ mu, int ldm,
                                        long long npix, int C, long long HW, float* __restrict__ y_hat, int ldyh) {
    const long long total = npix * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const long long n = pix / HW, p = pix - n * HW;
        y_hat[pix * ldyh + c] = (float)symbols[(n * C + c) * HW + p] + mu[pix * ldm + c];
    }
}

// indexes only (decoder side): raw2bit.py:2011
__global__ void build_indexes_kernel(const float* __restrict__ scale, int lds, long long npix, int C, long long HW,
                                     const float* __restrict__ table, int ntable, float scale_bound, int* __restrict__ indexes) {
    __shared__ float tab[64];
    for (int i = threadIdx.x; i < ntable && i < 64; i += blockDim.x) tab[i] = table[i];
    __syncthreads();
    const long long total = npix * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const float s = fmaxf(scale[pix * lds + c], scale_bound);
        int idx = ntable - 1;
        for (int t = 0; t < ntable - 1; ++t) idx -= (s <= tab[t]) ? 1 : 0;
        const long long n = pix / HW, p = pix - n * HW;
        indexes[(n * C + c) * HW + p] = idx;
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

extern "C" int rcn_eb_forward(const float* z, int ldz, int N, long long HW, int C, const float* params, const float* medians,
                              float* z_hat, int ldzh, float* lik, int ldl, int* symbols, float lik_bound, void* stream) {
    RCN_CHECK_ARG(z && params && medians, "rcn_eb_forward: null pointer");
    const long long npix = (long long)N * HW;
    eb_forward_kernel<<<ew_blocks(npix * C), 256, 0, (cudaStream_t)stream>>>(z, ldz, npix, C, HW, params, medians, z_hat, ldzh,
                                                                             lik, ldl, symbols, lik_bound);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_eb_forward");
    return RCN_OK;
}

extern "C" int rcn_eb_dequantize(const int* symbols, int N, long long HW, int C, const float* medians, float* z_hat, int ldzh, void* stream) {
    RCN_CHECK_ARG(symbols && medians && z_hat, "rcn_eb_dequantize: null pointer");
    const long long npix = (long long)N * HW;
    eb_dequant_kernel<<<ew_blocks(npix * C), 256, 0, (cudaStream_t)stream>>>(symbols, npix, C, HW, medians, z_hat, ldzh);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_eb_dequantize");
    return RCN_OK;
}

extern "C" int rcn_gaussian_conditional(const float* y, int ldy, const float* mu, int ldm, const float* scale, int lds, int N,
                                        long long HW, int C, const float* table, int ntable, float scale_bound, float lik_bound,
                                        float* y_hat, int ldyh, float* lik, int ldl, int* symbols, int* indexes, void* stream) {
    RCN_CHECK_ARG(y && mu && scale && table, "rcn_gaussian_conditional: null pointer");
    RCN_CHECK_ARG(ntable >= 2 && ntable <= 64, "rcn_gaussian_conditional: scale table must have 2..64 entries");
    RCN_CHECK_ARG((symbols == nullptr) == (indexes == nullptr), "rcn_gaussian_conditional: symbols and indexes come together");
    const long long npix = (long long)N * HW;
    gaussian_kernel<<<ew_blocks(npix * C), 256, 0, (cudaStream_t)stream>>>(y, ldy, mu, ldm, scale, lds, npix, C, HW, table, ntable,
                                                                           scale_bound, lik_bound, y_hat, ldyh, lik, ldl, symbols, indexes,
                                                                           RansPrep{});
    count_launch();
    RCN_CHECK_LAUNCH("rcn_gaussian_conditional");
    return RCN_OK;
}

extern "C" int rcn_gaussian_conditional_coded(const float* y, int ldy, const float* mu, int ldm, const float* scale, int lds, int N,
                                              long long HW, int C, const float* table, int ntable, float scale_bound,
                                              float lik_bound, float* y_hat, int ldyh, float* lik, int ldl, int* symbols,
                                              int* indexes, const int* cdf, int cdf_stride, const int* cdf_len, const int* cdf_off,
                                              unsigned* packed, unsigned* raw, unsigned char* flags, void* stream) {
    RCN_CHECK_ARG(y && mu && scale && table && symbols && indexes, "rcn_gaussian_conditional_coded: null pointer");
    RCN_CHECK_ARG(cdf && cdf_len && cdf_off && packed && raw && flags, "rcn_gaussian_conditional_coded: coder tables / outputs missing");
    RCN_CHECK_ARG(ntable >= 2 && ntable <= 64, "rcn_gaussian_conditional_coded: scale table must have 2..64 entries");
    const long long npix = (long long)N * HW;
    RansPrep rp{cdf, cdf_stride, cdf_len, cdf_off, packed, raw, flags};
    gaussian_kernel<<<ew_blocks(npix * C), 256, 0, (cudaStream_t)stream>>>(y, ldy, mu, ldm, scale, lds, npix, C, HW, table, ntable,
                                                                           scale_bound, lik_bound, y_hat, ldyh, lik, ldl, symbols, indexes, rp);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_gaussian_conditional_coded");
    return RCN_OK;
}

extern "C" int rcn_gaussian_dequantize(const int* symbols, const float* mu, int ldm, int N, long long HW, int C, float* y_hat,
                                       int ldyh, void* stream) {
    RCN_CHECK_ARG(symbols && mu && y_hat, "rcn_gaussian_dequantize: null pointer");
    const long long npix = (long long)N * HW;
    gaussian_dequant_kernel<<<ew_blocks(npix * C), 256, 0, (cudaStream_t)stream>>>(symbols, mu, ldm, npix, C, HW, y_hat, ldyh);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_gaussian_dequantize");
    return RCN_OK;
}

extern "C" int rcn_build_indexes(const float* scale, int lds, int N, long long HW, int C, const float* table, int ntable,
                                 float scale_bound, int* indexes, void* stream) {
    RCN_CHECK_ARG(scale && table && indexes && ntable >= 2 && ntable <= 64, "rcn_build_indexes: bad arguments");
    const long long npix = (long long)N * HW;
    build_indexes_kernel<<<ew_blocks(npix * C), 256, 0, (cudaStream_t)stream>>>(scale, lds, npix, C, HW, table, ntable, scale_bound, indexes);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_build_indexes");
    return RCN_OK;
}
