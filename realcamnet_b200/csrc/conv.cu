// fp32 implicit-GEMM convolution with fused epilogues (CUDA-core FFMA path).
//
// This is the exact-arithmetic engine: every conv / linear on the RAW->bitstream path can run
// through it with fp32 accumulation in the same order class as the reference's F.conv2d, so the
// 1e-3 parity bar holds through ~100 layers and a hard round().  Dense 3x3 / 1x1 layers with
// Cin % 64 == 0 are additionally served by the tcgen05 engine in conv_tc.cu.
//
// GEMM view:  M = N*Ho*Wo output pixels,  N = Cout,  K = k*k*Cin  (A gathered on the fly, NHWC).
// Tile 128 x BN x 16, 256 threads, 8 x (BN/16) register tile, double-buffered shared memory.
#include "common.cuh"

namespace rcn {

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int APAD = 4;

struct RowInfo {
    int n, hi0, wi0;  // hi0 = ho*stride - pad ; n < 0 marks a row past M
};

template <int BN>
__global__ void __launch_bounds__(256) conv2d_kernel(const rcn_conv_desc p, int Ho, int Wo, long long M, int K, int vecA) {
    constexpr int TN = BN / 16;  // 8, 4, 2, 1
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN];
    __shared__ RowInfo rows[BM];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int pad = p.k >> 1;

    for (int r = tid; r < BM; r += 256) {
        long long m = m0 + r;
        RowInfo ri;
        if (m < M) {
            int wo = (int)(m % Wo);
            long long t = m / Wo;
            int ho = (int)(t % Ho);
            ri.n = (int)(t / Ho);
            ri.hi0 = ho * p.stride - pad;
            ri.wi0 = wo * p.stride - pad;
        } else {
            ri.n = -1; ri.hi0 = 0; ri.wi0 = 0;
        }
        rows[r] = ri;
    }
    __syncthreads();

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (K + BK - 1) / BK;
    const bool vecB = (p.Cout & 3) == 0;

    float4 ra[2];
    float rs[8];
    float4 rb[2];
    float rbs[8];

    auto load_tiles = [&](int kt) {
        const int k0 = kt * BK;
        if (vecA) {
            // whole 16-chunk lies inside one tap: k0 = tap*Cin + c0
            const int tap = k0 / p.Cin, c0 = k0 - tap * p.Cin;
            const int ky = tap / p.k, kx = tap - ky * p.k;
            const int kv = tid & 3;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = (tid >> 2) + 64 * j;
                const RowInfo ri = rows[r];
                const int hi = ri.hi0 + ky, wi = ri.wi0 + kx;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ri.n >= 0 && (unsigned)hi < (unsigned)p.H && (unsigned)wi < (unsigned)p.W) {
                    const float* src = p.x + ((long long)(ri.n * p.H + hi) * p.W + wi) * p.ldx + c0 + kv * 4;
                    v = *reinterpret_cast<const float4*>(src);
                    if (p.in_square) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
                }
                ra[j] = v;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int e = tid + 256 * j;
                const int kk = e & 15, r = e >> 4;
                const int k = k0 + kk;
                float v = 0.f;
                if (k < K) {
                    const int tap = k / p.Cin, c = k - tap * p.Cin;
                    const int ky = tap / p.k, kx = tap - ky * p.k;
                    const RowInfo ri = rows[r];
                    const int hi = ri.hi0 + ky, wi = ri.wi0 + kx;
                    if (ri.n >= 0 && (unsigned)hi < (unsigned)p.H && (unsigned)wi < (unsigned)p.W) {
                        v = p.x[((long long)(ri.n * p.H + hi) * p.W + wi) * p.ldx + c];
                        if (p.in_square) v *= v;
                    }
                }
                rs[j] = v;
            }
        }
        if (vecB) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int e = tid + 256 * j;
                if (e < BK * BN / 4) {
                    const int kk = e / (BN / 4), nn = (e - kk * (BN / 4)) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k0 + kk < K && n0 + nn < p.Cout)
                        v = *reinterpret_cast<const float4*>(p.w + (long long)(k0 + kk) * p.Cout + n0 + nn);
                    rb[j] = v;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int e = tid + 256 * j;
                if (e < BK * BN) {
                    const int kk = e / BN, nn = e - kk * BN;
                    float v = 0.f;
                    if (k0 + kk < K && n0 + nn < p.Cout) v = p.w[(long long)(k0 + kk) * p.Cout + n0 + nn];
                    rbs[j] = v;
                }
            }
        }
    };

    auto store_tiles = [&](int buf) {
        if (vecA) {
            const int kv = tid & 3;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = (tid >> 2) + 64 * j;
                As[buf][kv * 4 + 0][r] = ra[j].x;
                As[buf][kv * 4 + 1][r] = ra[j].y;
                As[buf][kv * 4 + 2][r] = ra[j].z;
                As[buf][kv * 4 + 3][r] = ra[j].w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int e = tid + 256 * j;
                As[buf][e & 15][e >> 4] = rs[j];
            }
        }
        if (vecB) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int e = tid + 256 * j;
                if (e < BK * BN / 4) {
                    const int kk = e / (BN / 4), nn = (e - kk * (BN / 4)) * 4;
                    *reinterpret_cast<float4*>(&Bs[buf][kk][nn]) = rb[j];
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int e = tid + 256 * j;
                if (e < BK * BN) Bs[buf][e / BN][e % BN] = rbs[j];
            }
        }
    };

    load_tiles(0);
    store_tiles(0);
    __syncthreads();

    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if constexpr (TN == 8) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            } else if constexpr (TN == 4) {
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            } else if constexpr (TN == 2) {
                const float2 b0 = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
                b[0] = b0.x; b[1] = b0.y;
            } else {
                b[0] = Bs[buf][kk][tx];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue
    const int Hs = (p.store == RCN_STORE_PS2 || p.store == RCN_STORE_PS2_NCHW) ? 2 * Ho : Ho;
    const int Ws = (p.store == RCN_STORE_PS2 || p.store == RCN_STORE_PS2_NCHW) ? 2 * Wo : Wo;
    const int Cs = (p.store == RCN_STORE_PS2 || p.store == RCN_STORE_PS2_NCHW) ? p.Cout / 4 : p.Cout;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
        const long long m = m0 + r;
        if (m >= M) continue;
        const RowInfo ri = rows[r];
        const int wo = (int)(m % Wo);
        const int ho = (int)((m / Wo) % Ho);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int col;
            if constexpr (TN == 8) col = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4));
            else col = tx * TN + j;
            const int c = n0 + col;
            if (c >= p.Cout) continue;
            float v = acc[i][j];
            if (p.bias) v += p.bias[c];
            if (p.cscale) v = v * (1.f + p.cscale[ri.n * p.Cout + c]) + p.cshift[ri.n * p.Cout + c];
            if (p.epi != RCN_EPI_NONE) {
                const float a = p.aux_nchw ? p.aux[(((long long)ri.n * p.Cout + c) * Ho + ho) * Wo + wo] : p.aux[m * p.ldaux + c];
                switch (p.epi) {
                    case RCN_EPI_GDN: v = a * rsqrtf(v); break;
                    case RCN_EPI_IGDN: v = a * sqrtf(v); break;
                    case RCN_EPI_MUL_AUXP1: v = v * (a + 1.f); break;
                    case RCN_EPI_MULP1_AUX: v = (v + 1.f) * a; break;
                    case RCN_EPI_SIGMOID_GATE: v = a * (1.f / (1.f + expf(-v))); break;
                    default: break;
                }
            }
            int hh = ho, ww = wo, cc = c;
            if (p.store == RCN_STORE_PS2 || p.store == RCN_STORE_PS2_NCHW) {
                cc = c >> 2;
                hh = 2 * ho + ((c >> 1) & 1);
                ww = 2 * wo + (c & 1);
            }
            const long long pix = ((long long)ri.n * Hs + hh) * Ws + ww;
            float rv = 0.f;
            if (p.res) rv = p.res_scale * p.res[pix * p.ldres + cc];
            if (p.res && p.res_pre) v += rv;
            v = act_apply(v, p.act, p.slope);
            if (p.res && !p.res_pre) v += rv;
            if (p.store == RCN_STORE_NCHW || p.store == RCN_STORE_PS2_NCHW)
                p.y[(((long long)ri.n * Cs + cc) * Hs + hh) * Ws + ww] = v;
            else
                p.y[pix * p.ldy + cc] = v;
        }
    }
}

__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int k, float* __restrict__ out) {
    // out[((ky*k+kx)*Cin + ci)*Cout + co] = w[((co*Cin + ci)*k + ky)*k + kx]
    const long long total = (long long)Cout * Cin * k * k;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        long long t = i / Cout;
        const int ci = (int)(t % Cin);
        const int tap = (int)(t / Cin);
        out[i] = w[((long long)co * Cin + ci) * k * k + tap];
    }
}

}  // namespace

}  // namespace rcn

extern "C" int rcn_conv2d(const rcn_conv_desc* d, void* stream) {
    using namespace rcn;
    RCN_CHECK_ARG(d && d->x && d->w && d->y, "rcn_conv2d: null pointer");
    RCN_CHECK_ARG(d->k == 1 || d->k == 3, "rcn_conv2d: kernel size %d unsupported (1 or 3)", d->k);
    RCN_CHECK_ARG(d->stride == 1 || d->stride == 2, "rcn_conv2d: stride %d unsupported", d->stride);
    RCN_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "rcn_conv2d: bad shape");
    RCN_CHECK_ARG(d->ldx >= d->Cin, "rcn_conv2d: ldx < Cin");
    RCN_CHECK_ARG(d->epi == RCN_EPI_NONE || d->aux, "rcn_conv2d: epilogue needs aux");
    RCN_CHECK_ARG((d->cscale == nullptr) == (d->cshift == nullptr), "rcn_conv2d: cscale/cshift must come together");
    const bool ps = d->store == RCN_STORE_PS2 || d->store == RCN_STORE_PS2_NCHW;
    RCN_CHECK_ARG(!ps || (d->Cout % 4 == 0), "rcn_conv2d: pixel shuffle needs Cout %% 4 == 0");
    const int pad = d->k / 2;
    const int Ho = (d->H + 2 * pad - d->k) / d->stride + 1;
    const int Wo = (d->W + 2 * pad - d->k) / d->stride + 1;
    const long long M = (long long)d->N * Ho * Wo;
    const int K = d->k * d->k * d->Cin;
    const int vecA = (d->Cin % 16 == 0) && (d->ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(d->x) & 15) == 0);
    RCN_CHECK_ARG((d->Cout % 4 != 0) || ((reinterpret_cast<uintptr_t>(d->w) & 15) == 0), "rcn_conv2d: weights must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(256);
    const long long gx = (M + 127) / 128;
    RCN_CHECK_ARG(gx < 2147483647LL, "rcn_conv2d: too many output pixels");
    if (d->Cout > 64) {
        dim3 grid((unsigned)gx, (d->Cout + 127) / 128);
        conv2d_kernel<128><<<grid, block, 0, s>>>(*d, Ho, Wo, M, K, vecA);
    } else if (d->Cout > 32) {
        dim3 grid((unsigned)gx, 1);
        conv2d_kernel<64><<<grid, block, 0, s>>>(*d, Ho, Wo, M, K, vecA);
    } else if (d->Cout > 16) {
        dim3 grid((unsigned)gx, 1);
        conv2d_kernel<32><<<grid, block, 0, s>>>(*d, Ho, Wo, M, K, vecA);
    } else {
        dim3 grid((unsigned)gx, 1);
        conv2d_kernel<16><<<grid, block, 0, s>>>(*d, Ho, Wo, M, K, vecA);
    }
    count_launch();
    RCN_CHECK_LAUNCH("rcn_conv2d");
    return RCN_OK;
}

extern "C" int rcn_pack_conv_weight(const float* w, int Cout, int Cin, int k, float* out, void* stream) {
    using namespace rcn;
    RCN_CHECK_ARG(w && out && Cout > 0 && Cin > 0 && k > 0, "rcn_pack_conv_weight: bad arguments");
    const long long total = (long long)Cout * Cin * k * k;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, k, out);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_pack_conv_weight");
    return RCN_OK;
}
