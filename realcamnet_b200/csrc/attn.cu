// LayerNorm over channels and Swin window attention (W-MSA / SW-MSA), NHWC fp32.
#include <cuda_bf16.h>

#include "common.cuh"

namespace rcn {
namespace {

// one warp per pixel; C <= 32*NIT
template <int NIT>
__global__ void layernorm_kernel(const float* __restrict__ x, long long npix, int C, int ldx,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 float* __restrict__ y, int ldy, int act, __nv_bfloat16* __restrict__ y_hi,
                                 __nv_bfloat16* __restrict__ y_lo, int ldp) {
    const int lane = threadIdx.x & 31;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= npix) return;
    const float* xr = x + pix * ldx;
    float v[NIT];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        const int c = lane + 32 * i;
        v[i] = (c < C) ? xr[c] : 0.f;
        s += v[i];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        const float d = (lane + 32 * i < C) ? v[i] - mean : 0.f;
        q += d * d;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        const int c = lane + 32 * i;
        if (c < C) {
            const float r = act_apply((v[i] - mean) * rstd * gamma[c] + beta[c], act, 0.f);
            if (y) y[pix * ldy + c] = r;
            if (y_hi) {   // the consumer's tcgen05 operand planes (same rounding as rcn_split_bf16)
                const __nv_bfloat16 h = __float2bfloat16_rn(r);
                y_hi[pix * ldp + c] = h;
                if (y_lo) y_lo[pix * ldp + c] = __float2bfloat16_rn(r - __bfloat162float(h));
            }
        }
    }
}

// C = 4*LPP channels (64 or 128): LPP lanes own one pixel with a float4 each, a warp covers 32/LPP pixels per step and PIX_IT
// steps whose loads are all issued before the first reduction (the one-pixel-per-warp kernel above has 256 B in flight per warp
// and is latency bound: 0.19 ms for a 1M x 64 map against 0.08 ms of HBM time).
template <int LPP, int PIX_IT>
__global__ void layernorm_vec_kernel(const float* __restrict__ x, long long npix, int ldx, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, float* __restrict__ y, int ldy, int act,
                                     __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, int ldp) {
    constexpr int PPW = 32 / LPP, C = 4 * LPP;
    const int lane = threadIdx.x & 31, sub = lane / LPP, l = lane % LPP;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long base = wg * (PPW * PIX_IT);
    if (base >= npix) return;
    float4 v[PIX_IT];
#pragma unroll
    for (int it = 0; it < PIX_IT; ++it) {
        const long long pix = base + it * PPW + sub;
        v[it] = pix < npix ? *reinterpret_cast<const float4*>(x + pix * ldx + 4 * l) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * l), b = *reinterpret_cast<const float4*>(beta + 4 * l);
#pragma unroll
    for (int it = 0; it < PIX_IT; ++it) {
        const long long pix = base + it * PPW + sub;
        float s = (v[it].x + v[it].y) + (v[it].z + v[it].w);
#pragma unroll
        for (int o = LPP / 2; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)C;
        const float dx = v[it].x - mean, dy = v[it].y - mean, dz = v[it].z - mean, dw = v[it].w - mean;
        float q = (dx * dx + dy * dy) + (dz * dz + dw * dw);
#pragma unroll
        for (int o = LPP / 2; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / (float)C + eps);
        if (pix >= npix) continue;
        float r[4] = {act_apply(dx * rstd * g.x + b.x, act, 0.f), act_apply(dy * rstd * g.y + b.y, act, 0.f),
                      act_apply(dz * rstd * g.z + b.z, act, 0.f), act_apply(dw * rstd * g.w + b.w, act, 0.f)};
        if (y) *reinterpret_cast<float4*>(y + pix * ldy + 4 * l) = make_float4(r[0], r[1], r[2], r[3]);
        if (y_hi) {
            __nv_bfloat16 hb[4], lb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                hb[e] = __float2bfloat16_rn(r[e]);
                lb[e] = __float2bfloat16_rn(r[e] - __bfloat162float(hb[e]));
            }
            *reinterpret_cast<uint2*>(y_hi + pix * ldp + 4 * l) = *reinterpret_cast<uint2*>(hb);
            if (y_lo) *reinterpret_cast<uint2*>(y_lo + pix * ldp + 4 * l) = *reinterpret_cast<uint2*>(lb);
        }
    }
}

// Window attention. One thread per (query token, head); blockDim = (P, HPB) with P = WS*WS tokens.
// K and V of the head are staged in shared memory (broadcast reads), scores live in registers.
template <int WS, int HD>
__global__ void wmsa_kernel(const float* __restrict__ qkv, int H, int W, int C, int ldq, int shifted,
                            const float* __restrict__ relpos, float* __restrict__ out, int ldo, int nheads,
                            __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int ldp) {
    constexpr int P = WS * WS;
    extern __shared__ float sm[];  // [HPB][2][P][HD]
    const int t = threadIdx.x;     // token in window
    const int hl = threadIdx.y;    // head within block
    const int head = blockIdx.y * blockDim.y + hl;
    const int nww = W / WS, nwh = H / WS;
    const int win = blockIdx.x;
    const int n = win / (nwh * nww);
    const int wrem = win - n * nwh * nww;
    const int wy = wrem / nww, wx = wrem - wy * nww;
    const int sh = shifted ? WS / 2 : 0;
    const int py = t / WS, px = t - py * WS;
    const int gy = (wy * WS + py + sh) % H, gx = (wx * WS + px + sh) % W;
    const bool active = head < nheads;
    const float* base = qkv + ((long long)(n * H + gy) * W + gx) * ldq;
    float* ks = sm + (size_t)hl * 2 * P * HD;
    float* vs = ks + P * HD;
    float q[HD];
    if (active) {
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 a = *reinterpret_cast<const float4*>(base + head * HD + d);
            const float4 b = *reinterpret_cast<const float4*>(base + C + head * HD + d);
            const float4 c = *reinterpret_cast<const float4*>(base + 2 * C + head * HD + d);
            q[d] = a.x; q[d + 1] = a.y; q[d + 2] = a.z; q[d + 3] = a.w;
            *reinterpret_cast<float4*>(ks + t * HD + d) = b;
            *reinterpret_cast<float4*>(vs + t * HD + d) = c;
        }
    }
    // relative-position bias of this block's heads, pre-multiplied by log2(e) (scores are kept in the log2 domain so that the
    // softmax needs one MUFU.EX2 per score)
    constexpr int RP = (2 * WS - 1) * (2 * WS - 1);
    constexpr float LOG2E = 1.4426950408889634f;
    float* rps = sm + (size_t)blockDim.y * 2 * P * HD;   // [HPB][RP]
    for (int i = threadIdx.y * P + t; i < (int)blockDim.y * RP; i += P * (int)blockDim.y) {
        const int hh = blockIdx.y * blockDim.y + i / RP;
        rps[i] = hh < nheads ? relpos[(size_t)hh * RP + (i % RP)] * LOG2E : 0.f;
    }
    __syncthreads();
    if (!active) return;
    const float qscale = rsqrtf((float)HD) * LOG2E;
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] *= qscale;
    const bool rq = shifted && (wy == nwh - 1) && (py >= WS - sh);
    const bool cq = shifted && (wx == nww - 1) && (px >= WS - sh);
    const bool edge_y = shifted && (wy == nwh - 1), edge_x = shifted && (wx == nww - 1);
    const float* rp = rps + hl * RP + (py + WS - 1) * (2 * WS - 1) + (px + WS - 1);
    // Online softmax over the P keys (log2 domain): running maximum mx, denominator den and the un-normalised output o[] --
    // no score array.  The first kernel kept all P scores in registers (s[64] + q + o: ~130 registers, 16 warps per SM, every
    // dependent FMA chain exposed); this one needs ~2 HD + 20 registers and runs at full occupancy.  A new maximum rescales
    // den and o[] (a handful of times per query); masked keys (shifted windows on the last row / column) are skipped.
    float mx = -INFINITY, den = 0.f;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll 4
    for (int j = 0; j < P; ++j) {
        const int jy = j / WS, jx = j - jy * WS;
        if (edge_y || edge_x) {   // block-uniform
            const bool rk = edge_y && (jy >= WS - sh);
            const bool ck = edge_x && (jx >= WS - sh);
            if (rk != rq || ck != cq) continue;
        }
        float a = rp[-(jy * (2 * WS - 1) + jx)];
        const float4* kr = reinterpret_cast<const float4*>(ks + j * HD);
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 kk = kr[d4];
            a = fmaf(q[4 * d4], kk.x, a); a = fmaf(q[4 * d4 + 1], kk.y, a); a = fmaf(q[4 * d4 + 2], kk.z, a); a = fmaf(q[4 * d4 + 3], kk.w, a);
        }
        float pj;
        if (a > mx) {
            const float sc = exp2f(mx - a);   // mx = -inf on the first key: sc = 0
            den *= sc;
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] *= sc;
            mx = a;
            pj = 1.f;
        } else {
            pj = exp2f(a - mx);
        }
        den += pj;
        const float4* vr = reinterpret_cast<const float4*>(vs + j * HD);
#pragma unroll
        for (int d4 = 0; d4 < HD / 4; ++d4) {
            const float4 vv = vr[d4];
            o[4 * d4] = fmaf(pj, vv.x, o[4 * d4]); o[4 * d4 + 1] = fmaf(pj, vv.y, o[4 * d4 + 1]);
            o[4 * d4 + 2] = fmaf(pj, vv.z, o[4 * d4 + 2]); o[4 * d4 + 3] = fmaf(pj, vv.w, o[4 * d4 + 3]);
        }
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= inv;
    const long long opix = (long long)(n * H + gy) * W + gx;
    if (out) {
        float* op = out + opix * ldo + head * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 4) *reinterpret_cast<float4*>(op + d) = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
    }
    if (out_hi) {   // the projection layer's tcgen05 operand planes
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            __nv_bfloat16 hb[4], lb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                hb[e] = __float2bfloat16_rn(o[d + e]);
                lb[e] = __float2bfloat16_rn(o[d + e] - __bfloat162float(hb[e]));
            }
            const long long po = opix * ldp + head * HD + d;
            *reinterpret_cast<uint2*>(out_hi + po) = *reinterpret_cast<uint2*>(hb);
            if (out_lo) *reinterpret_cast<uint2*>(out_lo + po) = *reinterpret_cast<uint2*>(lb);
        }
    }
}

template <int WS, int HD>
int launch_wmsa(const float* qkv, int N, int H, int W, int C, int ldq, int shifted, const float* relpos,
                float* out, int ldo, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int ldp, cudaStream_t s) {
    constexpr int P = WS * WS;
    const int nheads = C / HD;
    int hpb = 256 / P;
    if (hpb > nheads) hpb = nheads;
    if (hpb < 1) hpb = 1;
    dim3 block(P, hpb);
    dim3 grid((unsigned)(N * (H / WS) * (W / WS)), (nheads + hpb - 1) / hpb);
    const size_t smem = ((size_t)hpb * 2 * P * HD + (size_t)hpb * (2 * WS - 1) * (2 * WS - 1)) * sizeof(float);
    auto kern = wmsa_kernel<WS, HD>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, block, smem, s>>>(qkv, H, W, C, ldq, shifted, relpos, out, ldo, nheads, out_hi, out_lo, ldp);
    return 0;
}

}  // namespace
}  // namespace rcn

extern "C" int rcn_layernorm(const float* x, long long npix, int C, int ldx, const float* gamma, const float* beta,
                             float eps, float* y, int ldy, int act, void* y_hi_, void* y_lo_, int ldp, void* stream) {
    using namespace rcn;
    __nv_bfloat16* y_hi = (__nv_bfloat16*)y_hi_;
    __nv_bfloat16* y_lo = (__nv_bfloat16*)y_lo_;
    RCN_CHECK_ARG(x && (y || y_hi) && gamma && beta && npix > 0, "rcn_layernorm: bad arguments");
    RCN_CHECK_ARG(!y_hi || ldp >= C, "rcn_layernorm: plane pixel stride %d < C", ldp);
    RCN_CHECK_ARG(C > 0 && C <= 1024, "rcn_layernorm: C=%d unsupported (1..1024)", C);
    const int wpb = 8;
    cudaStream_t s = (cudaStream_t)stream;
    const bool al = (ldx % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0) &&
                    (!y || (ldy % 4 == 0 && (uintptr_t)y % 16 == 0)) &&
                    (!y_hi || (ldp % 4 == 0 && (uintptr_t)y_hi % 8 == 0 && (!y_lo || (uintptr_t)y_lo % 8 == 0)));
    // NOT conditioned on npix: the two kernels sum in different orders, and a result that depended on how many pixels a call
    // happens to carry would make a tile's bitstream depend on the batch it was compressed in
    if (al && (C == 64 || C == 128)) {
        constexpr int PIX_IT = 4;
        if (C == 64) {
            const int g2 = cdiv(npix, (long long)wpb * 2 * PIX_IT);
            layernorm_vec_kernel<16, PIX_IT><<<g2, wpb * 32, 0, s>>>(x, npix, ldx, gamma, beta, eps, y, ldy, act, y_hi, y_lo, ldp);
        } else {
            const int g2 = cdiv(npix, (long long)wpb * 1 * PIX_IT);
            layernorm_vec_kernel<32, PIX_IT><<<g2, wpb * 32, 0, s>>>(x, npix, ldx, gamma, beta, eps, y, ldy, act, y_hi, y_lo, ldp);
        }
        count_launch();
        RCN_CHECK_LAUNCH("rcn_layernorm");
        return RCN_OK;
    }
    const int grid = cdiv(npix, wpb);
    if (C <= 32) layernorm_kernel<1><<<grid, wpb * 32, 0, s>>>(x, npix, C, ldx, gamma, beta, eps, y, ldy, act, y_hi, y_lo, ldp);
    else if (C <= 64) layernorm_kernel<2><<<grid, wpb * 32, 0, s>>>(x, npix, C, ldx, gamma, beta, eps, y, ldy, act, y_hi, y_lo, ldp);
    else if (C <= 128) layernorm_kernel<4><<<grid, wpb * 32, 0, s>>>(x, npix, C, ldx, gamma, beta, eps, y, ldy, act, y_hi, y_lo, ldp);
    else if (C <= 256) layernorm_kernel<8><<<grid, wpb * 32, 0, s>>>(x, npix, C, ldx, gamma, beta, eps, y, ldy, act, y_hi, y_lo, ldp);
    else layernorm_kernel<32><<<grid, wpb * 32, 0, s>>>(x, npix, C, ldx, gamma, beta, eps, y, ldy, act, y_hi, y_lo, ldp);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_layernorm");
    return RCN_OK;
}

extern "C" int rcn_wmsa(const float* qkv, int N, int H, int W, int C, int ldq, int head_dim, int ws, int shifted,
                        const float* relpos, float* out, int ldo, void* out_hi_, void* out_lo_, int ldp, void* stream) {
    using namespace rcn;
    __nv_bfloat16* out_hi = (__nv_bfloat16*)out_hi_;
    __nv_bfloat16* out_lo = (__nv_bfloat16*)out_lo_;
    RCN_CHECK_ARG(qkv && relpos && (out || out_hi), "rcn_wmsa: null pointer");
    RCN_CHECK_ARG(!out_hi || (ldp >= C && ldp % 4 == 0 && ((uintptr_t)out_hi & 7) == 0 && ((uintptr_t)out_lo & 7) == 0),
                  "rcn_wmsa: operand planes need a pixel stride >= C (multiple of 4) and 8-byte aligned bases");
    RCN_CHECK_ARG(H % ws == 0 && W % ws == 0, "rcn_wmsa: map %dx%d not divisible by window %d", H, W, ws);
    RCN_CHECK_ARG(C % head_dim == 0 && ldq % 4 == 0 && ldo % 4 == 0, "rcn_wmsa: bad channel layout");
    RCN_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)out & 15) == 0, "rcn_wmsa: pointers must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = -1;
    if (ws == 8 && head_dim == 8) rc = launch_wmsa<8, 8>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
    else if (ws == 8 && head_dim == 16) rc = launch_wmsa<8, 16>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
    else if (ws == 8 && head_dim == 32) rc = launch_wmsa<8, 32>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
    else if (ws == 4 && head_dim == 32) rc = launch_wmsa<4, 32>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
    else if (ws == 4 && head_dim == 16) rc = launch_wmsa<4, 16>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
    else if (ws == 4 && head_dim == 8) rc = launch_wmsa<4, 8>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
    if (rc != 0) {
        set_error("rcn_wmsa: (window %d, head_dim %d) unsupported", ws, head_dim);
        return RCN_ERR_UNSUPPORTED;
    }
    count_launch();
    RCN_CHECK_LAUNCH("rcn_wmsa");
    return RCN_OK;
}
