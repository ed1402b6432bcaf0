// LayerNorm over channels and Swin window attention (W-MSA / SW-MSA), NHWC fp32.
#include <cuda_bf16.h>

#include <stdlib.h>

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

// Any C % 4 == 0 up to 128 * NV (GroupMix: C = 80, 200, 16): one warp per pixel and step, NV float4 per lane (lanes past C / 4
// idle), PIX_IT pixels in flight per warp.  The scalar kernel above issues one 4-byte load per lane and channel.
template <int NV, int PIX_IT>
__global__ void layernorm_vecg_kernel(const float* __restrict__ x, long long npix, int C, int ldx, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float eps, float* __restrict__ y, int ldy, int act,
                                      __nv_bfloat16* __restrict__ y_hi, __nv_bfloat16* __restrict__ y_lo, int ldp) {
    const int lane = threadIdx.x & 31;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long base = wg * PIX_IT;
    if (base >= npix) return;
    float4 v[PIX_IT][NV], g[NV], b[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int c = 4 * lane + 128 * k;
        g[k] = c < C ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        b[k] = c < C ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < PIX_IT; ++it) {
        const long long pix = base + it;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = 4 * lane + 128 * k;
            v[it][k] = (pix < npix && c < C) ? *reinterpret_cast<const float4*>(x + pix * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int it = 0; it < PIX_IT; ++it) {
        const long long pix = base + it;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) s += (v[it][k].x + v[it][k].y) + (v[it][k].z + v[it][k].w);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)C;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (4 * lane + 128 * k < C) {
                const float dx = v[it][k].x - mean, dy = v[it][k].y - mean, dz = v[it][k].z - mean, dw = v[it][k].w - mean;
                q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / (float)C + eps);
        if (pix >= npix) continue;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = 4 * lane + 128 * k;
            if (c >= C) continue;
            float r[4] = {act_apply((v[it][k].x - mean) * rstd * g[k].x + b[k].x, act, 0.f), act_apply((v[it][k].y - mean) * rstd * g[k].y + b[k].y, act, 0.f),
                          act_apply((v[it][k].z - mean) * rstd * g[k].z + b[k].z, act, 0.f), act_apply((v[it][k].w - mean) * rstd * g[k].w + b[k].w, act, 0.f)};
            if (y) *reinterpret_cast<float4*>(y + pix * ldy + c) = make_float4(r[0], r[1], r[2], r[3]);
            if (y_hi) {
                __nv_bfloat16 hb[4], lb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    hb[e] = __float2bfloat16_rn(r[e]);
                    lb[e] = __float2bfloat16_rn(r[e] - __bfloat162float(hb[e]));
                }
                *reinterpret_cast<uint2*>(y_hi + pix * ldp + c) = *reinterpret_cast<uint2*>(hb);
                if (y_lo) *reinterpret_cast<uint2*>(y_lo + pix * ldp + c) = *reinterpret_cast<uint2*>(lb);
            }
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

// ---------------------------------------------------------------- window attention on tensor cores (8x8 windows)
// One warp per (window, head, 16-query m-tile); S = Q K^T and O = P V run as mma.sync m16n8k8 TF32 tiles with 3xTF32 split
// operands (x = hi + lo, hi*hi + lo*hi + hi*lo accumulated in fp32: ~2^-21 relative, the parity bar of the Swin blocks is 1e-4).
// The FFMA kernels above need 2*64*HD FMAs + ~6 element-wise instructions per (query, head, key); here the contractions are
// 48*HD/8 warp-level MMAs per 16 queries and only the softmax stays element-wise.
//   fragments (PTX m16n8k8 .tf32, g = lane / 4, t = lane % 4):
//     A (16x8, row): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);   B (8x8, col): b0 (k=t, n=g) b1 (k=t+4, n=g)
//     C (16x8):      c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
//   The probabilities never leave registers: for P V the k index of A is re-labelled (k=t <-> key 2t, k=t+4 <-> key 2t+1 inside
//   each 8-key block), which makes the C fragment of S exactly the A fragment of P V (a0=c0, a1=c2, a2=c1, a3=c3); the B
//   fragment of V uses the same labelling (rows 2t and 2t+1 of the block).
//   Shared memory per head: Q (pre-scaled by hd^-1/2 * log2 e) and K, V pre-split into hi / lo, rows padded to HD + 4 floats
//   (conflict-free fragment loads: (HD+4) g + t and 2 (HD+4) t + g hit 32 distinct banks for HD = 8, 16, 32).
// hi part of the 3xTF32 split by TRUNCATION (one LOP3 on the ALU pipe; cvt.rna.tf32 runs on the 16-lane XU pipe, which the
// exponentials already load): x = hi + lo exactly, |lo| < 2^-10 |x|, and the tensor core's own truncation of lo costs 2^-20 |x|
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// Persistent blocks: a block walks windows w = blockIdx.x, += gridDim.x for its HPB heads; the q / k / v rows of the NEXT window
// stream into the second shared-memory buffer with cp.async while the current one is computed (the first version of this
// kernel staged, synchronised and computed one window per block: 16 K cycles per block for ~2 K cycles of issue, i.e. bound by
// the exposed global-load latency at 32 warps per SM).  The hi / lo split happens at fragment-load time.
template <int HD, int HPB>
__global__ void __launch_bounds__(HPB * 128) wmsa_mma_kernel(const float* __restrict__ qkv, int H, int W, int C, int ldq, int shifted,
                                                             const float* __restrict__ relpos, float* __restrict__ out, int ldo, int nheads,
                                                             __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int ldp,
                                                             int nwin) {
    constexpr int WS = 8, P = 64, LD = HD + 4, RP = (2 * WS - 1) * (2 * WS - 1);
    constexpr float LOG2E = 1.4426950408889634f;
    extern __shared__ __align__(16) float sm[];
    // [2 buffers][HPB heads][q | k | v][P][LD] fp32, then [HPB][RP] bias
    constexpr int HEAD_FLOATS = 3 * P * LD, BUF_FLOATS = HPB * HEAD_FLOATS;
    float* bias_all = sm + 2 * BUF_FLOATS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hl = warp >> 2, mt = warp & 3;            // head within the block, 16-query tile
    const int head0 = blockIdx.y * HPB;
    const int nww = W / WS, nwh = H / WS;
    const int sh = shifted ? WS / 2 : 0;
    const float qscale = rsqrtf((float)HD) * LOG2E;
    constexpr int V4_PER_TOK = 3 * HPB * HD / 4;

    // thread -> (token, 16-byte chunk of the block's HPB*HD contiguous floats), the same for q, k and v: three cp.async per token
    constexpr int F4 = HPB * HD / 4, TOK_STEP = HPB * 128 / F4;
    const int pf_c = tid % F4, pf_t0 = tid / F4;
    const int pf_h = (4 * pf_c) / HD, pf_d = (4 * pf_c) % HD;
    const bool pf_live = head0 + pf_h < nheads;
    auto prefetch = [&](int win, int buf) {
        const int n = win / (nwh * nww);
        const int wrem = win - n * nwh * nww;
        const int wy = wrem / nww, wx = wrem - wy * nww;
        float* dstb = sm + buf * BUF_FLOATS + pf_h * HEAD_FLOATS + pf_d;
        if (pf_live) {
#pragma unroll
            for (int tok = pf_t0; tok < P; tok += TOK_STEP) {
                const int py = tok >> 3, px = tok & 7;
                int gy = wy * WS + py + sh, gx = wx * WS + px + sh;
                if (gy >= H) gy -= H;
                if (gx >= W) gx -= W;
                const float* src = qkv + ((long long)(n * H + gy) * W + gx) * ldq + head0 * HD + 4 * pf_c;
                float* dst = dstb + tok * LD;
                cp_async16(dst, src);
                cp_async16(dst + P * LD, src + C);
                cp_async16(dst + 2 * P * LD, src + 2 * C);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (int i = tid; i < HPB * RP; i += HPB * 128) {
        const int h = i / RP, r = i - h * RP;
        bias_all[i] = (head0 + h < nheads) ? relpos[(size_t)(head0 + h) * RP + r] * LOG2E : 0.f;
    }
    const int head = head0 + hl;
    const bool live = head < nheads;
    const float* bias = bias_all + hl * RP;
    const int g = lane >> 2, t = lane & 3;

    int buf = 0;
    if ((int)blockIdx.x < nwin) prefetch(blockIdx.x, 0);
    for (int win = blockIdx.x; win < nwin; win += gridDim.x, buf ^= 1) {
        const int next = win + gridDim.x;
        if (next < nwin) {
            prefetch(next, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (live) {
            const int n = win / (nwh * nww);
            const int wrem = win - n * nwh * nww;
            const int wy = wrem / nww, wx = wrem - wy * nww;
            const float* Qs = sm + buf * BUF_FLOATS + hl * HEAD_FLOATS;
            const float* Ks = Qs + P * LD;
            const float* Vs = Ks + P * LD;
            // ---- Q fragments of this warp's 16 queries (rows 16 mt + g and + 8), scaled and split hi / lo
            uint32_t qh[HD / 8][4], ql[HD / 8][4];
#pragma unroll
            for (int ks = 0; ks < HD / 8; ++ks) {
                const float f[4] = {Qs[(16 * mt + g) * LD + 8 * ks + t], Qs[(16 * mt + g + 8) * LD + 8 * ks + t],
                                    Qs[(16 * mt + g) * LD + 8 * ks + t + 4], Qs[(16 * mt + g + 8) * LD + 8 * ks + t + 4]};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float fs = f[j] * qscale;
                    qh[ks][j] = tf32_hi(fs);
                    ql[ks][j] = __float_as_uint(fs - __uint_as_float(qh[ks][j]));
                }
            }
            // ---- S = bias (+ mask) + Q K^T: query rows i0 = 16 mt + g -> (py, px) = (2 mt, g) and i1 = i0 + 8 -> (2 mt + 1, g);
            //      key 8 nt + 2 t (+1) -> (jy, jx) = (nt, 2 t (+1))
            const bool edge_y = shifted && (wy == nwh - 1), edge_x = shifted && (wx == nww - 1);
            const bool cq = edge_x && (g >= WS - sh);
            const bool rq0 = edge_y && (2 * mt >= WS - sh), rq1 = edge_y && (2 * mt + 1 >= WS - sh);
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int b0i = (2 * mt - nt + WS - 1) * (2 * WS - 1) + (g - 2 * t + WS - 1);    // (i0, key 2t); key 2t+1 is one entry to the left
                s[nt][0] = bias[b0i];
                s[nt][1] = bias[b0i - 1];
                s[nt][2] = bias[b0i + (2 * WS - 1)];
                s[nt][3] = bias[b0i + (2 * WS - 1) - 1];
            }
            if (edge_y || edge_x) {   // block-uniform: only the last window row / column of a shifted map pays for the mask
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const bool rk = edge_y && (nt >= WS - sh);
                    const bool ck0 = edge_x && (2 * t >= WS - sh), ck1 = edge_x && (2 * t + 1 >= WS - sh);
                    if (rk != rq0 || ck0 != cq) s[nt][0] = -INFINITY;
                    if (rk != rq0 || ck1 != cq) s[nt][1] = -INFINITY;
                    if (rk != rq1 || ck0 != cq) s[nt][2] = -INFINITY;
                    if (rk != rq1 || ck1 != cq) s[nt][3] = -INFINITY;
                }
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int ks = 0; ks < HD / 8; ++ks) {
                    const int o0 = (8 * nt + g) * LD + 8 * ks + t;
                    const float k0 = Ks[o0], k1 = Ks[o0 + 4];
                    const uint32_t kh0 = tf32_hi(k0), kh1 = tf32_hi(k1);
                    const uint32_t kl0 = __float_as_uint(k0 - __uint_as_float(kh0)), kl1 = __float_as_uint(k1 - __uint_as_float(kh1));
                    mma_tf32(s[nt], ql[ks], kh0, kh1);
                    mma_tf32(s[nt], qh[ks], kl0, kl1);
                    mma_tf32(s[nt], qh[ks], kh0, kh1);
                }
            }
            // ---- softmax over the 64 keys of rows i0 (elements 0, 1) and i1 (elements 2, 3): quad-wide reductions
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
                m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = ex2_approx(s[nt][0] - m0); s[nt][1] = ex2_approx(s[nt][1] - m0);
                s[nt][2] = ex2_approx(s[nt][2] - m1); s[nt][3] = ex2_approx(s[nt][3] - m1);
                d0 += s[nt][0] + s[nt][1];
                d1 += s[nt][2] + s[nt][3];
            }
            d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
            d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
            // ---- O = P V (un-normalised), keys in blocks of 8 with the re-labelled k index
            float o[HD / 8][4];
#pragma unroll
            for (int dt = 0; dt < HD / 8; ++dt) { o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const float pf[4] = {s[kk][0], s[kk][2], s[kk][1], s[kk][3]};     // a0 = c0, a1 = c2, a2 = c1, a3 = c3
                uint32_t ph[4], pl[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    ph[j] = tf32_hi(pf[j]);
                    pl[j] = __float_as_uint(pf[j] - __uint_as_float(ph[j]));
                }
#pragma unroll
                for (int dt = 0; dt < HD / 8; ++dt) {
                    const int o0 = (8 * kk + 2 * t) * LD + 8 * dt + g;
                    const float v0 = Vs[o0], v1 = Vs[o0 + LD];
                    const uint32_t vh0 = tf32_hi(v0), vh1 = tf32_hi(v1);
                    const uint32_t vl0 = __float_as_uint(v0 - __uint_as_float(vh0)), vl1 = __float_as_uint(v1 - __uint_as_float(vh1));
                    mma_tf32(o[dt], pl, vh0, vh1);
                    mma_tf32(o[dt], ph, vl0, vl1);
                    mma_tf32(o[dt], ph, vh0, vh1);
                }
            }
            const float inv0 = 1.f / d0, inv1 = 1.f / d1;
            // ---- store: rows i0, i1 -> pixels of the (shifted) window; each lane holds channels head*HD + 8 dt + 2 t, +1
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int py = 2 * mt + r, px = g;
                int gy = wy * WS + py + sh, gx = wx * WS + px + sh;
                if (gy >= H) gy -= H;
                if (gx >= W) gx -= W;
                const long long opix = (long long)(n * H + gy) * W + gx;
                const float inv = r ? inv1 : inv0;
#pragma unroll
                for (int dt = 0; dt < HD / 8; ++dt) {
                    const float v0 = o[dt][2 * r] * inv, v1 = o[dt][2 * r + 1] * inv;
                    const int ch = head * HD + 8 * dt + 2 * t;
                    if (out) *reinterpret_cast<float2*>(out + opix * ldo + ch) = make_float2(v0, v1);
                    if (out_hi) {
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                        __nv_bfloat162 hp; hp.x = h0; hp.y = h1;
                        *reinterpret_cast<__nv_bfloat162*>(out_hi + opix * ldp + ch) = hp;
                        if (out_lo) {
                            __nv_bfloat162 lp;
                            lp.x = __float2bfloat16_rn(v0 - __bfloat162float(h0));
                            lp.y = __float2bfloat16_rn(v1 - __bfloat162float(h1));
                            *reinterpret_cast<__nv_bfloat162*>(out_lo + opix * ldp + ch) = lp;
                        }
                    }
                }
            }
        }
        __syncthreads();   // every warp is done with this buffer before the next iteration's prefetch overwrites it
    }
}

template <int HD>
int launch_wmsa_mma(const float* qkv, int N, int H, int W, int C, int ldq, int shifted, const float* relpos, float* out, int ldo,
                    __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int ldp, cudaStream_t s) {
    constexpr int HPB = (HD == 32) ? 1 : 2;           // heads per block (4 warps each)
    constexpr int P = 64, LD = HD + 4, RP = 225;
    const int nheads = C / HD;
    const int nwin = N * (H / 8) * (W / 8);
    const int groups = (nheads + HPB - 1) / HPB;
    const size_t smem = ((size_t)2 * HPB * 3 * P * LD + (size_t)HPB * RP) * sizeof(float);
    auto kern = wmsa_mma_kernel<HD, HPB>;
    static bool attr[64] = {};
    static int resident[64] = {};      // blocks of this instantiation that fit on the device at once
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr[dev]) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, HPB * 128, smem);
        resident[dev] = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);
        attr[dev] = true;
    }
    // persistent: exactly one wave of resident blocks, shared by the head groups
    int gx = resident[dev] / groups;
    if (gx > nwin) gx = nwin;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)groups);
    kern<<<grid, HPB * 128, smem, s>>>(qkv, H, W, C, ldq, shifted, relpos, out, ldo, nheads, out_hi, out_lo, ldp, nwin);
    return 0;
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
    if (al && C % 4 == 0 && C <= 256) {   // (C = 64 / 128 took the fixed-width kernel above)
        constexpr int PIX_IT = 4;
        const int g3 = cdiv(npix, (long long)wpb * PIX_IT);
        if (C <= 128) layernorm_vecg_kernel<1, PIX_IT><<<g3, wpb * 32, 0, s>>>(x, npix, C, ldx, gamma, beta, eps, y, ldy, act, y_hi, y_lo, ldp);
        else layernorm_vecg_kernel<2, PIX_IT><<<g3, wpb * 32, 0, s>>>(x, npix, C, ldx, gamma, beta, eps, y, ldy, act, y_hi, y_lo, ldp);
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
    static const bool use_mma = !(getenv("RCN_WMSA_FFMA") && atoi(getenv("RCN_WMSA_FFMA")));   // triage: 1 = CUDA-core kernel for 8x8 windows too
    const bool mma_ok = use_mma && ws == 8 && (!out_hi || ((uintptr_t)out_hi % 4 == 0 && (!out_lo || (uintptr_t)out_lo % 4 == 0) && ldp % 2 == 0)) &&
                        (!out || ((uintptr_t)out % 8 == 0 && ldo % 2 == 0));
    if (mma_ok && head_dim == 8) rc = launch_wmsa_mma<8>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
    else if (mma_ok && head_dim == 16) rc = launch_wmsa_mma<16>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
    else if (mma_ok && head_dim == 32) rc = launch_wmsa_mma<32>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
    else if (ws == 8 && head_dim == 8) rc = launch_wmsa<8, 8>(qkv, N, H, W, C, ldq, shifted, relpos, out, ldo, out_hi, out_lo, ldp, s);
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
