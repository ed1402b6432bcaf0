// GroupMix "efficient attention" reductions (models/groupmix.py:186-189):
//   k_softmax = softmax over ALL N = H*W tokens (per batch, head, channel)
//   kv        = k_softmax^T v              (B, heads, Ch, Ch)
//   eff       = q kv                        (B, heads, N, Ch)
// Two passes over the token map: pass 1 (stats + kv) reads k and v once, pass 2 reads q (+crpe)
// and writes the result -- 3*C*4 B/token of compulsory traffic.  Per-lane serial accumulation over
// pixels (NHWC keeps lanes on consecutive channels => coalesced), cross-chunk merges of the
// online-softmax pairs (m, s) and of the kv partials use warp-shuffle reductions; fixed chunking
// keeps the result deterministic.
#include "common.cuh"

namespace rcn {
namespace {

// ---- pass 1a: per-(b, channel) online-softmax partials over a chunk of pixels
// grid (chunks, B), block (CX, RY): lanes along channels
__global__ void gm_kstats_partial(const float* __restrict__ k, int ldk, long long HW, int C, int chunks,
                                  float* __restrict__ pm, float* __restrict__ ps) {
    extern __shared__ float red[];  // [2][RY][CX]
    const int b = blockIdx.y, chunk = blockIdx.x;
    const long long per = (HW + chunks - 1) / chunks;
    const long long p0 = chunk * per, p1 = (p0 + per < HW) ? p0 + per : HW;
    const float* kb = k + (long long)b * HW * ldk;
    float* rm = red;
    float* rs = red + blockDim.x * blockDim.y;
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        float m = -INFINITY, s = 0.f;
        if (c < C)
            for (long long p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
                const float v = kb[p * ldk + c];
                if (v > m) { s = s * expf(m - v) + 1.f; m = v; }
                else s += expf(v - m);
            }
        rm[threadIdx.y * blockDim.x + threadIdx.x] = m;
        rs[threadIdx.y * blockDim.x + threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.y == 0 && c < C) {
            float M = -INFINITY;
            for (int r = 0; r < blockDim.y; ++r) M = fmaxf(M, rm[r * blockDim.x + threadIdx.x]);
            float S = 0.f;
            for (int r = 0; r < blockDim.y; ++r) {
                const float mr = rm[r * blockDim.x + threadIdx.x];
                if (mr > -INFINITY) S += rs[r * blockDim.x + threadIdx.x] * expf(mr - M);
            }
            pm[((long long)b * chunks + chunk) * C + c] = M;
            ps[((long long)b * chunks + chunk) * C + c] = S;
        }
        __syncthreads();
    }
}

// ---- pass 1b: merge chunk partials: one warp per (b, channel), shuffle reduction over chunks
__global__ void gm_kstats_final(const float* __restrict__ pm, const float* __restrict__ ps, int chunks, int C, int total,
                                float* __restrict__ kmax, float* __restrict__ ksum) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= total) return;
    const int b = gw / C, c = gw - b * C;
    float m = -INFINITY;
    for (int i = lane; i < chunks; i += 32) m = fmaxf(m, pm[((long long)b * chunks + i) * C + c]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int i = lane; i < chunks; i += 32) {
        const float mi = pm[((long long)b * chunks + i) * C + c];
        if (mi > -INFINITY) s += ps[((long long)b * chunks + i) * C + c] * expf(mi - m);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { kmax[gw] = m; ksum[gw] = s; }
}

// ---- pass 1c: kv partials.  grid (chunks, B); thread t owns entry (h, kc, vc) of the kv matrix;
// pixels are staged through shared memory in tiles of TP tokens.
template <int TP>
__global__ void gm_kv_partial(const float* __restrict__ k, int ldk, const float* __restrict__ v, int ldv, long long HW,
                              int heads, int Ch, int chunks, const float* __restrict__ kmax, float* __restrict__ part) {
    extern __shared__ float sm[];  // [TP][Ct] k-exp, [TP][Ct] v
    const int Ct = heads * Ch;
    float* sk = sm;
    float* sv = sm + TP * Ct;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const long long per = (HW + chunks - 1) / chunks;
    const long long p0 = chunk * per, p1 = (p0 + per < HW) ? p0 + per : HW;
    const float* kb = k + (long long)b * HW * ldk;
    const float* vb = v + (long long)b * HW * ldv;
    const int nent = heads * Ch * Ch;
    // each thread may own several entries (Ch=20: 3200 entries)
    constexpr int MAXE = 8;
    float acc[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) acc[e] = 0.f;
    for (long long t0 = p0; t0 < p1; t0 += TP) {
        const int np = (int)((p1 - t0 < TP) ? (p1 - t0) : TP);
        for (int i = threadIdx.x; i < np * Ct; i += blockDim.x) {
            const int pp = i / Ct, c = i - pp * Ct;
            sk[i] = expf(kb[(t0 + pp) * ldk + c] - kmax[b * Ct + c]);
            sv[i] = vb[(t0 + pp) * ldv + c];
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int ent = threadIdx.x + e * blockDim.x;
            if (ent < nent) {
                const int h = ent / (Ch * Ch), r = ent - h * Ch * Ch;
                const int kc = r / Ch, vc = r - kc * Ch;
                float a = acc[e];
                for (int pp = 0; pp < np; ++pp) a = fmaf(sk[pp * Ct + h * Ch + kc], sv[pp * Ct + h * Ch + vc], a);
                acc[e] = a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
        const int ent = threadIdx.x + e * blockDim.x;
        if (ent < nent) part[((long long)b * chunks + chunk) * nent + ent] = acc[e];
    }
}

// ---- pass 1d: reduce kv partials over chunks (warp per entry, shuffle reduction) and divide by ksum
__global__ void gm_kv_final(const float* __restrict__ part, int chunks, int heads, int Ch, int total,
                            const float* __restrict__ ksum, float* __restrict__ kv) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= total) return;
    const int nent = heads * Ch * Ch;
    const int b = gw / nent, ent = gw - b * nent;
    float s = 0.f;
    for (int i = lane; i < chunks; i += 32) s += part[((long long)b * chunks + i) * nent + ent];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const int h = ent / (Ch * Ch), kc = (ent - h * Ch * Ch) / Ch;
        kv[gw] = s / ksum[b * heads * Ch + h * Ch + kc];
    }
}

// ---- pass 2: out[p, h*Ch+vc] = scale * sum_kc q[p,h,kc] kv[b,h,kc,vc] + crpe[p, h*Ch+vc]
__global__ void gm_apply(const float* __restrict__ q, int ldq, const float* __restrict__ kv, const float* __restrict__ crpe,
                         int ldc, long long HW, int heads, int Ch, float scale, long long total, float* __restrict__ out, int ldo) {
    extern __shared__ float skv[];  // [heads][Ch][Ch] of this block's batch item
    const int Ct = heads * Ch;
    const int b = blockIdx.y;
    const int nent = heads * Ch * Ch;
    for (int i = threadIdx.x; i < nent; i += blockDim.x) skv[i] = kv[(long long)b * nent + i];
    __syncthreads();
    const long long per_b = HW * Ct;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_b; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Ct);
        const long long p = (long long)b * HW + i / Ct;
        const int h = c / Ch, vc = c - h * Ch;
        const float* qp = q + p * ldq + h * Ch;
        float a = 0.f;
        for (int kc = 0; kc < Ch; ++kc) a = fmaf(qp[kc], skv[(h * Ch + kc) * Ch + vc], a);
        float r = scale * a;
        if (crpe) r += crpe[p * ldc + c];
        out[p * ldo + c] = r;
    }
}

}  // namespace
}  // namespace rcn

using namespace rcn;

extern "C" long long rcn_groupmix_workspace_floats(int B, long long HW, int heads, int Ch) {
    long long chunks = HW / 256;
    if (chunks < 1) chunks = 1;
    if (chunks > 512) chunks = 512;
    const long long Ct = (long long)heads * Ch;
    return 2 * B * chunks * Ct + 2 * B * Ct + (long long)B * chunks * heads * Ch * Ch + 64;
}

// q, k, v: NHWC channel slices with Ct = heads*Ch channels (head-major).  kv_out: [B][heads][Ch][Ch].
extern "C" int rcn_groupmix_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                      const float* crpe, int ldc, int B, long long HW, int heads, int Ch, float scale,
                                      float* out, int ldo, float* kv_out, float* workspace, long long workspace_floats,
                                      void* stream) {
    RCN_CHECK_ARG(q && k && v && out && kv_out && workspace, "rcn_groupmix_attention: null pointer");
    RCN_CHECK_ARG(heads > 0 && Ch > 0 && heads * Ch * Ch <= 8 * 512, "rcn_groupmix_attention: heads*Ch*Ch=%d too large", heads * Ch * Ch);
    RCN_CHECK_ARG(workspace_floats >= rcn_groupmix_workspace_floats(B, HW, heads, Ch), "rcn_groupmix_attention: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const int Ct = heads * Ch;
    long long chunks = HW / 256;
    if (chunks < 1) chunks = 1;
    if (chunks > 512) chunks = 512;
    float* pm = workspace;
    float* ps = pm + (long long)B * chunks * Ct;
    float* kmax = ps + (long long)B * chunks * Ct;
    float* ksum = kmax + (long long)B * Ct;
    float* part = ksum + (long long)B * Ct;
    int cx = ((Ct + 31) / 32) * 32;
    if (cx > 256) cx = 256;
    const int ry = 256 / cx > 0 ? 256 / cx : 1;
    gm_kstats_partial<<<dim3((unsigned)chunks, B), dim3(cx, ry), 2 * cx * ry * sizeof(float), s>>>(k, ldk, HW, Ct, (int)chunks, pm, ps);
    gm_kstats_final<<<cdiv((long long)B * Ct * 32, 256), 256, 0, s>>>(pm, ps, (int)chunks, Ct, B * Ct, kmax, ksum);
    constexpr int TP = 32;
    const size_t smem = 2 * TP * Ct * sizeof(float);
    gm_kv_partial<TP><<<dim3((unsigned)chunks, B), 512, smem, s>>>(k, ldk, v, ldv, HW, heads, Ch, (int)chunks, kmax, part);
    const int nent = heads * Ch * Ch;
    gm_kv_final<<<cdiv((long long)B * nent * 32, 256), 256, 0, s>>>(part, (int)chunks, heads, Ch, B * nent, ksum, kv_out);
    long long blocks = (HW * Ct + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gm_apply<<<dim3((unsigned)blocks, B), 256, nent * sizeof(float), s>>>(q, ldq, kv_out, crpe, ldc, HW, heads, Ch, scale,
                                                                        HW * Ct, out, ldo);
    count_launch(5);
    RCN_CHECK_LAUNCH("rcn_groupmix_attention");
    return RCN_OK;
}
