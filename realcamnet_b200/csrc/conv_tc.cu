// tcgen05 / TMA implicit-GEMM convolution engine (sm_100a).
//
// GEMM view per CTA: D[128 pixels x Ntile channels] += A[128 x 64] * B[Ntile x 64]^T over
// (tap, 64-channel chunk) k-iterations.  A tiles are 8x16-pixel boxes of the NHWC bf16 activation
// planes fetched by TMA (one shifted box per filter tap -- im2col-free; out-of-image pixels are
// zero-filled by the TMA unit = the conv's zero padding); B tiles come from the K-major packed
// weights.  Both land in shared memory in the 128B-swizzled K-major canonical layout that
// tcgen05.mma consumes directly; accumulators live in TMEM (fp32) and are drained with tcgen05.ld
// into the same fused epilogue as the fp32 engine (conv.cu).
//
// Precision modes (passes): 1 = bf16 x bf16 (fp32 accumulate); 3 = "bf16x3": activations and
// weights are split as x = hi + lo (two bf16 planes) and hi*hi + lo*hi + hi*lo is accumulated in
// fp32 -- ~16 mantissa bits, which keeps the 1e-3 parity bar through the ~100-layer path.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  mbarrier ring: full[s] (TMA -> MMA),
// empty[s] (tcgen05.commit -> TMA), tmem_full (last commit -> epilogue).
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"

namespace rcn {
namespace {

constexpr int TILE_H = 8, TILE_W = 16, BLOCK_K = 64;
constexpr int A_BYTES = 128 * BLOCK_K * 2;  // 16 KB
constexpr uint32_t SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SPIN_LIMIT) __trap();
    }
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcParams {
    rcn_conv_desc d;
    int Cp;       // padded input channels (multiple of 64) of the bf16 planes / packed weights
    int Ntile;    // output channels per tile (multiple of 16, <= 128)
    int tiles_x, tiles_y, tiles_n;
    long long total_tiles;
    int passes;   // 1 or 3
    int stages;
    int s2;       // stride-2 conv: A planes are the 4 polyphase components stacked on the batch axis ((py*2+px)*N + n)
    int dbg;      // RCN_TC_DEBUG bit mask (perf triage only): 1 no stores, 2 no MMA, 4 no A loads, 8 no epilogue math
};

constexpr int EPI_WARPS = 16;            // 4 warps per TMEM lane quarter: one warp per scheduler cannot hide any latency
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int TC_THREADS = 64 + EPI_THREADS;
constexpr int STG_COLS = 32;             // accumulator columns staged per epilogue round
constexpr int STG_PITCH = STG_COLS + 4;  // floats; (pitch/4) odd -> conflict-free float4 rows
constexpr int STG_BYTES = 128 * STG_PITCH * 4;

template <int ACT>
__device__ __forceinline__ float act_ct(float v, int act, float slope) {
    if constexpr (ACT < 0) return act_apply(v, act, slope);
    else return act_apply(v, ACT, slope);  // constant-folds to the single selected branch
}
template <int EPI>
__device__ __forceinline__ float epi_ct(float v, float a, int epi) {
    const int e = (EPI < 0) ? epi : EPI;
    switch (e) {
        case RCN_EPI_GDN: return a * rsqrtf(v);
        case RCN_EPI_IGDN: return a * sqrtf(v);
        case RCN_EPI_MUL_AUXP1: return v * (a + 1.f);
        case RCN_EPI_MULP1_AUX: return (v + 1.f) * a;
        case RCN_EPI_SIGMOID_GATE: return a * (1.f / (1.f + expf(-v)));
        default: return v;
    }
}

// One staged slab (128 pixels x up to 32 channels starting at channel n0 + c0) -> fused element-wise -> global.
// ACT / EPI are compile-time so the loop body only holds the math of the selected variant (a runtime switch
// inside the loop gets if-converted into ALL branches executing predicated: ~1000 instructions per 4 elements).
template <int ACT, int EPI>
__device__ __forceinline__ void epilogue_slab(const rcn_conv_desc& p, int dbg, const float* __restrict__ stg, int cols, int n,
                                              int c_base, int x0, int y0, int et, bool vec) {
    const int Ho = p.H, Wo = p.W;
    if (vec) {
        // Loads-first, fully unrolled: every thread issues ALL of its global loads for the slab (residual / aux,
        // 16 B each) before the first dependent instruction, so ~8 requests per thread are in flight instead of 1-2
        // (the epilogue was latency-bound on HBM round trips).  Items are float4s of the STORED tensor:
        //   NHWC store : item = (row, 4 consecutive channels)
        //   PS2 store  : item = (row, sub-pixel (i,j), 4 consecutive shuffled channels) -- conv channels cc*4 + sub,
        //                gathered with stride 4 from the slab, so the shuffled tensor is also written 16 B at a time.
        const bool ps = (p.store == RCN_STORE_PS2);
        const int groups = cols >> 2;           // float4 items per row (both layouts)
        const int total = 128 * groups;
        const int Hs = ps ? 2 * Ho : Ho, Ws = ps ? 2 * Wo : Wo;
        constexpr int IT = (128 * 8) / EPI_THREADS;  // 128 rows * 8 groups over the epilogue threads
        float4 accv[IT], resv[IT], auxv[IT], biasv[IT];
        long long opix[IT];
        int och[IT];
        bool ok[IT];
#pragma unroll
        for (int u = 0; u < IT; ++u) {
            const int e = et + EPI_THREADS * u;
            const int row = e / groups, g = e - row * groups;
            const int ho = y0 + row / TILE_W, wo = x0 + (row % TILE_W);
            ok[u] = (e < total) && (ho < Ho) && (wo < Wo);
            resv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            auxv[u] = resv[u];
            accv[u] = resv[u];
            biasv[u] = resv[u];
            opix[u] = 0;
            och[u] = 0;
            if (!ok[u]) continue;
            const float* srow = stg + row * STG_PITCH;
            if (!ps) {
                och[u] = c_base + 4 * g;
                opix[u] = ((long long)n * Ho + ho) * Wo + wo;
                accv[u] = *reinterpret_cast<const float4*>(srow + 4 * g);
                if (EPI != 0) auxv[u] = *reinterpret_cast<const float4*>(p.aux + opix[u] * p.ldaux + och[u]);
                if (p.bias) biasv[u] = __ldg(reinterpret_cast<const float4*>(p.bias + och[u]));
            } else {
                const int sub = g & 3, ccg = g >> 2;          // groups = 4 sub-pixels x (cols/16) channel groups
                const int cc0 = (c_base >> 2) + 4 * ccg;      // first shuffled channel of the item
                och[u] = cc0;
                opix[u] = ((long long)n * Hs + 2 * ho + (sub >> 1)) * Ws + 2 * wo + (sub & 1);
                const float* sp = srow + 16 * ccg + sub;      // conv channel (cc0 + q)*4 + sub  ->  slab column 16*ccg + 4*q + sub
                accv[u] = make_float4(sp[0], sp[4], sp[8], sp[12]);
                if (p.bias) {
                    const float* bp = p.bias + (cc0 << 2) + sub;   // conv channel of element q: (cc0 + q)*4 + sub
                    biasv[u] = make_float4(__ldg(bp), __ldg(bp + 4), __ldg(bp + 8), __ldg(bp + 12));
                }
            }
            if (p.res) resv[u] = *reinterpret_cast<const float4*>(p.res + opix[u] * p.ldres + och[u]);
        }
#pragma unroll
        for (int u = 0; u < IT; ++u) {
            if (!ok[u]) continue;
            float val[4] = {accv[u].x + biasv[u].x, accv[u].y + biasv[u].y, accv[u].z + biasv[u].z, accv[u].w + biasv[u].w};
            const float ax[4] = {auxv[u].x, auxv[u].y, auxv[u].z, auxv[u].w};
            const float rv[4] = {p.res_scale * resv[u].x, p.res_scale * resv[u].y, p.res_scale * resv[u].z, p.res_scale * resv[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // conv channel of element j (bias / per-channel affine are indexed by it)
                const int c = ps ? ((och[u] + j) << 2) + (int)((opix[u] / Ws) & 1) * 2 + (int)(opix[u] % Ws & 1) : och[u] + j;
                if (p.cscale) val[j] = val[j] * (1.f + __ldg(p.cscale + n * p.Cout + c)) + __ldg(p.cshift + n * p.Cout + c);
                if (EPI != 0) val[j] = epi_ct<EPI>(val[j], ax[j], p.epi);
                if (p.res && p.res_pre) val[j] += rv[j];
                val[j] = act_ct<ACT>(val[j], p.act, p.slope);
                if (p.res && !p.res_pre) val[j] += rv[j];
            }
            if (!(dbg & 1)) {
                if (p.y) *reinterpret_cast<float4*>(p.y + opix[u] * p.ldy + och[u]) = make_float4(val[0], val[1], val[2], val[3]);
                if (p.y_hi) {
                    // the consumer's tcgen05 operand planes: x = hi + lo in bf16 (same rounding as rcn_split_bf16)
                    __nv_bfloat16 hh[4], ll[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        hh[j] = __float2bfloat16_rn(val[j]);
                        ll[j] = __float2bfloat16_rn(val[j] - __bfloat162float(hh[j]));
                    }
                    const long long po = opix[u] * p.Cp_out + och[u];
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.y_hi) + po) = *reinterpret_cast<uint2*>(hh);
                    if (p.y_lo) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.y_lo) + po) = *reinterpret_cast<uint2*>(ll);
                }
            }
        }
    } else {
        const bool ps = (p.store == RCN_STORE_PS2 || p.store == RCN_STORE_PS2_NCHW);
        const int Hs = ps ? 2 * Ho : Ho, Ws = ps ? 2 * Wo : Wo, Cs = ps ? p.Cout / 4 : p.Cout;
        const int total = 128 * cols;
#pragma unroll 2
        for (int e = et; e < total; e += EPI_THREADS) {
            const int row = e / cols, col = e - row * cols;
            const int ho = y0 + row / TILE_W, wo = x0 + (row % TILE_W);
            if (ho >= Ho || wo >= Wo) continue;
            const int c = c_base + col;
            const long long mpix = ((long long)n * Ho + ho) * Wo + wo;
            float val = stg[row * STG_PITCH + col];
            if (p.bias) val += __ldg(p.bias + c);
            if (p.cscale) val = val * (1.f + __ldg(p.cscale + n * p.Cout + c)) + __ldg(p.cshift + n * p.Cout + c);
            if (EPI != 0) val = epi_ct<EPI>(val, p.aux[mpix * p.ldaux + c], p.epi);
            int hh = ho, ww = wo, cc = c;
            if (ps) { cc = c >> 2; hh = 2 * ho + ((c >> 1) & 1); ww = 2 * wo + (c & 1); }
            const long long pix = ((long long)n * Hs + hh) * Ws + ww;
            float rv = 0.f;
            if (p.res) rv = p.res_scale * p.res[pix * p.ldres + cc];
            if (p.res && p.res_pre) val += rv;
            val = act_ct<ACT>(val, p.act, p.slope);
            if (p.res && !p.res_pre) val += rv;
            if (!(dbg & 1)) {
                if (p.store == RCN_STORE_NCHW || p.store == RCN_STORE_PS2_NCHW)
                    p.y[(((long long)n * Cs + cc) * Hs + hh) * Ws + ww] = val;
                else
                    p.y[pix * p.ldy + cc] = val;
            }
        }
    }
}

__device__ __forceinline__ void epilogue_dispatch(const rcn_conv_desc& p, int dbg, const float* stg, int cols, int n, int c_base,
                                                  int x0, int y0, int et, bool vec) {
#define RCN_EP(A, E) epilogue_slab<A, E>(p, dbg, stg, cols, n, c_base, x0, y0, et, vec)
    if (p.epi == RCN_EPI_NONE) {
        switch (p.act) {
            case RCN_ACT_NONE: RCN_EP(RCN_ACT_NONE, 0); break;
            case RCN_ACT_RELU: RCN_EP(RCN_ACT_RELU, 0); break;
            case RCN_ACT_LRELU: RCN_EP(RCN_ACT_LRELU, 0); break;
            case RCN_ACT_GELU: RCN_EP(RCN_ACT_GELU, 0); break;
            case RCN_ACT_SIGMOID: RCN_EP(RCN_ACT_SIGMOID, 0); break;
            case RCN_ACT_HALF_TANH: RCN_EP(RCN_ACT_HALF_TANH, 0); break;
            case RCN_ACT_HSWISH: RCN_EP(RCN_ACT_HSWISH, 0); break;
            default: RCN_EP(-1, 0); break;
        }
    } else if (p.act == RCN_ACT_NONE) {
        switch (p.epi) {
            case RCN_EPI_GDN: RCN_EP(RCN_ACT_NONE, RCN_EPI_GDN); break;
            case RCN_EPI_IGDN: RCN_EP(RCN_ACT_NONE, RCN_EPI_IGDN); break;
            case RCN_EPI_MUL_AUXP1: RCN_EP(RCN_ACT_NONE, RCN_EPI_MUL_AUXP1); break;
            case RCN_EPI_MULP1_AUX: RCN_EP(RCN_ACT_NONE, RCN_EPI_MULP1_AUX); break;
            default: RCN_EP(RCN_ACT_NONE, RCN_EPI_SIGMOID_GATE); break;
        }
    } else {
        RCN_EP(-1, -1);
    }
#undef RCN_EP
}

// Persistent kernel: one CTA per SM walks tiles t = blockIdx.x, +gridDim.x, ...; tile t -> (m-tile, n-tile).
// TMEM holds two 128-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const rcn_conv_desc& p = P.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B_BYTES = P.Ntile * BLOCK_K * 2;
    const int stage_bytes = (P.passes == 3 ? 2 : 1) * (A_BYTES + B_BYTES);  // multiple of 1024 (Ntile % 16 == 0)
    float* stg = reinterpret_cast<float*>(smem + (size_t)P.stages * stage_bytes);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * stage_bytes + STG_BYTES);
    uint64_t* empty_bar = full_bar + P.stages;
    uint64_t* tmem_full = empty_bar + P.stages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int pad = p.k >> 1;
    const int chunks = P.Cp / BLOCK_K;
    const int kiters = p.k * p.k * chunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            const bool loadA = !(P.dbg & 4);
            const uint32_t tx_bytes = loadA ? (uint32_t)stage_bytes : (uint32_t)((P.passes == 3 ? 2 : 1) * B_BYTES);
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
                const int nt = (int)(t % P.tiles_n);
                long long mt = t / P.tiles_n;
                const int tx = (int)(mt % P.tiles_x); mt /= P.tiles_x;
                const int ty = (int)(mt % P.tiles_y);
                const int n = (int)(mt / P.tiles_y);
                const int x0 = tx * TILE_W, y0 = ty * TILE_H, n0 = nt * P.Ntile;
                for (int it = 0; it < kiters; ++it) {
                    const int tap = it / chunks, ch = it - tap * chunks;
                    const int ky = tap / p.k, kx = tap - ky * p.k;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], tx_bytes);
                    // input box origin for this tap.  stride 1: shifted box of the same plane.  stride 2 (pad k/2):
                    // input row 2y+ky-pad lives in polyphase plane py = (ky-pad)&1 at row y + floor((ky-pad)/2).
                    int ax = x0 + kx - pad, ay = y0 + ky - pad, an = n;
                    if (P.s2) {
                        const int oy = ky - pad, ox = kx - pad;
                        const int py = oy & 1, px = ox & 1;
                        ay = y0 + ((oy - py) >> 1);
                        ax = x0 + ((ox - px) >> 1);
                        an = (py * 2 + px) * p.N + n;
                    }
                    if (loadA) tma_load_4d(sa, &map_a_hi, &full_bar[stage], ch * BLOCK_K, ax, ay, an);
                    tma_load_2d(sa + A_BYTES, &map_w_hi, &full_bar[stage], tap * P.Cp + ch * BLOCK_K, n0);
                    if (P.passes == 3) {
                        if (loadA) tma_load_4d(sa + A_BYTES + B_BYTES, &map_a_lo, &full_bar[stage], ch * BLOCK_K, ax, ay, an);
                        tma_load_2d(sa + 2 * A_BYTES + B_BYTES, &map_w_lo, &full_bar[stage], tap * P.Cp + ch * BLOCK_K, n0);
                    }
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t local = 0;
            for (long long t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++local) {
                const int n0 = (int)(t % P.tiles_n) * P.Ntile;
                int nact = p.Cout - n0;
                if (nact > P.Ntile) nact = P.Ntile;
                nact = (nact + 15) & ~15;
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nact >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint32_t ab = local & 1;
                mbar_wait(&tmem_empty[ab], ((local >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + ab * 128;
                uint32_t acc = 0;
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t a_hi = make_sw128_desc(sa), b_hi = make_sw128_desc(sa + A_BYTES);
                    const uint64_t a_lo = make_sw128_desc(sa + A_BYTES + B_BYTES), b_lo = make_sw128_desc(sa + 2 * A_BYTES + B_BYTES);
#pragma unroll
                    for (int j = 0; j < ((P.dbg & 2) ? 0 : BLOCK_K / 16); ++j) {
                        const uint64_t adv = (uint64_t)((j * 32) >> 4);  // 16 bf16 = 32 B along K inside the swizzle atom
                        if (P.passes == 3) {
                            umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, acc);
                            acc = 1;
                            umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
                        }
                        umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, acc);
                        acc = 1;
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[ab]);
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue: TMEM -> smem slab -> fused element-wise -> coalesced global =================
        const int q = warp & 3;          // TMEM lane quarter this warp may access (hardware: warp_id % 4)
        const int jsub = (warp - 2) >> 2;  // which 8-column block of the 32-column slab this warp drains (0..3)
        const int m = q * 32 + lane;
        const int et = threadIdx.x - 64;
        const bool vec = !(P.dbg & 8) && ((p.store == RCN_STORE_NHWC && (p.Cout & 3) == 0) ||
                                          (p.store == RCN_STORE_PS2 && (p.Cout & 15) == 0 && p.epi == RCN_EPI_NONE)) &&
                         ((p.ldy & 3) == 0) &&
                         (!p.y || (reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                         (!p.res || (((p.ldres & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0))) &&
                         (p.epi == RCN_EPI_NONE || (((p.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0))) &&
                         (!p.bias || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0));
        uint32_t local = 0;
        for (long long t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++local) {
            const int nt = (int)(t % P.tiles_n);
            long long mt = t / P.tiles_n;
            const int tx = (int)(mt % P.tiles_x); mt /= P.tiles_x;
            const int ty = (int)(mt % P.tiles_y);
            const int n = (int)(mt / P.tiles_y);
            const int x0 = tx * TILE_W, y0 = ty * TILE_H, n0 = nt * P.Ntile;
            int ncols = p.Cout - n0;
            if (ncols > P.Ntile) ncols = P.Ntile;
            const uint32_t ab = local & 1;
            mbar_wait(&tmem_full[ab], (local >> 1) & 1);
            tc_fence_after();
            for (int c0 = 0; c0 < ncols; c0 += STG_COLS) {
                uint32_t v[8];
                __syncwarp();  // tcgen05.ld is warp-collective (.sync.aligned)
                tmem_ld8(tmem_base + ab * 128 + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + 8 * jsub), v);
                if (c0 + STG_COLS >= ncols) tc_fence_before();  // last TMEM read of this tile by this warp
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");  // previous slab consumed (+ TMEM drained on the last slab)
                if (c0 + STG_COLS >= ncols && et == 0) {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[ab])) : "memory");
                }
                float4* dst = reinterpret_cast<float4*>(stg + m * STG_PITCH + 8 * jsub);
                dst[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
                dst[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
                asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");  // slab visible to all epilogue warps
                int cols = ncols - c0;
                if (cols > STG_COLS) cols = STG_COLS;
                if (!(P.dbg & 8)) epilogue_dispatch(p, P.dbg, stg, cols, n, n0 + c0, x0, y0, et, vec);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
    }
}


// ---------------------------------------------------------------- operand preparation
// fp32 NHWC (ld) -> bf16 hi / lo planes (npix, Cp), zero-padded channels; optional x*x (GDN)
__global__ void split_bf16_kernel(const float* __restrict__ x, int ldx, long long npix, int C, int Cp, int square,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const long long total = npix * (Cp / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % (Cp / 4)) * 4;
        const long long pix = i / (Cp / 4);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c4 + j;
            float t = (c < C) ? x[pix * ldx + c] : 0.f;
            v[j] = square ? t * t : t;
        }
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2bfloat16_rn(v[j]);
            l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
        }
        *reinterpret_cast<uint2*>(hi + pix * Cp + c4) = *reinterpret_cast<uint2*>(h);
        if (lo) *reinterpret_cast<uint2*>(lo + pix * Cp + c4) = *reinterpret_cast<uint2*>(l);
    }
}

// polyphase split for stride-2 convs: out[((py*2+px)*N + n), i, j, c] = x[n, 2i+py, 2j+px, c]
__global__ void split_bf16_s2_kernel(const float* __restrict__ x, int ldx, int N, int H, int W, int C, int Cp,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const int H2 = H / 2, W2 = W / 2;
    const long long total = (long long)N * H * W * (Cp / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % (Cp / 4)) * 4;
        long long t = i / (Cp / 4);
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (c4 + j < C) ? x[(((long long)n * H + h) * W + w) * ldx + c4 + j] : 0.f;
        __nv_bfloat16 hh[4], ll[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hh[j] = __float2bfloat16_rn(v[j]);
            ll[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hh[j]));
        }
        const long long plane = (long long)((h & 1) * 2 + (w & 1)) * N + n;
        const long long o = ((plane * H2 + (h >> 1)) * W2 + (w >> 1)) * Cp + c4;
        *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<uint2*>(hh);
        if (lo) *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<uint2*>(ll);
    }
}

// OIHW fp32 -> [Cout][k*k][Cp] bf16 hi / lo (K-major rows for the B operand)
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int Cp,
                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const long long total = (long long)Cout * k * k * Cp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        long long t = i / Cp;
        const int tap = (int)(t % (k * k));
        const int co = (int)(t / (k * k));
        const float v = (c < Cin) ? w[((long long)co * Cin + c) * k * k + tap] : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

bool make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int Cp) {
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
    cuuint32_t box[4] = {BLOCK_K, TILE_W, TILE_H, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_w_map(CUtensorMap* m, const void* base, int Cout, long long Ktot, int Ntile) {
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)Ntile};
    cuuint32_t es[2] = {1, 1};
    return get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
}  // namespace rcn

using namespace rcn;

extern "C" int rcn_split_bf16(const float* x, int ldx, long long npix, int C, int Cp, int square, void* hi, void* lo, void* stream) {
    RCN_CHECK_ARG(x && hi && npix > 0 && C > 0 && Cp >= C && Cp % 64 == 0, "rcn_split_bf16: bad arguments");
    const long long total = npix * (Cp / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    split_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, npix, C, Cp, square, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_split_bf16");
    return RCN_OK;
}

// stride-2 operand: the four polyphase planes (py,px) of x, each (N, H/2, W/2, Cp), stacked on the batch axis
extern "C" int rcn_split_bf16_s2(const float* x, int ldx, int N, int H, int W, int C, int Cp, void* hi, void* lo, void* stream) {
    RCN_CHECK_ARG(x && hi && N > 0 && C > 0 && Cp >= C && Cp % 64 == 0, "rcn_split_bf16_s2: bad arguments");
    RCN_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "rcn_split_bf16_s2: H and W must be even");
    const long long total = (long long)N * H * W * (Cp / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    split_bf16_s2_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, N, H, W, C, Cp, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_split_bf16_s2");
    return RCN_OK;
}

extern "C" int rcn_pack_conv_weight_tc(const float* w_oihw, int Cout, int Cin, int k, int Cp, void* hi, void* lo, void* stream) {
    RCN_CHECK_ARG(w_oihw && hi && lo && Cp >= Cin && Cp % 64 == 0, "rcn_pack_conv_weight_tc: bad arguments");
    const long long total = (long long)Cout * k * k * Cp;
    long long blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    pack_weight_tc_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, k, Cp, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_pack_conv_weight_tc");
    return RCN_OK;
}

extern "C" int rcn_conv2d_tc(const rcn_conv_desc* d, const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, int Cp,
                             int passes, void* stream) {
    RCN_CHECK_ARG(d && (d->y || d->y_hi) && x_hi && w_hi, "rcn_conv2d_tc: null pointer");
    if (d->y_hi) {
        const bool ps2 = d->store == RCN_STORE_PS2;
        const int cs = ps2 ? d->Cout / 4 : d->Cout;
        RCN_CHECK_ARG((d->store == RCN_STORE_NHWC || ps2) && d->Cp_out == cs && (cs % 64) == 0 && (!ps2 || d->epi == RCN_EPI_NONE) &&
                          (!d->y || ((d->ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 15) == 0)) &&
                          (!d->res || ((d->ldres & 3) == 0 && (reinterpret_cast<uintptr_t>(d->res) & 15) == 0)) &&
                          (d->epi == RCN_EPI_NONE || ((d->ldaux & 3) == 0 && (reinterpret_cast<uintptr_t>(d->aux) & 15) == 0)) &&
                          (!d->bias || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0),
                      "rcn_conv2d_tc: operand-plane emission needs an NHWC / pixel-shuffle store with a multiple of 64 stored "
                      "channels and 16-byte aligned tensors");
    }
    RCN_CHECK_ARG(passes == 1 || (passes == 3 && x_lo && w_lo), "rcn_conv2d_tc: passes must be 1 or 3 (3 needs the lo planes)");
    RCN_CHECK_ARG(d->k == 1 || d->k == 3, "rcn_conv2d_tc: kernel size %d unsupported", d->k);
    RCN_CHECK_ARG(d->stride == 1 || d->stride == 2, "rcn_conv2d_tc: stride %d unsupported", d->stride);
    RCN_CHECK_ARG(d->stride == 1 || (d->H % 2 == 0 && d->W % 2 == 0), "rcn_conv2d_tc: stride 2 needs even H and W");
    RCN_CHECK_ARG(Cp % 64 == 0 && Cp >= d->Cin, "rcn_conv2d_tc: Cp must be a multiple of 64 >= Cin");
    RCN_CHECK_ARG(d->epi == RCN_EPI_NONE || d->aux, "rcn_conv2d_tc: epilogue needs aux");
    RCN_CHECK_ARG(get_encode() != nullptr, "rcn_conv2d_tc: cuTensorMapEncodeTiled is not available from the driver");
    const bool ps = d->store == RCN_STORE_PS2 || d->store == RCN_STORE_PS2_NCHW;
    RCN_CHECK_ARG(!ps || (d->Cout % 4 == 0), "rcn_conv2d_tc: pixel shuffle needs Cout %% 4 == 0");
    TcParams P;
    P.d = *d;
    P.s2 = (d->stride == 2);
    if (P.s2) { P.d.H = d->H / 2; P.d.W = d->W / 2; }   // the kernel works in output geometry (== polyphase plane geometry)
    P.Cp = Cp;
    P.passes = passes;
    int nt = d->Cout >= 128 ? 128 : ((d->Cout + 15) & ~15);
    P.Ntile = nt;
    P.tiles_x = (P.d.W + TILE_W - 1) / TILE_W;
    P.tiles_y = (P.d.H + TILE_H - 1) / TILE_H;
    const int stage_bytes = (passes == 3 ? 2 : 1) * (A_BYTES + nt * BLOCK_K * 2);
    int stages = (226 * 1024 - STG_BYTES - 1024 - 256) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    P.stages = stages;
    {
        const char* e = getenv("RCN_TC_DEBUG");
        P.dbg = e ? atoi(e) : 0;
        const char* st = getenv("RCN_TC_STAGES");
        if (st && atoi(st) >= 2 && atoi(st) <= stages) P.stages = stages = atoi(st);
    }
    const size_t smem = (size_t)stages * stage_bytes + STG_BYTES + 1024 + 256;
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    const long long Ktot = (long long)d->k * d->k * Cp;
    const int planes = P.s2 ? 4 * d->N : d->N;
    bool ok = make_act_map(&ma_hi, x_hi, planes, P.d.H, P.d.W, Cp) && make_w_map(&mw_hi, w_hi, d->Cout, Ktot, nt);
    if (passes == 3) ok = ok && make_act_map(&ma_lo, x_lo, planes, P.d.H, P.d.W, Cp) && make_w_map(&mw_lo, w_lo, d->Cout, Ktot, nt);
    else { ma_lo = ma_hi; mw_lo = mw_hi; }
    RCN_CHECK_ARG(ok, "rcn_conv2d_tc: cuTensorMapEncodeTiled failed");
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    P.tiles_n = (d->Cout + nt - 1) / nt;
    P.total_tiles = (long long)P.tiles_x * P.tiles_y * d->N * P.tiles_n;
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    const unsigned grid = (unsigned)(P.total_tiles < num_sms ? P.total_tiles : num_sms);  // persistent: one CTA per SM
    conv_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(ma_hi, ma_lo, mw_hi, mw_lo, P);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_conv2d_tc");
    return RCN_OK;
}
