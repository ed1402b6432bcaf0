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
// Warp roles (352 threads): warp 0 = TMA producer (one lane); warps 1-2 = MMA issuers (one lane each; warp 1 also owns the TMEM
// allocation).  Long-K layers (3x3) alternate their pipeline stages between the two issuers, each accumulating into its own
// TMEM accumulator -- a single issuing thread, not the tensor pipe, bounded those layers -- and the epilogue adds the two partial
// accumulators in a fixed order.  Warps 3-10 = epilogue (TMEM lane quarter = warp_id % 4; each warp drains Ntile/2 or 64 columns).
// mbarrier ring: full[s] (TMA -> MMA warps), empty[s] (tcgen05.commit of every issuer -> TMA), tmem_full[2] (last commits of a
// tile -> epilogue), tmem_empty[2] (8 epilogue warps -> MMA issuers): the accumulators are double-buffered in TMEM (4 x 128 columns).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rcn {
namespace {

constexpr int TILE_H = 8, TILE_W = 16;
// K chunk per pipeline stage: 64 channels (128-byte swizzled rows) in general; 32 / 16 channels (64- / 32-byte swizzled rows) for
// layers with <= 32 / <= 16 input channels, whose operand planes are then only that wide (the packed-Bayer ingest convs and the
// condition UNet would otherwise move 2-16x zero padding through L2 and shared memory for each of the 9 taps).

struct TcParams {
    rcn_conv_desc d;
    int Cp;       // padded input channels of the bf16 planes / packed weights (multiple of bk)
    int bk;       // K chunk per stage: 64, 32 or 16 channels
    int a_bytes, b_bytes;   // shared-memory footprint of one A / B tile (1024-byte multiples)
    int Ntile;    // output channels per tile (multiple of 16, <= 128)
    int tiles_x, tiles_y, tiles_n;
    long long total_tiles;
    int passes;   // 1 or 3
    int stages;
    int s2;       // stride-2 conv: A planes are the 4 polyphase components stacked on the batch axis ((py*2+px)*N + n)
    int nmma;     // MMA-issuing warps: 2 for long K loops (3x3 layers), 1 otherwise
    int wcw;      // accumulator columns per epilogue warp: 64, or Ntile / 2 when that keeps all 8 warps busy (Ntile = 64, 96)
    int halo;     // halo-tile kernel (3x3 stride-1 layers with 64-channel K chunks): A operand = one (16+2) x (8+2)-pixel box per chunk
    int na;       // halo kernel: number of A (halo box) buffers in the ring
    int tps;      // halo kernel: filter taps per weight stage (9 when the nine weight tiles of a chunk fit in one stage, else 1)
    int rv;       // epilogue: row-vector (256-bit, no transposition) instead of the transposing one (VEC kernels only)
    int in_f16;   // operand planes and packed weights are fp16 (kind::f16 with f16 A/B formats) instead of bf16
    int dbg;      // RCN_TC_DEBUG bit mask (perf triage only): 1 no stores, 2 no MMA, 4 no A loads, 8 no epilogue math, 16 centre-tap A loads only
};

#ifndef RCN_TC_EPI_WARPS
#define RCN_TC_EPI_WARPS 8
#endif
constexpr int EPI_WARPS = RCN_TC_EPI_WARPS;   // EPI_WARPS / 4 warps per TMEM lane quarter (warp_id % 4), each draining WCOLS accumulator columns
constexpr int WCOLS = 128 / (EPI_WARPS / 4);  // 64 (8 warps) or 32 (16 warps)
constexpr int QN = WCOLS / 16;                // 16-column quarters per warp and tile
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int MMA_WARPS = 2;             // warps 1, 2: each issues half of a stage's k16 steps into its own TMEM accumulator
// Warp groups: warps 0-3 = TMA producer, MMA issuer(s), one idle warp; warps 4-11 = epilogue.  The roles are aligned to warp
// groups so that the register file can be re-split after the prologue (setmaxnreg): the three single-lane warps need ~40
// registers, the epilogue (64 accumulator columns, prefetched residual / aux operands, 256-bit stores) wants > 168 -- with one
// budget for all 12 warps ptxas spilled loop counters of the epilogue to local memory (ncu: LDL round trips through L2 on every
// tile, ~4 % of the launch).  384 threads x 168 = 64512 registers at launch; 128 x 96 + 256 x 200 afterwards (40 for the light warps made ptxas spill inside the halo kernel's MMA issue loop: 3x slower).
constexpr int EPI_WARP0 = 4;             // first epilogue warp
constexpr int TC_THREADS = 32 * EPI_WARP0 + EPI_THREADS;
constexpr int REGS_LIGHT = 96, REGS_EPI = 200;
// Called at the top of each role's branch (every warp of a warp group executes the same one), so that ptxas knows the budget
// of the code that follows: issued once before the role dispatch, the epilogue of the halo kernel was still compiled against
// the smaller budget and spilled 2 KB per thread.
__device__ __forceinline__ void regs_light() {
#if RCN_TC_EPI_WARPS == 8
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_LIGHT));
#endif
}
__device__ __forceinline__ void regs_epilogue() {
#if RCN_TC_EPI_WARPS == 8
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
#endif
}
constexpr int SLAB_FLOATS = 32 * 16;     // warp-private transposition slab: 32 pixel rows x 16 floats, 16-byte chunks XOR-swizzled
constexpr int STG_BYTES = EPI_WARPS * SLAB_FLOATS * 4;   // 16 KB
// The bias vector is staged in shared memory once per CTA: with ~220 KB of the SM's 228 KB carved out as shared memory there
// is practically no L1 left, and per-item bias loads from every SM hammer the same few L2 lines (measured: ~7 us per
// 128x128 tile, the whole epilogue cost of the 1x1 layers).
constexpr int BIAS_MAX = 2048;
constexpr int BIAS_BYTES = BIAS_MAX * 4;

// The generic (runtime-selected) variants are real function calls: inlining the full activation switch (erff, tanhf, expf)
// into the unrolled per-element code multiplies the kernel size and the register pressure for combinations that are rare.
__device__ __noinline__ float act_generic(float v, int act, float slope) { return act_apply(v, act, slope); }
__device__ __noinline__ float epi_generic(float v, float a, int epi) {
    switch (epi) {
        case RCN_EPI_GDN: return a * rsqrtf(v);
        case RCN_EPI_IGDN: return a * sqrtf(v);
        case RCN_EPI_MUL_AUXP1: return v * (a + 1.f);
        case RCN_EPI_MULP1_AUX: return (v + 1.f) * a;
        case RCN_EPI_SIGMOID_GATE: return a * (1.f / (1.f + expf(-v)));
        default: return v;
    }
}
template <int ACT>
__device__ __forceinline__ float act_ct(float v, int act, float slope) {
    if constexpr (ACT < 0) return act_generic(v, act, slope);
    else return act_apply(v, ACT, slope);  // constant-folds to the single selected branch
}
template <int EPI>
__device__ __forceinline__ float epi_ct(float v, float a, int epi) {
    if constexpr (EPI < 0) return epi_generic(v, a, epi);
    else if constexpr (EPI == RCN_EPI_GDN) return a * rsqrtf(v);
    else if constexpr (EPI == RCN_EPI_IGDN) return a * sqrtf(v);
    else if constexpr (EPI == RCN_EPI_MUL_AUXP1) return v * (a + 1.f);
    else if constexpr (EPI == RCN_EPI_MULP1_AUX) return (v + 1.f) * a;
    else if constexpr (EPI == RCN_EPI_SIGMOID_GATE) return a * (1.f / (1.f + expf(-v)));
    else return v;
}

// perf triage (RCN_TC_DEBUG & 128): cycles one epilogue warp of CTA 0 spends waiting for accumulators / working, tiles
__device__ unsigned long long g_tcprof[16];
// The accumulator of a tile is the sum of the two MMA warps' partial accumulators (columns +0 and +128 of the tile's TMEM
// buffer): a fixed summation order, so results stay bit-reproducible (the codec needs encoder == decoder arithmetic).
__device__ __forceinline__ void tmem_ld16x2_async(uint32_t taddr, uint32_t* v, uint32_t* w, bool dual) {
    tmem_ld16_async(taddr, v);
    if (dual) tmem_ld16_async(taddr + 128, w);
}
__device__ __forceinline__ void tmem_wait_sum16(uint32_t* v, uint32_t* w, bool dual) {
    tmem_wait_ld16(v);
    if (dual) {
        tmem_wait_ld16(w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
    }
}
__device__ __forceinline__ void epi_release(uint64_t* empty_bar, int lane) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty_bar)) : "memory");
}

// Launch-invariant epilogue parameters held in REGISTERS.  Reading them from the kernel-parameter constant bank inside the
// per-element code costs a uniform load + compare + branch chain (~40 cycles each, a dozen per float4: measured ~2000 cycles
// per 16-column quarter, i.e. the whole cost of the 1x1 layers); the empty asm statements make the values opaque so the
// compiler cannot rematerialise them from the constant bank.
struct EpiRegs {
    float* y; const float* res; const float* aux; const float* cscale; const float* cshift;
    __nv_bfloat16* y_hi; __nv_bfloat16* y_lo;
    int H, W, N, Cout, ldy, ldres, ldaux, cpo;
    float slope, rpre, rpost;   // residual weights: v = act(v + rpre*res) + rpost*res  (0 when absent)
    uint32_t flags;
};
enum { EF_Y = 1, EF_HI = 2, EF_LO = 4, EF_CS = 8, EF_RES = 16, EF_PS = 32, EF_NOSTORE = 64, EF_SKIP = 128, EF_PROF = 256, EF_S2OUT = 512,
       EF_DUAL = 1024, EF_F16OUT = 2048, EF_SQOUT = 4096, EF_YNCHW = 8192, EF_AUXNCHW = 16384 };
__device__ __forceinline__ EpiRegs make_epi_regs(const rcn_conv_desc& p, int dbg, bool dual) {
    EpiRegs r;
    r.y = p.y; r.res = p.res; r.aux = p.aux; r.cscale = p.cscale; r.cshift = p.cshift;
    r.y_hi = reinterpret_cast<__nv_bfloat16*>(p.y_hi); r.y_lo = reinterpret_cast<__nv_bfloat16*>(p.y_lo);
    r.H = p.H; r.W = p.W; r.N = p.N; r.Cout = p.Cout; r.ldy = p.ldy; r.ldres = p.ldres; r.ldaux = p.ldaux; r.cpo = p.Cp_out;
    r.slope = p.slope;
    r.rpre = (p.res && p.res_pre) ? p.res_scale : 0.f;
    r.rpost = (p.res && !p.res_pre) ? p.res_scale : 0.f;
    r.flags = (p.y ? EF_Y : 0) | (p.y_hi ? EF_HI : 0) | (p.y_lo ? EF_LO : 0) | (p.cscale ? EF_CS : 0) | (p.res ? EF_RES : 0) |
              (p.store == RCN_STORE_PS2 ? EF_PS : 0) | ((dbg & 1) ? EF_NOSTORE : 0) | ((dbg & 8) ? EF_SKIP : 0) |
              ((dbg & 128) ? EF_PROF : 0) | (p.planes_s2 ? EF_S2OUT : 0) | (dual ? EF_DUAL : 0) | (p.out_fmt ? EF_F16OUT : 0) | (p.planes_square ? EF_SQOUT : 0) |
              (p.store == RCN_STORE_NCHW ? EF_YNCHW : 0) | (p.aux_nchw ? EF_AUXNCHW : 0);
    opaque_ptr(r.y); opaque_ptr(r.res); opaque_ptr(r.aux); opaque_ptr(r.cscale); opaque_ptr(r.cshift); opaque_ptr(r.y_hi); opaque_ptr(r.y_lo);
    opaque(r.H); opaque(r.W); opaque(r.N); opaque(r.Cout); opaque(r.ldy); opaque(r.ldres); opaque(r.ldaux); opaque(r.cpo);
    opaque(r.slope); opaque(r.rpre); opaque(r.rpost); opaque(r.flags);
    return r;
}

// Vector epilogue of one tile for one warp (NHWC and pixel-shuffle stores, 16-byte aligned tensors).
// The warp owns pixel rows [32q, 32q+32) of the 8x16 tile (image rows y0+2q, y0+2q+1) and accumulator columns
// [cb, cb + wcols), wcols <= 64, processed 16 columns ("quarter" hh) at a time.  Accumulators arrive one pixel row per lane
// (TMEM lane = pixel) and are transposed through a warp-private swizzled slab -- no block-wide barrier -- so that 4 lanes hold
// the 4 float4s of one pixel's 16 columns and every global access of the warp covers 8 pixels x 64 contiguous bytes:
//   item `it` of a lane = pixel (y0 + 2q + (it >> 1), x0 + (lane >> 2) + 8 (it & 1)), float4 `lane & 3` of the slab row.
//   NHWC store : slab row = 16 consecutive channels; float4 c = channels cb + 16hh + 4c ..+3 of that pixel.
//   pixel shuffle (weights packed with ps_perm: accumulator column 64g + 16s + c holds conv channel 64g + 4c + s): quarter hh is
//                sub-pixel s = hh, i.e. OUTPUT pixel (2y + (hh >> 1), 2x + (hh & 1)), and its slab row = 16 consecutive shuffled
//                channels cb/4 ..+15 of that pixel -- the same 64-byte-contiguous store pattern as the NHWC case.
// The residual (or aux) operand of all 16 items is requested BEFORE the wait on the accumulator barrier, so its HBM latency
// hides behind the MMAs of this tile; the TMEM load of quarter hh+1 is in flight while quarter hh is finished.
template <int ACT, int EPI, bool DUAL, int TW>
__device__ __forceinline__ void epilogue_tile_vec(const EpiRegs& r, int act, int epi, uint32_t slab, uint32_t sbias, uint32_t taddr,
                                                  uint64_t* full_bar, uint32_t parity, uint64_t* empty_bar, int n, int x0, int y0,
                                                  int cb, int wcols, int q, int lane) {
    const bool ps = (r.flags & EF_PS) != 0;
    const int chunk = lane & 3, r0 = lane >> 2;
    // tile geometry TW x (128 / TW): accumulator row m = pixel (m / TW, m % TW).  Item `it` of a lane is row 32q + r0 + 8 it:
    // TW = 16 -> pixel (2q + (it >> 1), r0 + 8 (it & 1));  TW = 8 (halo-tile kernel) -> pixel (4q + it, r0)
#define RCN_DY(it) (TW == 16 ? ((it) >> 1) : (it))
#define RCN_DX8(it) (TW == 16 ? ((it) & 1) : 0)
    const int hy = y0 + (TW == 16 ? 2 : 4) * q, wx = x0 + r0;
    bool okp[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) okp[it] = (hy + RCN_DY(it) < r.H) && (wx + 8 * RCN_DX8(it) < r.W);
    // stored-tensor geometry of the lane's items
    long long pix0;
    int sx, sy, chb, chs, Wst;      // pixel steps for x + 8 / y + 1; first stored channel of quarter 0 and its step per quarter
    if (!ps) {
        pix0 = ((long long)n * r.H + hy) * r.W + wx;
        sx = 8; sy = r.W; chb = cb + 4 * chunk; chs = 16; Wst = 0;
    } else {
        Wst = 2 * r.W;
        pix0 = ((long long)n * 2 * r.H + 2 * hy) * Wst + 2 * wx;
        sx = 16; sy = 2 * Wst; chb = (cb >> 2) + 4 * chunk; chs = 0;
    }
    const int lane_cols = 4 * chunk;   // columns of a quarter that must exist for this lane's float4 to be valid
    // stored pixel of quarter hh relative to ip[it]: pixel shuffle -> sub-pixel (hh >> 1, hh & 1)
#define RCN_QOFF(hh) (ps ? (long long)((hh) >> 1) * Wst + ((hh) & 1) : 0ll)
    long long ip[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) ip[it] = pix0 + RCN_DX8(it) * sx + (long long)RCN_DY(it) * sy;
    const bool has_res = (r.flags & EF_RES) != 0;
    constexpr bool has_aux = (EPI != 0);
    // operand prefetched ahead of the accumulator: the residual when there is one, else aux
    const float* pre_ptr = has_res ? r.res : (has_aux ? r.aux : nullptr);
    const int pre_ld = has_res ? r.ldres : r.ldaux;
    float4 pre[QN][4];
#pragma unroll
    for (int hh = 0; hh < QN; ++hh) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            pre[hh][it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pre_ptr && (16 * hh + lane_cols < wcols) && okp[it])
                pre[hh][it] = ldg4(pre_ptr + (ip[it] + RCN_QOFF(hh)) * pre_ld + chb + chs * hh);
        }
    }
    const bool prof = (r.flags & EF_PROF) && blockIdx.x == 0 && threadIdx.x == 32 * EPI_WARP0;
    constexpr bool dual = DUAL;
    long long tp0 = 0, tp1 = 0;
    if (prof) tp0 = clock64();
    mbar_wait(full_bar, parity);
    tc_fence_after();
    if (prof) { tp1 = clock64(); atomicAdd(&g_tcprof[0], (unsigned long long)(tp1 - tp0)); atomicAdd(&g_tcprof[2], 1ull); }
    if (wcols <= 0 || (r.flags & EF_SKIP)) { epi_release(empty_bar, lane); return; }
    uint32_t v[16], w2[16];
    tmem_ld16x2_async(taddr, v, w2, dual);
    const int wsw = (lane >> 1) & 3;   // write-side swizzle of row `lane`
    const int rsw = (r0 >> 1) & 3;     // read-side swizzle of rows r0 + 8*it
#pragma unroll
    for (int hh = 0; hh < QN; ++hh) {
        if (16 * hh >= wcols) break;   // warp-uniform
        const bool ah = 16 * hh + lane_cols < wcols;
        const int ch = chb + chs * hh;
        // per-channel parameters of this lane's 4 elements (conv channel index of element e: c0 + ce * e)
        const int c0 = ps ? (cb + 16 * chunk + hh) : ch, ce = ps ? 4 : 1;
        const long long qoff = RCN_QOFF(hh);
        float bi[4] = {0.f, 0.f, 0.f, 0.f}, cs[4] = {1.f, 1.f, 1.f, 1.f}, csh[4] = {0.f, 0.f, 0.f, 0.f};
        if (ah) {
            if (sbias) {
#pragma unroll
                for (int e = 0; e < 4; ++e) bi[e] = lds1(sbias + 4u * (uint32_t)(c0 + ce * e));
            }
            if (r.flags & EF_CS) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    cs[e] = 1.f + __ldg(r.cscale + n * r.Cout + c0 + ce * e);
                    csh[e] = __ldg(r.cshift + n * r.Cout + c0 + ce * e);
                }
            }
        }
        // second operand (aux when a residual is also present)
        float4 sec[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            sec[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_aux && has_res && ah && okp[it]) sec[it] = ldg4(r.aux + (ip[it] + qoff) * r.ldaux + ch);
        }
        // ---- transpose: row `lane` <- the 16 accumulator columns of this quarter
        tmem_wait_sum16(v, w2, dual);
        const uint32_t wrow = slab + (uint32_t)lane * 64u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 t = make_float4(__uint_as_float(v[4 * k]), __uint_as_float(v[4 * k + 1]), __uint_as_float(v[4 * k + 2]), __uint_as_float(v[4 * k + 3]));
            sts4(wrow + 16u * (uint32_t)(k ^ wsw), t);
        }
        if (16 * (hh + 1) < wcols) tmem_ld16x2_async(taddr + 16 * (hh + 1), v, w2, dual);   // next quarter in flight
        else epi_release(empty_bar, lane);                                      // last TMEM read of this tile by this warp
        __syncwarp();
        float4 acc[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) acc[it] = lds4(slab + (uint32_t)(r0 + 8 * it) * 64u + 16u * (uint32_t)(chunk ^ rsw));
        __syncwarp();
        if (!ah) continue;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            if (!okp[it]) continue;
            float val[4] = {acc[it].x + bi[0], acc[it].y + bi[1], acc[it].z + bi[2], acc[it].w + bi[3]};
            const float4 rr = has_res ? pre[hh][it] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 aa = has_res ? sec[it] : pre[hh][it];
            const float ax[4] = {aa.x, aa.y, aa.z, aa.w};
            const float rv[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                val[e] = fmaf(val[e], cs[e], csh[e]);                       // Res_GFM modulation (identity when absent)
                if (EPI != 0) val[e] = epi_ct<EPI>(val[e], ax[e], epi);
                val[e] = fmaf(r.rpre, rv[e], val[e]);                       // residual before / after the activation
                val[e] = act_ct<ACT>(val[e], act, r.slope);
                val[e] = fmaf(r.rpost, rv[e], val[e]);
            }
            if (r.flags & EF_NOSTORE) continue;
            if (r.flags & EF_Y) *reinterpret_cast<float4*>(r.y + (ip[it] + qoff) * r.ldy + ch) = make_float4(val[0], val[1], val[2], val[3]);
            if (r.flags & EF_HI) {
                // the consumer's tcgen05 operand planes: x = hi + lo in bf16 / fp16 (same rounding as rcn_split_bf16)
                const int f16 = (r.flags & EF_F16OUT) ? 1 : 0;
                uint16_t hb[4], lb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float pv = (r.flags & EF_SQOUT) ? val[e] * val[e] : val[e];
                    hb[e] = to_plane(pv, f16);
                    lb[e] = to_plane(pv - from_plane(hb[e], f16), f16);
                }
                long long po = (ip[it] + qoff) * r.cpo + ch;
                if (r.flags & EF_S2OUT) {
                    // polyphase layout of a stride-2 consumer: pixel (h, w) -> plane (h&1)*2 + (w&1), position (h>>1, w>>1)
                    const int h = hy + RCN_DY(it), w = wx + 8 * RCN_DX8(it);
                    const long long plane = (long long)(((h & 1) * 2 + (w & 1)) * r.N + n);
                    po = ((plane * (r.H >> 1) + (h >> 1)) * (r.W >> 1) + (w >> 1)) * r.cpo + ch;
                }
                *reinterpret_cast<uint2*>(r.y_hi + po) = *reinterpret_cast<uint2*>(hb);
                if (r.flags & EF_LO) *reinterpret_cast<uint2*>(r.y_lo + po) = *reinterpret_cast<uint2*>(lb);
            }
        }
    }
    if (prof) atomicAdd(&g_tcprof[1], (unsigned long long)(clock64() - tp1));
#undef RCN_QOFF
#undef RCN_DY
#undef RCN_DX8
}

// 256-bit global accesses (sm_100: LDG / STG .E.ENL2.256): one full 32-byte sector per lane and instruction
__device__ __forceinline__ void ldg256(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void stg256u(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                 "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// Row-vector epilogue of one tile for one warp: NO transposition.  A lane keeps the accumulator row it gets from TMEM (lane =
// pixel) and handles 16 columns = 16 consecutive stored channels of its pixel at a time: 64 contiguous bytes of the fp32 map
// (two 256-bit stores), 32 contiguous bytes of each 16-bit operand plane (one 256-bit store) -- every access a full sector, no
// shared-memory slab, no warp barriers, and ~4x fewer instructions than the transposing epilogue below (measured with ncu on the
// 128 -> 128 layers: 2077 warp instructions per warp and tile, 29 of them local-memory spills waiting on L2).  The pixel-shuffle
// store has the same shape: with sub-pixel-grouped weights, block hh of a warp's 64 columns is sub-pixel hh and holds 16
// consecutive shuffled channels of OUTPUT pixel (2y + (hh >> 1), 2x + (hh & 1)).
// Requirements (host-checked, else the transposing epilogue runs): Cout % 16 == 0, pixel strides % 8 == 0 (fp32) / % 16 == 0
// (planes), 32-byte aligned bases.
// FEAT: compile-time mask of the RARE options (bit 0 NCHW aux / NCHW store, bit 1 squared planes, bit 2 polyphase planes, bit 3
// Res_GFM scale / shift).  They occur in ~15 of the ~1000 launches of a step, but with their code inlined behind run-time flags
// the common layers paid for them: measured on one box, 1x1 128 -> 128 0.949 ms with all four compiled in vs 0.785 ms without,
// GELU 64 -> 256 0.604 vs 0.487 ms (the epilogue is bound by instruction issue / fetch).  Kernels are instantiated lean (FEAT = 0)
// and, for the few (activation, combinator) pairs that need them, full.
template <int ACT, int EPI, bool DUAL, int TW, int FEAT>
__device__ __forceinline__ void epilogue_tile_rv(const EpiRegs& r, int act, int epi, uint32_t sbias, uint32_t taddr, uint64_t* full_bar,
                                                 uint32_t parity, uint64_t* empty_bar, int n, int x0, int y0, int cb, int wcols, int q,
                                                 int lane) {
    const bool ps = (r.flags & EF_PS) != 0;
    const int m = q * 32 + lane;
    const int ho = y0 + m / TW, wo = x0 + (m % TW);
    const bool ok = ho < r.H && wo < r.W;
    const bool has_res = (r.flags & EF_RES) != 0;
    constexpr bool has_aux = (EPI != 0);
    long long pix0;        // stored pixel of block 0
    int Wst = 0, chb, chs;
    if (!ps) { pix0 = ((long long)n * r.H + ho) * r.W + wo; chb = cb; chs = 16; }
    else { Wst = 2 * r.W; pix0 = ((long long)n * 2 * r.H + 2 * ho) * Wst + 2 * wo; chb = cb >> 2; chs = 0; }
#define RCN_QOFF(hh) (ps ? (long long)((hh) >> 1) * Wst + ((hh) & 1) : 0ll)
    // residual (else aux) of every block, requested BEFORE the wait on the accumulator barrier: its HBM latency hides behind the MMAs
    const float* pre_ptr = has_res ? r.res : (has_aux ? r.aux : nullptr);
    const int pre_ld = has_res ? r.ldres : r.ldaux;
    float pre[QN][16];
#pragma unroll
    for (int hh = 0; hh < QN; ++hh) {
#pragma unroll
        for (int j = 0; j < 16; ++j) pre[hh][j] = 0.f;
        if (pre_ptr && 16 * hh < wcols && ok) {
            if ((FEAT & 1) && !has_res && (r.flags & EF_AUXNCHW)) {
                // aux is an NCHW tensor (an API-facing map such as the lens-shading features): lanes are consecutive pixels of an
                // image row, so each channel's load is 32 / 64 contiguous bytes per image row of the tile
                const float* pp = r.aux + (((long long)n * r.Cout + cb + 16 * hh) * r.H + ho) * r.W + wo;
                const long long cst = (long long)r.H * r.W;
#pragma unroll
                for (int j = 0; j < 16; ++j) pre[hh][j] = __ldg(pp + j * cst);
            } else {
                const float* pp = pre_ptr + (pix0 + RCN_QOFF(hh)) * pre_ld + chb + chs * hh;
                ldg256(pp, &pre[hh][0]);
                ldg256(pp + 8, &pre[hh][8]);
            }
        }
    }
    static_assert(!DUAL, "the row-vector epilogue reads one accumulator per tile");
    mbar_wait(full_bar, parity);
    tc_fence_after();
    if (wcols <= 0 || (r.flags & EF_SKIP)) { epi_release(empty_bar, lane); return; }
    // TMEM -> registers two 16-column blocks at a time: tcgen05.wait::ld waits for EVERY outstanding load, so with one block in
    // flight each of the four waits of a tile exposed the load latency (ncu: ~12 % of the epilogue's samples on the instruction
    // after the wait).  Blocks (0,1) are requested together; while they are processed, (2,3) are in flight.
    uint32_t va[16], vb[16];
    tmem_ld16_async(taddr, va);
    if (16 < wcols) tmem_ld16_async(taddr + 16, vb);
    tmem_wait_ld16(va);
    tmem_wait_ld16(vb);
    if (wcols <= 32) epi_release(empty_bar, lane);      // every TMEM read of this tile by this warp has landed
#pragma unroll
    for (int hh = 0; hh < QN; ++hh) {
        if (16 * hh >= wcols) break;   // warp-uniform
        uint32_t* v = (hh & 1) ? vb : va;
        if (hh == 2) {                 // second pair: requested while blocks 0 and 1 were processed
            tmem_wait_ld16(va);
            tmem_wait_ld16(vb);
            epi_release(empty_bar, lane);
        }
        const long long spix = pix0 + RCN_QOFF(hh);
        const int ch = chb + chs * hh;                    // first stored channel of this block
        // conv channel of column j: plain store cb + 16 hh + j; pixel shuffle (sub-pixel-grouped rows) cb + 4 j + hh
        const int c0 = ps ? cb + hh : cb + 16 * hh, ce = ps ? 4 : 1;
        float sec[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sec[j] = 0.f;
        if (has_aux && has_res && ok) {
            const float* ap = r.aux + spix * r.ldaux + ch;
            ldg256(ap, &sec[0]);
            ldg256(ap + 8, &sec[8]);
        }
        float val[16];
        if (sbias) {
            if (!ps) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 b4 = lds4(sbias + 4u * (uint32_t)(c0 + 4 * k4));   // same address in every lane: broadcast
                    val[4 * k4] = __uint_as_float(v[4 * k4]) + b4.x; val[4 * k4 + 1] = __uint_as_float(v[4 * k4 + 1]) + b4.y;
                    val[4 * k4 + 2] = __uint_as_float(v[4 * k4 + 2]) + b4.z; val[4 * k4 + 3] = __uint_as_float(v[4 * k4 + 3]) + b4.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) val[j] = __uint_as_float(v[j]) + lds1(sbias + 4u * (uint32_t)(c0 + 4 * j));
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) val[j] = __uint_as_float(v[j]);
        }
        // this buffer's registers are consumed: request block hh + 2 into it; after the last request the accumulator is free
        if (16 * (hh + 2) < wcols) tmem_ld16_async(taddr + 16 * (hh + 2), v);
        if (!ok) continue;
        if ((FEAT & 8) && (r.flags & EF_CS)) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                val[j] = fmaf(val[j], 1.f + __ldg(r.cscale + n * r.Cout + c0 + ce * j), __ldg(r.cshift + n * r.Cout + c0 + ce * j));
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float rv = has_res ? pre[hh][j] : 0.f;
            const float ax = has_res ? sec[j] : pre[hh][j];
            if (EPI != 0) val[j] = epi_ct<EPI>(val[j], ax, epi);
            val[j] = fmaf(r.rpre, rv, val[j]);                       // residual before / after the activation
            val[j] = act_ct<ACT>(val[j], act, r.slope);
            val[j] = fmaf(r.rpost, rv, val[j]);
        }
        if (r.flags & EF_NOSTORE) continue;
        if (r.flags & EF_Y) {
            if ((FEAT & 1) && (r.flags & EF_YNCHW)) {      // API-facing NCHW map: 16 channel planes, each store coalesced along the image row
                float* yp = r.y + (((long long)n * r.Cout + cb + 16 * hh) * r.H + ho) * r.W + wo;
                const long long cst = (long long)r.H * r.W;
#pragma unroll
                for (int j = 0; j < 16; ++j) yp[j * cst] = val[j];
            } else {
                float* yp = r.y + spix * r.ldy + ch;
                stg256(yp, &val[0]);
                stg256(yp + 8, &val[8]);
            }
        }
        if (r.flags & EF_HI) {
            // the consumer's tcgen05 operand planes: x = hi + lo in bf16 / fp16 (same rounding as rcn_split_bf16), converted two
            // elements per instruction (F2FP pack); the format branch is warp-uniform and sits outside the element loop
            uint32_t hp[8], lp[8];
            if ((FEAT & 2) && (r.flags & EF_SQOUT)) {   // the consumer is a GDN norm pool: its operand is x^2 (rcn_split_bf16 with square = 1)
#pragma unroll
                for (int j = 0; j < 16; ++j) val[j] *= val[j];
            }
            if (r.flags & EF_F16OUT) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const __half2 h2 = __floats2half2_rn(val[2 * j], val[2 * j + 1]);
                    hp[j] = *reinterpret_cast<const uint32_t*>(&h2);
                    if (r.flags & EF_LO) {
                        const float2 hf = __half22float2(h2);
                        const __half2 l2 = __floats2half2_rn(val[2 * j] - hf.x, val[2 * j + 1] - hf.y);
                        lp[j] = *reinterpret_cast<const uint32_t*>(&l2);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(val[2 * j], val[2 * j + 1]);
                    hp[j] = *reinterpret_cast<const uint32_t*>(&h2);
                    if (r.flags & EF_LO) {
                        // bf16 -> fp32 is a 16-bit shift: low half << 16, high half masked
                        const float f0 = __uint_as_float(hp[j] << 16), f1 = __uint_as_float(hp[j] & 0xFFFF0000u);
                        const __nv_bfloat162 l2 = __floats2bfloat162_rn(val[2 * j] - f0, val[2 * j + 1] - f1);
                        lp[j] = *reinterpret_cast<const uint32_t*>(&l2);
                    }
                }
            }
            long long po = spix * r.cpo + ch;
            if ((FEAT & 4) && (r.flags & EF_S2OUT)) {
                // polyphase layout of a stride-2 consumer: pixel (h, w) -> plane (h&1)*2 + (w&1), position (h>>1, w>>1)
                const long long plane = (long long)(((ho & 1) * 2 + (wo & 1)) * r.N + n);
                po = ((plane * (r.H >> 1) + (ho >> 1)) * (r.W >> 1) + (wo >> 1)) * r.cpo + ch;
            }
            stg256u(r.y_hi + po, hp);
            if (r.flags & EF_LO) stg256u(r.y_lo + po, lp);
        }
    }
#undef RCN_QOFF
}

// Generic epilogue straight from the accumulator registers (one pixel row per lane): NCHW / pixel-shuffle-to-NCHW stores
// (lanes = consecutive pixels of an image row, i.e. the coalesced direction of an NCHW tensor) and odd / unaligned shapes.
template <int ACT, int EPI>
__device__ __forceinline__ void rows_block(const EpiRegs& r, int act, int epi, int store, uint32_t sbias, const uint32_t* v, int n, int ho,
                                           int wo, int cb, int wc) {
    const int Ho = r.H, Wo = r.W;
    const bool ps = (store == RCN_STORE_PS2 || store == RCN_STORE_PS2_NCHW);
    const bool nchw = (store == RCN_STORE_NCHW || store == RCN_STORE_PS2_NCHW);
    const int Hs = ps ? 2 * Ho : Ho, Ws = ps ? 2 * Wo : Wo, Cs = ps ? r.Cout / 4 : r.Cout;
    const long long mpix = ((long long)n * Ho + ho) * Wo + wo;
    const bool has_res = (r.flags & EF_RES) != 0, has_cs = (r.flags & EF_CS) != 0;
    const bool has_aux = EPI > 0 || (EPI < 0 && epi != RCN_EPI_NONE);
    float outv[16];
#pragma unroll
    for (int col = 0; col < 16; ++col) {
        outv[col] = 0.f;
        if (col >= wc) continue;
        const int c = cb + col;
        float val = __uint_as_float(v[col]);
        if (sbias) val += lds1(sbias + 4u * (uint32_t)c);
        if (has_cs) val = val * (1.f + __ldg(r.cscale + n * r.Cout + c)) + __ldg(r.cshift + n * r.Cout + c);
        if (has_aux)
            val = epi_ct<EPI>(val, (r.flags & EF_AUXNCHW) ? r.aux[(((long long)n * r.Cout + c) * Ho + ho) * Wo + wo] : r.aux[mpix * r.ldaux + c], epi);
        int hh = ho, ww = wo, cc = c;
        if (ps) { cc = c >> 2; hh = 2 * ho + ((c >> 1) & 1); ww = 2 * wo + (c & 1); }
        float rv = 0.f;
        if (has_res) rv = r.res[(((long long)n * Hs + hh) * Ws + ww) * r.ldres + cc];
        val = fmaf(r.rpre, rv, val);
        val = act_ct<ACT>(val, act, r.slope);
        val = fmaf(r.rpost, rv, val);
        outv[col] = val;
    }
    if (r.flags & EF_NOSTORE) return;
    if (store == RCN_STORE_PS2_NCHW) {
        // columns (4cc + 2i + j): the pair j = 0,1 is two adjacent output pixels -> one 8-byte store, 128 B per image row per warp
#pragma unroll
        for (int col = 0; col < 16; col += 2) {
            if (col >= wc) continue;
            const int c = cb + col, cc = c >> 2, i = (c >> 1) & 1;
            float* dst = r.y + (((long long)n * Cs + cc) * Hs + 2 * ho + i) * Ws + 2 * wo;
            if (col + 1 < wc) *reinterpret_cast<float2*>(dst) = make_float2(outv[col], outv[col + 1]);
            else dst[0] = outv[col];
        }
    } else {
#pragma unroll
        for (int col = 0; col < 16; ++col) {
            if (col >= wc) continue;
            const int c = cb + col;
            int hh = ho, ww = wo, cc = c;
            if (ps) { cc = c >> 2; hh = 2 * ho + ((c >> 1) & 1); ww = 2 * wo + (c & 1); }
            if (nchw) r.y[(((long long)n * Cs + cc) * Hs + hh) * Ws + ww] = outv[col];
            else r.y[(((long long)n * Hs + hh) * Ws + ww) * r.ldy + cc] = outv[col];
        }
    }
}
template <int ACT, int EPI, bool DUAL, int TW>
__device__ __forceinline__ void epilogue_tile_rows(const EpiRegs& r, int act, int epi, int store, uint32_t sbias, uint32_t taddr,
                                                   uint64_t* full_bar, uint32_t parity, uint64_t* empty_bar, int n, int x0, int y0, int cb,
                                                   int wcols, int q, int lane) {
    mbar_wait(full_bar, parity);
    tc_fence_after();
    if (wcols <= 0 || (r.flags & EF_SKIP)) { epi_release(empty_bar, lane); return; }
    const int m = q * 32 + lane;
    const int ho = y0 + m / TW, wo = x0 + (m % TW);
    const bool ok = ho < r.H && wo < r.W;
#pragma unroll 1
    for (int c0 = 0; c0 < wcols; c0 += 16) {
        uint32_t v[16], w2[16];
        __syncwarp();
        tmem_ld16x2_async(taddr + c0, v, w2, DUAL);
        tmem_wait_sum16(v, w2, DUAL);
        if (c0 + 16 >= wcols) epi_release(empty_bar, lane);
        if (ok) rows_block<ACT, EPI>(r, act, epi, store, sbias, v, n, ho, wo, cb + c0, wcols - c0 > 16 ? 16 : wcols - c0);
    }
}

// The MMA issue loop of one issuing warp, run by the whole (converged) warp with the tcgen05 instructions behind elect.sync.
// Everything it needs is hoisted into registers first: reading kernel parameters from the constant bank inside the loop cost an
// LDC / LDCU round trip per use (~100 of the ~830 cycles per stage the old `if (lane == 0)` loop took at 4 MMAs per stage).
// Descriptors are advanced arithmetically: the start-address field (bits 0-13, 16-byte units) of the stage-0 descriptor plus
// stage * (stage_bytes / 16) never carries out of the field (shared memory is < 256 KB).
struct MmaLoopArgs {
    uint32_t total_tiles, tiles_n, grid, first_tile;
    int Ntile, Cout, kiters, stages, mw, owner_mode;   // owner_mode 0: every stage (all of it, or this warp's half of its k16 steps); 1: stages with gstage % 2 == mw
    uint32_t sb16, oB, oAlo, oBlo, koff;                // descriptor offsets in 16-byte units
    uint32_t full_bar, empty_bar, tmem_full, tmem_empty;   // shared-space addresses of the barrier arrays
    uint32_t tmem_base, idesc_fmt;
    uint64_t desc0;                                      // A_hi descriptor of stage 0
    int skip_mma;
};
template <int PASSES, int SPW>
__device__ __forceinline__ void mma_issue_loop(const MmaLoopArgs& A) {
    uint32_t stage = 0, phase = 0, local = 0, gstage = 0;
    for (uint32_t t = A.first_tile; t < A.total_tiles; t += A.grid, ++local) {
        const int n0 = (int)(t % A.tiles_n) * A.Ntile;
        int nact = A.Cout - n0;
        if (nact > A.Ntile) nact = A.Ntile;
        nact = (nact + 15) & ~15;
        // instruction descriptor: D = f32 (bit 4), A / B format at bits 7 / 10 (0 = f16, 1 = bf16), N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | A.idesc_fmt | ((uint32_t)(nact >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t ab = local & 1;
        mbar_wait_a(A.tmem_empty + 8u * ab, ((local >> 1) & 1) ^ 1);  // epilogue has drained this accumulator pair
        tc_fence_after();
        const uint32_t tmem_d = A.tmem_base + ab * 256 + (uint32_t)A.mw * 128;
        uint32_t acc = 0;
        for (int it = 0; it < A.kiters; ++it, ++gstage) {
            if (A.owner_mode == 0 || (int)(gstage & 1u) == A.mw) {
                mbar_wait_a(A.full_bar + 8u * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t a_hi = A.desc0 + (uint64_t)(stage * A.sb16 + A.koff);
                    if (!A.skip_mma) issue_stage<PASSES, SPW>(tmem_d, a_hi, a_hi + A.oB, a_hi + A.oAlo, a_hi + A.oBlo, idesc, acc);
                    umma_commit_a(A.empty_bar + 8u * stage);  // frees the smem slot once this warp's MMAs on it have retired
                }
                __syncwarp();
                acc = 1;
            }
            if (++stage == (uint32_t)A.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_a(A.tmem_full + 8u * ab);
        __syncwarp();
    }
}

// Persistent kernel: one CTA per SM walks tiles t = blockIdx.x, +gridDim.x, ...; tile t -> (m-tile, n-tile).
// TMEM holds two 128-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
// One instantiation per (activation, combinator, store path): the epilogue of a single variant is ~3k SASS instructions; with all
// variants inlined behind a runtime switch the kernel was 50k instructions and a quarter of the epilogue warps' stall samples were
// instruction-cache misses (ncu: stall_no_inst).
#if RCN_TC_EPI_WARPS == 16
#define RCN_TC_BOUNDS __maxnreg__(RCN_TC_MAXNREG)
#else
#define RCN_TC_BOUNDS __launch_bounds__(TC_THREADS, 1)
#endif
template <int ACT, int EPI, bool VEC, bool DUAL, int FEAT>
__global__ void RCN_TC_BOUNDS
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const rcn_conv_desc& p = P.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int BLOCK_K = P.bk, A_BYTES = P.a_bytes, B_BYTES = P.b_bytes;
    const int stage_bytes = (P.passes == 3 ? 2 : 1) * (A_BYTES + B_BYTES);  // multiple of 1024
    float* stg = reinterpret_cast<float*>(smem + (size_t)P.stages * stage_bytes);
    float* sbias_mem = reinterpret_cast<float*>(smem + (size_t)P.stages * stage_bytes + STG_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * stage_bytes + STG_BYTES + BIAS_BYTES);
    uint64_t* empty_bar = full_bar + P.stages;
    uint64_t* tmem_full = empty_bar + P.stages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int pad = p.k >> 1;
    const int ksteps = BLOCK_K / 16;
    constexpr int nmma = DUAL ? MMA_WARPS : 1;
    // How two issuers share the K loop.  Stages of >= 2 k16 steps: both consume EVERY stage, half of its k16 steps each (slot barriers
    // count two commits).  Single-step stages (16-channel layers): stages alternate by the GLOBAL stage counter, with an even
    // ring size (host-enforced) so that every slot has a fixed owner.  Either way no warp ever skips a phase of a barrier it waits
    // on -- a parity wait cannot tell phase n from phase n+2, and an earlier scheme that alternated stages within a tile produced
    // stale-slot reads / spin-limit traps at exactly the tile counts where one issuer ran a ring ahead of the other.
    const bool splitk = DUAL && ksteps >= 2;
    const int chunks = P.Cp / BLOCK_K;
    const int kiters = p.k * p.k * chunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], splitk ? 2 : 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], nmma); mbar_init(&tmem_empty[b], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.bias) {
        for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) sbias_mem[i] = __ldg(p.bias + i);
    }
    const uint32_t sbias = p.bias ? smem_u32(sbias_mem) : 0u;   // shared-space address (0 = no bias)
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        regs_light();
        // ================= TMA producer =================
        // Same shape as the MMA issue loops: the whole warp walks the loop (uniform control flow), one elected lane issues the
        // TMA instructions; loop invariants live in registers; (tap, chunk) advance incrementally instead of by division.
        {
            int dbg = P.dbg, k = p.k, chunks_ = chunks, kit = kiters, stages = P.stages, s2 = P.s2, Nimg = p.N, Cp = P.Cp, Ntile = P.Ntile;
            int three = (P.passes == 3);
            uint32_t sbytes = (uint32_t)stage_bytes, abytes = (uint32_t)A_BYTES, bbytes = (uint32_t)B_BYTES, bk = (uint32_t)BLOCK_K;
            uint32_t total_tiles = (uint32_t)P.total_tiles, tiles_n = (uint32_t)P.tiles_n, tiles_x = (uint32_t)P.tiles_x, tiles_y = (uint32_t)P.tiles_y;
            opaque(dbg); opaque(k); opaque(chunks_); opaque(kit); opaque(stages); opaque(s2); opaque(Nimg); opaque(Cp); opaque(Ntile);
            opaque(three); opaque(sbytes); opaque(abytes); opaque(bbytes); opaque(bk); opaque(total_tiles); opaque(tiles_n);
            opaque(tiles_x); opaque(tiles_y);
            const bool loadA_all = !(dbg & 4);
            const uint32_t b_tx = (uint32_t)Ntile * bk * 2u, a_tx = 128u * bk * 2u;   // bytes the boxes really carry
            const uint32_t smem0 = smem_u32(smem), fullb = smem_u32(full_bar), emptyb = smem_u32(empty_bar);
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int nt = (int)(t % tiles_n);
                uint32_t mt = t / tiles_n;
                const int tx = (int)(mt % tiles_x); mt /= tiles_x;
                const int ty = (int)(mt % tiles_y);
                const int n = (int)(mt / tiles_y);
                const int x0 = tx * TILE_W, y0 = ty * TILE_H, n0 = nt * Ntile;
                int tap = 0, ky = 0, kx = 0, ch = 0;
                for (int it = 0; it < kit; ++it) {
                    mbar_wait_a(emptyb + 8u * stage, phase ^ 1);
                    if (elect_one()) {
                        const uint32_t sa = smem0 + stage * sbytes, fb = fullb + 8u * stage;
                        // perf triage: dbg & 16 fetches the A boxes of the centre tap only; the other taps multiply stale but finite
                        // shared-memory contents
                        const bool loadA = loadA_all && (!(dbg & 16) || tap == (k * k) / 2);
                        const uint32_t tx_bytes = (three ? 2u : 1u) * (b_tx + (loadA ? a_tx : 0u));
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(tx_bytes) : "memory");
                        // input box origin for this tap.  stride 1: shifted box of the same plane.  stride 2 (pad k/2):
                        // input row 2y+ky-pad lives in polyphase plane py = (ky-pad)&1 at row y + floor((ky-pad)/2).
                        int ax = x0 + kx - pad, ay = y0 + ky - pad, an = n;
                        if (s2) {
                            const int oy = ky - pad, ox = kx - pad;
                            const int py = oy & 1, px = ox & 1;
                            ay = y0 + ((oy - py) >> 1);
                            ax = x0 + ((ox - px) >> 1);
                            an = (py * 2 + px) * Nimg + n;
                        }
                        const int c0 = ch * (int)bk, w0 = tap * Cp + c0;
                        if (loadA) tma_load_4d_a(sa, &map_a_hi, fb, c0, ax, ay, an);
                        tma_load_2d_a(sa + abytes, &map_w_hi, fb, w0, n0);
                        if (three) {
                            if (loadA) tma_load_4d_a(sa + abytes + bbytes, &map_a_lo, fb, c0, ax, ay, an);
                            tma_load_2d_a(sa + 2u * abytes + bbytes, &map_w_lo, fb, w0, n0);
                        }
                    }
                    __syncwarp();
                    if (++ch == chunks_) {
                        ch = 0; ++tap;
                        if (++kx == k) { kx = 0; ++ky; }
                    }
                    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < EPI_WARP0) {
        regs_light();
        // ================= MMA issuers (warps 1 and 2; warp 3 idles) =================
        // The single thread that issues tcgen05.mma is the bottleneck of every layer with a long K loop: ptxas wraps each UTCHMMA
        // issued from divergent code in an ELECT / PLOP3 / BRA.U.ANY sequence, ~120 cycles per MMA against the 64-90 the tensor
        // pipe needs for M128 x N128 x K16 (ncu: producer waiting on free stages, MMA warp never waiting on data, tensor pipe 56 %
        // active).  So TWO warps issue -- warp 1 the even pipeline stages of a tile, warp 2 the odd ones -- each into its own
        // TMEM accumulator (the epilogue adds the two in a fixed order).  Everything read from memory is first made warp-uniform
        // (REDUX results live in uniform registers) so that descriptors and addresses stay in the uniform datapath.
        const int mw = warp - 1;
        if (mw < nmma) {
            MmaLoopArgs A;
            A.total_tiles = (uint32_t)P.total_tiles; A.tiles_n = (uint32_t)P.tiles_n; A.grid = gridDim.x; A.first_tile = blockIdx.x;
            A.Ntile = P.Ntile; A.Cout = p.Cout; A.kiters = kiters; A.stages = P.stages; A.mw = mw;
            A.owner_mode = (nmma == 1 || splitk) ? 0 : 1;
            A.sb16 = (uint32_t)stage_bytes >> 4;
            A.oB = (uint32_t)A_BYTES >> 4; A.oAlo = (uint32_t)(A_BYTES + B_BYTES) >> 4; A.oBlo = (uint32_t)(2 * A_BYTES + B_BYTES) >> 4;
            A.koff = splitk ? (uint32_t)(mw * (ksteps >> 1)) * 2u : 0u;   // this warp's k16 steps (32 B each)
            A.full_bar = smem_u32(full_bar); A.empty_bar = smem_u32(empty_bar);
            A.tmem_full = smem_u32(tmem_full); A.tmem_empty = smem_u32(tmem_empty);
            A.tmem_base = tmem_base;
            A.idesc_fmt = P.in_f16 ? 0u : ((1u << 7) | (1u << 10));
            A.desc0 = make_kmajor_desc(smem_u32(smem), BLOCK_K);
            A.skip_mma = (P.dbg & 2) != 0;
            // opaque: keep the loop invariants in registers instead of re-reading the constant bank at every use
            opaque(A.total_tiles); opaque(A.tiles_n); opaque(A.Ntile); opaque(A.Cout); opaque(A.kiters); opaque(A.stages);
            opaque(A.owner_mode); opaque(A.sb16); opaque(A.oB); opaque(A.oAlo); opaque(A.oBlo); opaque(A.koff); opaque(A.idesc_fmt);
            opaque(A.skip_mma); asm volatile("" : "+l"(A.desc0));
            const int spw = splitk ? (ksteps >> 1) : ksteps;
            if (P.passes == 3) {
                if (spw == 4) mma_issue_loop<3, 4>(A);
                else if (spw == 2) mma_issue_loop<3, 2>(A);
                else mma_issue_loop<3, 1>(A);
            } else {
                if (spw == 4) mma_issue_loop<1, 4>(A);
                else if (spw == 2) mma_issue_loop<1, 2>(A);
                else mma_issue_loop<1, 1>(A);
            }
        }
    } else {
        regs_epilogue();
        // ================= epilogue: TMEM -> registers -> (warp-private transposition slab) -> fused element-wise -> global
        const int q = warp & 3;            // TMEM lane quarter this warp may access (hardware: warp_id % 4)
        const int jsub = (warp - EPI_WARP0) >> 2;  // which WCOLS-column slice of the accumulator this warp drains
        const uint32_t slab = smem_u32(stg + (warp - EPI_WARP0) * SLAB_FLOATS);
        const EpiRegs er = make_epi_regs(p, P.dbg, DUAL);
        int act = p.act, epi = p.epi, store = p.store;
        opaque(act); opaque(epi); opaque(store);
        uint32_t tiles_n = (uint32_t)P.tiles_n, tiles_x = (uint32_t)P.tiles_x, tiles_y = (uint32_t)P.tiles_y;
        opaque(tiles_n); opaque(tiles_x); opaque(tiles_y);
        int Ntile = P.Ntile, wcw = P.wcw, rvflag = P.rv;
        opaque(Ntile); opaque(wcw); opaque(rvflag);
        uint32_t local = 0;
        for (uint32_t t = blockIdx.x; t < (uint32_t)P.total_tiles; t += gridDim.x, ++local) {   // total_tiles < 2^31 (host-checked)
            const int nt = (int)(t % tiles_n);
            uint32_t mt = t / tiles_n;
            const int tx = (int)(mt % tiles_x); mt /= tiles_x;
            const int ty = (int)(mt % tiles_y);
            const int n = (int)(mt / tiles_y);
            const int x0 = tx * TILE_W, y0 = ty * TILE_H, n0 = nt * Ntile;
            int ncols = er.Cout - n0;
            if (ncols > Ntile) ncols = Ntile;
            int wcols = ncols - wcw * jsub;
            if (wcols > wcw) wcols = wcw;
            const uint32_t ab = local & 1;
            const uint32_t taddr = tmem_base + ab * 256 + ((uint32_t)(q * 32) << 16) + (uint32_t)(wcw * jsub);
            if constexpr (VEC) {
                bool done = false;
                if constexpr (!DUAL) {
                    if (rvflag) {
                        epilogue_tile_rv<ACT, EPI, false, TILE_W, FEAT>(er, act, epi, sbias, taddr, &tmem_full[ab], (local >> 1) & 1, &tmem_empty[ab], n, x0,
                                                                  y0, n0 + wcw * jsub, wcols, q, lane);
                        done = true;
                    }
                }
                if (!done)
                    epilogue_tile_vec<ACT, EPI, DUAL, TILE_W>(er, act, epi, slab, sbias, taddr, &tmem_full[ab], (local >> 1) & 1, &tmem_empty[ab], n,
                                                              x0, y0, n0 + wcw * jsub, wcols, q, lane);
            }
            else
                epilogue_tile_rows<ACT, EPI, DUAL, TILE_W>(er, act, epi, store, sbias, taddr, &tmem_full[ab], (local >> 1) & 1, &tmem_empty[ab], n, x0, y0,
                                             n0 + wcw * jsub, wcols, q, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}


// ---------------------------------------------------------------- halo-tile kernel for 3x3 stride-1 layers
// The tap-by-tap kernel above re-fetches a shifted 128-pixel box for each of the 9 taps: per 128 x 128 tile of a 128 -> 128 layer
// it moves 1152 KB (bf16x3) from L2 into shared memory for 4.2 MFLOP-equivalents of tensor work, and ncu shows BOTH engines pinned
// at the same 13.7 TB/s of L2 -> SM traffic (l1tex__m_xbar2l1tex_read_bytes: 39.1 GB / 2.86 ms at 3 passes, 19.8 GB / 1.45 ms at
// 1 pass) -- the layer is bound by the L2 fabric, not by the tensor pipe or the issuing threads.  Here the A operand of a tile is
// fetched ONCE per 64-channel chunk as a (16+2) x (8+2)-pixel halo box (180 rows of 128 B; out-of-image rows zero-filled by TMA =
// the conv padding), and each tap's A tile is a VIEW of it: the tile is 16 image rows x 8 pixels, so the 8 rows of every UMMA
// core-matrix group are the 8 pixels of one image row (consecutive 128-byte rows of the box) and groups are one box row apart
// (stride-byte-offset = 10 * 128 B); the tap (ky, kx) only moves the descriptor's start address by (ky * 10 + kx) * 128 B.  The
// 128B swizzle XORs address bits [4,7) with bits [7,10) of the ABSOLUTE shared-memory address on both the TMA write and the UMMA
// read (the box buffer is 1024-byte aligned), so a start address that is not a multiple of 1024 stays consistent with a ZERO
// base-offset field (measured: every conv parity test passes with base offset 0 and fails with (start >> 7) & 7 in the field).
// A traffic drops 6.4x; B (weights) is unchanged.
// Loop order is chunk-major (A box outermost), one issuing warp, one 128-column accumulator per tile (double-buffered).
constexpr int HT_H = 16, HT_W = 8;                          // tile: 16 image rows x 8 pixels
constexpr int HB_H = HT_H + 2, HB_W = HT_W + 2;             // halo box
// The same construction serves the few-channel layers (K chunk bk = 16 / 32 channels: box rows of 32 / 64 bytes, 32- / 64-byte
// swizzle): the packed-Bayer ingest convs and the condition UNet used to run 9 single-MMA pipeline stages per tile.  Where all
// nine weight tiles of a chunk fit in one stage (small bk or few output channels) they are fetched on ONE barrier ("taps per
// stage" = 9) and the issuer runs the chunk's 9 x bk/16 x passes MMAs back to back: 2 barrier hand-shakes per tile and chunk
// instead of 10.
__host__ __device__ constexpr int halo_tx(int bk) { return HB_H * HB_W * bk * 2; }                    // bytes one box carries
__host__ __device__ constexpr int halo_bytes(int bk) { return (halo_tx(bk) + 1023) & ~1023; }         // buffers stay 1024-byte aligned

__device__ __forceinline__ uint64_t make_halo_desc(uint32_t smem_addr, uint32_t row_shift, int bk, int base_off_mode) {
    const uint32_t start = smem_addr + row_shift * (uint32_t)(bk * 2);
    uint64_t d = 0;
    d |= (uint64_t)((start & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((uint32_t)(HB_W * bk * 2) >> 4) << 32;  // 8-row groups are one box row (10 pixels) apart
    d |= (uint64_t)1 << 46;
    if (base_off_mode) d |= (uint64_t)((start >> 7) & 7u) << 49;
    d |= (uint64_t)(bk == 64 ? 2 : (bk == 32 ? 4 : 6)) << 61;   // SWIZZLE_128B / 64B / 32B
    return d;
}

// issue loop of the halo kernel for one (passes, k16 steps per tap) combination
struct HaloLoopArgs {
    uint32_t total_tiles, tiles_n, grid, first_tile;
    int Ntile, Cout, nch, stages, na, tps, bk, skip_mma, bo_mode;
    uint32_t abytes, bsbytes, bbytes, halo_b;
    uint32_t a0, afull, aempty, bfull, bempty, tfull, tempty, tmem_base, idesc_fmt;
    uint64_t bdesc0;
};
// TPS (taps per weight stage) and the chunk width are compile-time: with them at run time the descriptor arithmetic of every tap
// (tap / 3, tap % 3, shifts by the chunk width, the swizzle mode) sat in the single issuing thread's way -- the fp16 x1 tail conv
// (4 MMAs per stage) went from 1.106 to 1.27 ms.
template <int PASSES, int KS, int TPS>
__device__ __forceinline__ void halo_issue_loop(const HaloLoopArgs& A) {
    constexpr int BKC = KS * 16;
    constexpr int groups = 9 / TPS;
    uint32_t ab = 0, aph = 0, st = 0, bph = 0, local = 0;
    for (uint32_t t = A.first_tile; t < A.total_tiles; t += A.grid, ++local) {
        const int n0 = (int)(t % A.tiles_n) * A.Ntile;
        int nact = A.Cout - n0;
        if (nact > A.Ntile) nact = A.Ntile;
        nact = (nact + 15) & ~15;
        const uint32_t idesc = (1u << 4) | A.idesc_fmt | ((uint32_t)(nact >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t acb = local & 1;
        mbar_wait_a(A.tempty + 8u * acb, ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = A.tmem_base + acb * 256;
        uint32_t acc = 0;
        for (int c = 0; c < A.nch; ++c) {
            mbar_wait_a(A.afull + 8u * ab, aph);
            const uint32_t abase = A.a0 + ab * A.abytes;
            // descriptors of tap 0; a tap moves the A start address by a compile-time number of 16-byte units (no carry out of the
            // 14-bit field: shared memory is < 256 KB)
            const uint64_t a_hi0 = make_halo_desc(abase, 0, BKC, A.bo_mode);
            const uint64_t a_lo0 = make_halo_desc(abase + A.halo_b, 0, BKC, A.bo_mode);
#pragma unroll
            for (int sg = 0; sg < groups; ++sg) {
                mbar_wait_a(A.bfull + 8u * st, bph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bst = A.bdesc0 + (uint64_t)((st * A.bsbytes) >> 4);
#pragma unroll
                    for (int tt = 0; tt < TPS; ++tt) {
                        constexpr int unit = BKC * 2 / 16;                 // 16-byte units per box pixel
                        const int tp = sg * TPS + tt;
                        const uint32_t shift = (uint32_t)(((tp / 3) * HB_W + (tp % 3)) * unit);
                        uint64_t a_hi = a_hi0 + shift, a_lo = a_lo0 + shift;
                        if (A.bo_mode) {                                   // triage only
                            a_hi = make_halo_desc(abase, (tp / 3) * HB_W + (tp % 3), BKC, 1);
                            a_lo = make_halo_desc(abase + A.halo_b, (tp / 3) * HB_W + (tp % 3), BKC, 1);
                        }
                        const uint64_t b_hi = bst + (uint64_t)((tt * A.bbytes) >> 4);
                        const uint64_t b_lo = b_hi + (uint64_t)((TPS * A.bbytes) >> 4);
                        if (!A.skip_mma) issue_stage<PASSES, KS>(tmem_d, a_hi, b_hi, a_lo, b_lo, idesc, acc);
                        acc = 1;
                    }
                    umma_commit_a(A.bempty + 8u * st);
                }
                __syncwarp();
                acc = 1;
                if (++st == (uint32_t)A.stages) { st = 0; bph ^= 1; }
            }
            if (elect_one()) umma_commit_a(A.aempty + 8u * ab);     // the box is free once every MMA reading it has retired
            __syncwarp();
            if (++ab == (uint32_t)A.na) { ab = 0; aph ^= 1; }
        }
        if (elect_one()) umma_commit_a(A.tfull + 8u * acb);
        __syncwarp();
    }
}

template <int ACT, int EPI, bool VEC, int FEAT>
__global__ void RCN_TC_BOUNDS
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const rcn_conv_desc& p = P.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int planes = P.passes == 3 ? 2 : 1;
    const int BK = P.bk, TPS = P.tps;
    const int abuf_bytes = planes * halo_bytes(BK), bstage_bytes = planes * TPS * P.b_bytes;
    uint8_t* bsm = smem + (size_t)P.na * abuf_bytes;
    uint8_t* tail = bsm + (size_t)P.stages * bstage_bytes;
    float* stg = reinterpret_cast<float*>(tail);
    float* sbias_mem = reinterpret_cast<float*>(tail + STG_BYTES);
    uint64_t* afull = reinterpret_cast<uint64_t*>(tail + STG_BYTES + BIAS_BYTES);
    uint64_t* aempty = afull + P.na;
    uint64_t* bfull = aempty + P.na;
    uint64_t* bempty = bfull + P.stages;
    uint64_t* tmem_full = bempty + P.stages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int chunks = P.Cp / BK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < P.na; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < P.stages; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.bias) {
        for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) sbias_mem[i] = __ldg(p.bias + i);
    }
    const uint32_t sbias = p.bias ? smem_u32(sbias_mem) : 0u;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // loop invariants of the producer / issuer loops, held in registers (declared inside each role's branch: values hoisted
    // above the role dispatch stay live in the epilogue warps too and pushed their code into local-memory spills)
#define RCN_HALO_INVARIANTS                                                                                                              \
    uint32_t total_tiles = (uint32_t)P.total_tiles, tiles_n = (uint32_t)P.tiles_n, tiles_x = (uint32_t)P.tiles_x,                         \
             tiles_y = (uint32_t)P.tiles_y;                                                                                              \
    int Ntile = P.Ntile, Cp = P.Cp, nch = chunks, stages = P.stages, na = P.na, three = (P.passes == 3), dbg = P.dbg, Cout = p.Cout;     \
    uint32_t abytes = (uint32_t)abuf_bytes, bsbytes = (uint32_t)bstage_bytes, bbytes = (uint32_t)P.b_bytes;                               \
    opaque(total_tiles); opaque(tiles_n); opaque(tiles_x); opaque(tiles_y); opaque(Ntile); opaque(Cp); opaque(nch); opaque(stages);      \
    opaque(na); opaque(three); opaque(dbg); opaque(Cout); opaque(abytes); opaque(bsbytes); opaque(bbytes);                               \
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(bsm);                                                                              \
    const uint32_t afull_a = smem_u32(afull), aempty_a = smem_u32(aempty), bfull_a = smem_u32(bfull), bempty_a = smem_u32(bempty);       \
    (void)tiles_x; (void)tiles_y; (void)Cp; (void)Cout; (void)dbg; (void)a0; (void)b0; (void)afull_a; (void)aempty_a; (void)bfull_a;     \
    (void)bempty_a; (void)abytes; (void)bsbytes; (void)bbytes; (void)na; (void)stages; (void)nch; (void)three; (void)Ntile; (void)tiles_n

    if (warp == 0) {
        regs_light();
        // ================= TMA producer: per tile and chunk one halo box (hi [+ lo]), then the 9 weight tiles of that chunk
        RCN_HALO_INVARIANTS;
        int bk = BK, tps = TPS;
        opaque(bk); opaque(tps);
        const uint32_t halo_b = (uint32_t)halo_bytes(bk);
        const uint32_t a_tx = (uint32_t)(three ? 2 : 1) * (uint32_t)halo_tx(bk);
        const uint32_t b_tx = (uint32_t)(three ? 2 : 1) * (uint32_t)tps * (uint32_t)Ntile * (uint32_t)(bk * 2);
        const int groups = 9 / tps;
        uint32_t ab = 0, aph = 0, st = 0, bph = 0;
        for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int nt = (int)(t % tiles_n);
            uint32_t mt = t / tiles_n;
            const int tx = (int)(mt % tiles_x); mt /= tiles_x;
            const int ty = (int)(mt % tiles_y);
            const int n = (int)(mt / tiles_y);
            const int x0 = tx * HT_W, y0 = ty * HT_H, n0 = nt * Ntile;
            for (int c = 0; c < nch; ++c) {
                mbar_wait_a(aempty_a + 8u * ab, aph ^ 1);
                if (elect_one()) {
                    const uint32_t fb = afull_a + 8u * ab, dst = a0 + ab * abytes;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(a_tx) : "memory");
                    tma_load_4d_a(dst, &map_a_hi, fb, c * bk, x0 - 1, y0 - 1, n);
                    if (three) tma_load_4d_a(dst + halo_b, &map_a_lo, fb, c * bk, x0 - 1, y0 - 1, n);
                }
                __syncwarp();
                if (++ab == (uint32_t)na) { ab = 0; aph ^= 1; }
                int tap = 0;
                for (int sg = 0; sg < groups; ++sg) {
                    mbar_wait_a(bempty_a + 8u * st, bph ^ 1);
                    if (elect_one()) {
                        const uint32_t fb = bfull_a + 8u * st, dst = b0 + st * bsbytes;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(b_tx) : "memory");
                        for (int tt = 0; tt < tps; ++tt) {       // stage layout: hi tiles of the taps, then lo tiles
                            const int w0 = (tap + tt) * Cp + c * bk;
                            tma_load_2d_a(dst + (uint32_t)tt * bbytes, &map_w_hi, fb, w0, n0);
                            if (three) tma_load_2d_a(dst + (uint32_t)(tps + tt) * bbytes, &map_w_lo, fb, w0, n0);
                        }
                    }
                    __syncwarp();
                    tap += tps;
                    if (++st == (uint32_t)stages) { st = 0; bph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        regs_light();
        // ================= MMA issuer (one warp, converged, tcgen05 instructions behind elect.sync)
        RCN_HALO_INVARIANTS;
        HaloLoopArgs A;
        A.total_tiles = total_tiles; A.tiles_n = tiles_n; A.grid = gridDim.x; A.first_tile = blockIdx.x;
        A.Ntile = Ntile; A.Cout = Cout; A.nch = nch; A.stages = stages; A.na = na; A.tps = TPS; A.bk = BK;
        A.skip_mma = (dbg & 2) != 0;
        A.bo_mode = (dbg & 64) ? 1 : 0;   // triage only: 1 sets the descriptor's base-offset field (WRONG on B200, see above)
        A.abytes = abytes; A.bsbytes = bsbytes; A.bbytes = bbytes; A.halo_b = (uint32_t)halo_bytes(BK);
        A.a0 = a0; A.afull = afull_a; A.aempty = aempty_a; A.bfull = bfull_a; A.bempty = bempty_a;
        A.tfull = smem_u32(tmem_full); A.tempty = smem_u32(tmem_empty); A.tmem_base = tmem_base;
        A.idesc_fmt = P.in_f16 ? 0u : ((1u << 7) | (1u << 10));
        A.bdesc0 = make_kmajor_desc(b0, BK);
        opaque(A.tps); opaque(A.bk); opaque(A.skip_mma); opaque(A.bo_mode); opaque(A.halo_b); opaque(A.idesc_fmt);
        asm volatile("" : "+l"(A.bdesc0));
        if (TPS == 9) {
            if (three) {
                if (BK == 64) halo_issue_loop<3, 4, 9>(A);
                else if (BK == 32) halo_issue_loop<3, 2, 9>(A);
                else halo_issue_loop<3, 1, 9>(A);
            } else {
                if (BK == 64) halo_issue_loop<1, 4, 9>(A);
                else if (BK == 32) halo_issue_loop<1, 2, 9>(A);
                else halo_issue_loop<1, 1, 9>(A);
            }
        } else if (three) {
            if (BK == 64) halo_issue_loop<3, 4, 1>(A);
            else if (BK == 32) halo_issue_loop<3, 2, 1>(A);
            else halo_issue_loop<3, 1, 1>(A);
        } else {
            if (BK == 64) halo_issue_loop<1, 4, 1>(A);
            else if (BK == 32) halo_issue_loop<1, 2, 1>(A);
            else halo_issue_loop<1, 1, 1>(A);
        }
    } else if (warp < EPI_WARP0) {
        regs_light();      // idle warps of the first warp group
    } else {
        regs_epilogue();
        // ================= epilogue (same code as the tap-by-tap kernel, tile geometry 16 x 8)
        const int q = warp & 3;
        const int jsub = (warp - EPI_WARP0) >> 2;
        const uint32_t slab = smem_u32(stg + (warp - EPI_WARP0) * SLAB_FLOATS);
        const EpiRegs er = make_epi_regs(p, P.dbg, false);
        int act = p.act, epi = p.epi, store = p.store;
        opaque(act); opaque(epi); opaque(store);
        int wcw = P.wcw, rvflag = P.rv, Ntile = P.Ntile;
        uint32_t total_tiles = (uint32_t)P.total_tiles, tiles_n = (uint32_t)P.tiles_n, tiles_x = (uint32_t)P.tiles_x, tiles_y = (uint32_t)P.tiles_y;
        opaque(wcw); opaque(rvflag); opaque(Ntile); opaque(total_tiles); opaque(tiles_n); opaque(tiles_x); opaque(tiles_y);
        uint32_t local = 0;
        for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++local) {
            const int nt = (int)(t % tiles_n);
            uint32_t mt = t / tiles_n;
            const int tx = (int)(mt % tiles_x); mt /= tiles_x;
            const int ty = (int)(mt % tiles_y);
            const int n = (int)(mt / tiles_y);
            const int x0 = tx * HT_W, y0 = ty * HT_H, n0 = nt * Ntile;
            int ncols = er.Cout - n0;
            if (ncols > Ntile) ncols = Ntile;
            int wcols = ncols - wcw * jsub;
            if (wcols > wcw) wcols = wcw;
            const uint32_t acb = local & 1;
            const uint32_t taddr = tmem_base + acb * 256 + ((uint32_t)(q * 32) << 16) + (uint32_t)(wcw * jsub);
            if constexpr (VEC) {
                if (rvflag)
                    epilogue_tile_rv<ACT, EPI, false, HT_W, FEAT>(er, act, epi, sbias, taddr, &tmem_full[acb], (local >> 1) & 1, &tmem_empty[acb], n,
                                                            x0, y0, n0 + wcw * jsub, wcols, q, lane);
                else
                    epilogue_tile_vec<ACT, EPI, false, HT_W>(er, act, epi, slab, sbias, taddr, &tmem_full[acb], (local >> 1) & 1, &tmem_empty[acb], n,
                                                             x0, y0, n0 + wcw * jsub, wcols, q, lane);
            }
            else
                epilogue_tile_rows<ACT, EPI, false, HT_W>(er, act, epi, store, sbias, taddr, &tmem_full[acb], (local >> 1) & 1, &tmem_empty[acb],
                                                          n, x0, y0, n0 + wcw * jsub, wcols, q, lane);
        }
    }
#undef RCN_HALO_INVARIANTS
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}


// ---------------------------------------------------------------- operand preparation
// fp32 NHWC (ld) -> bf16 hi / lo planes (npix, Cp), zero-padded channels; optional x*x (GDN)
__global__ void split_bf16_kernel(const float* __restrict__ x, int ldx, long long npix, int C, int Cp, int square, int f16,
                                  uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
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
        uint16_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = to_plane(v[j], f16);
            l[j] = to_plane(v[j] - from_plane(h[j], f16), f16);
        }
        *reinterpret_cast<uint2*>(hi + pix * Cp + c4) = *reinterpret_cast<uint2*>(h);
        if (lo) *reinterpret_cast<uint2*>(lo + pix * Cp + c4) = *reinterpret_cast<uint2*>(l);
    }
}

// polyphase split for stride-2 convs: out[((py*2+px)*N + n), i, j, c] = x[n, 2i+py, 2j+px, c]
__global__ void split_bf16_s2_kernel(const float* __restrict__ x, int ldx, int N, int H, int W, int C, int Cp, int f16,
                                     uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
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
        uint16_t hh[4], ll[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hh[j] = to_plane(v[j], f16);
            ll[j] = to_plane(v[j] - from_plane(hh[j], f16), f16);
        }
        const long long plane = (long long)((h & 1) * 2 + (w & 1)) * N + n;
        const long long o = ((plane * H2 + (h >> 1)) * W2 + (w >> 1)) * Cp + c4;
        *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<uint2*>(hh);
        if (lo) *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<uint2*>(ll);
    }
}

// OIHW fp32 -> [Cout][k*k][Cp] bf16 hi / lo (K-major rows for the B operand)
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int Cp, int ps_perm, int f16,
                                      uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const long long total = (long long)Cout * k * k * Cp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        long long t = i / Cp;
        const int tap = (int)(t % (k * k));
        int co = (int)(t / (k * k));
        if (ps_perm) {   // packed row 64g + 16s + cc  <-  conv channel 64g + 4cc + s (pixel-shuffle sub-pixel s, shuffled channel cc)
            const int g = co >> 6, rr = co & 63;
            co = (g << 6) + ((rr & 15) << 2) + (rr >> 4);
        }
        const float v = (c < Cin) ? w[((long long)co * Cin + c) * k * k + tap] : 0.f;
        const uint16_t h = to_plane(v, f16);
        hi[i] = h;
        if (lo) lo[i] = to_plane(v - from_plane(h, f16), f16);
    }
}


// planes (N,H,W,Cp) with pixel stride ldp >= Cp elements (a channel slice of a wider plane buffer when ldp > Cp)
bool make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int Cp, int ldp, int bk, int box_w = TILE_W, int box_h = TILE_H) {
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ldp * 2, (cuuint64_t)W * ldp * 2, (cuuint64_t)H * W * ldp * 2};
    cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}



typedef void (*TcKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);


template <int ACT, int EPI, bool VEC, bool DUAL, int FEAT>
TcKernel tc_variant_d() {
    static bool attr_set[MAX_DEVICES] = {};   // function attributes are per device: once per (instantiation, device)
    TcKernel k = conv_tc_kernel<ACT, EPI, VEC, DUAL, FEAT>;
    const int dev = current_device();
    if (!attr_set[dev]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set[dev] = true;
    }
    return k;
}

// perf / fault triage knobs, read from the environment ONCE per process (not on every launch)
struct TcDebugEnv {
    int dbg = 0, nmma = 0, stages = 0, halo = 1, rv = 1;
    TcDebugEnv() {
        const char* h = getenv("RCN_TC_HALO");       // 0: keep 3x3 stride-1 layers on the tap-by-tap kernel; 64: halo kernel for 64-channel chunks only
        if (h) halo = atoi(h);
        const char* rvs = getenv("RCN_TC_RV");       // 0: transposing epilogue everywhere
        if (rvs) rv = atoi(rvs) != 0;
        const char* e = getenv("RCN_TC_DEBUG");
        dbg = e ? atoi(e) : 0;
        const char* nm = getenv("RCN_TC_NMMA");      // force the number of MMA-issuing warps
        if (nm && (atoi(nm) == 1 || atoi(nm) == 2)) nmma = atoi(nm);
        const char* st = getenv("RCN_TC_STAGES");
        if (st && atoi(st) >= 2) stages = atoi(st);
    }
};
const TcDebugEnv& tc_debug_env() {
    static const TcDebugEnv env;
    return env;
}


template <int ACT, int EPI, bool VEC, int FEAT>
TcKernel tc_variant_h() {
    static bool attr_set[MAX_DEVICES] = {};
    TcKernel k = conv_tc_halo_kernel<ACT, EPI, VEC, FEAT>;
    const int dev = current_device();
    if (!attr_set[dev]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set[dev] = true;
    }
    return k;
}

// mode: 0 = tap-by-tap kernel, one issuing warp; 1 = tap-by-tap, two issuing warps (triage); 2 = halo-tile kernel
template <int ACT, int EPI, bool VEC, int FEAT>
TcKernel tc_variant(int mode) {
    if (mode == 2) return tc_variant_h<ACT, EPI, VEC, FEAT>();
    if constexpr (FEAT == 0) {
        if (mode == 1) return tc_variant_d<ACT, EPI, VEC, true, 0>();
    }
    return tc_variant_d<ACT, EPI, VEC, false, FEAT>();
}

constexpr int FEAT_ALL = 0xF;

// (act, epi, 16-byte path?, kernel mode, rare epilogue options needed?) -> instantiation; combinations the path never uses share the
// generic (-1, -1) variant
TcKernel select_kernel(int act, int epi, bool vec, int mode, bool full) {
    if (vec && full) {
        // the (activation, combinator) pairs that occur with NCHW aux / store, squared planes, polyphase planes or Res_GFM modulation
        if (epi == RCN_EPI_NONE) {
            switch (act) {
                case RCN_ACT_NONE: return tc_variant<RCN_ACT_NONE, 0, true, FEAT_ALL>(mode);
                case RCN_ACT_RELU: return tc_variant<RCN_ACT_RELU, 0, true, FEAT_ALL>(mode);
                case RCN_ACT_LRELU: return tc_variant<RCN_ACT_LRELU, 0, true, FEAT_ALL>(mode);
                default: break;
            }
        }
        if (act == RCN_ACT_NONE && epi == RCN_EPI_MUL_AUXP1) return tc_variant<RCN_ACT_NONE, RCN_EPI_MUL_AUXP1, true, FEAT_ALL>(mode);
        return tc_variant<-1, -1, true, FEAT_ALL>(mode);
    }
    if (vec) {
        if (epi == RCN_EPI_NONE) {
            switch (act) {
                case RCN_ACT_NONE: return tc_variant<RCN_ACT_NONE, 0, true, 0>(mode);
                case RCN_ACT_RELU: return tc_variant<RCN_ACT_RELU, 0, true, 0>(mode);
                case RCN_ACT_LRELU: return tc_variant<RCN_ACT_LRELU, 0, true, 0>(mode);
                case RCN_ACT_GELU: return tc_variant<RCN_ACT_GELU, 0, true, 0>(mode);
                case RCN_ACT_SIGMOID: return tc_variant<RCN_ACT_SIGMOID, 0, true, 0>(mode);
                case RCN_ACT_HALF_TANH: return tc_variant<RCN_ACT_HALF_TANH, 0, true, 0>(mode);
                case RCN_ACT_HSWISH: return tc_variant<RCN_ACT_HSWISH, 0, true, 0>(mode);
                default: return tc_variant<-1, -1, true, 0>(mode);
            }
        }
        if (act == RCN_ACT_NONE) {
            switch (epi) {
                case RCN_EPI_GDN: return tc_variant<RCN_ACT_NONE, RCN_EPI_GDN, true, 0>(mode);
                case RCN_EPI_IGDN: return tc_variant<RCN_ACT_NONE, RCN_EPI_IGDN, true, 0>(mode);
                case RCN_EPI_MUL_AUXP1: return tc_variant<RCN_ACT_NONE, RCN_EPI_MUL_AUXP1, true, 0>(mode);
                case RCN_EPI_MULP1_AUX: return tc_variant<RCN_ACT_NONE, RCN_EPI_MULP1_AUX, true, 0>(mode);
                case RCN_EPI_SIGMOID_GATE: return tc_variant<RCN_ACT_NONE, RCN_EPI_SIGMOID_GATE, true, 0>(mode);
                default: break;
            }
        }
        return tc_variant<-1, -1, true, 0>(mode);
    }
    if (epi == RCN_EPI_NONE && act == RCN_ACT_NONE) return tc_variant<RCN_ACT_NONE, 0, false, 0>(mode);
    if (epi == RCN_EPI_NONE && act == RCN_ACT_CLAMP01) return tc_variant<RCN_ACT_CLAMP01, 0, false, 0>(mode);
    return tc_variant<-1, -1, false, 0>(mode);
}

}  // namespace
}  // namespace rcn

using namespace rcn;

extern "C" int rcn_tc_prof(unsigned long long* out16, int reset) {
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_tcprof, z, sizeof(z)); return RCN_OK; }
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out16, g_tcprof, 16 * sizeof(unsigned long long)) == cudaSuccess ? RCN_OK : RCN_ERR_CUDA;
}

static inline bool cp_ok(int Cp) { return Cp == 16 || Cp == 32 || (Cp > 0 && Cp % 64 == 0); }

extern "C" int rcn_split_bf16(const float* x, int ldx, long long npix, int C, int Cp, int square, int fmt, void* hi, void* lo, void* stream) {
    RCN_CHECK_ARG(x && hi && npix > 0 && C > 0 && Cp >= C && cp_ok(Cp) && (fmt == RCN_PLANE_BF16 || fmt == RCN_PLANE_F16), "rcn_split_bf16: bad arguments");
    const long long total = npix * (Cp / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    split_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, npix, C, Cp, square, fmt, (uint16_t*)hi, (uint16_t*)lo);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_split_bf16");
    return RCN_OK;
}

// stride-2 operand: the four polyphase planes (py,px) of x, each (N, H/2, W/2, Cp), stacked on the batch axis
extern "C" int rcn_split_bf16_s2(const float* x, int ldx, int N, int H, int W, int C, int Cp, int fmt, void* hi, void* lo, void* stream) {
    RCN_CHECK_ARG(x && hi && N > 0 && C > 0 && Cp >= C && cp_ok(Cp) && (fmt == RCN_PLANE_BF16 || fmt == RCN_PLANE_F16), "rcn_split_bf16_s2: bad arguments");
    RCN_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "rcn_split_bf16_s2: H and W must be even");
    const long long total = (long long)N * H * W * (Cp / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    split_bf16_s2_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, N, H, W, C, Cp, fmt, (uint16_t*)hi, (uint16_t*)lo);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_split_bf16_s2");
    return RCN_OK;
}

extern "C" int rcn_pack_conv_weight_tc(const float* w_oihw, int Cout, int Cin, int k, int Cp, int ps_perm, int fmt, void* hi, void* lo, void* stream) {
    RCN_CHECK_ARG(w_oihw && hi && Cp >= Cin && cp_ok(Cp) && (fmt == RCN_PLANE_BF16 || fmt == RCN_PLANE_F16), "rcn_pack_conv_weight_tc: bad arguments");
    RCN_CHECK_ARG(!ps_perm || Cout % 64 == 0, "rcn_pack_conv_weight_tc: the pixel-shuffle row order needs Cout %% 64 == 0");
    const long long total = (long long)Cout * k * k * Cp;
    long long blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    pack_weight_tc_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, k, Cp, ps_perm, fmt, (uint16_t*)hi, (uint16_t*)lo);
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
        RCN_CHECK_ARG((d->store == RCN_STORE_NHWC || ps2) && d->Cp_out >= cs && (d->Cp_out % 8) == 0 && (cs % 4) == 0 && (!ps2 || d->epi == RCN_EPI_NONE) &&
                          (!d->y || ((d->ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(d->y) & 15) == 0)) &&
                          (!d->res || ((d->ldres & 3) == 0 && (reinterpret_cast<uintptr_t>(d->res) & 15) == 0)) &&
                          (d->epi == RCN_EPI_NONE || ((d->ldaux & 3) == 0 && (reinterpret_cast<uintptr_t>(d->aux) & 15) == 0)) &&
                          (!d->bias || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0),
                      "rcn_conv2d_tc: operand-plane emission needs an NHWC / pixel-shuffle store, a plane pixel stride Cp_out >= the "
                      "stored channels (multiple of 8) and 16-byte aligned tensors");
    }
    // the polyphase layout is that of the OUTPUT map (the epilogue works in output geometry): a stride-2 layer feeding another one
    // (CondNet2 / CondNet3 of the condition module, raw2bit.py:842-849) needs H, W divisible by 4
    RCN_CHECK_ARG(!d->planes_s2 || (d->y_hi && d->store == RCN_STORE_NHWC && d->H % (2 * d->stride) == 0 && d->W % (2 * d->stride) == 0),
                  "rcn_conv2d_tc: polyphase plane emission needs an NHWC layer whose output map has even H and W");
    RCN_CHECK_ARG(passes == 1 || (passes == 3 && x_lo && w_lo), "rcn_conv2d_tc: passes must be 1 or 3 (3 needs the lo planes)");
    RCN_CHECK_ARG(d->k == 1 || d->k == 3, "rcn_conv2d_tc: kernel size %d unsupported", d->k);
    RCN_CHECK_ARG(d->Cout <= BIAS_MAX, "rcn_conv2d_tc: Cout %d > %d unsupported", d->Cout, BIAS_MAX);
    RCN_CHECK_ARG(d->stride == 1 || d->stride == 2, "rcn_conv2d_tc: stride %d unsupported", d->stride);
    RCN_CHECK_ARG(d->stride == 1 || (d->H % 2 == 0 && d->W % 2 == 0), "rcn_conv2d_tc: stride 2 needs even H and W");
    RCN_CHECK_ARG(cp_ok(Cp) && Cp >= d->Cin, "rcn_conv2d_tc: Cp must be 16, 32 or a multiple of 64, and >= Cin");
    const int ldp = d->ldp_in > 0 ? d->ldp_in : Cp;
    RCN_CHECK_ARG(ldp >= Cp && ldp % 8 == 0 && (reinterpret_cast<uintptr_t>(x_hi) & 15) == 0 && (!x_lo || (reinterpret_cast<uintptr_t>(x_lo) & 15) == 0),
                  "rcn_conv2d_tc: operand planes need a pixel stride >= Cp that is a multiple of 8 and 16-byte aligned bases");
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
    P.in_f16 = d->in_fmt == RCN_PLANE_F16;
    RCN_CHECK_ARG((d->in_fmt == RCN_PLANE_BF16 || d->in_fmt == RCN_PLANE_F16) && (d->out_fmt == RCN_PLANE_BF16 || d->out_fmt == RCN_PLANE_F16),
                  "rcn_conv2d_tc: unknown operand plane format");
    int nt = d->Cout >= 128 ? 128 : ((d->Cout + 15) & ~15);
    P.Ntile = nt;
    P.tiles_x = (P.d.W + TILE_W - 1) / TILE_W;
    P.tiles_y = (P.d.H + TILE_H - 1) / TILE_H;
    const int bk = Cp >= 64 ? 64 : Cp;
    P.bk = bk;
    const TcDebugEnv& env = tc_debug_env();
    P.dbg = env.dbg;
    // One issuing warp: with the lean issue loops (converged warp + elect.sync) a second issuer never helped any layer and cost
    // the epilogue a second accumulator to add; RCN_TC_NMMA=2 keeps the dual-issuer variant reachable for triage.
    P.nmma = env.nmma == 2 ? MMA_WARPS : 1;
    P.wcw = (!ps && nt >= 64 && nt % 32 == 0) ? nt / 2 : WCOLS;
    P.a_bytes = 128 * bk * 2;                          // 16 / 8 / 4 KB
    P.b_bytes = (nt * bk * 2 + 1023) & ~1023;
    // halo-tile kernel: 3x3, stride 1 (RCN_TC_HALO=0: off; =64: 64-channel K chunks only)
    P.halo = (env.halo && d->k == 3 && d->stride == 1 && (bk == 64 || env.halo != 64)) ? 1 : 0;
    P.na = 0;
    P.tps = 1;
    const int planes_n = passes == 3 ? 2 : 1;
    int stages;
    size_t smem;
    if (P.halo) {
        P.nmma = 1;
        P.tiles_x = (P.d.W + HT_W - 1) / HT_W;
        P.tiles_y = (P.d.H + HT_H - 1) / HT_H;
        P.na = planes_n == 2 ? 2 : 3;
        const int fixed = P.na * planes_n * halo_bytes(bk) + STG_BYTES + BIAS_BYTES + 1024 + 512;
        const int avail = 226 * 1024 - fixed;
        if (avail / (planes_n * 9 * P.b_bytes) >= 2) P.tps = 9;     // all nine weight tiles of a chunk behind one barrier
        stages = avail / (planes_n * P.tps * P.b_bytes);
        if (stages > (P.tps == 9 ? 4 : 12)) stages = (P.tps == 9 ? 4 : 12);
        RCN_CHECK_ARG(stages >= 2, "rcn_conv2d_tc: halo kernel does not fit in shared memory");
        if (env.stages && env.stages <= stages) stages = env.stages;
        P.stages = stages;
        smem = (size_t)fixed + (size_t)stages * planes_n * P.tps * P.b_bytes;
    } else {
        const int stage_bytes = planes_n * (P.a_bytes + P.b_bytes);
        stages = (226 * 1024 - STG_BYTES - BIAS_BYTES - 1024 - 512) / stage_bytes;
        if (stages > 12) stages = 12;
        if (stages < 2) stages = 2;
        if (P.nmma == 2 && bk == 16 && stages > 2) stages &= ~1;   // single-step stages alternate between the issuers: fixed slot owners need an even ring
        if (env.stages && env.stages <= stages) stages = env.stages;
        P.stages = stages;
        smem = (size_t)stages * stage_bytes + STG_BYTES + BIAS_BYTES + 1024 + 512;
    }
    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    const long long Ktot = (long long)d->k * d->k * Cp;
    const int planes = P.s2 ? 4 * d->N : d->N;
    const int box_w = P.halo ? HB_W : TILE_W, box_h = P.halo ? HB_H : TILE_H;
    bool ok = make_act_map(&ma_hi, x_hi, planes, P.d.H, P.d.W, Cp, ldp, bk, box_w, box_h) && make_w_map(&mw_hi, w_hi, d->Cout, Ktot, nt, bk);
    if (passes == 3) ok = ok && make_act_map(&ma_lo, x_lo, planes, P.d.H, P.d.W, Cp, ldp, bk, box_w, box_h) && make_w_map(&mw_lo, w_lo, d->Cout, Ktot, nt, bk);
    else { ma_lo = ma_hi; mw_lo = mw_hi; }
    RCN_CHECK_ARG(ok, "rcn_conv2d_tc: cuTensorMapEncodeTiled failed");
    // epilogue variant (mirrors the alignment rules of the 16-byte path)
    // epilogue variant.  Row-vector (16-column blocks, 32-byte accesses; also serves NCHW aux reads and NCHW fp32 stores, coalesced
    // along image rows), else the transposing 16-byte epilogue (NHWC / pixel-shuffle stores), else the generic rows path.
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    const bool nchw_io = d->aux_nchw || d->store == RCN_STORE_NCHW;
    const int cs_store = (d->store == RCN_STORE_PS2) ? d->Cout / 4 : d->Cout;
    const bool store_ok = d->store == RCN_STORE_NHWC || (d->store == RCN_STORE_NCHW && !d->res && !d->y_hi) ||
                          (d->store == RCN_STORE_PS2 && d->ps_perm && (d->Cout & 63) == 0 && d->epi == RCN_EPI_NONE && !d->cscale);
    const bool rv_ok = env.rv && P.nmma == 1 && store_ok && (d->Cout % 16) == 0 && (cs_store % 16) == 0 &&
                       (!d->y || d->store == RCN_STORE_NCHW || ((d->ldy & 7) == 0 && al32(d->y))) &&
                       (!d->res || ((d->ldres & 7) == 0 && al32(d->res))) &&
                       (d->epi == RCN_EPI_NONE || d->aux_nchw || ((d->ldaux & 7) == 0 && al32(d->aux))) &&
                       (!d->y_hi || ((d->Cp_out & 15) == 0 && al32(d->y_hi) && (!d->y_lo || al32(d->y_lo)))) &&
                       (!d->bias || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0);
    const bool vec16_ok = !nchw_io && store_ok && (d->store != RCN_STORE_NHWC || (d->Cout & 3) == 0) && ((d->ldy & 3) == 0) &&
                          (!d->y || (reinterpret_cast<uintptr_t>(d->y) & 15) == 0) &&
                          (!d->res || (((d->ldres & 3) == 0) && ((reinterpret_cast<uintptr_t>(d->res) & 15) == 0))) &&
                          (d->epi == RCN_EPI_NONE || (((d->ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(d->aux) & 15) == 0)));
    const bool vec = rv_ok || vec16_ok;
    P.rv = rv_ok ? 1 : 0;
    RCN_CHECK_ARG(vec || (d->y && !d->y_hi), "rcn_conv2d_tc: this store / alignment combination cannot emit operand planes (y_hi) and needs y");
    // rare epilogue options -> the full instantiation; they are served by the row-vector epilogue only
    const bool full = d->cscale || d->planes_s2 || d->planes_square || nchw_io;
    const TcKernel kern = select_kernel(d->act, d->epi, vec, P.halo ? 2 : (P.nmma == 2 ? 1 : 0), full);
    P.tiles_n = (d->Cout + nt - 1) / nt;
    P.total_tiles = (long long)P.tiles_x * P.tiles_y * d->N * P.tiles_n;
    RCN_CHECK_ARG(P.total_tiles < (1ll << 31), "rcn_conv2d_tc: too many tiles");
    const int num_sms = sm_count();
    const unsigned grid = (unsigned)(P.total_tiles < num_sms ? P.total_tiles : num_sms);  // persistent: one CTA per SM
    kern<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(ma_hi, ma_lo, mw_hi, mw_lo, P);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_conv2d_tc");
    return RCN_OK;
}
