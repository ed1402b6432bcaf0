// Fused Swin MLP (sm_100a): x + fc2(GELU(fc1(LN(x)))) of the transformer half of a ConvTransBlock as ONE kernel.
//
// Reference: models/tcm.py:225-236 (Block: `self.mlp = Sequential(Linear(C, 4C), GELU(), Linear(4C, C))`,
// `x = x + self.drop_path(self.mlp(self.ln2(x)))`), used by ConvTransBlock (tcm.py:242-268) and ConvTransBlock_mzj
// (raw2bit.py:292-328) at C = 64.
//
// Layer by layer the 4C-wide hidden map travelled through HBM as bf16 hi/lo operand planes: 1 KB written and 1 KB read per
// pixel, against 0.5 KB of compulsory traffic (LN planes in, residual in, result out).  Here a tile of 128 pixels keeps its
// hidden activations in TENSOR MEMORY (same construction as csrc/ingest.cu):
//
//   fc1 runs in two 128-channel halves (A = LN planes from shared memory via TMA, B = W1 rows of the half): the epilogue warps
//   read the half's fp32 accumulator H (tcgen05.ld, lane = pixel), apply bias + GELU, split into bf16 hi + lo and write the pair
//   planes back in place (tcgen05.st); fc2 then consumes H as its A operand straight from tensor memory
//   (`tcgen05.mma [d], [a_tmem], b_desc`), accumulating both halves into one 64-column accumulator D2; the final epilogue adds
//   bias + residual and writes the block output (fp32 NHWC and / or the consumer's operand planes).
//
// W1 and W2 (hi + lo, 128 KB) stay resident in shared memory; the LN planes of the next tiles stream through a 2-stage TMA ring.
// Tensor memory: 2 tile slots x (H: 128 columns + D2: 64 columns); one epilogue warp group per slot, so the tensor pipe works
// for one tile while the other tile's GELU epilogue runs.  bf16x3 arithmetic in the k-step order of the conv engine.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rcn {
namespace {

constexpr int MF_C = 64, MF_HID = 256;
#ifndef RCN_MF_EWG
#define RCN_MF_EWG 2
#endif
constexpr int MF_EWG = RCN_MF_EWG;       // epilogue warp groups per tile slot (1 or 2)
constexpr int MF_THREADS = 128 + 256 * MF_EWG;   // warp 0: TMA producer, warp 1: MMA issuer, warps 4..: MF_EWG epilogue warp groups per slot
constexpr int MF_W1 = 2 * 32768;         // hi, lo: 256 rows x 128 B (one 64-channel K chunk)
constexpr int MF_W2 = 2 * 32768;         // hi, lo: 4 K chunks x (64 rows x 128 B)
constexpr int MF_A = 2 * 16384;          // one stage: hi, lo tile of 128 pixels x 64 channels
constexpr int MF_NA = 2;
constexpr int MF_CONST_FLOATS = MF_HID + 3 * MF_C;    // b1[256], b2[64], LayerNorm gamma[64], beta[64]
constexpr int MF_REGS_LIGHT = MF_EWG == 1 ? 96 : 64, MF_REGS_EPI = MF_EWG == 1 ? 200 : 104;   // 384 x 168 -> 128 x 96 + 256 x 200; 640 x 96 -> 128 x 64 + 512 x 104

struct MlpParams {
    rcn_mlp_desc d;
    uint32_t total_tiles;
    int wide;      // outputs are 32-byte aligned rows: 256-bit stores (full sectors) instead of 128-bit ones
};

// LN: the kernel also applies the LayerNorm in front of fc1 (models/tcm.py:234): each epilogue thread normalises its pixel's fp32 row in
// registers and writes the bf16 hi/lo A operand of fc1 into tensor memory (A1: 64 columns), so fc1 runs in the TS form as well and
// neither the LayerNorm launch nor its planes exist.  TMEM slot: [A1 64] H 128, D2 64.
template <bool LN>
__global__ void __launch_bounds__(MF_THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap ma_hi, const __grid_constant__ CUtensorMap ma_lo, const __grid_constant__ CUtensorMap m1h,
                 const __grid_constant__ CUtensorMap m1l, const __grid_constant__ CUtensorMap m2h, const __grid_constant__ CUtensorMap m2l,
                 const MlpParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const rcn_mlp_desc& p = P.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* w1s = smem;
    uint8_t* w2s = smem + MF_W1;
    uint8_t* asm_ = smem + MF_W1 + MF_W2;
    float* cst = reinterpret_cast<float*>(asm_ + (LN ? 0 : MF_NA * MF_A));      // b1[256], b2[64], gamma[64], beta[64]
    constexpr uint32_t SLOT = LN ? 256u : 192u, HOFF = LN ? 64u : 0u;        // TMEM columns per tile slot, offset of H in it
    uint64_t* wfull = reinterpret_cast<uint64_t*>(cst + MF_CONST_FLOATS);
    uint64_t* afull = wfull + 1;          // [MF_NA] TMA -> issuer
    uint64_t* aempty = afull + MF_NA;     // [MF_NA] issuer (tcgen05.commit) -> TMA
    uint64_t* a_ready = aempty + MF_NA;   // [2] epilogue warp group s -> issuer (4 warps arrive)
    uint64_t* d_ready = a_ready + 2;      // [2] issuer -> epilogue warp group s
    uint64_t* h_free = d_ready + 2;       // [2] issuer -> issuer: fc2's first half has finished reading H (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_free + 2);

    if (threadIdx.x == 0) {
        mbar_init(wfull, 1);
        for (int i = 0; i < MF_NA; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&a_ready[s], 4 * MF_EWG); mbar_init(&d_ready[s], 1); mbar_init(&h_free[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < MF_CONST_FLOATS; i += MF_THREADS) {
        float v = 0.f;
        if (i < MF_HID) v = __ldg(p.b1 + i);
        else if (i < MF_HID + MF_C) v = __ldg(p.b2 + i - MF_HID);
        else if (LN) v = i < MF_HID + 2 * MF_C ? __ldg(p.gamma + i - MF_HID - MF_C) : __ldg(p.beta + i - MF_HID - 2 * MF_C);
        cst[i] = v;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t cnt = (P.total_tiles > blockIdx.x) ? (P.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MF_REGS_LIGHT));
        // ================= TMA producer: weights once, then the LN planes of tile i into stage i % MF_NA
        const uint32_t fbw = smem_u32(wfull), w1a = smem_u32(w1s), w2a = smem_u32(w2s), a0 = smem_u32(asm_);
        const uint32_t af = smem_u32(afull), ae = smem_u32(aempty);
        if (elect_one()) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fbw), "r"((uint32_t)(MF_W1 + MF_W2)) : "memory");
            tma_load_2d_a(w1a, &m1h, fbw, 0, 0);
            tma_load_2d_a(w1a + 32768u, &m1l, fbw, 0, 0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tma_load_2d_a(w2a + (uint32_t)(c * 8192), &m2h, fbw, c * 64, 0);
                tma_load_2d_a(w2a + 32768u + (uint32_t)(c * 8192), &m2l, fbw, c * 64, 0);
            }
        }
        __syncwarp();
        uint32_t st = 0, phs = 0;
        const uint32_t nload = LN ? 0u : cnt;      // LN: the epilogue groups build the A operand themselves
        for (uint32_t i = 0; i < nload; ++i) {
            const uint32_t t = blockIdx.x + i * gridDim.x;
            mbar_wait_a(ae + 8u * st, phs ^ 1u);
            if (elect_one()) {
                const uint32_t fb = af + 8u * st, dst = a0 + st * (uint32_t)MF_A;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)MF_A) : "memory");
                tma_load_2d_a(dst, &ma_hi, fb, 0, (int)(t * 128u));
                tma_load_2d_a(dst + 16384u, &ma_lo, fb, 0, (int)(t * 128u));
            }
            __syncwarp();
            if (++st == MF_NA) { st = 0; phs ^= 1u; }
        }
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MF_REGS_LIGHT));
        // ================= MMA issuer
        const uint32_t ar = smem_u32(a_ready), dr = smem_u32(d_ready), af = smem_u32(afull), ae = smem_u32(aempty);
        const uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t w1d = make_kmajor_desc(smem_u32(w1s), 64), w2d = make_kmajor_desc(smem_u32(w2s), 64);
        const uint64_t ad0 = make_kmajor_desc(smem_u32(asm_), 64);
        mbar_wait_a(smem_u32(wfull), 0);
        const uint32_t hf = smem_u32(h_free);
        uint32_t ph = 0, hph = 0;  // bit s: parity of the next a_ready[s] / h_free[s] completion
        uint32_t ast = 0, aph = 0; // A ring position of the next tile
        uint32_t stage_of[2] = {0, 0};
        for (uint32_t pr = 0; 2 * pr < cnt; ++pr) {
            // step A: fc1, hidden channels 0..127 -> H
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (2 * pr + s >= cnt) continue;
                if (LN || pr > 0) {   // LN: the slot's A1 operand is written (which the epilogue groups do after they have drained the
                                      // previous tile); else: the slot's previous tile has left tensor memory (final epilogue read D2)
                    mbar_wait_a(ar + 8u * s, (ph >> s) & 1u);
                    ph ^= 1u << s;
                }
                if (!LN) mbar_wait_a(af + 8u * ast, aph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t A1 = tmem_base + (uint32_t)s * SLOT, H = A1 + HOFF;
                    const uint64_t ah = ad0 + (uint64_t)((ast * (uint32_t)MF_A) >> 4), al = ah + (uint64_t)(16384 >> 4);
                    const uint64_t bh = w1d, bl = w1d + (uint64_t)(32768 >> 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (LN) {
                            umma_ts(H, A1 + 16u * j + 8u, bh + 2u * j, idesc128, j > 0 ? 1u : 0u);
                            umma_ts(H, A1 + 16u * j, bl + 2u * j, idesc128, 1u);
                            umma_ts(H, A1 + 16u * j, bh + 2u * j, idesc128, 1u);
                        } else {
                            umma_bf16(H, al + 2u * j, bh + 2u * j, idesc128, j > 0 ? 1u : 0u);
                            umma_bf16(H, ah + 2u * j, bl + 2u * j, idesc128, 1u);
                            umma_bf16(H, ah + 2u * j, bh + 2u * j, idesc128, 1u);
                        }
                    }
                    umma_commit_a(dr + 8u * s);
                }
                __syncwarp();
                stage_of[s] = ast;
                if (++ast == MF_NA) { ast = 0; aph ^= 1u; }
            }
            // step B: fc2 over hidden 0..127 (A = H in tensor memory) -> D2, then fc1 for hidden channels 128..255 -> H again
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (2 * pr + s >= cnt) continue;
                mbar_wait_a(ar + 8u * s, (ph >> s) & 1u);
                ph ^= 1u << s;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t H = tmem_base + (uint32_t)s * SLOT + HOFF, D2 = H + 128u;
                    const uint64_t ch = w2d, cl = w2d + (uint64_t)(32768 >> 4);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint64_t off = (uint64_t)(((j >> 2) * 8192) >> 4) + (uint64_t)((j & 3) * 2);
                        umma_ts(D2, H + 16u * j + 8u, ch + off, idesc64, j > 0 ? 1u : 0u);
                        umma_ts(D2, H + 16u * j, cl + off, idesc64, 1u);
                        umma_ts(D2, H + 16u * j, ch + off, idesc64, 1u);
                    }
                    umma_commit_a(hf + 8u * s);
                }
                __syncwarp();
                // H is overwritten next.  tcgen05.mma instructions are only guaranteed to execute in issue order when they share
                // the accumulator and shape, so the read of H (as A operand, N = 64 into D2) is fenced off from the write of H
                // (N = 128) by a commit / wait; the tensor pipe has slack here (3072 cycles of MMA per tile against a GELU epilogue
                // of ~2x that).
                mbar_wait_a(hf + 8u * s, (hph >> s) & 1u);
                hph ^= 1u << s;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t A1 = tmem_base + (uint32_t)s * SLOT, H = A1 + HOFF;
                    const uint32_t sg = stage_of[s];
                    const uint64_t ah = ad0 + (uint64_t)((sg * (uint32_t)MF_A) >> 4), al = ah + (uint64_t)(16384 >> 4);
                    const uint64_t bh = w1d + (uint64_t)(16384 >> 4), bl = bh + (uint64_t)(32768 >> 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (LN) {
                            umma_ts(H, A1 + 16u * j + 8u, bh + 2u * j, idesc128, j > 0 ? 1u : 0u);
                            umma_ts(H, A1 + 16u * j, bl + 2u * j, idesc128, 1u);
                            umma_ts(H, A1 + 16u * j, bh + 2u * j, idesc128, 1u);
                        } else {
                            umma_bf16(H, al + 2u * j, bh + 2u * j, idesc128, j > 0 ? 1u : 0u);
                            umma_bf16(H, ah + 2u * j, bl + 2u * j, idesc128, 1u);
                            umma_bf16(H, ah + 2u * j, bh + 2u * j, idesc128, 1u);
                        }
                    }
                    umma_commit_a(dr + 8u * s);
                    if (!LN) umma_commit_a(ae + 8u * sg);       // the LN planes of this tile are consumed
                }
                __syncwarp();
            }
            // step C: fc2 over hidden 128..255 -> D2 complete
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (2 * pr + s >= cnt) continue;
                mbar_wait_a(ar + 8u * s, (ph >> s) & 1u);
                ph ^= 1u << s;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t H = tmem_base + (uint32_t)s * SLOT + HOFF, D2 = H + 128u;
                    const uint64_t ch = w2d + (uint64_t)(16384 >> 4), cl = ch + (uint64_t)(32768 >> 4);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint64_t off = (uint64_t)(((j >> 2) * 8192) >> 4) + (uint64_t)((j & 3) * 2);
                        umma_ts(D2, H + 16u * j + 8u, ch + off, idesc64, 1u);
                        umma_ts(D2, H + 16u * j, cl + off, idesc64, 1u);
                        umma_ts(D2, H + 16u * j, ch + off, idesc64, 1u);
                    }
                    umma_commit_a(dr + 8u * s);
                }
                __syncwarp();
            }
        }
    } else if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MF_REGS_LIGHT));
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MF_REGS_EPI));
        // ================= epilogue: MF_EWG warp groups per slot; warp q of a group owns TMEM lanes 32q .. 32q+31 (lane = pixel);
        // group h handles the blocks [h * NB, (h + 1) * NB) of 16 hidden channels of each fc1 half and [h * NB2, ..) of the output
        const int e = warp - 4;
        const int s = (e >> 2) & 1, h = e >> 3, q = warp & 3;
        constexpr int NB = 8 / MF_EWG, NB2 = 4 / MF_EWG;
        const uint32_t A1 = tmem_base + (uint32_t)s * SLOT + ((uint32_t)(q * 32) << 16), H = A1 + HOFF, D2 = H + 128u;
        const uint32_t arb = smem_u32(&a_ready[s]), drb = smem_u32(&d_ready[s]);
        const uint32_t c_b1 = smem_u32(cst), c_b2 = c_b1 + 4u * MF_HID, c_g = c_b2 + 4u * MF_C, c_be = c_g + 4u * MF_C;
        const float* xln = p.x_ln;
        int ldx = p.ldx;
        float eps = p.eps;
        opaque_ptr(xln); opaque(ldx); opaque(eps);
        const float* res = p.res; float* y = p.y;
        uint16_t* yhi = reinterpret_cast<uint16_t*>(p.y_hi); uint16_t* ylo = reinterpret_cast<uint16_t*>(p.y_lo);
        int ldres = p.ldres, ldy = p.ldy, cpo = p.Cp_out, wide = P.wide;
        opaque(wide);
        long long npix = p.npix;
        opaque_ptr(res); opaque_ptr(y); opaque_ptr(yhi); opaque_ptr(ylo); opaque(ldres); opaque(ldy); opaque(cpo);
        uint32_t dph = 0;
        for (uint32_t i = (uint32_t)s; i < cnt; i += 2) {
            const long long pix = (long long)(blockIdx.x + i * gridDim.x) * 128 + q * 32 + lane;
            const bool ok = pix < npix;
            if (LN) {
                // ---- LayerNorm of this pixel's row in registers (two-pass statistics like rcn_layernorm) -> A operand of fc1 in A1;
                // group h writes the k-steps [j0, j0 + nj).  The next tile's rows are pulled into L2 meanwhile.
                if (i + 2 < cnt) {
                    const long long pn = pix + (long long)2 * gridDim.x * 128;
                    if (pn < npix) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xln + pn * ldx));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xln + pn * ldx + 32));
                    }
                }
                float4 xv[16];
#pragma unroll
                for (int g = 0; g < 16; ++g) xv[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) {
                    const float4* xp = reinterpret_cast<const float4*>(xln + pix * ldx);
#pragma unroll
                    for (int g = 0; g < 16; ++g) xv[g] = __ldg(xp + g);
                }
                float sm = 0.f;
#pragma unroll
                for (int g = 0; g < 16; ++g) sm += (xv[g].x + xv[g].y) + (xv[g].z + xv[g].w);
                const float mean = sm * (1.f / 64.f);
                float sq = 0.f;
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    const float dx = xv[g].x - mean, dy = xv[g].y - mean, dz = xv[g].z - mean, dw = xv[g].w - mean;
                    sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                }
                const float rstd = rsqrtf(sq * (1.f / 64.f) + eps);
                const int j0 = (MF_EWG == 2 && h == 1) ? 2 : 0, nj = 4 / MF_EWG;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j >= j0 && j < j0 + nj) {
                        float val[16];
                        uint32_t pk[16];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 gg = lds4(c_g + 64u * j + 16u * g), be = lds4(c_be + 64u * j + 16u * g);
                            const float4 xx = xv[4 * j + g];
                            val[4 * g + 0] = (xx.x - mean) * rstd * gg.x + be.x;
                            val[4 * g + 1] = (xx.y - mean) * rstd * gg.y + be.y;
                            val[4 * g + 2] = (xx.z - mean) * rstd * gg.z + be.z;
                            val[4 * g + 3] = (xx.w - mean) * rstd * gg.w + be.w;
                        }
                        split_pack16(val, pk);
                        tmem_st16(A1 + 16u * j, pk);
                    }
                }
                tmem_wait_st();
                chain_arrive(arb, lane);
            }
            // ---- fc1 halves: accumulator -> bias + GELU -> hi/lo pairs, in place
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                mbar_wait_a(drb, dph);
                dph ^= 1u;
                tc_fence_after();
                const uint32_t D = H + 16u * (uint32_t)(h * NB);
                const uint32_t cb = c_b1 + 512u * half + 64u * (uint32_t)(h * NB);
                uint32_t v[2][16];
                tmem_ld16_async(D, v[0]);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    tmem_wait_ld16(v[b & 1]);
                    if (b < NB - 1) tmem_ld16_async(D + 16u * (b + 1), v[(b + 1) & 1]);
                    float val[16];
                    uint32_t pk[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 bb = lds4(cb + 64u * b + 16u * g);
                        val[4 * g + 0] = gelu_as(__uint_as_float(v[b & 1][4 * g + 0]) + bb.x);
                        val[4 * g + 1] = gelu_as(__uint_as_float(v[b & 1][4 * g + 1]) + bb.y);
                        val[4 * g + 2] = gelu_as(__uint_as_float(v[b & 1][4 * g + 2]) + bb.z);
                        val[4 * g + 3] = gelu_as(__uint_as_float(v[b & 1][4 * g + 3]) + bb.w);
                    }
                    split_pack16(val, pk);
                    tmem_st16(D + 16u * b, pk);
                }
                tmem_wait_st();
                chain_arrive(arb, lane);
            }
            // ---- residual channels of this pixel and group, fetched while fc2 finishes
            const int c0 = 16 * NB2 * h;
            float4 r[4 * NB2];
#pragma unroll
            for (int g = 0; g < 4 * NB2; ++g) r[g] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (res && ok) {
                const float4* rp = reinterpret_cast<const float4*>(res + pix * ldres + c0);
#pragma unroll
                for (int g = 0; g < 4 * NB2; ++g) r[g] = __ldg(rp + g);
            }
            mbar_wait_a(drb, dph);
            dph ^= 1u;
            tc_fence_after();
            {
                const uint32_t D = D2 + (uint32_t)c0;
                uint32_t v[2][16];
                tmem_ld16_async(D, v[0]);
#pragma unroll
                for (int b = 0; b < NB2; ++b) {
                    tmem_wait_ld16(v[b & 1]);
                    if (b < NB2 - 1) tmem_ld16_async(D + 16u * (b + 1), v[(b + 1) & 1]);
                    float val[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 bb = lds4(c_b2 + 4u * (uint32_t)c0 + 64u * b + 16u * g);
                        const float4 rr = r[4 * b + g];
                        val[4 * g + 0] = __uint_as_float(v[b & 1][4 * g + 0]) + bb.x + rr.x;
                        val[4 * g + 1] = __uint_as_float(v[b & 1][4 * g + 1]) + bb.y + rr.y;
                        val[4 * g + 2] = __uint_as_float(v[b & 1][4 * g + 2]) + bb.z + rr.z;
                        val[4 * g + 3] = __uint_as_float(v[b & 1][4 * g + 3]) + bb.w + rr.w;
                    }
                    if (ok) {
                        if (y) {
                            float* yp = y + pix * ldy + c0 + 16 * b;
                            if (wide) {
                                uint32_t u[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) u[j] = __float_as_uint(val[j]);
                                stg256(yp, u);
                                stg256(yp + 8, u + 8);
                            } else {
#pragma unroll
                                for (int g = 0; g < 4; ++g)
                                    reinterpret_cast<float4*>(yp)[g] = make_float4(val[4 * g], val[4 * g + 1], val[4 * g + 2], val[4 * g + 3]);
                            }
                        }
                        if (yhi) {
                            uint32_t pk[16];
                            split_pack16(val, pk);
                            const long long po = pix * cpo + c0 + 16 * b;
                            if (wide) {
                                stg256(yhi + po, pk);
                                stg256(ylo + po, pk + 8);
                            } else {
                                stg128(yhi + po, pk[0], pk[1], pk[2], pk[3]);
                                stg128(yhi + po + 8, pk[4], pk[5], pk[6], pk[7]);
                                stg128(ylo + po, pk[8], pk[9], pk[10], pk[11]);
                                stg128(ylo + po + 8, pk[12], pk[13], pk[14], pk[15]);
                            }
                        }
                    }
                }
            }
            // D2 is drained: the slot may take its next tile.  With LN the groups' A1 arrivals of the next tile say so (program order)
            if (!LN) chain_arrive(arb, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// rows x cols bf16 matrix with a row stride (elements), box {64, box_rows}, 128-byte swizzle
bool make_rows_map(CUtensorMap* m, const void* base, long long rows, int cols, long long ld, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
}  // namespace rcn

using namespace rcn;

extern "C" int rcn_mlp_fused(const rcn_mlp_desc* d, void* stream) {
    const bool ln = d && d->x_ln != nullptr;
    RCN_CHECK_ARG(d && (ln || (d->x_hi && d->x_lo)) && d->w1_hi && d->w1_lo && d->w2_hi && d->w2_lo && d->b1 && d->b2, "rcn_mlp_fused: null pointer");
    RCN_CHECK_ARG(d->C == MF_C && d->hidden == MF_HID, "rcn_mlp_fused: only C = 64, hidden = 256 is built (got %d, %d)", d->C, d->hidden);
    RCN_CHECK_ARG(d->npix > 0 && d->npix < (1ll << 31), "rcn_mlp_fused: bad pixel count");
    RCN_CHECK_ARG(d->y || d->y_hi, "rcn_mlp_fused: no output");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    RCN_CHECK_ARG(ln || (d->ldp_in >= MF_C && d->ldp_in % 8 == 0 && al16(d->x_hi) && al16(d->x_lo)),
                  "rcn_mlp_fused: input planes need a pixel stride >= 64 (multiple of 8), 16-byte aligned");
    RCN_CHECK_ARG(!ln || (d->gamma && d->beta && d->ldx >= MF_C && d->ldx % 4 == 0 && al16(d->x_ln)),
                  "rcn_mlp_fused: the LayerNorm input needs gamma, beta and 16-byte aligned rows");
    RCN_CHECK_ARG(!d->res || (d->ldres >= MF_C && d->ldres % 4 == 0 && al16(d->res)), "rcn_mlp_fused: residual must be 16-byte aligned rows");
    RCN_CHECK_ARG(!d->y || (d->ldy >= MF_C && d->ldy % 4 == 0 && al16(d->y)), "rcn_mlp_fused: output must be 16-byte aligned rows");
    RCN_CHECK_ARG(!d->y_hi || (d->y_lo && d->Cp_out >= MF_C && d->Cp_out % 8 == 0 && al16(d->y_hi) && al16(d->y_lo)),
                  "rcn_mlp_fused: output planes need hi and lo, a pixel stride >= 64 (multiple of 8), 16-byte aligned");
    RCN_CHECK_ARG(get_encode() != nullptr, "rcn_mlp_fused: cuTensorMapEncodeTiled is not available from the driver");
    MlpParams P;
    P.d = *d;
    const long long tiles = (d->npix + 127) / 128;
    P.total_tiles = (uint32_t)tiles;
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    P.wide = (!d->y || (al32(d->y) && d->ldy % 8 == 0)) && (!d->y_hi || (al32(d->y_hi) && al32(d->y_lo) && d->Cp_out % 16 == 0));
    CUtensorMap ma_hi, ma_lo, m1h, m1l, m2h, m2l;
    bool ok = true;
    if (!ln) ok = make_rows_map(&ma_hi, d->x_hi, d->npix, MF_C, d->ldp_in, 128) && make_rows_map(&ma_lo, d->x_lo, d->npix, MF_C, d->ldp_in, 128);
    ok = ok && make_rows_map(&m1h, d->w1_hi, MF_HID, MF_C, MF_C, 256) && make_rows_map(&m1l, d->w1_lo, MF_HID, MF_C, MF_C, 256) &&
                    make_rows_map(&m2h, d->w2_hi, MF_C, MF_HID, MF_HID, 64) && make_rows_map(&m2l, d->w2_lo, MF_C, MF_HID, MF_HID, 64);
    RCN_CHECK_ARG(ok, "rcn_mlp_fused: cuTensorMapEncodeTiled failed");
    if (ln) { ma_hi = m1h; ma_lo = m1l; }
    static bool attr_set[MAX_DEVICES] = {};
    const int dev = current_device();
    if (!attr_set[dev]) {
        cudaFuncSetAttribute(mlp_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(mlp_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set[dev] = true;
    }
    const size_t smem = (size_t)MF_W1 + MF_W2 + (ln ? 0 : MF_NA * MF_A) + MF_CONST_FLOATS * 4 + 128 + 1024;
    const int sms = sm_count();
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    if (ln) mlp_fused_kernel<true><<<grid, MF_THREADS, smem, (cudaStream_t)stream>>>(ma_hi, ma_lo, m1h, m1l, m2h, m2l, P);
    else mlp_fused_kernel<false><<<grid, MF_THREADS, smem, (cudaStream_t)stream>>>(ma_hi, ma_lo, m1h, m1l, m2h, m2l, P);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_mlp_fused");
    return RCN_OK;
}
