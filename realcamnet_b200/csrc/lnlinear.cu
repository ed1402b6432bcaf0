// Fused LayerNorm + Linear (sm_100a): y = Linear(LayerNorm(x)) for C = 64 rows, e.g. the qkv embedding of the Swin blocks.
//
// Reference: models/tcm.py:233 (`x = x + self.msa(self.ln1(x))`) with models/tcm.py:193 (`qkv = self.embedding_layer(x)`,
// Linear(C, 3C)) -- the LayerNorm used to be its own launch that wrote bf16 hi/lo operand planes (256 B per pixel written and
// read again) for the projection to fetch by TMA.  Here every epilogue thread normalises its pixel's fp32 row in registers
// (two-pass statistics like rcn_layernorm), writes the bf16 hi/lo A operand straight into tensor memory and the projection
// runs in the TS form (`tcgen05.mma [d_tmem], [a_tmem], b_desc`), as in csrc/mlp.cu and csrc/ingest.cu.  Weights (hi + lo)
// stay resident in shared memory.  Tensor memory: 2 tile slots x (A1: 64 columns + D: up to 192 columns); two epilogue warp
// groups per slot.  bf16x3 arithmetic in the k-step order of the conv engine.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rcn {
namespace {

constexpr int LL_C = 64, LL_NMAX = 192;
constexpr int LL_EWG = 2;                        // epilogue warp groups per tile slot
constexpr int LL_THREADS = 128 + 256 * LL_EWG;   // warp 0: weight TMA, warp 1: MMA issuer, warps 4..: epilogue
constexpr int LL_W = 2 * LL_NMAX * 128;          // hi, lo: up to 192 rows x 128 B (one 64-channel K chunk)
constexpr int LL_CONST_FLOATS = LL_NMAX + 2 * LL_C;   // bias[192], gamma[64], beta[64]
constexpr int LL_REGS_LIGHT = 64, LL_REGS_EPI = 104;

struct LnLinearParams {
    rcn_lnlinear_desc d;
    uint32_t total_tiles;
    int wide;
};

__global__ void __launch_bounds__(LL_THREADS, 1)
ln_linear_kernel(const __grid_constant__ CUtensorMap mwh, const __grid_constant__ CUtensorMap mwl, const LnLinearParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const rcn_lnlinear_desc& p = P.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Cout = p.Cout;
    float* cst = reinterpret_cast<float*>(smem + LL_W);
    uint64_t* wfull = reinterpret_cast<uint64_t*>(cst + LL_CONST_FLOATS);
    uint64_t* a_ready = wfull + 1;    // [2] epilogue groups of slot s -> issuer: A1 is written (and the slot's previous tile is drained)
    uint64_t* d_ready = a_ready + 2;  // [2] issuer -> epilogue groups of slot s
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_ready + 2);

    if (threadIdx.x == 0) {
        mbar_init(wfull, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&a_ready[s], 4 * LL_EWG); mbar_init(&d_ready[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < LL_CONST_FLOATS; i += LL_THREADS) {
        float v = 0.f;
        if (i < LL_NMAX) v = (i < Cout && p.bias) ? __ldg(p.bias + i) : 0.f;
        else if (i < LL_NMAX + LL_C) v = __ldg(p.gamma + i - LL_NMAX);
        else v = __ldg(p.beta + i - LL_NMAX - LL_C);
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
    const uint32_t lo_off = (uint32_t)Cout * 128u;     // the lo weights follow the hi weights

    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LL_REGS_LIGHT));
        if (elect_one()) {
            const uint32_t fb = smem_u32(wfull), wa = smem_u32(smem);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(2u * lo_off) : "memory");
            tma_load_2d_a(wa, &mwh, fb, 0, 0);
            tma_load_2d_a(wa + lo_off, &mwl, fb, 0, 0);
        }
        __syncwarp();
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LL_REGS_LIGHT));
        // ================= MMA issuer: one contraction per tile, slots in round-robin order
        const uint32_t ar = smem_u32(a_ready), dr = smem_u32(d_ready);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t wh = make_kmajor_desc(smem_u32(smem), 64), wl = wh + (uint64_t)(lo_off >> 4);
        mbar_wait_a(smem_u32(wfull), 0);
        uint32_t ph = 0;
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t s = i & 1u;
            mbar_wait_a(ar + 8u * s, (ph >> s) & 1u);
            ph ^= 1u << s;
            tc_fence_after();
            if (elect_one()) {
                const uint32_t A1 = tmem_base + s * 256u, D = A1 + 64u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    umma_ts(D, A1 + 16u * j + 8u, wh + 2u * j, idesc, j > 0 ? 1u : 0u);
                    umma_ts(D, A1 + 16u * j, wl + 2u * j, idesc, 1u);
                    umma_ts(D, A1 + 16u * j, wh + 2u * j, idesc, 1u);
                }
                umma_commit_a(dr + 8u * s);
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LL_REGS_LIGHT));
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(LL_REGS_EPI));
        // ================= epilogue: group h of slot s; warp q owns TMEM lanes 32q .. 32q+31 (lane = pixel)
        const int e = warp - 4;
        const int s = (e >> 2) & 1, h = e >> 3, q = warp & 3;
        const uint32_t A1 = tmem_base + (uint32_t)s * 256u + ((uint32_t)(q * 32) << 16), D = A1 + 64u;
        const uint32_t arb = smem_u32(&a_ready[s]), drb = smem_u32(&d_ready[s]);
        const uint32_t c_b = smem_u32(cst), c_g = c_b + 4u * LL_NMAX, c_be = c_g + 4u * LL_C;
        const float* x = p.x; float* y = p.y;
        int ldx = p.ldx, ldy = p.ldy, wide = P.wide;
        float eps = p.eps;
        long long npix = p.npix;
        opaque_ptr(x); opaque_ptr(y); opaque(ldx); opaque(ldy); opaque(wide); opaque(eps);
        const int nblk = Cout >> 4;                                    // 16-channel output blocks
        const int bper = (nblk + LL_EWG - 1) / LL_EWG, bfirst = h * bper, blast = min(nblk, bfirst + bper);
        uint32_t dph = 0;
        for (uint32_t i = (uint32_t)s; i < cnt; i += 2) {
            const long long pix = (long long)(blockIdx.x + i * gridDim.x) * 128 + q * 32 + lane;
            const bool ok = pix < npix;
            // ---- LayerNorm of this pixel's row -> A operand in A1 (group h writes k-steps [2h, 2h + 2))
            {
                if (i + 2 < cnt) {
                    const long long pn = pix + (long long)2 * gridDim.x * 128;
                    if (pn < npix) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(x + pn * ldx));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(x + pn * ldx + 32));
                    }
                }
                float4 xv[16];
#pragma unroll
                for (int g = 0; g < 16; ++g) xv[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) {
                    const float4* xp = reinterpret_cast<const float4*>(x + pix * ldx);
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
                // only this group's 32 channels stay live past the statistics (selects, not a run-time array index)
                float4 mine[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) mine[g] = h ? xv[8 + g] : xv[g];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const uint32_t j = 2u * (uint32_t)h + jj;
                    float val[16];
                    uint32_t pk[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 gg = lds4(c_g + 64u * j + 16u * g), be = lds4(c_be + 64u * j + 16u * g);
                        const float4 xx = mine[4 * jj + g];
                        val[4 * g + 0] = (xx.x - mean) * rstd * gg.x + be.x;
                        val[4 * g + 1] = (xx.y - mean) * rstd * gg.y + be.y;
                        val[4 * g + 2] = (xx.z - mean) * rstd * gg.z + be.z;
                        val[4 * g + 3] = (xx.w - mean) * rstd * gg.w + be.w;
                    }
                    split_pack16(val, pk);
                    tmem_st16(A1 + 16u * j, pk);
                }
                tmem_wait_st();
                chain_arrive(arb, lane);
            }
            // ---- accumulator -> + bias -> fp32 rows
            mbar_wait_a(drb, dph);
            dph ^= 1u;
            tc_fence_after();
            if (bfirst < blast) {
                // two register buffers with compile-time names (a run-time parity index would put them in local memory)
                uint32_t v0[16], v1[16];
                auto emit = [&](int b, const uint32_t (&v)[16]) {
                    uint32_t u[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 bb = lds4(c_b + 64u * (uint32_t)b + 16u * g);
                        u[4 * g + 0] = __float_as_uint(__uint_as_float(v[4 * g + 0]) + bb.x);
                        u[4 * g + 1] = __float_as_uint(__uint_as_float(v[4 * g + 1]) + bb.y);
                        u[4 * g + 2] = __float_as_uint(__uint_as_float(v[4 * g + 2]) + bb.z);
                        u[4 * g + 3] = __float_as_uint(__uint_as_float(v[4 * g + 3]) + bb.w);
                    }
                    if (ok) {
                        float* yp = y + pix * ldy + 16 * b;
                        if (wide) {
                            stg256(yp, u);
                            stg256(yp + 8, u + 8);
                        } else {
                            stg128(yp, u[0], u[1], u[2], u[3]);
                            stg128(yp + 4, u[4], u[5], u[6], u[7]);
                            stg128(yp + 8, u[8], u[9], u[10], u[11]);
                            stg128(yp + 12, u[12], u[13], u[14], u[15]);
                        }
                    }
                };
                tmem_ld16_async(D + 16u * (uint32_t)bfirst, v0);
#pragma unroll 1
                for (int b = bfirst; b < blast; b += 2) {
                    tmem_wait_ld16(v0);
                    if (b + 1 < blast) tmem_ld16_async(D + 16u * (uint32_t)(b + 1), v1);
                    emit(b, v0);
                    if (b + 1 < blast) {
                        tmem_wait_ld16(v1);
                        if (b + 2 < blast) tmem_ld16_async(D + 16u * (uint32_t)(b + 2), v0);
                        emit(b + 1, v1);
                    }
                }
            }
            // the next tile of this slot starts with these warps' A1 writes and arrivals: D is drained by then (program order)
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

}  // namespace
}  // namespace rcn

using namespace rcn;

extern "C" int rcn_ln_linear_fused(const rcn_lnlinear_desc* d, void* stream) {
    RCN_CHECK_ARG(d && d->x && d->gamma && d->beta && d->w_hi && d->w_lo && d->y, "rcn_ln_linear_fused: null pointer");
    RCN_CHECK_ARG(d->C == LL_C && d->Cout >= 16 && d->Cout <= LL_NMAX && d->Cout % 16 == 0,
                  "rcn_ln_linear_fused: only C = 64 and Cout in {16, 32, .., 192} are built (got %d -> %d)", d->C, d->Cout);
    RCN_CHECK_ARG(d->npix > 0 && d->npix < (1ll << 31), "rcn_ln_linear_fused: bad pixel count");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    RCN_CHECK_ARG(d->ldx >= LL_C && d->ldx % 4 == 0 && al16(d->x), "rcn_ln_linear_fused: input rows must be 16-byte aligned");
    RCN_CHECK_ARG(d->ldy >= d->Cout && d->ldy % 4 == 0 && al16(d->y), "rcn_ln_linear_fused: output rows must be 16-byte aligned");
    RCN_CHECK_ARG(get_encode() != nullptr, "rcn_ln_linear_fused: cuTensorMapEncodeTiled is not available from the driver");
    LnLinearParams P;
    P.d = *d;
    const long long tiles = (d->npix + 127) / 128;
    P.total_tiles = (uint32_t)tiles;
    P.wide = al32(d->y) && d->ldy % 8 == 0;
    CUtensorMap mwh, mwl;
    const bool ok = make_w_map(&mwh, d->w_hi, d->Cout, LL_C, d->Cout, 64) && make_w_map(&mwl, d->w_lo, d->Cout, LL_C, d->Cout, 64);
    RCN_CHECK_ARG(ok, "rcn_ln_linear_fused: cuTensorMapEncodeTiled failed");
    static bool attr_set[MAX_DEVICES] = {};
    const int dev = current_device();
    if (!attr_set[dev]) {
        cudaFuncSetAttribute(ln_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_set[dev] = true;
    }
    const size_t smem = (size_t)LL_W + LL_CONST_FLOATS * 4 + 128 + 1024;
    const int sms = sm_count();
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    ln_linear_kernel<<<grid, LL_THREADS, smem, (cudaStream_t)stream>>>(mwh, mwl, P);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_ln_linear_fused");
    return RCN_OK;
}
