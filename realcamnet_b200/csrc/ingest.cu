// Fused packed-Bayer ingest (sm_100a): the lens-shading coordinate MLP and conv_first * (lsc + 1) as ONE kernel.
//
// Reference: models/raw2bit.py:1771-1780 (`lsc_fea = self.lsc(coord)`, `fea = self.conv_first(raw) * (lsc_fea + 1)`),
// models/LiteISP.py:363-378 (Lens_Shading_Correction: 1x1 conv 2->C, LeakyReLU(0.1), 1x1 C->C, LReLU, 1x1 C->C, LReLU, 1x1 C->C).
//
// Per-layer launches moved every hidden map through HBM as 16-bit hi/lo operand planes (2.1 GB written + 2.1 GB read per
// layer at 2048^2) and the lens-shading map a second time as the aux operand of conv_first.  Here a tile of 128 pixels stays
// on the SM from the two coordinate channels to the finished outputs:
//
//   * layer 0 (K = 2) is two FFMAs per value on the CUDA cores;
//   * layers 1-3 (C = 128 -> 128) run on tcgen05 with the ACTIVATIONS AS THE A OPERAND IN TENSOR MEMORY (`tcgen05.mma
//     [d_tmem], [a_tmem], b_desc`): the epilogue warps read a layer's fp32 accumulator (tcgen05.ld, lane = pixel), apply
//     bias + LeakyReLU, split the value into bf16 hi + lo (same rounding as rcn_split_bf16) and write the pair planes back IN
//     PLACE over the accumulator columns they came from (tcgen05.st): 16 fp32 columns of channels 16j..16j+15 become the 8 hi
//     and 8 lo columns of k-step j.  No shared-memory round trip, no swizzle arithmetic, no HBM traffic between layers;
//   * the weights of the three layers (hi + lo, 192 KB) are fetched ONCE per CTA by TMA and stay in shared memory;
//   * conv_first (3x3, 4 -> 128) joins as one more contraction of the same tile: each thread gathers its pixel's 36 im2col
//     values (9 coalesced 16-byte loads), writes them as a 48-wide hi/lo A operand into tensor memory, and the final
//     epilogue multiplies the two accumulators, `(conv + b) * (lsc + 1)`, writing the lens-shading map once (NCHW fp32, an
//     output of forward()) and the product once (the 16-bit operand planes of conv_down, polyphase layout).
//
// Tensor memory: 2 tile slots x 256 columns.  A slot's regions X and Y (128 columns each) alternate between "A operand" and
// "accumulator" from layer to layer; two tiles are in flight per CTA, each served by IG_EWG epilogue warp groups (which split the
// channel blocks of every phase), so the tensor pipe works on one tile while the other tile's epilogue runs.  The conv accumulator has no 128 free columns left in a slot: it is produced
// in two 64-channel halves into X[64..128) after layer 3 has consumed X.
//
// Arithmetic = the bf16x3 engine's (a_lo*w_hi + a_hi*w_lo + a_hi*w_hi, fp32 accumulate); layer 0 is plain fp32.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace rcn {
namespace {

constexpr int IG_C = 128;                      // hidden / output width
#ifndef RCN_IG_EWG
#define RCN_IG_EWG 2
#endif
constexpr int IG_EWG = RCN_IG_EWG;             // epilogue warp groups per tile slot (1 or 2): 2 halves every epilogue phase's latency
constexpr int IG_THREADS = 128 + 256 * IG_EWG; // warp 0: weight TMA, warp 1: MMA issuer, warps 4..: IG_EWG epilogue warp groups per slot
constexpr int IG_W_LAYER = 4 * 16384;          // one layer's B operand: hi chunk 0, hi chunk 1, lo chunk 0, lo chunk 1 (64-channel K chunks)
constexpr int IG_W_MLP = 3 * IG_W_LAYER;       // 196608
constexpr int IG_KC = 48;                      // conv_first K: 9 taps x 4 channels = 36, zero-padded to three k16 steps
constexpr int IG_W_CONV = 2 * 3 * 4096;        // hi chunks 0-2, lo chunks 0-2 (16-channel K chunks, 32-byte rows)
constexpr int IG_CONST_FLOATS = 7 * IG_C;      // w0[:,0], w0[:,1], b0, b1, b2, b3, b_conv
// register re-split after the prologue: 384 x 168 -> 128 x 96 + 256 x 200; 640 x 96 -> 128 x 64 + 512 x 104
constexpr int IG_REGS_LIGHT = IG_EWG == 1 ? 96 : 64, IG_REGS_EPI = IG_EWG == 1 ? 200 : 104;

struct IngestParams {
    rcn_ingest_desc d;
    int tiles_x, tiles_y;
    uint32_t total_tiles;
};

// CONV: also conv_first * (lsc + 1) -> operand planes
template <bool CONV>
__global__ void __launch_bounds__(IG_THREADS, 1)
ingest_kernel(const __grid_constant__ CUtensorMap m1h, const __grid_constant__ CUtensorMap m1l, const __grid_constant__ CUtensorMap m2h,
              const __grid_constant__ CUtensorMap m2l, const __grid_constant__ CUtensorMap m3h, const __grid_constant__ CUtensorMap m3l,
              const __grid_constant__ CUtensorMap mch, const __grid_constant__ CUtensorMap mcl, const IngestParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const rcn_ingest_desc& p = P.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* wsm = smem;                                      // MLP weights, then conv weights
    float* cst = reinterpret_cast<float*>(smem + IG_W_MLP + (CONV ? IG_W_CONV : 0));
    uint64_t* wfull = reinterpret_cast<uint64_t*>(cst + IG_CONST_FLOATS);
    uint64_t* a_ready = wfull + 1;    // [2] epilogue warp group s -> issuer: the slot's A operand is in tensor memory (4 warps arrive)
    uint64_t* d_ready = a_ready + 2;  // [2] issuer -> epilogue warp group s: the slot's accumulator is complete (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_ready + 2);

    if (threadIdx.x == 0) {
        mbar_init(wfull, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&a_ready[s], 4 * IG_EWG); mbar_init(&d_ready[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < IG_C; i += IG_THREADS) {
        cst[i] = __ldg(p.w0 + 2 * i);
        cst[IG_C + i] = __ldg(p.w0 + 2 * i + 1);
        cst[2 * IG_C + i] = __ldg(p.b0 + i);
        cst[3 * IG_C + i] = __ldg(p.b1 + i);
        cst[4 * IG_C + i] = __ldg(p.b2 + i);
        cst[5 * IG_C + i] = __ldg(p.b3 + i);
        cst[6 * IG_C + i] = CONV ? __ldg(p.bc + i) : 0.f;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // tiles of this CTA: t = blockIdx.x + i * gridDim.x, i = 0 .. cnt-1; tile i lives in slot i & 1
    const uint32_t cnt = (P.total_tiles > blockIdx.x) ? (P.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(IG_REGS_LIGHT));
        // ================= weights: once per CTA
        if (elect_one()) {
            const uint32_t fb = smem_u32(wfull), w0a = smem_u32(wsm);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)(IG_W_MLP + (CONV ? IG_W_CONV : 0)))
                         : "memory");
            const CUtensorMap* maps[6] = {&m1h, &m1l, &m2h, &m2l, &m3h, &m3l};
#pragma unroll
            for (int l = 0; l < 3; ++l)
#pragma unroll
                for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        tma_load_2d_a(w0a + (uint32_t)(l * IG_W_LAYER + pl * 32768 + c * 16384), maps[2 * l + pl], fb, c * 64, 0);
            if (CONV) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    tma_load_2d_a(w0a + (uint32_t)(IG_W_MLP + c * 4096), &mch, fb, c * 16, 0);
                    tma_load_2d_a(w0a + (uint32_t)(IG_W_MLP + 12288 + c * 4096), &mcl, fb, c * 16, 0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(IG_REGS_LIGHT));
        // ================= MMA issuer: events of the two slots in fixed round-robin order
        const uint32_t ar = smem_u32(a_ready), dr = smem_u32(d_ready);
        // instruction descriptor: D = f32 (bit 4), A / B = bf16 (bits 7, 10), N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t wdesc0 = make_kmajor_desc(smem_u32(wsm), 64);
        const uint64_t cdesc0 = make_kmajor_desc(smem_u32(wsm) + IG_W_MLP, 16);
        mbar_wait_a(smem_u32(wfull), 0);
        uint32_t ph = 0;   // bit s: parity of the next a_ready[s] completion
        constexpr int EVENTS = CONV ? 5 : 3;
        for (uint32_t pr = 0; 2 * pr < cnt; ++pr) {
            for (int e = 0; e < EVENTS; ++e) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if (2 * pr + s >= cnt) continue;
                    mbar_wait_a(ar + 8u * s, (ph >> s) & 1u);
                    ph ^= 1u << s;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t X = tmem_base + (uint32_t)s * 256u, Y = X + 128u;
                        if (e < 3) {
                            const uint32_t A = (e == 1) ? Y : X, D = (e == 1) ? X : Y;
                            const uint64_t wh = wdesc0 + (uint64_t)((e * IG_W_LAYER) >> 4), wl = wh + (uint64_t)(32768 >> 4);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const uint64_t off = (uint64_t)(((j >> 2) * 16384) >> 4) + (uint64_t)((j & 3) * 2);
                                umma_ts(D, A + 16u * j + 8u, wh + off, idesc128, j > 0 ? 1u : 0u);   // a_lo * w_hi
                                umma_ts(D, A + 16u * j, wl + off, idesc128, 1u);                     // a_hi * w_lo
                                umma_ts(D, A + 16u * j, wh + off, idesc128, 1u);                     // a_hi * w_hi
                            }
                        } else {
                            // conv_first, output channels 64 * (e - 3) .. +63: A = X[0..48), D = X[64..128)
                            const uint64_t ch = cdesc0 + (uint64_t)(((e - 3) * 64 * 32) >> 4), cl = ch + (uint64_t)(12288 >> 4);
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const uint64_t off = (uint64_t)((j * 4096) >> 4);
                                umma_ts(X + 64u, X + 16u * j + 8u, ch + off, idesc64, j > 0 ? 1u : 0u);
                                umma_ts(X + 64u, X + 16u * j, cl + off, idesc64, 1u);
                                umma_ts(X + 64u, X + 16u * j, ch + off, idesc64, 1u);
                            }
                        }
                        umma_commit_a(dr + 8u * s);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(IG_REGS_LIGHT));
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(IG_REGS_EPI));
        // ================= epilogue: IG_EWG warp groups per slot.  Warp q of a group owns TMEM lanes 32q .. 32q+31 (lane = pixel);
        // group h of a slot handles the channel blocks [h * NB, (h + 1) * NB) of every phase (NB = 8 / IG_EWG blocks of 16 channels).
        const int e = warp - 4;
        const int s = (e >> 2) & 1, h = e >> 3, q = warp & 3;
        constexpr int NB = 8 / IG_EWG;
        const int b0 = h * NB;
        const uint32_t X = tmem_base + (uint32_t)s * 256u + ((uint32_t)(q * 32) << 16), Y = X + 128u;
        const uint32_t arb = smem_u32(&a_ready[s]), drb = smem_u32(&d_ready[s]);
        const uint32_t c_w0x = smem_u32(cst), c_w0y = c_w0x + 4 * IG_C, c_b0 = c_w0x + 8 * IG_C;
        int H = p.H, W = p.W, N = p.N;
        float slope = p.slope;
        const float* coord = p.coord; float* lsc = p.lsc; const float* raw = p.raw;
        uint16_t* fhi = reinterpret_cast<uint16_t*>(p.fea_hi); uint16_t* flo = reinterpret_cast<uint16_t*>(p.fea_lo);
        opaque(H); opaque(W); opaque(N); opaque(slope); opaque_ptr(coord); opaque_ptr(lsc); opaque_ptr(raw); opaque_ptr(fhi); opaque_ptr(flo);
        const long long HW = (long long)H * W;
        uint32_t dph = 0;
        const uint32_t tiles_x = (uint32_t)P.tiles_x, tiles_y = (uint32_t)P.tiles_y;
        // this lane's pixel of tile t: tile = 2 image rows x 64 pixels
        auto locate = [&](uint32_t t, int& n, int& yy, int& xx) {
            const int tx = (int)(t % tiles_x); t /= tiles_x;
            const int ty = (int)(t % tiles_y);
            n = (int)(t / tiles_y);
            yy = 2 * ty + (q >> 1);
            xx = 64 * tx + 32 * (q & 1) + lane;
        };
        auto load_coord = [&](int n, int yy, int xx, float& cx, float& cy) {
            const float* cp = coord + (long long)n * p.coord_bs + ((long long)yy * W + xx) * p.coord_ps;
            cx = __ldg(cp);
            cy = __ldg(cp + p.coord_cs);
        };
        float ncx = 0.f, ncy = 0.f;     // coordinates of the NEXT tile, fetched one tile ahead (a global-load latency per tile otherwise)
        if ((uint32_t)s < cnt) {
            int n, yy, xx;
            locate(blockIdx.x + (uint32_t)s * gridDim.x, n, yy, xx);
            load_coord(n, yy, xx, ncx, ncy);
        }
        for (uint32_t i = (uint32_t)s; i < cnt; i += 2) {
            int n, yy, xx;
            locate(blockIdx.x + i * gridDim.x, n, yy, xx);
            const long long pix = (long long)yy * W + xx;
            // ---- layer 0 on the CUDA cores -> A operand of layer 1 in X
            {
                const float cx = ncx, cy = ncy;
                if (i + 2 < cnt) {
                    int n2, y2, x2;
                    locate(blockIdx.x + (i + 2) * gridDim.x, n2, y2, x2);
                    load_coord(n2, y2, x2, ncx, ncy);
                }
#pragma unroll 2
                for (int bb = 0; bb < NB; ++bb) {
                    const int b = b0 + bb;
                    float val[16];
                    uint32_t pk[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 wx = lds4(c_w0x + 64u * b + 16u * g), wy = lds4(c_w0y + 64u * b + 16u * g),
                                     bv = lds4(c_b0 + 64u * b + 16u * g);
                        val[4 * g + 0] = fmaf(wy.x, cy, fmaf(wx.x, cx, bv.x));
                        val[4 * g + 1] = fmaf(wy.y, cy, fmaf(wx.y, cx, bv.y));
                        val[4 * g + 2] = fmaf(wy.z, cy, fmaf(wx.z, cx, bv.z));
                        val[4 * g + 3] = fmaf(wy.w, cy, fmaf(wx.w, cx, bv.w));
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) val[j] = fmaxf(val[j], val[j] * slope);   // LeakyReLU for 0 <= slope <= 1 (host-checked)
                    split_pack16(val, pk);
                    tmem_st16(X + 16u * b, pk);
                }
                tmem_wait_st();
                chain_arrive(arb, lane);
            }
            // ---- layers 1, 2: accumulator -> bias + LeakyReLU -> hi/lo pairs, in place
#pragma unroll 1
            for (int l = 0; l < 2; ++l) {
                const uint32_t D = ((l == 1) ? X : Y) + 16u * (uint32_t)b0;
                const uint32_t c_b = smem_u32(cst) + 4u * (uint32_t)((3 + l) * IG_C) + 64u * (uint32_t)b0;
                mbar_wait_a(drb, dph);
                dph ^= 1u;
                tc_fence_after();
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
                        const float4 bv = lds4(c_b + 64u * b + 16u * g);
                        val[4 * g + 0] = __uint_as_float(v[b & 1][4 * g + 0]) + bv.x;
                        val[4 * g + 1] = __uint_as_float(v[b & 1][4 * g + 1]) + bv.y;
                        val[4 * g + 2] = __uint_as_float(v[b & 1][4 * g + 2]) + bv.z;
                        val[4 * g + 3] = __uint_as_float(v[b & 1][4 * g + 3]) + bv.w;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) val[j] = fmaxf(val[j], val[j] * slope);
                    split_pack16(val, pk);
                    tmem_st16(D + 16u * b, pk);
                }
                tmem_wait_st();
                chain_arrive(arb, lane);
            }
            // ---- im2col operand of conv_first, gathered while layer 3 runs.  K index = tap * 4 + channel; k-step j holds taps 4j .. 4j+3
            // (zero beyond tap 8); group h writes the k-steps [j0, j0 + nj)
            const int j0 = (IG_EWG == 2 && h == 1) ? 2 : 0, nj = (IG_EWG == 1) ? 3 : (h == 0 ? 2 : 1);
            float4 tap[9];
            if (CONV) {
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const int y2 = yy + k / 3 - 1, x2 = xx + k % 3 - 1;
                    tap[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((k >> 2) >= j0 && (k >> 2) < j0 + nj && y2 >= 0 && y2 < H && x2 >= 0 && x2 < W)
                        tap[k] = __ldg(reinterpret_cast<const float4*>(raw + ((long long)n * HW + (long long)y2 * W + x2) * p.ldraw));
                }
            }
            // ---- layer 3 accumulator (lsc, in Y)
            mbar_wait_a(drb, dph);
            dph ^= 1u;
            tc_fence_after();
            const uint32_t c_b3 = smem_u32(cst) + 4u * (uint32_t)(5 * IG_C), c_bc = smem_u32(cst) + 4u * (uint32_t)(6 * IG_C);
            float* lbase = lsc + ((long long)n * IG_C) * HW + pix;
            if (!CONV) {
                float* lp = lbase + (long long)(16 * b0) * HW;
                const uint32_t D = Y + 16u * (uint32_t)b0;
                uint32_t v[2][16];
                tmem_ld16_async(D, v[0]);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    tmem_wait_ld16(v[b & 1]);
                    if (b < NB - 1) tmem_ld16_async(D + 16u * (b + 1), v[(b + 1) & 1]);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 bv = lds4(c_b3 + 64u * (uint32_t)(b0 + b) + 16u * g);
                        lp[0] = __uint_as_float(v[b & 1][4 * g + 0]) + bv.x; lp += HW;
                        lp[0] = __uint_as_float(v[b & 1][4 * g + 1]) + bv.y; lp += HW;
                        lp[0] = __uint_as_float(v[b & 1][4 * g + 2]) + bv.z; lp += HW;
                        lp[0] = __uint_as_float(v[b & 1][4 * g + 3]) + bv.w; lp += HW;
                    }
                }
                // the next tile of this slot starts by overwriting X; Y is rewritten by its layer 1, issued after the groups'
                // next a_ready arrivals -- program order of these warps covers both
            } else {
                // X is free (layer 3 has read it): conv A operand into X[0..48)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (j >= j0 && j < j0 + nj) {
                        float val[16];
                        uint32_t pk[16];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int k = 4 * j + g;
                            const float4 tv = k < 9 ? tap[k] : make_float4(0.f, 0.f, 0.f, 0.f);
                            val[4 * g + 0] = tv.x; val[4 * g + 1] = tv.y; val[4 * g + 2] = tv.z; val[4 * g + 3] = tv.w;
                        }
                        split_pack16(val, pk);
                        tmem_st16(X + 16u * j, pk);
                    }
                }
                tmem_wait_st();
                chain_arrive(arb, lane);
                // plane address of this pixel
                long long po;
                if (p.planes_s2) {
                    const long long plane = (long long)(((yy & 1) * 2 + (xx & 1)) * N + n);
                    po = ((plane * (H >> 1) + (yy >> 1)) * (W >> 1) + (xx >> 1)) * IG_C;
                } else {
                    po = ((long long)n * HW + pix) * IG_C;
                }
                constexpr int NBH = 4 / IG_EWG;      // blocks per group and 64-channel half
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    mbar_wait_a(drb, dph);
                    dph ^= 1u;
                    tc_fence_after();
                    const int cb = 64 * half + 16 * NBH * h;          // first channel of this group in this half
                    float* lp = lbase + (long long)cb * HW;
                    const uint32_t Dl = Y + (uint32_t)cb, Dc = X + 64u + (uint32_t)(16 * NBH * h);
                    uint32_t va[2][16], vc[2][16];
                    tmem_ld16_async(Dl, va[0]);
                    tmem_ld16_async(Dc, vc[0]);
#pragma unroll
                    for (int b = 0; b < NBH; ++b) {
                        tmem_wait_ld16(va[b & 1]);
                        tmem_wait_ld16(vc[b & 1]);
                        if (b < NBH - 1) {
                            tmem_ld16_async(Dl + 16u * (b + 1), va[(b + 1) & 1]);
                            tmem_ld16_async(Dc + 16u * (b + 1), vc[(b + 1) & 1]);
                        }
                        const int c0 = cb + 16 * b;
                        float val[16];
                        uint32_t pk[16];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 b3 = lds4(c_b3 + 4u * (uint32_t)c0 + 16u * g), bc = lds4(c_bc + 4u * (uint32_t)c0 + 16u * g);
                            const float a0 = __uint_as_float(va[b & 1][4 * g + 0]) + b3.x, a1 = __uint_as_float(va[b & 1][4 * g + 1]) + b3.y,
                                        a2 = __uint_as_float(va[b & 1][4 * g + 2]) + b3.z, a3 = __uint_as_float(va[b & 1][4 * g + 3]) + b3.w;
                            lp[0] = a0; lp += HW;
                            lp[0] = a1; lp += HW;
                            lp[0] = a2; lp += HW;
                            lp[0] = a3; lp += HW;
                            val[4 * g + 0] = (__uint_as_float(vc[b & 1][4 * g + 0]) + bc.x) * (a0 + 1.f);
                            val[4 * g + 1] = (__uint_as_float(vc[b & 1][4 * g + 1]) + bc.y) * (a1 + 1.f);
                            val[4 * g + 2] = (__uint_as_float(vc[b & 1][4 * g + 2]) + bc.z) * (a2 + 1.f);
                            val[4 * g + 3] = (__uint_as_float(vc[b & 1][4 * g + 3]) + bc.w) * (a3 + 1.f);
                        }
                        split_pack16(val, pk);
                        stg256(fhi + po + c0, pk);          // 16 channels = one full 32-byte sector per plane
                        stg256(flo + po + c0, pk + 8);
                    }
                    if (half == 0) chain_arrive(arb, lane);    // X[64..128) is drained: the second half may overwrite it
                }
                // The conv accumulator X[64..128) lies in the columns the slot's SECOND group overwrites first thing in the next tile
                // (layer 0 of its channel blocks): both groups of the slot must have drained it.
                if (IG_EWG == 2) {
                    if (s == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
                    else asm volatile("bar.sync 2, 256;" ::: "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// conv_first weight (128, 4, 3, 3) OIHW -> [128][48] bf16 hi / lo, K index = (ky*3 + kx) * 4 + c, zero beyond 36
__global__ void pack_ingest_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= IG_C * IG_KC) return;
    const int co = i / IG_KC, k = i % IG_KC;
    float v = 0.f;
    if (k < 36) v = w[(co * 4 + (k & 3)) * 9 + (k >> 2)];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = __bfloat16_as_ushort(h);
    lo[i] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
}

template <bool CONV>
auto ingest_variant() {
    static bool attr_set[MAX_DEVICES] = {};
    auto k = ingest_kernel<CONV>;
    const int dev = current_device();
    if (!attr_set[dev]) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set[dev] = true;
    }
    return k;
}

}  // namespace
}  // namespace rcn

using namespace rcn;

extern "C" int rcn_pack_ingest_weight(const float* w_oihw, void* hi, void* lo, void* stream) {
    RCN_CHECK_ARG(w_oihw && hi && lo, "rcn_pack_ingest_weight: null pointer");
    pack_ingest_weight_kernel<<<(IG_C * IG_KC + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_oihw, (uint16_t*)hi, (uint16_t*)lo);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_pack_ingest_weight");
    return RCN_OK;
}

extern "C" int rcn_ingest_fused(const rcn_ingest_desc* d, void* stream) {
    RCN_CHECK_ARG(d && d->coord && d->w0 && d->b0 && d->b1 && d->b2 && d->b3 && d->lsc, "rcn_ingest_fused: null pointer");
    RCN_CHECK_ARG(d->w1_hi && d->w1_lo && d->w2_hi && d->w2_lo && d->w3_hi && d->w3_lo, "rcn_ingest_fused: null weight plane");
    RCN_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->H % 2 == 0 && d->W % 64 == 0,
                  "rcn_ingest_fused: needs even H and W %% 64 == 0 (got %d x %d)", d->H, d->W);
    RCN_CHECK_ARG(d->slope >= 0.f && d->slope <= 1.f, "rcn_ingest_fused: LeakyReLU slope must be in [0, 1]");
    const bool conv = d->raw != nullptr;
    if (conv) {
        RCN_CHECK_ARG(d->wc_hi && d->wc_lo && d->bc && d->fea_hi && d->fea_lo, "rcn_ingest_fused: the fused conv_first needs its weights, bias and output planes");
        RCN_CHECK_ARG(d->ldraw >= 4 && d->ldraw % 4 == 0 && (reinterpret_cast<uintptr_t>(d->raw) & 15) == 0,
                      "rcn_ingest_fused: raw must be NHWC with 4 channels, 16-byte aligned pixels");
        RCN_CHECK_ARG((reinterpret_cast<uintptr_t>(d->fea_hi) & 31) == 0 && (reinterpret_cast<uintptr_t>(d->fea_lo) & 31) == 0,
                      "rcn_ingest_fused: output planes must be 32-byte aligned");
    }
    RCN_CHECK_ARG(get_encode() != nullptr, "rcn_ingest_fused: cuTensorMapEncodeTiled is not available from the driver");
    IngestParams P;
    P.d = *d;
    P.tiles_x = d->W / 64;
    P.tiles_y = d->H / 2;
    const long long total = (long long)P.tiles_x * P.tiles_y * d->N;
    RCN_CHECK_ARG(total < (1ll << 31), "rcn_ingest_fused: too many tiles");
    P.total_tiles = (uint32_t)total;
    CUtensorMap m[8];
    const void* wp[6] = {d->w1_hi, d->w1_lo, d->w2_hi, d->w2_lo, d->w3_hi, d->w3_lo};
    bool ok = true;
    for (int i = 0; i < 6; ++i) ok = ok && make_w_map(&m[i], wp[i], IG_C, IG_C, IG_C, 64);
    if (conv) ok = ok && make_w_map(&m[6], d->wc_hi, IG_C, IG_KC, IG_C, 16) && make_w_map(&m[7], d->wc_lo, IG_C, IG_KC, IG_C, 16);
    else { m[6] = m[0]; m[7] = m[1]; }
    RCN_CHECK_ARG(ok, "rcn_ingest_fused: cuTensorMapEncodeTiled failed");
    const size_t smem = (size_t)IG_W_MLP + (conv ? IG_W_CONV : 0) + IG_CONST_FLOATS * 4 + 128 + 1024;
    const int sms = sm_count();
    const unsigned grid = (unsigned)(total < sms ? total : sms);
    if (conv) ingest_variant<true>()<<<grid, IG_THREADS, smem, (cudaStream_t)stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], P);
    else ingest_variant<false>()<<<grid, IG_THREADS, smem, (cudaStream_t)stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], P);
    count_launch();
    RCN_CHECK_LAUNCH("rcn_ingest_fused");
    return RCN_OK;
}
