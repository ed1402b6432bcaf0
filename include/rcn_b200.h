/* rcn_b200.h -- C ABI of the B200-native RealCamNet hot path (librcn_b200.so).
 *
 * The reference (kepengxu/RealCamNet) has no FFI/plugin registry: its boundary is the Python
 * nn.Module API.  This header is therefore the interface a maintainer
 * binds from Python (ctypes) so that the reference model classes keep
 * their forward()/compress()/decompress() signatures while every tensor op below runs as a
 * hand-written sm_100a kernel.  Each entry cites the reference code it replaces
 * (paths relative to the reference repo root).
 *
 * Conventions
 *   - every pointer named x/y/w/... is a DEVICE pointer unless the comment says "host";
 *   - activations are NHWC fp32: element (n,h,w,c) at  p[((n*H + h)*W + w)*ld + c],
 *     where ld >= C is the pixel stride in elements (lets an op read/write a channel slice
 *     of a wider concat buffer -- torch.cat/torch.split on the reference side);
 *   - `stream` is a cudaStream_t passed as void* (the caller -- PyTorch -- owns it);
 *   - return value: 0 on success, negative RCN_ERR_* otherwise; rcn_last_error() gives a
 *     thread-local message.  Nothing throws, nothing allocates device memory.
 *   - re-entrant per stream; the only global state is the launch counter.
 */
#ifndef RCN_B200_H
#define RCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCN_OK 0
#define RCN_ERR_INVALID (-1)
#define RCN_ERR_CUDA (-2)
#define RCN_ERR_NOMEM (-3)
#define RCN_ERR_UNSUPPORTED (-4)

const char* rcn_last_error(void);
int rcn_version(void);
/* number of kernels this library has launched so far in this process (bench.py gpu_launches) */
unsigned long long rcn_launch_count(void);

/* ---- activations applied in epilogues ------------------------------------------------ */
enum {
    RCN_ACT_NONE = 0,
    RCN_ACT_RELU = 1,      /* nn.ReLU */
    RCN_ACT_LRELU = 2,     /* nn.LeakyReLU(slope) */
    RCN_ACT_GELU = 3,      /* nn.GELU (erf form) */
    RCN_ACT_SIGMOID = 4,
    RCN_ACT_HALF_TANH = 5, /* 0.5*tanh(v): LRP, models/raw2bit.py:1837 */
    RCN_ACT_HSWISH = 6,    /* nn.Hardswish, models/groupmix.py:65-77 */
    RCN_ACT_CLAMP01 = 7    /* .clamp_(0, 1) of decompress(), models/raw2bit.py:2025 */
};

/* ---- epilogue combinators of rcn_conv2d ------------------------------------------------ */
enum {
    RCN_EPI_NONE = 0,
    RCN_EPI_GDN = 1,         /* v = aux * rsqrt(v)  (GDN, compressai.layers.GDN; raw2bit.py:1642) */
    RCN_EPI_IGDN = 2,        /* v = aux * sqrt(v)   (inverse GDN in ResidualBlockUpsample) */
    RCN_EPI_MUL_AUXP1 = 3,   /* v = v * (aux + 1)   (fea*(lsc_fea+1), raw2bit.py:1780) */
    RCN_EPI_MULP1_AUX = 4,   /* v = (v + 1) * aux   (x*scale + x of SpatialFeatureTransform, raw2bit.py:878-886) */
    RCN_EPI_SIGMOID_GATE = 5 /* v = aux * sigmoid(v) (AttentionBlock a*sigmoid(b), tcm.py:286-288) */
};

enum {
    RCN_STORE_NHWC = 0,
    RCN_STORE_PS2 = 1,      /* fused nn.PixelShuffle(2): channel co -> (c=co/4, i=(co%4)/2, j=co%2) */
    RCN_STORE_NCHW = 2,     /* plain NCHW contiguous output (API boundary tensors) */
    RCN_STORE_PS2_NCHW = 3  /* pixel shuffle then NCHW (final x_hat, raw2bit.py:1682) */
};

/* One 2-D convolution / linear layer with a fused epilogue.
 * Replaces every nn.Conv2d(k in {1,3}, stride in {1,2}, padding=k//2) and nn.Linear on the path
 * (F.conv2d call sites: models/raw2bit.py:1641-1754, models/tcm.py:130-137,252-253,
 * models/LiteISP.py:23-30,363-378,537-559, models/networks.py:146-221) plus the element-wise
 * op that follows it in the reference graph.
 *   v = sum_k x*w + bias[c]
 *   if (cscale) v = v*(1 + cscale[n][c]) + cshift[n][c]      (Res_GFM, LiteISP.py:552-555)
 *   v = EPI(v, aux)
 *   if (res && res_pre)  v += rs*res ;  v = ACT(v) ;  if (res && !res_pre) v += rs*res     (rs = res_scale)
 */
typedef struct rcn_conv_desc {
    const float* x; int N, H, W, Cin, ldx;
    const float* w;      /* repacked weights [k*k*Cin][Cout], row = (ky*k+kx)*Cin + ci */
    const float* bias;   /* [Cout] or NULL */
    int k, stride, Cout;
    int in_square;       /* feed x*x to the contraction (GDN norm pool) */
    float* y; int ldy;   /* pixel stride of the OUTPUT tensor (after pixel shuffle if any) */
    int store;           /* RCN_STORE_* */
    int epi; const float* aux; int ldaux;      /* aux has the conv's own output geometry (N,Ho,Wo,Cout) */
    const float* cscale; const float* cshift;  /* [N][Cout] or NULL */
    const float* res; int ldres; int res_pre;  /* res has the geometry of the stored output */
    int act; float slope;
    float res_scale;     /* multiplies res (2.0 for `conv_block(x) + x` of ConvTransBlock, models/tcm.py:262) */
    /* optional (tcgen05 engine, RCN_STORE_NHWC / RCN_STORE_PS2 with Cp_out % 64 == 0 stored channels): also write the result
     * as the NEXT layer's bf16 hi/lo operand planes (pixel stride Cp_out), so that layer needs no rcn_split_bf16 pass.
     * y may then be NULL (planes only: the fp32 tensor of a conv->conv chain is never materialised). */
    void* y_hi; void* y_lo; int Cp_out;   /* Cp_out = pixel stride of the emitted planes (>= stored channels: a channel slice of a
                                           * wider plane buffer, e.g. one half of a concat the next layer reads) */
    int ldp_in;          /* rcn_conv2d_tc: pixel stride of the INPUT planes x_hi/x_lo in elements (0 = Cp, dense planes) */
    int planes_s2;       /* emitted planes y_hi/y_lo use the polyphase layout of rcn_split_bf16_s2 ((4N, H/2, W/2, Cp_out): the next layer
                          * is a stride-2 conv); stride-1 NHWC layers with even H, W only */
    int ps_perm;         /* rcn_conv2d_tc with RCN_STORE_PS2: w_hi/w_lo were packed with ps_perm = 1 (rows grouped by sub-pixel),
                          * which lets the pixel-shuffle store write 64 contiguous bytes per pixel like the NHWC store */
    int aux_nchw;        /* aux is a contiguous NCHW tensor (N,Cout,Ho,Wo) instead of NHWC with pixel stride ldaux: lets an API-facing
                          * map (the lens-shading features returned by forward(), models/raw2bit.py:1776,1853) be written once, in
                          * the layout the caller gets, and still feed `fea * (lsc + 1)` */
    int planes_square;   /* emitted planes hold the SQUARE of the result (the next layer is a GDN norm pool, in_square) */
    int in_fmt;          /* rcn_conv2d_tc: element format RCN_PLANE_* of the INPUT planes x_hi/x_lo and of the packed weights */
    int out_fmt;         /* element format RCN_PLANE_* of the emitted planes y_hi/y_lo (what the NEXT layer's in_fmt will be) */
} rcn_conv_desc;

/* 16-bit element formats of the tcgen05 operand planes.  bf16 hi+lo (3 MMA passes) is the parity engine everywhere a
 * quantisation decision depends on the result; fp16 (1 pass, 11 significand bits) is used by the per-stage precision policy for
 * the post-quantisation synthesis tail, where only the 1e-3 bar on x_hat applies (models/raw2bit.py:1680-1682). */
enum { RCN_PLANE_BF16 = 0, RCN_PLANE_F16 = 1 };

int rcn_conv2d(const rcn_conv_desc* d, void* stream);

/* Repack an OIHW nn.Conv2d / (O,I) nn.Linear weight into the [k*k*Cin][Cout] layout above. */
int rcn_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int k, float* out, void* stream);

/* ---- normalisation / attention --------------------------------------------------------- */
/* nn.LayerNorm(C, eps) over the channel dim of an NHWC/token tensor (+ optional activation:
 * Agg_0 applies Hardswish right after its LayerNorm, models/groupmix.py:47-53).
 * models/tcm.py:223,227,233-234; models/groupmix.py:279,285,291,296. */
int rcn_layernorm(const float* x, long long npix, int C, int ldx, const float* gamma, const float* beta,
                  float eps, float* y, int ldy, int act, void* y_hi, void* y_lo, int ldp, void* stream);
/* y_hi / y_lo (optional, pixel stride ldp elements): also / only write the result as the next layer's bf16 operand planes
 * (LayerNorm feeds embedding_layer / mlp[0], models/tcm.py:233-234); y may be NULL when y_hi is given. */

/* Swin window attention core of WMSA.forward (models/tcm.py:179-207): cyclic shift, window
 * partition, per-head softmax(q k^T * hd^-0.5 + relative-position bias [+ shift mask]) v,
 * un-partition and shift back.  qkv is the output of embedding_layer laid out
 * [q: head0..|k: head0..|v: head0..] per pixel (models/tcm.py:192-193); relpos is the
 * (heads, 2ws-1, 2ws-1) relative_position_params tensor (models/tcm.py:155,209-212).
 * The mask of generate_mask (models/tcm.py:160-177) is computed from indices, never stored. */
int rcn_wmsa(const float* qkv, int N, int H, int W, int C, int ldq, int head_dim, int ws, int shifted,
             const float* relpos, float* out, int ldo, void* out_hi, void* out_lo, int ldp, void* stream);
/* out_hi / out_lo (optional): operand planes of the attention output for WMSA.linear (models/tcm.py:206); out may be NULL. */

/* ---- layout changes at the API boundary (reference tensors are NCHW) ------------------------ */
int rcn_nchw_to_nhwc(const float* x, int N, int C, int H, int W, float* y, int ldy, void* stream);
int rcn_nhwc_to_nchw(const float* x, int ldx, int N, int C, int H, int W, float* y, void* stream);
/* torch.cat / torch.split helper: copy a channel slice */
int rcn_copy_channels(const float* x, int ldx, long long npix, int C, float* y, int ldy, void* stream);

/* ---- pooling / gating / statistics ----------------------------------------------------------- */
/* nn.AdaptiveAvgPool2d(1): mean[n][c] (CALayer models/networks.py:259,268, models/raw2bit.py:241,251;
 * Color_Condition_GFM models/LiteISP.py:357).  Deterministic two-stage sum; workspace >= N*1024*C floats
 * is always enough (fewer chunks are used when it is smaller). */
int rcn_channel_mean(const float* x, int N, long long HW, int C, int ldx, float* mean, float* workspace,
                     long long workspace_floats, void* stream);
/* per-(n,c) mean and biased variance (nn.InstanceNorm2d statistics, models/LiteISP.py:28-29) */
int rcn_channel_meanvar(const float* x, int N, long long HW, int C, int ldx, float* mean, float* var, void* stream);
/* y = (x - mean[n,c]) * rsqrt(var[n,c] + eps) * gamma[c] + beta[c] */
int rcn_norm_apply(const float* x, int ldx, int N, long long HW, int C, const float* mean, const float* var,
                   const float* gamma, const float* beta, float eps, float* y, int ldy, void* stream);
/* y = act(x * g + b) + r ; g,b indexed [n][c] (per_n=1: CALayer "x * y", models/networks.py:270) or [c]
 * (per_n=0: folded eval-mode SyncBatchNorm + Hardswish of the GroupMix Aggregator, models/groupmix.py:96-99) */
int rcn_scale_add(const float* x, int ldx, int N, long long HW, int C, const float* g, const float* b, int per_n,
                  const float* r, int ldr, float* y, int ldy, int act, void* stream);
/* nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=True) + nn.LeakyReLU(slope): color_block, models/LiteISP.py:23-30 */
int rcn_avgpool3s2_lrelu(const float* x, int N, int H, int W, int C, int ldx, float slope, float* y, int ldy, void* stream);
/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True): HyCondModDecBlock, models/raw2bit.py:793-796 */
int rcn_upsample_bilinear2x(const float* x, int N, int H, int W, int C, int ldx, float* y, int ldy, void* stream);
/* same, written as the bf16 hi/lo operand planes (pixel stride ldp) of the conv that follows (HyCondModDecBlock.up[1]) */
int rcn_upsample_bilinear2x_planes(const float* x, int N, int H, int W, int C, int ldx, void* y_hi, void* y_lo, int ldp, void* stream);
/* Haar DWTForward / DWTInverse as fixed 2x2 stride-2 (transposed) grouped convs, channel order
 * [LL,LH,HL,HH] per input channel: models/networks.py:224-249 */
int rcn_dwt_forward(const float* x, int N, int H, int W, int C, int ldx, float* y, int ldy, void* stream);
int rcn_dwt_inverse(const float* x, int N, int H, int W, int C4, int ldx, float* y, int ldy, void* stream);
/* space-to-depth by 2: y[n, ho, wo, (i*2+j)*C + c] = x[n, 2ho+i, 2wo+j, c].  With weights re-ordered to (O, i, j, c) the learned
 * nn.Conv2d(C, O, 2, 2) down-samplers of ISPUNet_GFM_LSC / ResUNet (models/LiteISP.py:1253,1265,1278,2056,2065,2075) become 1x1
 * contractions on the conv engines. */
int rcn_space_to_depth2(const float* x, int N, int H, int W, int C, int ldx, float* y, int ldy, void* stream);
/* depthwise k x k conv (k odd, padding k//2), weights [k*k][C]; add_input: "+ feat" of ConvPosEnc
 * (models/groupmix.py:213-215); mul: "q * conv(v)" of ConvRelPosEnc (models/groupmix.py:146-154) */
int rcn_depthwise_conv(const float* x, int N, int H, int W, int C, int ldx, const float* w, const float* bias, int k,
                       int add_input, const float* mul, int ldm, float* y, int ldy, void* stream);

/* ---- entropy models ---------------------------------------------------------------------------- */
/* EntropyBottleneck.forward in eval mode + quantize("symbols") (compressai.entropy_models, call sites
 * models/raw2bit.py:1803-1807,1906): z_hat = round(z - median) + median, likelihood =
 * max(|sigmoid(s*upper) - sigmoid(s*lower)|, lik_bound) with the per-channel 1-3-3-3-3-1 cumulative.
 * params: [C][58] = per layer softplus(matrix), bias, tanh(factor) (see entropy.cu); symbols are
 * written in NCHW order (the order EntropyModel.compress flattens them).  Any output may be NULL. */
int rcn_eb_forward(const float* z, int ldz, int N, long long HW, int C, const float* params, const float* medians,
                   float* z_hat, int ldzh, float* lik, int ldl, int* symbols, float lik_bound, void* stream);
int rcn_eb_dequantize(const int* symbols, int N, long long HW, int C, const float* medians, float* z_hat, int ldzh, void* stream);
/* GaussianConditional on one latent slice (models/raw2bit.py:1829-1831 forward; 1939-1941 compress):
 * y_hat = round(y - mu) + mu ; likelihood = max(Phi((.5-|y_hat-mu|)/s) - Phi((-.5-|y_hat-mu|)/s), lik_bound),
 * s = max(scale, scale_bound), Phi via erfc ; symbols = int(round(y - mu)) ; indexes =
 * (ntable-1) - #{t in table[:-1] : s <= t} (build_indexes).  symbols/indexes in NCHW order. */
int rcn_gaussian_conditional(const float* y, int ldy, const float* mu, int ldm, const float* scale, int lds, int N,
                             long long HW, int C, const float* table, int ntable, float scale_bound, float lik_bound,
                             float* y_hat, int ldyh, float* lik, int ldl, int* symbols, int* indexes, void* stream);
/* Same kernel, plus the range coder's per-symbol front end executed on the GPU: for every symbol the CDF row
 * `indexes[o]` of the (device-resident) integer tables is looked up and packed[o] = (start << 16) | (freq - 1) is
 * written in coding order; symbols outside the row's range take the sentinel bin, get flags[o] = 1 and leave their
 * bypass payload in raw[o].  This is the work BufferedRansEncoder.encode_with_indexes does per symbol on the host
 * (models/raw2bit.py:1956); only the serial state chain (rcn_rans_encode_packed) is left for the CPU. */
int rcn_gaussian_conditional_coded(const float* y, int ldy, const float* mu, int ldm, const float* scale, int lds, int N,
                                   long long HW, int C, const float* table, int ntable, float scale_bound,
                                   float lik_bound, float* y_hat, int ldyh, float* lik, int ldl, int* symbols,
                                   int* indexes, const int* cdf, int cdf_stride, const int* cdf_len, const int* cdf_off,
                                   unsigned* packed, unsigned* raw, unsigned char* flags, void* stream);
/* decoder side: indexes only (models/raw2bit.py:2011) and y_hat = symbol + mu (models/raw2bit.py:2014-2015) */
int rcn_build_indexes(const float* scale, int lds, int N, long long HW, int C, const float* table, int ntable,
                      float scale_bound, int* indexes, void* stream);
int rcn_gaussian_dequantize(const int* symbols, const float* mu, int ldm, int N, long long HW, int C, float* y_hat,
                            int ldyh, void* stream);

/* ---- range coder (HOST pointers) ------------------------------------------------------------------ */
/* BufferedRansEncoder.encode_with_indexes + flush as ONE call (models/raw2bit.py:1921,1956-1957):
 * returns the number of bytes written to out (little-endian uint32 words) or a negative error. */
long long rcn_rans_encode(const int32_t* symbols, const int32_t* indexes, long long n, const int32_t* cdfs,
                          int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* out,
                          long long out_cap);
/* Back end of the same coder for symbols pre-digested by rcn_gaussian_conditional_coded: HOST arrays packed[n],
 * raw[n], flags[n].  Produces the identical byte stream as rcn_rans_encode on (symbols, indexes). */
long long rcn_rans_encode_packed(const uint32_t* packed, const uint32_t* raw, const uint8_t* flags, long long n,
                                 uint8_t* out, long long out_cap);
/* RansDecoder.set_stream / decode_stream (models/raw2bit.py:1996-1997,2013): state persists across calls */
typedef struct rcn_rans_decoder rcn_rans_decoder;
rcn_rans_decoder* rcn_rans_decoder_create(const uint8_t* stream, long long nbytes);
/* indexes[i] must lie in [0, n_rows); a corrupt / truncated stream or an out-of-range index returns RCN_ERR_INVALID */
int rcn_rans_decode(rcn_rans_decoder* d, const int32_t* indexes, long long n, const int32_t* cdfs, int cdf_stride,
                    int n_rows, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* out);
void rcn_rans_decoder_destroy(rcn_rans_decoder* d);
/* compressai._CXX.pmf_to_quantized_cdf: cdf has n+1 entries (used by update(), models/raw2bit.py:1759-1764) */
int rcn_pmf_to_quantized_cdf(const float* pmf, int n, int precision, int32_t* cdf);

/* ---- GroupMix efficient attention (models/groupmix.py:186-196) ------------------------------------ */
/* q, k, v: NHWC channel slices with heads*Ch channels (head-major, as produced by Aggregator.forward,
 * models/groupmix.py:103).  Computes k.softmax(dim=N) over ALL tokens, kv = softmax(k)^T v per head,
 * out = scale * q kv + crpe  (crpe = q * dwconv(v), ConvRelPosEnc; may be NULL).  kv_out: [B][heads][Ch][Ch].
 * Two passes over the token map (k,v then q): 3*heads*Ch*4 bytes/token of compulsory traffic. */
long long rcn_groupmix_workspace_floats(int B, long long HW, int heads, int Ch);
int rcn_groupmix_attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                           const float* crpe, int ldc, int B, long long HW, int heads, int Ch, float scale,
                           float* out, int ldo, float* kv_out, float* workspace, long long workspace_floats,
                           void* stream);

/* ---- tcgen05 / TMA convolution engine ---------------------------------------------------------------- */
/* Same contraction + fused epilogue as rcn_conv2d (d->x, d->w, d->in_square are ignored), but the operands
 * are bf16 planes fetched by TMA and multiplied by tcgen05.mma with fp32 accumulators in TMEM.
 *   x_hi/x_lo : (N,H,W,Cp) bf16 planes from rcn_split_bf16 (x = hi + lo; lo may be NULL when passes == 1)
 *   w_hi/w_lo : [Cout][k*k][Cp] bf16 from rcn_pack_conv_weight_tc
 *   passes    : 1 = bf16 (hi*hi) ; 3 = "bf16x3" (hi*hi + lo*hi + hi*lo, ~fp32-grade products)
 * stride 1: planes from rcn_split_bf16; stride 2 (even H, W): planes from rcn_split_bf16_s2.
 * Cp is 16 (Cin <= 16), 32 (Cin <= 32) or a multiple of 64: it is also the K chunk the kernel stages per pipeline step. */
int rcn_conv2d_tc(const rcn_conv_desc* d, const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                  int Cp, int passes, void* stream);
/* fp32 NHWC (pixel stride ldx) -> zero-padded bf16 hi/lo planes; square != 0 feeds x*x (GDN norm pool) */
/* fmt = RCN_PLANE_BF16 | RCN_PLANE_F16: element format of the planes written (lo may be NULL: single-pass consumers) */
int rcn_split_bf16(const float* x, int ldx, long long npix, int C, int Cp, int square, int fmt, void* hi, void* lo, void* stream);
/* stride-2 layers: the four polyphase planes x[:, py::2, px::2, :] as (4N, H/2, W/2, Cp) bf16 hi/lo, plane index
 * (py*2+px)*N + n -- each filter tap of a stride-2 conv then reads ONE plane at unit stride (TMA box per tap). */
int rcn_split_bf16_s2(const float* x, int ldx, int N, int H, int W, int C, int Cp, int fmt, void* hi, void* lo, void* stream);
/* OIHW fp32 weight -> [Cout][k*k][Cp] bf16 hi/lo (K-major rows of the B operand).  ps_perm != 0 (Cout % 64 == 0): row
 * 64g + 16s + c holds conv channel 64g + 4c + s, i.e. the four PixelShuffle(2) sub-pixels s of 16 shuffled channels c. */
int rcn_pack_conv_weight_tc(const float* w_oihw, int Cout, int Cin, int k, int Cp, int ps_perm, int fmt, void* hi, void* lo, void* stream);

/* ---- fused packed-Bayer ingest ------------------------------------------------------------------------- */
/* models/raw2bit.py:1771-1780 with models/LiteISP.py:363-378 as ONE kernel (csrc/ingest.cu):
 *   lsc = Lens_Shading_Correction(coord)   4-layer per-pixel MLP 2 -> 128 -> 128 -> 128 -> 128, LeakyReLU(slope) between layers
 *   fea = conv_first(raw) * (lsc + 1)      3x3, 4 -> 128, padding 1   (optional: raw == NULL computes lsc only)
 * Hidden maps never leave the SM (activations are the tcgen05 A operand in tensor memory); lsc is written once as the
 * contiguous NCHW map forward() returns (raw2bit.py:1853) and fea once, as the bf16 hi/lo operand planes of its consumer
 * conv_down (planes_s2 != 0: polyphase layout of rcn_split_bf16_s2; else (N,H,W,128)).  bf16x3 arithmetic.
 * Requirements: 128-wide layers, even H, W % 64 == 0 (the codec's tiles are multiples of 64). */
typedef struct rcn_ingest_desc {
    const float* coord;            /* 2 channels; element (n, y, x, c) at coord[n*coord_bs + (y*W + x)*coord_ps + c*coord_cs] */
    long long coord_bs; int coord_ps, coord_cs;
    const float* w0; const float* b0;                 /* layer 0: nn.Conv2d(2, 128, 1) weight [128][2] and bias */
    const void* w1_hi; const void* w1_lo;             /* layers 1-3: [128][128] bf16 hi/lo from rcn_pack_conv_weight_tc (Cp = 128) */
    const void* w2_hi; const void* w2_lo;
    const void* w3_hi; const void* w3_lo;
    const float* b1; const float* b2; const float* b3;
    float slope;
    float* lsc;                    /* out: (N,128,H,W) fp32, contiguous */
    int N, H, W;
    const float* raw; int ldraw;   /* NHWC packed-Bayer tile (N,H,W,4), pixel stride ldraw floats; NULL: lens-shading map only */
    const void* wc_hi; const void* wc_lo;             /* conv_first: [128][48] bf16 hi/lo from rcn_pack_ingest_weight */
    const float* bc;
    void* fea_hi; void* fea_lo;    /* out: bf16 operand planes of fea, pixel stride 128 */
    int planes_s2;
} rcn_ingest_desc;
int rcn_ingest_fused(const rcn_ingest_desc* d, void* stream);
/* conv_first weight (128,4,3,3) OIHW -> [128][48] bf16 hi/lo, K index = (ky*3 + kx)*4 + c (zero beyond 36) */
int rcn_pack_ingest_weight(const float* w_oihw, void* hi, void* lo, void* stream);

/* ---- fused Swin MLP ------------------------------------------------------------------------------------- */
/* models/tcm.py:225-236: y = res + fc2(GELU(fc1(x))) for C = 64, hidden = 256 as ONE kernel (csrc/mlp.cu): x comes as the bf16
 * hi/lo operand planes the LayerNorm kernel emits, the 4C-wide hidden activations live in tensor memory only (fc2 reads them
 * as its tcgen05 A operand), the result is written as fp32 rows (y, pixel stride ldy) and / or as the consumer's operand planes
 * (y_hi / y_lo, pixel stride Cp_out).  With x_ln the LayerNorm in front (self.ln2) is part of the kernel too.  bf16x3 arithmetic. */
typedef struct rcn_mlp_desc {
    const void* x_hi; const void* x_lo; int ldp_in;   /* (npix, C) bf16 planes, pixel stride ldp_in elements */
    long long npix;
    int C, hidden;
    const void* w1_hi; const void* w1_lo; const float* b1;   /* fc1: [hidden][C] bf16 hi/lo (rcn_pack_conv_weight_tc, Cp = C), bias */
    const void* w2_hi; const void* w2_lo; const float* b2;   /* fc2: [C][hidden] */
    const float* res; int ldres;                      /* residual rows (may be NULL) */
    float* y; int ldy;                                /* may be NULL when y_hi is given */
    void* y_hi; void* y_lo; int Cp_out;               /* may be NULL */
    /* optional: x_ln != NULL makes the kernel apply nn.LayerNorm(C, eps) itself (models/tcm.py:234: self.ln2) on fp32 rows x_ln
     * (pixel stride ldx) instead of reading x_hi / x_lo: no rcn_layernorm launch, no LayerNorm planes in HBM */
    const float* x_ln; int ldx; const float* gamma; const float* beta; float eps;
} rcn_mlp_desc;
int rcn_mlp_fused(const rcn_mlp_desc* d, void* stream);

/* ---- fused LayerNorm + Linear ----------------------------------------------------------------------------- */
/* y = Linear(nn.LayerNorm(C)(x)) for C = 64 as ONE kernel (csrc/lnlinear.cu): models/tcm.py:233 + 193, the qkv embedding of the
 * Swin blocks (`self.msa(self.ln1(x))` -> `self.embedding_layer`).  The normalised row never leaves the SM: it is written as the
 * bf16 hi/lo tcgen05 A operand into tensor memory.  x, y: fp32 rows (pixel strides ldx, ldy); w_hi / w_lo: [Cout][64] bf16
 * (rcn_pack_conv_weight_tc, Cp = 64); Cout a multiple of 16, <= 192; bias may be NULL.  bf16x3 arithmetic. */
typedef struct rcn_lnlinear_desc {
    const float* x; int ldx; long long npix;
    int C, Cout;
    const float* gamma; const float* beta; float eps;
    const void* w_hi; const void* w_lo; const float* bias;
    float* y; int ldy;
} rcn_lnlinear_desc;
int rcn_ln_linear_fused(const rcn_lnlinear_desc* d, void* stream);

/* perf triage only (RCN_TC_DEBUG bit 128): cycles one epilogue warp of CTA 0 spent {waiting for accumulators, working},
 * tiles seen, 0.  reset != 0 clears the counters. */
int rcn_tc_prof(unsigned long long* out16, int reset);

#ifdef __cplusplus
}
#endif
#endif /* RCN_B200_H */
