"""Tensor-level wrappers over the C ABI (include/rcn_b200.h).

PyTorch is used for device memory (torch.empty), the current CUDA stream and parameter storage
only; every arithmetic op on the path is a kernel of librcn_b200.so.  Internal activations are
NHWC fp32 torch tensors (N,H,W,C) -- possibly channel-slice views of a wider buffer, which is
how torch.cat/torch.split of the reference are expressed without copies.
"""
from __future__ import annotations

import ctypes
import os
import weakref

import torch

from . import _C

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_GELU, ACT_SIGMOID, ACT_HALF_TANH, ACT_HSWISH, ACT_CLAMP01 = range(8)
EPI_NONE, EPI_GDN, EPI_IGDN, EPI_MUL_AUXP1, EPI_MULP1_AUX, EPI_SIGMOID_GATE = range(6)
STORE_NHWC, STORE_PS2, STORE_NCHW, STORE_PS2_NCHW = range(4)

_vp = ctypes.c_void_p


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _on_current_device(t, name="tensor"):
    """Kernels are launched on the CURRENT device's current stream (one process per GPU): a tensor of another device would be
    dereferenced by the wrong GPU, so that is an error here rather than an illegal address (or a silent peer access) later."""
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"{name}: tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()} "
                           "(use torch.cuda.set_device / torch.cuda.device around the call)")


def _chk(t, name="tensor"):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise TypeError(f"{name}: expected a CUDA float32 tensor, got {t.device} {t.dtype}")
    _on_current_device(t, name)


def geom(t, name="tensor"):
    """(N, H, W, C, ld) of an NHWC tensor / channel-slice view."""
    _chk(t, name)
    if t.dim() != 4:
        raise ValueError(f"{name}: expected 4 dims (N,H,W,C), got {tuple(t.shape)}")
    N, H, W, C = t.shape
    if C > 1 and t.stride(3) != 1:
        raise ValueError(f"{name}: channel dim must be contiguous")
    if W > 1:
        ld = t.stride(2)
    elif H > 1:
        ld = t.stride(1)
    elif N > 1:
        ld = t.stride(0)
    else:
        ld = C
    if (H > 1 and t.stride(1) != W * ld) or (N > 1 and t.stride(0) != H * W * ld) or ld < C:
        raise ValueError(f"{name}: not an NHWC tensor or channel slice (shape {tuple(t.shape)}, strides {t.stride()})")
    return N, H, W, C, ld


def empty(N, H, W, C, like=None, device=None):
    dev = like.device if like is not None else device
    return torch.empty((N, H, W, C), device=dev, dtype=torch.float32)


_replayed = 0


def launch_count() -> int:
    """Kernels launched by the library in this process, plus the launches replayed through CUDA graphs."""
    return int(_C.lib().rcn_launch_count()) + _replayed


def count_replayed(n: int):
    global _replayed
    _replayed += int(n)


# ----------------------------------------------------------------------------- conv engine selection
# "fp32"   : CUDA-core FFMA implicit GEMM (rcn_conv2d) for every layer -- exact-parity engine
# "bf16x3" : tcgen05 engine with bf16 hi/lo split operands (3 MMAs per product, ~fp32-grade) where eligible
# "fp16"   : tcgen05 engine, single fp16 pass (11 significand bits).  NOT for layers whose result feeds a quantisation decision;
#            the per-stage precision policy (raw2bit._g_s / tcm._g_s) scopes it to the post-quantisation synthesis tail
# "bf16"   : tcgen05 engine, single bf16 pass (fast mode; does NOT meet the 1e-3 parity bar end to end)
FMT_BF16, FMT_F16 = 0, 1
_MODES = {"fp32": (0, FMT_BF16), "bf16x3": (3, FMT_BF16), "bf16": (1, FMT_BF16), "fp16": (1, FMT_F16)}   # name -> (MMA passes, plane format)
_PLANE_DTYPE = {FMT_BF16: torch.bfloat16, FMT_F16: torch.float16}
_ENGINE = os.environ.get("RCN_CONV_ENGINE", "bf16x3")
_TC_MIN_CIN = 1   # every k in {1,3} layer is tcgen05-eligible (tiny Cin is zero-padded to one 64-channel K chunk)


def set_engine(name: str):
    global _ENGINE
    if name not in _MODES:
        raise ValueError(f"unknown conv engine {name!r}")
    _ENGINE = name


def get_engine() -> str:
    return _ENGINE


class engine_scope:
    """with ops.engine_scope("fp16"): ...  -- the layers launched inside use that engine (None = leave as is).  Layers are
    launched (or captured into a CUDA graph) synchronously from Python, so a scope is just a save / restore of the global."""

    def __init__(self, name):
        if name is not None and name not in _MODES:
            raise ValueError(f"unknown conv engine {name!r}")
        self.name, self.old = name, None

    def __enter__(self):
        global _ENGINE
        self.old = _ENGINE
        if self.name is not None:
            _ENGINE = self.name
        return self

    def __exit__(self, *exc):
        global _ENGINE
        _ENGINE = self.old
        return False


def _mode(eng=None):
    """(MMA passes, plane format) of an engine name"""
    return _MODES[eng or _ENGINE]


# ----------------------------------------------------------------------------- weights
def plane_channels(c: int) -> int:
    """Channel count of the bf16 operand planes of a c-channel tensor: 16, 32 or a multiple of 64."""
    return 16 if c <= 16 else (32 if c <= 32 else (c + 63) // 64 * 64)


class PackedConv:
    __slots__ = ("w", "bias", "k", "cin", "cout", "cp", "w_hi", "w_lo", "w_src", "w_hi_ps", "w_lo_ps", "_alt")

    def __init__(self, w, bias, k, cin, cout, cp=0, w_hi=None, w_lo=None):
        self.w, self.bias, self.k, self.cin, self.cout = w, bias, k, cin, cout
        self.cp, self.w_hi, self.w_lo = cp, w_hi, w_lo
        self.w_src = self.w_hi_ps = self.w_lo_ps = None
        self._alt = {}

    def ps_weights(self):
        """bf16 planes with rows grouped by PixelShuffle(2) sub-pixel (packed on first use by a pixel-shuffle store)."""
        if self.w_hi_ps is None:
            self.w_hi_ps, self.w_lo_ps = torch.empty_like(self.w_hi), torch.empty_like(self.w_lo)
            _C.check(_C.lib().rcn_pack_conv_weight_tc(_ptr(self.w_src), self.cout, self.cin, self.k, self.cp, 1, FMT_BF16,
                                                      _ptr(self.w_hi_ps), _ptr(self.w_lo_ps), _stream()), "rcn_pack_conv_weight_tc")
        return self.w_hi_ps, self.w_lo_ps

    def tc_weights(self, fmt, ps_perm=False):
        """(w_hi, w_lo) operand weights in plane format `fmt`, optionally with sub-pixel-grouped rows; other formats than bf16
        are packed on first use (single-pass engines: hi only)."""
        if fmt == FMT_BF16:
            return self.ps_weights() if ps_perm else (self.w_hi, self.w_lo)
        key = (fmt, bool(ps_perm))
        hit = self._alt.get(key)
        if hit is None:
            hi = torch.empty(self.w_hi.shape, device=self.w_hi.device, dtype=_PLANE_DTYPE[fmt])
            _C.check(_C.lib().rcn_pack_conv_weight_tc(_ptr(self.w_src), self.cout, self.cin, self.k, self.cp, int(ps_perm), fmt,
                                                      _ptr(hi), _vp(0), _stream()), "rcn_pack_conv_weight_tc")
            hit = self._alt[key] = (hi, None)
        return hit


_pack_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def pack_weight(weight: torch.Tensor, bias=None) -> PackedConv:
    """OIHW conv weight or (O,I) linear weight -> [k*k*Cin][Cout] on the device."""
    _chk(weight, "weight")
    w = weight.detach()
    if w.dim() == 2:
        cout, cin, k = w.shape[0], w.shape[1], 1
    else:
        cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
        assert w.shape[2] == w.shape[3]
    w = w.contiguous()
    out = torch.empty((k * k * cin, cout), device=w.device, dtype=torch.float32)
    _C.check(_C.lib().rcn_pack_conv_weight(_ptr(w), cout, cin, k, _ptr(out), _stream()), "rcn_pack_conv_weight")
    b = bias.detach().contiguous() if bias is not None else None
    pc = PackedConv(out, b, k, cin, cout)
    if k in (1, 3) and cin >= _TC_MIN_CIN:
        # tcgen05 operand: [Cout][k*k][Cp] bf16 hi/lo, Cp = Cin rounded up to the K chunk (16 / 32 for the few-channel layers)
        cp = plane_channels(cin)
        pc.cp = cp
        pc.w_hi = torch.empty((cout, k * k * cp), device=w.device, dtype=torch.bfloat16)
        pc.w_lo = torch.empty_like(pc.w_hi)
        _C.check(_C.lib().rcn_pack_conv_weight_tc(_ptr(w), cout, cin, k, cp, 0, FMT_BF16, _ptr(pc.w_hi), _ptr(pc.w_lo), _stream()),
                 "rcn_pack_conv_weight_tc")
        pc.w_src = w
    return pc


def pack(module) -> PackedConv:
    """Cached packing of an nn.Conv2d / nn.Linear parameter holder (re-packed if the weights change)."""
    w, b = module.weight, getattr(module, "bias", None)
    # identity (not address) of the source tensors + their in-place version counters: a replaced parameter that happens to
    # reuse a freed address at version 0 must not hit
    key = (w._version, w.data_ptr(), None if b is None else (b._version, b.data_ptr()), str(w.device))
    hit = _pack_cache.get(module)
    if hit is not None and hit[0] == key and hit[2]() is w and (b is None or hit[3]() is b):
        return hit[1]
    if isinstance(module, torch.nn.Conv2d):
        k = module.kernel_size[0]
        if (module.padding[0] != k // 2 or module.padding[1] != k // 2 or module.dilation != (1, 1) or module.groups != 1 or
                module.kernel_size[0] != module.kernel_size[1] or module.stride[0] != module.stride[1] or
                module.padding_mode != "zeros"):
            raise NotImplementedError(f"conv engine: only square kernels with padding k//2, dilation 1, groups 1 are on the path, got {module}")
    pc = pack_weight(w, b)
    _pack_cache[module] = (key, pc, weakref.ref(w), weakref.ref(b) if b is not None else None)
    return pc


# ----------------------------------------------------------------------------- conv / linear
class SplitOperand:
    """bf16 hi/lo planes of one activation tensor for the tcgen05 engine (shareable between the convs that read it).
    hi / lo are (N,H,W,Cp) bf16 tensors, possibly channel-slice views of a wider plane buffer (pixel stride = ld)."""
    __slots__ = ("hi", "lo", "key", "fmt")

    def __init__(self, hi, lo, key, fmt=FMT_BF16):
        self.hi, self.lo, self.key, self.fmt = hi, lo, key, fmt

    @property
    def ld(self):
        return plane_ld(self.hi)

    def channels(self, c0, c1):
        """Planes of the channel slice [c0, c1) -- what torch.split / torch.cat views are for the fp32 tensors."""
        if self.key[7] != 1 or self.key[8]:
            raise ValueError("only stride-1 operand planes can be sliced")
        cp = plane_channels(c1 - c0)
        if cp != c1 - c0:
            raise ValueError(f"a {c1 - c0}-channel slice is not a valid plane width")
        k = self.key
        return SplitOperand(self.hi[..., c0:c1], None if self.lo is None else self.lo[..., c0:c1],
                            (None, k[1], k[2], k[3], c1 - c0, None, cp, 1, False), self.fmt)


def plane_ld(t):
    """pixel stride (elements) of an (N,H,W,Cp) plane tensor / channel-slice view"""
    N, H, W, C = t.shape
    if C > 1 and t.stride(3) != 1:
        raise ValueError("operand planes: channel dim must be contiguous")
    ld = t.stride(2) if W > 1 else (t.stride(1) if H > 1 else (t.stride(0) if N > 1 else C))
    if (H > 1 and t.stride(1) != W * ld) or (N > 1 and t.stride(0) != H * W * ld) or ld < C:
        raise ValueError("operand planes: not an NHWC tensor or channel slice")
    return ld


def alloc_planes(N, H, W, C, device, passes=None, stride=1, fmt=None) -> SplitOperand:
    """Uninitialised operand planes of an (N,H,W,C) tensor for producers to write (conv2d(split_out=...), layernorm / wmsa
    emit); C must be a valid plane width.  stride=2: the polyphase layout a stride-2 consumer reads (4N, H/2, W/2, C).
    passes / fmt default to what the CURRENT engine's consumers read."""
    if plane_channels(C) != C:
        raise ValueError(f"{C} channels is not a valid plane width")
    if passes is None:
        passes = _mode()[0]
    if fmt is None:
        fmt = _mode()[1]
    shape = (N, H, W, C) if stride == 1 else (4 * N, H // 2, W // 2, C)
    hi = torch.empty(shape, device=device, dtype=_PLANE_DTYPE[fmt])
    lo = torch.empty_like(hi) if passes == 3 else None
    return SplitOperand(hi, lo, (None, N, H, W, C, None, C, stride, False), fmt)


def sp_align32(cs: int) -> bool:
    """plane pixel strides of freshly allocated planes are the channel count: the row-vector epilogue wants multiples of 16"""
    return cs % 16 == 0


def planes_enabled() -> bool:
    return _ENGINE != "fp32"


def bf16_planes_enabled() -> bool:
    """LayerNorm / window-attention kernels emit bf16 planes only"""
    return _ENGINE != "fp32" and _mode()[1] == FMT_BF16


def split_operand(x, cp, stride=1, in_square=False, passes=None, fmt=None) -> SplitOperand:
    """fp32 NHWC -> zero-padded 16-bit planes; stride 2 -> the four polyphase planes stacked on the batch axis."""
    N, H, W, C, ldx = geom(x, "split_operand.x")
    if passes is None:
        passes = _mode()[0] or 3
    if fmt is None:
        fmt = _mode()[1]
    if stride == 2:
        hi = torch.empty((4 * N, H // 2, W // 2, cp), device=x.device, dtype=_PLANE_DTYPE[fmt])
        lo = torch.empty_like(hi) if passes == 3 else None
        if in_square:
            raise ValueError("in_square is not used with stride 2")
        _C.check(_C.lib().rcn_split_bf16_s2(_ptr(x), ldx, N, H, W, C, cp, fmt, _ptr(hi), _ptr(lo), _stream()), "rcn_split_bf16_s2")
    else:
        hi = torch.empty((N, H, W, cp), device=x.device, dtype=_PLANE_DTYPE[fmt])
        lo = torch.empty_like(hi) if passes == 3 else None
        _C.check(_C.lib().rcn_split_bf16(_ptr(x), ldx, N * H * W, C, cp, int(in_square), fmt, _ptr(hi), _ptr(lo), _stream()),
                 "rcn_split_bf16")
    return SplitOperand(hi, lo, (x.data_ptr(), N, H, W, C, ldx, cp, stride, bool(in_square)), fmt)


def shared_split(x, pcs, stride=1):
    """One split for several convs reading the same tensor with the same stride (None when the engine is fp32
    or a layer is not tcgen05-eligible)."""
    if _ENGINE == "fp32" or any(pc.w_hi is None for pc in pcs) or len({pc.cp for pc in pcs}) != 1:
        return None
    if stride == 2 and (x.shape[1] % 2 or x.shape[2] % 2):
        return None
    return split_operand(x, pcs[0].cp, stride)


def conv2d(x, pc: PackedConv, stride=1, act=ACT_NONE, slope=0.0, out=None, store=STORE_NHWC, epi=EPI_NONE, aux=None,
           cscale=None, cshift=None, res=None, res_pre=False, in_square=False, bias=True, res_scale=1.0, engine=None, presplit=None,
           emit_split=False, keep_fp32=True, split_out=None, emit_stride=1, emit_square=False, aux_nchw=False):
    """One conv / linear layer with its fused epilogue.

    x may be None when `presplit` carries operand planes that a previous tcgen05 conv emitted (conv->conv chains never
    materialise the fp32 intermediate).  emit_split=True returns (out, SplitOperand | None): the epilogue additionally
    writes the NEXT layer's bf16 hi/lo planes; with keep_fp32=False `out` is None when that was possible.
    split_out: planes allocated by the caller (e.g. one half of a concat's planes) to emit into; implies emit_split.
    emit_stride=2: the emitted planes use the polyphase layout of a stride-2 consumer (no rcn_split_bf16_s2 pass).
    emit_square=True: the emitted planes hold the square of the result (operand of a GDN norm pool, conv2d(..., in_square=True))."""
    if split_out is not None:
        emit_split = True
    eng = engine or _ENGINE
    if x is None:
        if presplit is None or presplit.key[7] != stride:
            raise ValueError("conv2d: x=None needs operand planes from a previous layer, emitted for this layer's stride")
        N, H, W = presplit.key[1:4]
        Cin, ldx = presplit.key[4], 0
        _on_current_device(presplit.hi, "conv2d.presplit")
    else:
        N, H, W, Cin, ldx = geom(x, "conv2d.x")
    if Cin != pc.cin:
        raise ValueError(f"conv2d: input has {Cin} channels, weight expects {pc.cin}")
    pad = pc.k // 2
    Ho = (H + 2 * pad - pc.k) // stride + 1
    Wo = (W + 2 * pad - pc.k) // stride + 1
    ps = store in (STORE_PS2, STORE_PS2_NCHW)
    Hs, Ws, Cs = (2 * Ho, 2 * Wo, pc.cout // 4) if ps else (Ho, Wo, pc.cout)
    use_tc = eng != "fp32" and pc.w_hi is not None and (stride == 1 or (H % 2 == 0 and W % 2 == 0))
    if x is None and not use_tc:
        raise ValueError("conv2d: operand planes can only feed the tcgen05 engine")
    dev = x.device if x is not None else presplit.hi.device
    can_emit = (emit_split and use_tc and store in (STORE_NHWC, STORE_PS2) and plane_channels(Cs) == Cs and
                (store != STORE_PS2 or (epi == EPI_NONE and cscale is None and pc.cout % 64 == 0)) and pc.cout % 16 == 0)
    if emit_stride == 2:
        can_emit = can_emit and store == STORE_NHWC and Ho % 2 == 0 and Wo % 2 == 0
    if split_out is not None and not can_emit:
        raise ValueError("conv2d: split_out given but this layer cannot emit operand planes")
    want_out = keep_fp32 or not can_emit or out is not None
    if want_out and out is None:
        if store in (STORE_NCHW, STORE_PS2_NCHW):
            out = torch.empty((N, Cs, Hs, Ws), device=dev, dtype=torch.float32)
        else:
            out = torch.empty((N, Hs, Ws, Cs), device=dev, dtype=torch.float32)
    ldy = 0
    if out is not None:
        if store in (STORE_NCHW, STORE_PS2_NCHW):
            if tuple(out.shape) != (N, Cs, Hs, Ws) or not out.is_contiguous():
                raise ValueError("conv2d: NCHW output must be contiguous with the right shape")
        else:
            oN, oH, oW, oC, ldy = geom(out, "conv2d.out")
            if (oN, oH, oW, oC) != (N, Hs, Ws, Cs):
                raise ValueError(f"conv2d: output shape {tuple(out.shape)} != {(N, Hs, Ws, Cs)}")
    d = _C.ConvDesc()
    d.x, d.N, d.H, d.W, d.Cin, d.ldx = (x.data_ptr() if x is not None else None), N, H, W, Cin, ldx
    d.w = pc.w.data_ptr()
    d.bias = pc.bias.data_ptr() if (bias and pc.bias is not None) else None
    d.k, d.stride, d.Cout, d.in_square = pc.k, stride, pc.cout, int(in_square)
    d.y, d.ldy, d.store = (out.data_ptr() if out is not None else None), ldy, store
    d.epi = epi
    lda = ldr = 0
    if epi != EPI_NONE:
        if aux_nchw:     # a contiguous NCHW map (N,Cout,Ho,Wo), e.g. an API-facing tensor written by conv2d(store=STORE_NCHW)
            _chk(aux, "conv2d.aux")
            if tuple(aux.shape) != (N, pc.cout, Ho, Wo) or not aux.is_contiguous():
                raise ValueError("conv2d: an NCHW aux must be contiguous with the conv output geometry")
            d.aux, d.ldaux, d.aux_nchw = aux.data_ptr(), 0, 1
        else:
            aN, aH, aW, aC, lda = geom(aux, "conv2d.aux")
            if (aN, aH, aW, aC) != (N, Ho, Wo, pc.cout):
                raise ValueError("conv2d: aux must have the conv output geometry")
            d.aux, d.ldaux = aux.data_ptr(), lda
    if cscale is not None:
        _chk(cscale), _chk(cshift)
        assert cscale.is_contiguous() and cshift.is_contiguous() and cscale.numel() == N * pc.cout == cshift.numel()
        d.cscale, d.cshift = cscale.data_ptr(), cshift.data_ptr()
    if res is not None:
        rN, rH, rW, rC, ldr = geom(res, "conv2d.res")
        if (rN, rH, rW, rC) != (N, Hs, Ws, Cs):
            raise ValueError(f"conv2d: residual shape {tuple(res.shape)} != {(N, Hs, Ws, Cs)}")
        d.res, d.ldres, d.res_pre = res.data_ptr(), ldr, int(res_pre)
    d.act, d.slope, d.res_scale = act, float(slope), float(res_scale)
    sp_out = None
    if can_emit:   # alignment requirements of the 16-byte epilogue path
        ok = (ldy % 4 == 0 and (out is None or out.data_ptr() % 16 == 0) and lda % 4 == 0 and ldr % 4 == 0 and
              (aux is None or aux.data_ptr() % 16 == 0) and (res is None or res.data_ptr() % 16 == 0))
        if aux_nchw:         # served by the row-vector epilogue only (32-byte rules)
            ok = ok and pc.cout % 16 == 0 and (out is None or (ldy % 8 == 0 and out.data_ptr() % 32 == 0)) and sp_align32(Cs)
        if ok:
            sp_out = split_out if split_out is not None else alloc_planes(N, Hs, Ws, Cs, dev, stride=emit_stride)
            want_shape = (N, Hs, Ws, Cs) if emit_stride == 1 else (4 * N, Hs // 2, Ws // 2, Cs)
            if tuple(sp_out.hi.shape) != want_shape or sp_out.key[7] != emit_stride:
                raise ValueError("conv2d: split_out planes do not have the stored output geometry")
            d.planes_s2 = int(emit_stride == 2)
            d.planes_square = int(bool(emit_square))
            if emit_square:
                if split_out is not None or emit_stride != 1:
                    raise ValueError("conv2d: emit_square works on freshly allocated stride-1 planes only")
                k_ = sp_out.key
                sp_out = SplitOperand(sp_out.hi, sp_out.lo, k_[:8] + (True,), sp_out.fmt)
            d.out_fmt = sp_out.fmt
            d.y_hi, d.y_lo, d.Cp_out = sp_out.hi.data_ptr(), (sp_out.lo.data_ptr() if sp_out.lo is not None else None), sp_out.ld
        elif out is None:
            raise ValueError("conv2d: cannot drop the fp32 output of a layer whose epilogue is not 16-byte aligned")
    if use_tc:
        # tcgen05 path: 16-bit operand planes (from the producer's epilogue, a shared split, or a split pass here)
        passes, fmt = _mode(eng)
        sp = presplit if presplit is not None else split_operand(x, pc.cp, stride, in_square, passes, fmt)
        k0 = sp.key
        same = (k0[1:5] == (N, H, W, Cin) and k0[6:] == (pc.cp, stride, bool(in_square)) and
                (k0[0] is None or (x is not None and k0[0] == x.data_ptr() and k0[5] == ldx)))
        if sp.fmt != fmt:
            raise ValueError(f"conv2d: operand planes are in format {sp.fmt} but the {eng!r} engine reads format {fmt} "
                             "(a producer outside the engine scope emitted them)")
        if not same or (passes == 3 and sp.lo is None):
            raise ValueError("conv2d: presplit operand does not belong to this input / layer geometry")
        d.ldp_in, d.in_fmt = sp.ld, fmt
        ps_perm = store == STORE_PS2 and pc.cout % 64 == 0 and epi == EPI_NONE and cscale is None
        w_hi, w_lo = pc.tc_weights(fmt, ps_perm)  # ps_perm: rows grouped by sub-pixel -> 64-byte contiguous pixel-shuffle stores
        if ps_perm:
            d.ps_perm = 1
        _C.check(_C.lib().rcn_conv2d_tc(ctypes.byref(d), _ptr(sp.hi), _ptr(sp.lo), _ptr(w_hi), _ptr(w_lo), pc.cp, passes,
                                        _stream()), "rcn_conv2d_tc")
    else:
        _C.check(_C.lib().rcn_conv2d(ctypes.byref(d), _stream()), "rcn_conv2d")
    return (out, sp_out) if emit_split else out


def layernorm(x, weight, bias, eps=1e-5, out=None, act=ACT_NONE, emit_split=False):
    """emit_split=True: returns (None, SplitOperand) -- the result only exists as the consumer conv's operand planes -- when the
    tcgen05 engine is active and C is a valid plane width; (out, None) otherwise."""
    N, H, W, C, ldx = geom(x, "layernorm.x")
    sp = None
    if emit_split and bf16_planes_enabled() and plane_channels(C) == C and out is None:
        sp = alloc_planes(N, H, W, C, x.device)
        _C.check(_C.lib().rcn_layernorm(_ptr(x), N * H * W, C, ldx, _ptr(weight), _ptr(bias), eps, _vp(0), 0, act,
                                        _ptr(sp.hi), _ptr(sp.lo), sp.ld, _stream()), "rcn_layernorm")
        return None, sp
    if out is None:
        out = empty(N, H, W, C, like=x)
    _, _, _, _, ldy = geom(out, "layernorm.out")
    _C.check(_C.lib().rcn_layernorm(_ptr(x), N * H * W, C, ldx, _ptr(weight), _ptr(bias), eps, _ptr(out), ldy, act,
                                    _vp(0), _vp(0), 0, _stream()), "rcn_layernorm")
    return (out, None) if emit_split else out


def wmsa(qkv, relpos, head_dim, ws, shifted, out=None, emit_split=False):
    """emit_split=True: (None, SplitOperand) when the attention output can be handed to the projection layer as planes."""
    N, H, W, C3, ldq = geom(qkv, "wmsa.qkv")
    C = C3 // 3
    _chk(relpos)
    assert relpos.is_contiguous() and tuple(relpos.shape) == (C // head_dim, 2 * ws - 1, 2 * ws - 1)
    if emit_split and bf16_planes_enabled() and plane_channels(C) == C and out is None:
        sp = alloc_planes(N, H, W, C, qkv.device)
        _C.check(_C.lib().rcn_wmsa(_ptr(qkv), N, H, W, C, ldq, head_dim, ws, int(shifted), _ptr(relpos), _vp(0), 0,
                                   _ptr(sp.hi), _ptr(sp.lo), sp.ld, _stream()), "rcn_wmsa")
        return None, sp
    if out is None:
        out = empty(N, H, W, C, like=qkv)
    _, _, _, _, ldo = geom(out, "wmsa.out")
    _C.check(_C.lib().rcn_wmsa(_ptr(qkv), N, H, W, C, ldq, head_dim, ws, int(shifted), _ptr(relpos), _ptr(out), ldo,
                               _vp(0), _vp(0), 0, _stream()), "rcn_wmsa")
    return (out, None) if emit_split else out


# ----------------------------------------------------------------------------- layout

# ----------------------------------------------------------------------------- CUDA-graph replay of a fixed launch sequence
class GraphReplay:
    """Mixin for nn.Modules whose inference forward is a fixed sequence of C-ABI launches (the ISP networks: ~250-1350 launches per
    tile, 20 us of host time each when issued eagerly -- a 256^2 tile is host-bound by 4x).  enable_cuda_graphs() makes
    `_graph_call(fn, inputs)` capture fn(inputs) once per (input shapes, engine) and replay it afterwards; a fingerprint of every
    parameter / buffer version invalidates the captures (in-place weight edits, load_state_dict), moving the module does too.
    The tensor returned in graph mode lives in the graph's memory pool and is OVERWRITTEN by the next call (clone what must survive)."""

    def enable_cuda_graphs(self, flag=True):
        self.__dict__["_use_graphs"] = bool(flag)
        self.__dict__["_graphs"] = {}
        return self

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_graphs"] = {}
        return super()._apply(fn, *args, **kwargs)

    def _graph_fingerprint(self):
        return sum(t._version for t in list(self.parameters()) + list(self.buffers()))

    def _graph_call(self, fn, inputs):
        if not self.__dict__.get("_use_graphs") or not inputs[0].is_cuda:
            return fn(inputs)
        cache = self.__dict__.setdefault("_graphs", {})
        fp = self._graph_fingerprint()
        if cache.get("_fp") != fp:
            cache.clear()
            cache["_fp"] = fp
        key = (_ENGINE,) + tuple((tuple(t.shape), str(t.device)) for t in inputs)
        e = cache.get(key)
        if e is None:
            static_in = [t.detach().clone() for t in inputs]
            fn(static_in)                                  # eager warm-up: weight packing and every other lazy initialisation
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = launch_count()
            with torch.cuda.graph(g):
                out = fn(static_in)
            e = cache[key] = (g, static_in, out, launch_count() - n0)
            cache["_fp"] = self._graph_fingerprint()
        g, static_in, out, n = e
        for dst, src in zip(static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        g.replay()
        count_replayed(n)
        return out


# ----------------------------------------------------------------------------- independent branches on two streams
# 0: branches run one after the other (triage / A-B)
_CONCURRENT_BRANCHES = os.environ.get("RCN_CONCURRENT_BRANCHES", "1") != "0"
_side_streams = {}
_fork_path = ""      # position in the tree of nested fork_join calls ("" = not inside one): one side stream per position


def fork_join(fa, fb, dev):
    """fa() on the current stream, fb() concurrently on a side stream, joined before returning (ra, rb).

    The entropy-parameter networks are chains of launches on 128^2 / 64^2 maps: one wave of <= 128 CTAs each, bound by launch
    latency and prologue / drain, not by throughput.  Independent chains (mean / scale halves, models/raw2bit.py:1812-1828; the two
    branches of an attention gate, compressai AttentionBlock) are issued on two streams so that the block scheduler starts one
    chain's next kernel while the other chain's CTAs drain.  Inside a CUDA-graph capture the fork / join events become parallel
    branches of the graph.  Same kernels, same arithmetic: results are bit-identical to the serial order.  Nested calls get a side
    stream of their own per position in the call tree."""
    global _fork_path
    # Only while a CUDA graph is being captured: eager passes are bound by the host's launch rate, where a second stream buys nothing and
    # the fork / join events cost host time (decode: 78 -> 82 ms per tile when forked eagerly).
    if not _CONCURRENT_BRANCHES or dev.type != "cuda" or not torch.cuda.is_current_stream_capturing():
        return fa(), fb()
    cur = torch.cuda.current_stream(dev)
    path = _fork_path
    key = (dev.index, path)
    side = _side_streams.get(key)
    if side is None:
        side = _side_streams[key] = torch.cuda.Stream(device=dev)
    fork, join = torch.cuda.Event(), torch.cuda.Event()
    fork.record(cur)
    side.wait_event(fork)
    try:
        _fork_path = path + "b"
        with torch.cuda.stream(side):
            rb = fb()
            join.record(side)
        _fork_path = path + "a"
        ra = fa()
    finally:
        _fork_path = path
    cur.wait_event(join)       # (tensors allocated on the side stream live in the capture's private pool, which never recycles mid-capture)
    return ra, rb


# ----------------------------------------------------------------------------- fused packed-Bayer ingest
_FUSED_INGEST = os.environ.get("RCN_FUSED_INGEST", "1") != "0"    # 0: the lens-shading MLP and conv_first run layer by layer (triage / A-B)


def fused_ingest_ok(lsc_layers, coord, conv_first=None) -> bool:
    """Shapes rcn_ingest_fused serves: the 2 -> 128 -> 128 -> 128 -> 128 lens-shading MLP (+ a 3x3 4 -> 128 conv_first) on the
    bf16x3 engine, even H, W % 64 == 0."""
    if not _FUSED_INGEST or _ENGINE != "bf16x3" or len(lsc_layers) != 4:
        return False
    shapes = [tuple(m.weight.shape) for m in lsc_layers]
    if shapes != [(128, 2, 1, 1)] + [(128, 128, 1, 1)] * 3 or any(m.bias is None for m in lsc_layers):
        return False
    if conv_first is not None and (tuple(conv_first.weight.shape) != (128, 4, 3, 3) or conv_first.bias is None or conv_first.stride[0] != 1):
        return False
    N, C, H, W = coord.shape
    return C == 2 and H % 2 == 0 and W % 64 == 0


def ingest_fused(coord_nchw, lsc_layers, slope, raw=None, conv_first=None, emit_stride=2):
    """models/raw2bit.py:1771-1780 as one kernel (rcn_ingest_fused): returns (lsc NCHW fp32, operand planes of
    conv_first(raw) * (lsc + 1) | None).  coord_nchw: the API tensor (N,2,H,W); raw: NHWC (N,H,W,4)."""
    _chk(coord_nchw, "ingest_fused.coord")
    coord = coord_nchw.contiguous()
    N, _, H, W = coord.shape
    d = _C.IngestDesc()
    d.coord, d.coord_bs, d.coord_ps, d.coord_cs = coord.data_ptr(), 2 * H * W, 1, H * W
    l0 = lsc_layers[0]
    w0, b0 = l0.weight.detach().contiguous(), l0.bias.detach().contiguous()
    d.w0, d.b0 = w0.data_ptr(), b0.data_ptr()
    pcs = [pack(m) for m in lsc_layers[1:]]
    (d.w1_hi, d.w1_lo), (d.w2_hi, d.w2_lo), (d.w3_hi, d.w3_lo) = [(pc.w_hi.data_ptr(), pc.w_lo.data_ptr()) for pc in pcs]
    d.b1, d.b2, d.b3 = [pc.bias.data_ptr() for pc in pcs]
    d.slope = float(slope)
    lsc = torch.empty((N, 128, H, W), device=coord.device, dtype=torch.float32)
    d.lsc, d.N, d.H, d.W = lsc.data_ptr(), N, H, W
    sp = None
    keep = [coord, w0, b0, pcs]
    if raw is not None:
        rN, rH, rW, rC, ldr = geom(raw, "ingest_fused.raw")
        if (rN, rH, rW, rC) != (N, H, W, 4):
            raise ValueError(f"ingest_fused: raw {tuple(raw.shape)} does not match coord {tuple(coord.shape)}")
        pc = pack(conv_first)
        hit = pc._alt.get("ingest")
        if hit is None:
            hi = torch.empty((128, 48), device=raw.device, dtype=torch.bfloat16)
            lo = torch.empty_like(hi)
            _C.check(_C.lib().rcn_pack_ingest_weight(_ptr(pc.w_src), _ptr(hi), _ptr(lo), _stream()), "rcn_pack_ingest_weight")
            hit = pc._alt["ingest"] = (hi, lo)
        sp = alloc_planes(N, H, W, 128, raw.device, passes=3, stride=emit_stride, fmt=FMT_BF16)
        d.raw, d.ldraw = raw.data_ptr(), ldr
        d.wc_hi, d.wc_lo, d.bc = hit[0].data_ptr(), hit[1].data_ptr(), pc.bias.data_ptr()
        d.fea_hi, d.fea_lo, d.planes_s2 = sp.hi.data_ptr(), sp.lo.data_ptr(), int(emit_stride == 2)
        keep.append(hit)
    _C.check(_C.lib().rcn_ingest_fused(ctypes.byref(d), _stream()), "rcn_ingest_fused")
    return lsc, sp

_FUSED_MLP = os.environ.get("RCN_FUSED_MLP", "1") != "0"    # 0: fc1 and fc2 of the Swin MLP run as two conv launches (triage / A-B)


def mlp_fused_ok(fc1, fc2, xsp=None, ln_x=None, ln=None) -> bool:
    """Shapes rcn_mlp_fused serves: Linear(64, 256) -> GELU -> Linear(256, 64) on the bf16x3 engine; the input either as bf16 hi/lo
    planes (xsp) or as the fp32 rows ln_x of the LayerNorm `ln` in front, which the kernel then applies itself."""
    if not (_FUSED_MLP and _ENGINE == "bf16x3" and tuple(fc1.weight.shape[:2]) == (256, 64) and tuple(fc2.weight.shape[:2]) == (64, 256) and
            fc1.bias is not None and fc2.bias is not None):
        return False
    if ln_x is not None:
        return (_FUSED_MLP_LN and ln is not None and tuple(ln.normalized_shape) == (64,) and ln.weight is not None and ln.bias is not None and
                ln_x.shape[-1] == 64 and ln_x.data_ptr() % 16 == 0 and geom(ln_x)[4] % 4 == 0)
    return (xsp is not None and xsp.fmt == FMT_BF16 and xsp.lo is not None and xsp.key[7] == 1 and xsp.hi.shape[-1] == 64)


# 1: the LayerNorm in front of the fused MLP runs inside it (A operand of fc1 built in tensor memory).  Correct and tested, but OFF by
# default: measured at step level it costs 0.5 ms (54.5 against 54.0 ms) -- the streaming LayerNorm kernel runs near the HBM roofline,
# while inside the chain kernel it adds a serial phase (row loads, statistics, tcgen05.st, hand-over) to every tile slot.
_FUSED_MLP_LN = os.environ.get("RCN_FUSED_MLP_LN", "0") != "0"


def mlp_fused(xsp, fc1, fc2, res=None, out=None, split_out=None, keep_fp32=True, ln_x=None, ln=None):
    """models/tcm.py:225-236 as one kernel (rcn_mlp_fused): res + fc2(GELU(fc1(x))); x given as operand planes (xsp), or as
    ln(ln_x) with the LayerNorm applied inside the kernel (tcm.py:234).  Returns (y | None, planes | None) like conv2d(emit_split=True)."""
    p1, p2 = pack(fc1), pack(fc2)
    d = _C.MlpDesc()
    if ln_x is not None:
        N, H, W, C, ldx = geom(ln_x, "mlp_fused.ln_x")
        g, b = ln.weight.detach(), ln.bias.detach()
        d.x_ln, d.ldx, d.gamma, d.beta, d.eps = ln_x.data_ptr(), ldx, g.data_ptr(), b.data_ptr(), float(ln.eps)
        dev_t = ln_x
    else:
        N, H, W, C = xsp.hi.shape
        d.x_hi, d.x_lo, d.ldp_in = xsp.hi.data_ptr(), xsp.lo.data_ptr(), plane_ld(xsp.hi)
        dev_t = xsp.hi
    d.npix, d.C, d.hidden = N * H * W, 64, 256
    d.w1_hi, d.w1_lo, d.b1 = p1.w_hi.data_ptr(), p1.w_lo.data_ptr(), p1.bias.data_ptr()
    d.w2_hi, d.w2_lo, d.b2 = p2.w_hi.data_ptr(), p2.w_lo.data_ptr(), p2.bias.data_ptr()
    _on_current_device(dev_t, "mlp_fused.x")
    if res is not None:
        rN, rH, rW, rC, ldr = geom(res, "mlp_fused.res")
        if (rN, rH, rW, rC) != (N, H, W, 64):
            raise ValueError("mlp_fused: residual geometry mismatch")
        d.res, d.ldres = res.data_ptr(), ldr
    want_out = keep_fp32 or split_out is None or out is not None
    if want_out:
        if out is None:
            out = torch.empty((N, H, W, 64), device=dev_t.device, dtype=torch.float32)
        oN, oH, oW, oC, ldy = geom(out, "mlp_fused.out")
        if (oN, oH, oW, oC) != (N, H, W, 64):
            raise ValueError("mlp_fused: output geometry mismatch")
        d.y, d.ldy = out.data_ptr(), ldy
    if split_out is not None:
        if tuple(split_out.hi.shape) != (N, H, W, 64) or split_out.lo is None or split_out.fmt != FMT_BF16 or split_out.key[7] != 1:
            raise ValueError("mlp_fused: split_out must be stride-1 bf16 hi/lo planes of the output geometry")
        d.y_hi, d.y_lo, d.Cp_out = split_out.hi.data_ptr(), split_out.lo.data_ptr(), plane_ld(split_out.hi)
    _C.check(_C.lib().rcn_mlp_fused(ctypes.byref(d), _stream()), "rcn_mlp_fused")
    return (out if want_out else None), split_out

# 1: LayerNorm + qkv embedding as one kernel (csrc/lnlinear.cu).  Correct and tested, OFF by default: no gain at step level (54.0 against
# 53.8 ms), for the same reason as _FUSED_MLP_LN.
_FUSED_LN_LINEAR = os.environ.get("RCN_FUSED_LN_LINEAR", "0") != "0"


def ln_linear_ok(x, ln, fc) -> bool:
    """Shapes rcn_ln_linear_fused serves: Linear(64, Cout <= 192, Cout % 16 == 0) of a LayerNorm(64), bf16x3 engine, fp32 rows."""
    if not (_FUSED_LN_LINEAR and _ENGINE == "bf16x3" and x is not None):
        return False
    w = fc.weight
    return (tuple(ln.normalized_shape) == (64,) and ln.weight is not None and ln.bias is not None and w.shape[1] == 64 and
            w.shape[0] % 16 == 0 and w.shape[0] <= 192 and x.shape[-1] == 64 and x.data_ptr() % 16 == 0 and geom(x)[4] % 4 == 0)


def ln_linear(x, ln, fc, out=None):
    """models/tcm.py:233 + 193 as one kernel (rcn_ln_linear_fused): fc(ln(x)) with the normalised rows as the tcgen05 A operand in
    tensor memory.  x: (N,H,W,64) fp32 NHWC / channel slice; returns (N,H,W,Cout) fp32."""
    N, H, W, C, ldx = geom(x, "ln_linear.x")
    pc = pack(fc)
    if out is None:
        out = torch.empty((N, H, W, pc.cout), device=x.device, dtype=torch.float32)
    oN, oH, oW, oC, ldy = geom(out, "ln_linear.out")
    if (oN, oH, oW, oC) != (N, H, W, pc.cout):
        raise ValueError("ln_linear: output geometry mismatch")
    d = _C.LnLinearDesc()
    d.x, d.ldx, d.npix, d.C, d.Cout = x.data_ptr(), ldx, N * H * W, 64, pc.cout
    g, b = ln.weight.detach(), ln.bias.detach()
    d.gamma, d.beta, d.eps = g.data_ptr(), b.data_ptr(), float(ln.eps)
    d.w_hi, d.w_lo = pc.w_hi.data_ptr(), pc.w_lo.data_ptr()
    d.bias = pc.bias.data_ptr() if pc.bias is not None else None
    d.y, d.ldy = out.data_ptr(), ldy
    _C.check(_C.lib().rcn_ln_linear_fused(ctypes.byref(d), _stream()), "rcn_ln_linear_fused")
    return out


def to_nhwc(x, out=None):
    _chk(x, "to_nhwc.x")
    x = x.contiguous()
    N, C, H, W = x.shape
    if out is None:
        out = empty(N, H, W, C, like=x)
    _, _, _, _, ldy = geom(out)
    _C.check(_C.lib().rcn_nchw_to_nhwc(_ptr(x), N, C, H, W, _ptr(out), ldy, _stream()), "rcn_nchw_to_nhwc")
    return out


def to_nchw(x):
    N, H, W, C, ldx = geom(x, "to_nchw.x")
    out = torch.empty((N, C, H, W), device=x.device, dtype=torch.float32)
    _C.check(_C.lib().rcn_nhwc_to_nchw(_ptr(x), ldx, N, C, H, W, _ptr(out), _stream()), "rcn_nhwc_to_nchw")
    return out


def copy_channels(x, out):
    N, H, W, C, ldx = geom(x)
    oN, oH, oW, oC, ldy = geom(out)
    assert (N, H, W, C) == (oN, oH, oW, oC)
    _C.check(_C.lib().rcn_copy_channels(_ptr(x), ldx, N * H * W, C, _ptr(out), ldy, _stream()), "rcn_copy_channels")
    return out


# ----------------------------------------------------------------------------- pooling / gating
def channel_mean(x):
    """AdaptiveAvgPool2d(1) -> (N,1,1,C)"""
    N, H, W, C, ldx = geom(x, "channel_mean.x")
    mean = empty(N, 1, 1, C, like=x)
    wsf = N * min(1024, max(1, (H * W) // 64)) * C
    ws = torch.empty((wsf,), device=x.device, dtype=torch.float32)
    _C.check(_C.lib().rcn_channel_mean(_ptr(x), N, H * W, C, ldx, _ptr(mean), _ptr(ws), wsf, _stream()), "rcn_channel_mean")
    return mean


def instance_norm(x, gamma, beta, eps=1e-5):
    N, H, W, C, ldx = geom(x, "instance_norm.x")
    mean = torch.empty((N, C), device=x.device, dtype=torch.float32)
    var = torch.empty((N, C), device=x.device, dtype=torch.float32)
    _C.check(_C.lib().rcn_channel_meanvar(_ptr(x), N, H * W, C, ldx, _ptr(mean), _ptr(var), _stream()), "rcn_channel_meanvar")
    out = empty(N, H, W, C, like=x)
    _C.check(_C.lib().rcn_norm_apply(_ptr(x), ldx, N, H * W, C, _ptr(mean), _ptr(var), _ptr(gamma), _ptr(beta), eps,
                                     _ptr(out), C, _stream()), "rcn_norm_apply")
    return out


def scale_add(x, g, b=None, per_n=True, res=None, out=None, act=ACT_NONE):
    """out = act(x * g + b) + res, g/b per (n,c) or per c."""
    N, H, W, C, ldx = geom(x, "scale_add.x")
    if out is None:
        out = empty(N, H, W, C, like=x)
    _, _, _, _, ldy = geom(out)
    ldr = geom(res)[4] if res is not None else 0
    assert g.is_contiguous() and g.numel() == (N * C if per_n else C)
    _C.check(_C.lib().rcn_scale_add(_ptr(x), ldx, N, H * W, C, _ptr(g), _ptr(b), int(per_n), _ptr(res), ldr, _ptr(out), ldy,
                                    act, _stream()), "rcn_scale_add")
    return out


def avgpool3s2_lrelu(x, slope):
    N, H, W, C, ldx = geom(x)
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    out = empty(N, Ho, Wo, C, like=x)
    _C.check(_C.lib().rcn_avgpool3s2_lrelu(_ptr(x), N, H, W, C, ldx, slope, _ptr(out), C, _stream()), "rcn_avgpool3s2_lrelu")
    return out


def upsample_bilinear2x(x, out=None, emit_split=False):
    """emit_split=True: returns (None, SplitOperand) -- the result only exists as the consumer conv's operand planes -- when the
    bf16 plane engine is active and C is a plane width; else (fp32 tensor, None)."""
    N, H, W, C, ldx = geom(x)
    if emit_split and out is None and bf16_planes_enabled() and _mode()[0] == 3 and plane_channels(C) == C and ldx % 4 == 0 and \
            x.data_ptr() % 16 == 0:
        sp = alloc_planes(N, 2 * H, 2 * W, C, x.device)
        _C.check(_C.lib().rcn_upsample_bilinear2x_planes(_ptr(x), N, H, W, C, ldx, _ptr(sp.hi), _ptr(sp.lo), plane_ld(sp.hi), _stream()),
                 "rcn_upsample_bilinear2x_planes")
        return None, sp
    if emit_split:
        return upsample_bilinear2x(x, out=out), None
    if out is None:
        out = empty(N, 2 * H, 2 * W, C, like=x)
    assert tuple(out.shape) == (N, 2 * H, 2 * W, C)
    ldy = geom(out)[4]
    _C.check(_C.lib().rcn_upsample_bilinear2x(_ptr(x), N, H, W, C, ldx, _ptr(out), ldy, _stream()), "rcn_upsample_bilinear2x")
    return out


def dwt_forward(x, out=None):
    N, H, W, C, ldx = geom(x)
    if out is None:
        out = empty(N, H // 2, W // 2, 4 * C, like=x)
    _C.check(_C.lib().rcn_dwt_forward(_ptr(x), N, H, W, C, ldx, _ptr(out), geom(out)[4], _stream()), "rcn_dwt_forward")
    return out


def dwt_inverse(x, out=None):
    N, H, W, C4, ldx = geom(x)
    if out is None:
        out = empty(N, 2 * H, 2 * W, C4 // 4, like=x)
    _C.check(_C.lib().rcn_dwt_inverse(_ptr(x), N, H, W, C4, ldx, _ptr(out), geom(out)[4], _stream()), "rcn_dwt_inverse")
    return out


def space_to_depth2(x, out=None):
    """(N,H,W,C) -> (N,H/2,W/2,4C), channel (i*2+j)*C + c <- pixel (2y+i, 2x+j)"""
    N, H, W, C, ldx = geom(x)
    if out is None:
        out = empty(N, H // 2, W // 2, 4 * C, like=x)
    _C.check(_C.lib().rcn_space_to_depth2(_ptr(x), N, H, W, C, ldx, _ptr(out), geom(out)[4], _stream()), "rcn_space_to_depth2")
    return out


def add(x, y, out=None):
    """x + y (scale_add with a unit gain)"""
    C = x.shape[-1]
    key = (C, str(x.device))
    ones = _ones_cache.get(key)
    if ones is None:
        ones = _ones_cache[key] = torch.ones(C, device=x.device, dtype=torch.float32)
    return scale_add(x, ones, per_n=False, res=y, out=out)


_ones_cache = {}


def depthwise_conv(x, w_taps, bias, k, add_input=False, mul=None, out=None):
    """w_taps: [k*k][C] (tap-major) device tensor."""
    N, H, W, C, ldx = geom(x)
    if out is None:
        out = empty(N, H, W, C, like=x)
    ldm = geom(mul)[4] if mul is not None else 0
    _C.check(_C.lib().rcn_depthwise_conv(_ptr(x), N, H, W, C, ldx, _ptr(w_taps), _ptr(bias), k, int(add_input), _ptr(mul), ldm,
                                         _ptr(out), geom(out)[4], _stream()), "rcn_depthwise_conv")
    return out


# ----------------------------------------------------------------------------- entropy
def eb_forward(z, params, medians, want_zhat=True, want_lik=True, want_symbols=False, lik_bound=1e-9):
    N, H, W, C, ldz = geom(z, "eb_forward.z")
    z_hat = empty(N, H, W, C, like=z) if want_zhat else None
    lik = empty(N, H, W, C, like=z) if want_lik else None
    sym = torch.empty((N, C, H, W), device=z.device, dtype=torch.int32) if want_symbols else None
    _C.check(_C.lib().rcn_eb_forward(_ptr(z), ldz, N, H * W, C, _ptr(params), _ptr(medians), _ptr(z_hat), C, _ptr(lik), C,
                                     _ptr(sym), lik_bound, _stream()), "rcn_eb_forward")
    return z_hat, lik, sym


def eb_dequantize(symbols, medians):
    N, C, H, W = symbols.shape
    assert symbols.dtype == torch.int32 and symbols.is_contiguous()
    z_hat = torch.empty((N, H, W, C), device=symbols.device, dtype=torch.float32)
    _C.check(_C.lib().rcn_eb_dequantize(_ptr(symbols), N, H * W, C, _ptr(medians), _ptr(z_hat), C, _stream()), "rcn_eb_dequantize")
    return z_hat


class CoderPrep:
    """Device-side buffers for the GPU front end of the range coder (rcn_gaussian_conditional_coded)."""

    def __init__(self, nsym, cdf, cdf_len, cdf_off):
        dev = cdf.device
        self.cdf, self.cdf_len, self.cdf_off = cdf.contiguous(), cdf_len.contiguous(), cdf_off.contiguous()
        assert self.cdf.dtype == torch.int32 and self.cdf_len.dtype == torch.int32 and self.cdf_off.dtype == torch.int32
        self.packed = torch.empty((nsym,), device=dev, dtype=torch.int32)   # bit pattern (start << 16) | (freq - 1)
        self.raw = torch.empty((nsym,), device=dev, dtype=torch.int32)      # bypass payload where flags != 0
        self.flags = torch.empty((nsym,), device=dev, dtype=torch.uint8)


def gaussian_conditional(y, mu, scale, table, y_hat=None, lik=None, symbols=None, indexes=None, scale_bound=0.11,
                         lik_bound=1e-9, coder: "CoderPrep" = None, pos_base=0):
    N, H, W, C, ldy = geom(y, "gaussian.y")
    ldm, lds = geom(mu)[4], geom(scale)[4]
    ldyh = geom(y_hat)[4] if y_hat is not None else 0
    ldl = geom(lik)[4] if lik is not None else 0
    if coder is not None:
        n = N * C * H * W
        sl_ = slice(pos_base, pos_base + n)
        _C.check(_C.lib().rcn_gaussian_conditional_coded(
            _ptr(y), ldy, _ptr(mu), ldm, _ptr(scale), lds, N, H * W, C, _ptr(table), table.numel(), scale_bound, lik_bound,
            _ptr(y_hat), ldyh, _ptr(lik), ldl, _ptr(symbols), _ptr(indexes), _ptr(coder.cdf), coder.cdf.shape[1],
            _ptr(coder.cdf_len), _ptr(coder.cdf_off), _ptr(coder.packed[sl_]), _ptr(coder.raw[sl_]), _ptr(coder.flags[sl_]),
            _stream()), "rcn_gaussian_conditional_coded")
        return
    _C.check(_C.lib().rcn_gaussian_conditional(
        _ptr(y), ldy, _ptr(mu), ldm, _ptr(scale), lds, N, H * W, C, _ptr(table), table.numel(), scale_bound, lik_bound,
        _ptr(y_hat), ldyh, _ptr(lik), ldl, _ptr(symbols), _ptr(indexes), _stream()), "rcn_gaussian_conditional")


def build_indexes(scale, table, indexes, scale_bound=0.11):
    N, H, W, C, lds = geom(scale)
    _C.check(_C.lib().rcn_build_indexes(_ptr(scale), lds, N, H * W, C, _ptr(table), table.numel(), scale_bound, _ptr(indexes),
                                        _stream()), "rcn_build_indexes")


def gaussian_dequantize(symbols, mu, y_hat):
    N, H, W, C, ldm = geom(mu)
    _C.check(_C.lib().rcn_gaussian_dequantize(_ptr(symbols), _ptr(mu), ldm, N, H * W, C, _ptr(y_hat), geom(y_hat)[4], _stream()),
             "rcn_gaussian_dequantize")
