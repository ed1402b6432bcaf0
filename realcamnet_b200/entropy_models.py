"""Host-side mirrors of the entropy models and range coder the reference imports from CompressAI
(``compressai.entropy_models.{EntropyBottleneck, GaussianConditional}``, ``compressai.ans.*``,
``compressai.models.CompressionModel``; imports at models/tcm.py:1-3, call sites
models/raw2bit.py:1756-1764,1803-1807,1829,1906-1907,1917-1921,1939-1957,1983-2015).

Split of work: per-element arithmetic (quantise, likelihood, index lookup) = CUDA kernels;
the range coder and the pmf->CDF quantiser = native C++ in librcn_b200.so (host); building the
float pmf tables in ``update()`` = host-side torch CPU code exactly as CompressAI does it.
Parameter / buffer names follow CompressAI so reference checkpoints keep their keys.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _C, ops
from .layers import LowerBound


# ----------------------------------------------------------------------------- native coder
def _ip(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i32(t):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.int32)


def pmf_to_quantized_cdf(pmf, precision: int = 16) -> torch.Tensor:
    p = np.ascontiguousarray(pmf.detach().cpu().numpy() if isinstance(pmf, torch.Tensor) else pmf, dtype=np.float32)
    out = np.empty(p.size + 1, dtype=np.int32)
    _C.check(_C.lib().rcn_pmf_to_quantized_cdf(_ip(p), p.size, precision, _ip(out)), "rcn_pmf_to_quantized_cdf")
    return torch.from_numpy(out)


def rans_encode(symbols, indexes, cdf, cdf_length, offset) -> bytes:
    symbols, indexes = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
    cdf, cdf_length, offset = _i32(cdf), _i32(cdf_length).reshape(-1), _i32(offset).reshape(-1)
    if symbols.size != indexes.size:
        raise ValueError("symbols and indexes must have the same length")
    if indexes.size and (indexes.min() < 0 or indexes.max() >= cdf.shape[0]):
        raise ValueError("index out of range of the CDF table")
    cap = 4 * symbols.size + 64
    out = np.empty(cap, dtype=np.uint8)
    lib = _C.lib()
    n = lib.rcn_rans_encode(_ip(symbols), _ip(indexes), symbols.size, _ip(cdf), cdf.shape[1], _ip(cdf_length), _ip(offset),
                            _ip(out), cap)
    if n == _C_ERR_NOMEM:  # pathological escape density: retry with the worst-case bound
        cap = 44 * symbols.size + 64
        out = np.empty(cap, dtype=np.uint8)
        n = lib.rcn_rans_encode(_ip(symbols), _ip(indexes), symbols.size, _ip(cdf), cdf.shape[1], _ip(cdf_length),
                                _ip(offset), _ip(out), cap)
    _C.check(n, "rcn_rans_encode")
    return out[:n].tobytes()


_C_ERR_NOMEM = -3


def rans_encode_packed(packed, raw, flags) -> bytes:
    """State chain only: `packed` = (start << 16) | (freq - 1) per symbol from the GPU front end
    (rcn_gaussian_conditional_coded); flags[i] != 0 marks an escaped symbol with bypass payload raw[i].
    Same bytes as rans_encode()."""
    packed = np.ascontiguousarray(packed).reshape(-1).view(np.uint32)
    raw = np.ascontiguousarray(raw).reshape(-1).view(np.uint32)
    flags = np.ascontiguousarray(flags, dtype=np.uint8).reshape(-1)
    if not (packed.size == raw.size == flags.size):
        raise ValueError("packed, raw and flags must have the same length")
    cap = 4 * packed.size + 48 * int(np.count_nonzero(flags)) + 64
    out = np.empty(cap, dtype=np.uint8)
    n = _C.lib().rcn_rans_encode_packed(_ip(packed), _ip(raw), _ip(flags), packed.size, _ip(out), cap)
    _C.check(n, "rcn_rans_encode_packed")
    return out[:n].tobytes()


class BufferedRansEncoder:
    """compressai.ans.BufferedRansEncoder (models/raw2bit.py:1921,1956-1957)."""

    def __init__(self):
        self._s, self._i, self._tab = [], [], None

    def encode_with_indexes(self, symbols, indexes, cdf, cdf_length, offset):
        self._s.append(_i32(symbols).reshape(-1))
        self._i.append(_i32(indexes).reshape(-1))
        self._tab = (cdf, cdf_length, offset)

    def flush(self) -> bytes:
        if self._tab is None:
            return rans_encode(np.zeros(0, np.int32), np.zeros(0, np.int32), np.array([[0, 65536]], np.int32), [2], [0])
        out = rans_encode(np.concatenate(self._s), np.concatenate(self._i), *self._tab)
        self._s, self._i = [], []
        return out


class RansDecoder:
    """compressai.ans.RansDecoder (models/raw2bit.py:1996-1997,2013)."""

    def __init__(self):
        self._h = None

    def set_stream(self, stream: bytes):
        self.close()
        buf = np.frombuffer(stream, dtype=np.uint8)
        h = _C.lib().rcn_rans_decoder_create(_ip(buf), buf.size)
        if not h:
            raise RuntimeError("rcn_rans_decoder_create failed: " + _C.lib().rcn_last_error().decode())
        self._h = ctypes.c_void_p(h)

    def decode_stream(self, indexes, cdf, cdf_length, offset) -> np.ndarray:
        if self._h is None:
            raise RuntimeError("set_stream() first")
        indexes = _i32(indexes).reshape(-1)
        cdf, cdf_length, offset = _i32(cdf), _i32(cdf_length).reshape(-1), _i32(offset).reshape(-1)
        out = np.empty(indexes.size, dtype=np.int32)
        _C.check(_C.lib().rcn_rans_decode(self._h, _ip(indexes), indexes.size, _ip(cdf), cdf.shape[1], cdf.shape[0], _ip(cdf_length),
                                          _ip(offset), _ip(out)), "rcn_rans_decode")
        return out

    def close(self):
        if self._h is not None:
            _C.lib().rcn_rans_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------- entropy models
class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.likelihood_bound = float(likelihood_bound)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._host_tables = None

    offset = property(lambda self: self._offset)
    quantized_cdf = property(lambda self: self._quantized_cdf)
    cdf_length = property(lambda self: self._cdf_length)

    def host_tables(self):
        """(cdf, cdf_length, offset) as contiguous int32 numpy arrays (cached)."""
        if self._offset.numel() == 0:
            raise RuntimeError("call update() before compress()/decompress() (models/raw2bit.py:1759-1764)")
        key = (self._quantized_cdf.data_ptr(), self._quantized_cdf._version)
        if self._host_tables is None or self._host_tables[0] != key or self._host_tables[2] is not self._quantized_cdf:
            self._host_tables = (key, (_i32(self._quantized_cdf), _i32(self._cdf_length).reshape(-1), _i32(self._offset).reshape(-1)),
                                 self._quantized_cdf)
        return self._host_tables[1]

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            row = pmf_to_quantized_cdf(prob, self.entropy_coder_precision)
            cdf[i, : row.size(0)] = row
        return cdf


class EntropyBottleneck(EntropyModel):
    """Factorized prior on z (compressai EntropyBottleneck, eval mode)."""

    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            raise NotImplementedError("kernel is specialised for filters=(3,3,3,3), the only setting the reference uses")
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(torch.full((channels, filters[i + 1], filters[i]), float(init))))
            self.register_parameter(f"_bias{i:d}", nn.Parameter(torch.empty(channels, filters[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(channels, filters[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(channels, 1, 1))
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))
        self._packed = None

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    # -- device parameters for rcn_eb_forward: [C][58]
    def _pack(self):
        ps = [getattr(self, f"_matrix{i}") for i in range(5)] + [getattr(self, f"_bias{i}") for i in range(5)] + \
             [getattr(self, f"_factor{i}") for i in range(4)] + [self.quantiles]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is None or self._packed[0] != key:
            with torch.no_grad():
                C = self.channels
                cols = []
                for i in range(5):
                    cols.append(F.softplus(getattr(self, f"_matrix{i}")).reshape(C, -1))
                    cols.append(getattr(self, f"_bias{i}").reshape(C, -1))
                    if i < 4:
                        cols.append(torch.tanh(getattr(self, f"_factor{i}")).reshape(C, -1))
                params = torch.cat(cols, dim=1).contiguous()
                assert params.shape[1] == 58
                med = self.quantiles[:, 0, 1].contiguous()
            self._packed = (key, params, med)
        return self._packed[1], self._packed[2]

    def _f(self, z, want_symbols=False, want_lik=True):
        """z NHWC -> (z_hat NHWC, likelihood NHWC, symbols NCHW int32 | None)"""
        params, med = self._pack()
        return ops.eb_forward(z, params, med, True, want_lik, want_symbols,
                              self.likelihood_bound if self.use_likelihood_bound else 0.0)

    def forward(self, x, training=None):
        if training or (training is None and self.training):
            raise NotImplementedError("inference path only: call .eval() (the reference adds uniform noise in train mode)")
        z_hat, lik, _ = self._f(ops.to_nhwc(x))
        return ops.to_nchw(z_hat), ops.to_nchw(lik)

    # -- host side (CPU torch, as CompressAI) -------------------------------------------------
    def _cpu_logits_cumulative(self, inputs):
        logits = inputs
        for i in range(5):
            logits = torch.matmul(F.softplus(getattr(self, f"_matrix{i}").detach().cpu()), logits)
            logits = logits + getattr(self, f"_bias{i}").detach().cpu()
            if i < 4:
                logits = logits + torch.tanh(getattr(self, f"_factor{i}").detach().cpu()) * torch.tanh(logits)
        return logits

    def update(self, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        q = self.quantiles.detach().cpu()
        medians = q[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max().item())
        samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
        lower = self._cpu_logits_cumulative(samples - 0.5)
        upper = self._cpu_logits_cumulative(samples + 0.5)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        dev = self.quantiles.device
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
        self._offset = (-minima).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)
        return True

    def compress(self, x):
        """x NCHW (device) -> list of byte strings, one stream per batch element."""
        _, _, sym = ops.eb_forward(ops.to_nhwc(x), *self._pack(), want_zhat=False, want_lik=False, want_symbols=True)
        return self.compress_symbols(sym)

    def compress_symbols(self, sym):
        cdf, length, offset = self.host_tables()
        s = sym.cpu().numpy()
        N, C, H, W = s.shape
        idx = np.ascontiguousarray(np.broadcast_to(np.arange(C, dtype=np.int32)[:, None, None], (C, H, W))).reshape(-1)
        return [rans_encode(s[i].reshape(-1), idx, cdf, length, offset) for i in range(N)]

    def decompress_symbols(self, strings, size):
        cdf, length, offset = self.host_tables()
        C = cdf.shape[0]
        H, W = int(size[0]), int(size[1])
        idx = np.ascontiguousarray(np.broadcast_to(np.arange(C, dtype=np.int32)[:, None, None], (C, H, W))).reshape(-1)
        out = np.empty((len(strings), C, H, W), dtype=np.int32)
        for i, s in enumerate(strings):
            d = RansDecoder()
            d.set_stream(s)
            out[i] = d.decode_stream(idx, cdf, length, offset).reshape(C, H, W)
            d.close()
        return out

    def _decompress_nhwc(self, strings, size):
        sym = torch.from_numpy(self.decompress_symbols(strings, size)).to(self.quantiles.device)
        return ops.eb_dequantize(sym, self._pack()[1])

    def decompress(self, strings, size):
        return ops.to_nchw(self._decompress_nhwc(strings, size))


class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))
        self._scale_bound = float(np.float32(scale_bound))

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        dev = self.scale_bound.device
        self.scale_table = self._prepare_scale_table(scale_table).to(dev)
        self.update()
        return True

    def update(self):
        import scipy.stats

        st = self.scale_table.detach().cpu()
        multiplier = -scipy.stats.norm.ppf(self.tail_mass / 2)
        pmf_center = torch.ceil(st * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        scale = st.unsqueeze(1).float()
        const = float(-(2 ** -0.5))
        upper = 0.5 * torch.erfc(const * ((0.5 - samples) / scale))
        lower = 0.5 * torch.erfc(const * ((-0.5 - samples) / scale))
        dev = self.scale_bound.device
        self._quantized_cdf = self._pmf_to_cdf(upper - lower, 2 * lower[:, :1], pmf_length, max_length).to(dev)
        self._offset = (-pmf_center).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)

    def table_device(self):
        if self.scale_table.numel() == 0:
            raise RuntimeError("scale table is empty: call model.update() first")
        return self.scale_table

    def forward(self, inputs, scales, means=None, training=None):
        """(y_hat, likelihood) for NCHW tensors, eval mode (models/raw2bit.py:1829)."""
        if training or (training is None and self.training):
            raise NotImplementedError("inference path only: call .eval()")
        y, s = ops.to_nhwc(inputs), ops.to_nhwc(scales)
        m = ops.to_nhwc(means) if means is not None else torch.zeros_like(y)
        y_hat, lik = torch.empty_like(y), torch.empty_like(y)
        table = self.scale_table if self.scale_table.numel() else self.scale_bound.new_tensor([self._scale_bound, 1e30])
        ops.gaussian_conditional(y, m, s, table, y_hat=y_hat, lik=lik, scale_bound=self._scale_bound,
                                 lik_bound=self.likelihood_bound if self.use_likelihood_bound else 0.0)
        return ops.to_nchw(y_hat), ops.to_nchw(lik)

    def build_indexes(self, scales):
        s = ops.to_nhwc(scales)
        N, H, W, C = s.shape
        idx = torch.empty((N, C, H, W), device=s.device, dtype=torch.int32)
        ops.build_indexes(s, self.table_device(), idx, self._scale_bound)
        return idx


class CompressionModel(nn.Module):
    """compressai.models.CompressionModel (>= 1.2: constructed without arguments, models/raw2bit.py:1617)."""

    def __init__(self, entropy_bottleneck_channels=None, init_weights=None):
        super().__init__()

    def update(self, scale_table=None, force=False):
        from .tcm import get_scale_table

        if scale_table is None:
            scale_table = get_scale_table()
        updated = False
        for _, m in self.named_modules():
            if isinstance(m, EntropyBottleneck):
                updated |= m.update(force=force)
            if isinstance(m, GaussianConditional):
                updated |= m.update_scale_table(scale_table, force=force)
        return updated
