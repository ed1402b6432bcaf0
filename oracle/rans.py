"""Pure-Python restatement of the range coder the reference calls (TEST ORACLE ONLY).

The reference never ships this arithmetic: it lives in the un-vendored, un-pinned
third-party dependency CompressAI (>= 1.2.0) as a C++ pybind
extension over ryg_rans' public-domain ``rans64.h``.  Call sites in the reference:
``models/raw2bit.py:1921,1956-1957`` (BufferedRansEncoder.encode_with_indexes/flush),
``models/raw2bit.py:1996-1997,2013`` (RansDecoder.set_stream/decode_stream),
``models/tcm.py:531,566-567,606-607,621``.

PARITY UNPINNED: the reference holds no golden vector for this path and the real
CompressAI wheel is not available offline; the published algorithm is restated here
from knowledge of upstream and pinned only by hand-derived known-answer tests
(tests/test_rans_oracle.py).

Published algorithm (rANS, 64-bit state, 32-bit renormalisation words):
  * state x starts at L = 2**31; precision = 16 bits; symbols are coded in REVERSE
    order so the decoder reads them forwards.
  * put(start, freq):  if x >= ((L >> 16) << 32) * freq: emit low 32 bits, x >>= 32
                       x = (x // freq << 16) + x % freq + start
  * put_bits(val, 4):  same with freq = 1 << 12 and x = (x << 4) | val
  * flush: emit the two 32-bit halves of x (low word first in the final stream).
  * a value outside [0, max_value) (max_value = cdf_length - 2) is coded as the
    sentinel bin ``max_value`` followed by a bypass sequence of 4-bit nibbles:
    first the nibble count n in "unary base 15" (15,15,...,r), then n nibbles of
    ``raw`` LSB first, with raw = -2v-1 for v < 0 and 2(v - max_value) otherwise.
Output = little-endian uint32 words.
"""
from __future__ import annotations

import numpy as np

PRECISION = 16
BYPASS_BITS = 4
MAX_BYPASS = (1 << BYPASS_BITS) - 1
RANS_L = 1 << 31
_M32 = 0xFFFFFFFF


def pmf_to_quantized_cdf(pmf, precision: int = PRECISION):
    """CompressAI ``_CXX.pmf_to_quantized_cdf`` restated (float32 pmf in, list[int] out).

    Frequencies are round(p * 2**precision) computed in float32, renormalised to the
    total, prefix-summed, last entry forced to 2**precision, and every zero-width bin
    steals one count from the narrowest bin wider than 1 (first such bin wins ties).
    """
    p32 = np.asarray(pmf, dtype=np.float32)
    if not np.all(np.isfinite(p32)) or np.any(p32 < 0):
        raise ValueError("invalid pmf")
    scale = np.float32(1 << precision)
    # C++: std::round on float (half away from zero), stored into uint32
    prod = (p32 * scale).astype(np.float32)
    freq = np.floor(prod.astype(np.float64) + 0.5).astype(np.uint64)
    cdf = np.zeros(len(p32) + 1, dtype=np.uint64)
    cdf[1:] = freq
    total = int(cdf.sum()) & _M32
    if total == 0:
        raise ValueError("pmf sums to zero")
    cdf = (cdf * np.uint64(1 << precision)) // np.uint64(total)
    cdf = np.cumsum(cdf).astype(np.int64)
    cdf[-1] = 1 << precision
    cdf = cdf.tolist()
    n = len(cdf)
    for i in range(n - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq, best = None, -1
            for j in range(n - 1):
                f = cdf[j + 1] - cdf[j]
                if f > 1 and (best_freq is None or f < best_freq):
                    best_freq, best = f, j
            if best < 0:
                raise ValueError("cannot steal frequency")
            if best < i:
                for j in range(best + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best + 1):
                    cdf[j] += 1
    assert cdf[0] == 0 and cdf[-1] == (1 << precision)
    for i in range(n - 1):
        assert cdf[i + 1] > cdf[i]
    return cdf


def _expand(symbols, indexes, cdfs, cdf_sizes, offsets):
    """symbol -> list of (start, freq, is_bypass) in coding order."""
    out = []
    for s, ci in zip(symbols, indexes):
        cdf = cdfs[ci]
        max_value = cdf_sizes[ci] - 2
        v = int(s) - offsets[ci]
        raw = 0
        if v < 0:
            raw = -2 * v - 1
            v = max_value
        elif v >= max_value:
            raw = 2 * (v - max_value)
            v = max_value
        out.append((int(cdf[v]), int(cdf[v + 1]) - int(cdf[v]), False))
        if v == max_value:
            n = 0
            while (raw >> (n * BYPASS_BITS)) != 0:
                n += 1
            val = n
            while val >= MAX_BYPASS:
                out.append((MAX_BYPASS, 0, True))
                val -= MAX_BYPASS
            out.append((val, 0, True))
            for j in range(n):
                out.append(((raw >> (j * BYPASS_BITS)) & MAX_BYPASS, 0, True))
    return out


def encode_with_indexes(symbols, indexes, cdfs, cdf_sizes, offsets) -> bytes:
    """BufferedRansEncoder.encode_with_indexes + flush (one stream)."""
    cdfs = [list(map(int, c)) for c in np.asarray(cdfs).tolist()] if not isinstance(cdfs, list) else cdfs
    cdf_sizes = [int(v) for v in np.asarray(cdf_sizes).reshape(-1)]
    offsets = [int(v) for v in np.asarray(offsets).reshape(-1)]
    syms = _expand([int(v) for v in np.asarray(symbols).reshape(-1)],
                   [int(v) for v in np.asarray(indexes).reshape(-1)], cdfs, cdf_sizes, offsets)
    x = RANS_L
    words = []  # emitted back to front
    for start, freq, bypass in reversed(syms):
        if bypass:
            f = 1 << (16 - BYPASS_BITS)
            x_max = ((RANS_L >> 16) << 32) * f
            if x >= x_max:
                words.append(x & _M32)
                x >>= 32
            x = (x << BYPASS_BITS) | start
        else:
            x_max = ((RANS_L >> PRECISION) << 32) * freq
            if x >= x_max:
                words.append(x & _M32)
                x >>= 32
            x = ((x // freq) << PRECISION) + (x % freq) + start
    words.append((x >> 32) & _M32)
    words.append(x & _M32)
    words.reverse()
    return np.asarray(words, dtype="<u4").tobytes()


class Decoder:
    """RansDecoder.set_stream / decode_stream restated."""

    def __init__(self, stream: bytes):
        self.w = np.frombuffer(stream, dtype="<u4").tolist()
        self.x = self.w[0] | (self.w[1] << 32)
        self.p = 2

    def _renorm(self):
        if self.x < RANS_L:
            self.x = (self.x << 32) | self.w[self.p]
            self.p += 1

    def _bits(self, n):
        val = self.x & ((1 << n) - 1)
        self.x >>= n
        self._renorm()
        return val

    def decode_stream(self, indexes, cdfs, cdf_sizes, offsets):
        cdf_sizes = [int(v) for v in np.asarray(cdf_sizes).reshape(-1)]
        offsets = [int(v) for v in np.asarray(offsets).reshape(-1)]
        out = []
        mask = (1 << PRECISION) - 1
        for ci in np.asarray(indexes).reshape(-1).tolist():
            cdf = cdfs[ci]
            max_value = cdf_sizes[ci] - 2
            cum = self.x & mask
            s = 0
            n = cdf_sizes[ci]
            while s + 1 < n and cdf[s + 1] <= cum:  # first entry > cum, minus one
                s += 1
            start, freq = int(cdf[s]), int(cdf[s + 1]) - int(cdf[s])
            self.x = freq * (self.x >> PRECISION) + cum - start
            self._renorm()
            v = s
            if v == max_value:
                val = self._bits(BYPASS_BITS)
                nb = val
                while val == MAX_BYPASS:
                    val = self._bits(BYPASS_BITS)
                    nb += val
                raw = 0
                for j in range(nb):
                    raw |= self._bits(BYPASS_BITS) << (j * BYPASS_BITS)
                v = raw >> 1
                v = -v - 1 if (raw & 1) else v + max_value
            out.append(v + offsets[ci])
        return out


def decode_with_indexes(stream, indexes, cdfs, cdf_sizes, offsets):
    cdfs = np.asarray(cdfs).tolist() if not isinstance(cdfs, list) else cdfs
    return Decoder(stream).decode_stream(indexes, cdfs, cdf_sizes, offsets)
