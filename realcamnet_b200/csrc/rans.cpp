// Host-side range coder of the product: rANS with a 64-bit state, 32-bit renormalisation words,
// 16-bit probabilities and 4-bit bypass nibbles -- stream-compatible with what the reference
// obtains from CompressAI's C++ coder (BufferedRansEncoder / RansDecoder, call sites
// models/raw2bit.py:1921,1956-1957,1996-1997,2013; models/tcm.py:531,566-567,606-607,621).
//
// The encoder walks the symbol list BACKWARDS and emits each symbol's coding events in reverse
// (raw nibbles high->low, nibble-count digits, then the CDF bin), writing words from the end of
// the buffer, so no intermediate event list is materialised.  The decoder keeps its state in an
// object so the five-slice loop can pull one slice at a time from a single stream.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <memory>
#include <new>
#include <chrono>
#include <thread>
#include <vector>

#include "../../include/rcn_b200.h"

namespace rcn {
void set_error(const char* fmt, ...);
}

namespace {

constexpr uint32_t kProbBits = 16;
constexpr uint32_t kNibbleBits = 4;
constexpr uint32_t kNibbleMax = 15;
constexpr uint64_t kLow = 1ull << 31;

struct Writer {
    uint32_t* cur;  // next free slot is cur - 1
    uint64_t x = kLow;
    inline void bin(uint32_t start, uint32_t freq) {
        const uint64_t limit = ((kLow >> kProbBits) << 32) * (uint64_t)freq;
        if (x >= limit) { *--cur = (uint32_t)x; x >>= 32; }
        x = ((x / freq) << kProbBits) + (x % freq) + start;
    }
    inline void nibble(uint32_t v) {
        const uint64_t limit = ((kLow >> kProbBits) << 32) << (kProbBits - kNibbleBits);
        if (x >= limit) { *--cur = (uint32_t)x; x >>= 32; }
        x = (x << kNibbleBits) | v;
    }
};

struct Classified {
    int bin;        // index into the CDF row
    uint32_t raw;   // bypass payload when bin == sentinel
    bool escaped;
};

inline Classified classify(int symbol, int offset, int sentinel) {
    Classified c;
    int v = symbol - offset;
    c.raw = 0;
    if (v < 0) { c.raw = (uint32_t)(-2 * (int64_t)v - 1); v = sentinel; }
    else if (v >= sentinel) { c.raw = (uint32_t)(2 * ((int64_t)v - sentinel)); v = sentinel; }
    c.bin = v;
    c.escaped = (v == sentinel);
    return c;
}

inline int nibble_count(uint32_t raw) {
    int n = 0;
    while (n < 8 && (raw >> (n * kNibbleBits)) != 0) ++n;
    return n;
}

}  // namespace

// Division-free bin update (ryg_rans "RansEncSymbol" identity): for 2 <= freq <= 65536
//   x / freq == mulhi(x, rcp[freq]) >> shift[freq]      for every x < 2^63
// so  ((x / freq) << 16) + (x % freq) + start == x + start + q * (65536 - freq).
struct RcpTable {
    uint64_t rcp[65537];
    uint8_t shift[65537];
    RcpTable() {
        rcp[0] = rcp[1] = ~0ull;
        shift[0] = shift[1] = 0;
        for (uint32_t f = 2; f <= 65536; ++f) {
            uint32_t sh = 0;
            while (f > (1u << sh)) ++sh;
            const unsigned __int128 num = ((unsigned __int128)1 << (sh + 63)) + f - 1;
            rcp[f] = (uint64_t)(num / f);
            shift[f] = (uint8_t)(sh - 1);
        }
    }
};
static const RcpTable& rcp_table() {
    static const RcpTable t;
    return t;
}

struct Escape { long long pos; uint32_t raw; };

// The inherently serial part: the rANS state chain over the symbols in reverse (escapes sorted by position).
static long long state_chain(const uint32_t* packed, long long n, const std::vector<Escape>& escapes, uint8_t* out, long long out_cap) {
    long long extra = 0;
    for (const Escape& e : escapes) { const int nb = nibble_count(e.raw); extra += nb + nb / (int)kNibbleMax + 1; }
    const size_t words = (size_t)(n + extra) + 2;
    std::unique_ptr<uint32_t[]> buf(new (std::nothrow) uint32_t[words]);
    if (!buf) {
        rcn::set_error("rcn_rans_encode: out of host memory");
        return RCN_ERR_NOMEM;
    }
    const RcpTable& R = rcp_table();
    Writer w;
    w.cur = buf.get() + words;
    long long ei = (long long)escapes.size() - 1;
    for (long long i = n - 1; i >= 0; --i) {
        if (ei >= 0 && escapes[(size_t)ei].pos == i) {
            const uint32_t raw = escapes[(size_t)ei].raw;
            --ei;
            const int nb = nibble_count(raw);
            for (int j = nb - 1; j >= 0; --j) w.nibble((raw >> (j * kNibbleBits)) & kNibbleMax);
            // count is written as 15,15,...,r in coding order -> reversed here: r first, then the 15s
            w.nibble((uint32_t)(nb % (int)kNibbleMax));
            for (int k = 0; k < nb / (int)kNibbleMax; ++k) w.nibble(kNibbleMax);
        }
        const uint32_t pk = packed[(size_t)i];
        const uint32_t start = pk >> 16, freq = (pk & 0xFFFFu) + 1;
        uint64_t x = w.x;
        const uint64_t limit = ((kLow >> kProbBits) << 32) * (uint64_t)freq;
        if (x >= limit) { *--w.cur = (uint32_t)x; x >>= 32; }
        if (freq == 1) {
            x = (x << kProbBits) + start;
        } else {
            const uint64_t q = (uint64_t)(((unsigned __int128)x * R.rcp[freq]) >> 64) >> R.shift[freq];
            x = x + start + q * (uint64_t)((1u << kProbBits) - freq);
        }
        w.x = x;
    }
    *--w.cur = (uint32_t)(w.x >> 32);
    *--w.cur = (uint32_t)w.x;
    const long long nbytes = (long long)(buf.get() + words - w.cur) * 4;
    if (nbytes > out_cap) {
        rcn::set_error("rcn_rans_encode: output buffer too small (%lld > %lld)", nbytes, out_cap);
        return RCN_ERR_NOMEM;
    }
    memcpy(out, w.cur, (size_t)nbytes);
    return nbytes;
}

extern "C" long long rcn_rans_encode(const int32_t* symbols, const int32_t* indexes, long long n, const int32_t* cdfs,
                                     int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* out,
                                     long long out_cap) {
    if ((n > 0 && (!symbols || !indexes)) || !cdfs || !cdf_sizes || !offsets || !out || n < 0) {
        rcn::set_error("rcn_rans_encode: bad arguments");
        return RCN_ERR_INVALID;
    }
    const bool timing = getenv("RCN_RANS_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    // ---- pass 1 (parallel over symbols): CDF lookups -> packed (start, freq-1) per symbol; escapes listed per chunk
    std::unique_ptr<uint32_t[]> packed(new (std::nothrow) uint32_t[(size_t)n + 1]);
    if (!packed) {
        rcn::set_error("rcn_rans_encode: out of host memory");
        return RCN_ERR_NOMEM;
    }
    int nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads > 16) nthreads = 16;
    if (nthreads < 1 || n < (1 << 16)) nthreads = 1;
    std::vector<std::vector<Escape>> esc((size_t)nthreads);
    auto lookup = [&](int tid) {
        const long long lo = n * tid / nthreads, hi = n * (tid + 1) / nthreads;
        std::vector<Escape>& mine = esc[(size_t)tid];
        for (long long i = lo; i < hi; ++i) {
            const int ci = indexes[i];
            const int32_t* row = cdfs + (long long)ci * cdf_stride;
            const int sentinel = cdf_sizes[ci] - 2;
            const Classified c = classify(symbols[i], offsets[ci], sentinel);
            const uint32_t st = (uint32_t)row[c.bin], fr = (uint32_t)(row[c.bin + 1] - row[c.bin]);
            packed[(size_t)i] = (st << 16) | ((fr - 1) & 0xFFFFu);
            if (c.escaped) mine.push_back({i, c.raw});
        }
    };
    if (nthreads == 1) lookup(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(lookup, t);
        for (auto& th : pool) th.join();
    }
    std::vector<Escape> escapes;
    for (auto& v : esc) escapes.insert(escapes.end(), v.begin(), v.end());  // chunks are position-ordered
    auto t1 = std::chrono::steady_clock::now();
    const long long nbytes = state_chain(packed.get(), n, escapes, out, out_cap);
    if (timing) {
        auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "rcn_rans_encode: n=%lld threads=%d lookup %.2f ms, state chain %.2f ms\n", n, nthreads,
                std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
    }
    return nbytes;
}

// Back end only: the per-symbol CDF lookups were done on the GPU (rcn_gaussian_conditional_coded); the host runs the
// serial state chain over packed = (start << 16) | (freq - 1), with flags[i] != 0 marking escaped symbols whose bypass
// payload is raw[i].
extern "C" long long rcn_rans_encode_packed(const uint32_t* packed, const uint32_t* raw, const uint8_t* flags, long long n,
                                            uint8_t* out, long long out_cap) {
    if ((n > 0 && (!packed || !raw || !flags)) || n < 0 || !out) {
        rcn::set_error("rcn_rans_encode_packed: bad arguments");
        return RCN_ERR_INVALID;
    }
    std::vector<Escape> escapes;
    for (long long i = 0; i < n; ++i)
        if (flags[i]) escapes.push_back({i, raw[i]});
    return state_chain(packed, n, escapes, out, out_cap);
}

struct rcn_rans_decoder {
    static constexpr int kLutShift = 7, kLutSize = 65536 >> kLutShift;
    std::vector<uint32_t> words;
    size_t pos;
    uint64_t x;
    // CDF search accelerator for the table this decoder is used with: lut[row * kLutSize + b] = symbol whose interval holds the
    // cumulative value b << kLutShift (512 buckets of 128: 1 KB per row, the walk that follows is a few steps even for the widest rows); sentinel[row] = size - 2 (the escape symbol).  Built -- and every row validated -- once per (table, stride, rows).
    const int32_t* lut_cdfs = nullptr;
    int lut_stride = 0, lut_rows = 0;
    std::vector<uint16_t> lut;
    std::vector<int32_t> sentinel;
    // returns the index of the first malformed row, or -1
    int prepare(const int32_t* cdfs, int stride, int n_rows, const int32_t* sizes) {
        if (cdfs == lut_cdfs && stride == lut_stride && n_rows == lut_rows) return -1;
        lut.assign((size_t)n_rows * kLutSize, 0);
        sentinel.assign((size_t)n_rows, 0);
        for (int r = 0; r < n_rows; ++r) {
            const int size = sizes[r];
            if (size < 2 || size > stride) { lut_cdfs = nullptr; return r; }
            const int32_t* row = cdfs + (long long)r * stride;
            sentinel[(size_t)r] = size - 2;
            int sym = 0;
            for (int b = 0; b < kLutSize; ++b) {
                const int32_t cum = b << kLutShift;
                while (sym + 2 < size && row[sym + 1] <= cum) ++sym;
                lut[(size_t)r * kLutSize + (size_t)b] = (uint16_t)sym;
            }
        }
        lut_cdfs = cdfs; lut_stride = stride; lut_rows = n_rows;
        return -1;
    }
    inline void refill() {
        if (x < kLow) { x = (x << 32) | (pos < words.size() ? words[pos] : 0u); ++pos; }
    }
    inline uint32_t nibble() {
        const uint32_t v = (uint32_t)(x & kNibbleMax);
        x >>= kNibbleBits;
        refill();
        return v;
    }
};

extern "C" rcn_rans_decoder* rcn_rans_decoder_create(const uint8_t* stream, long long nbytes) {
    if (!stream || nbytes < 8 || (nbytes & 3)) {
        rcn::set_error("rcn_rans_decoder_create: stream must hold at least the 8-byte state and be word aligned in length");
        return nullptr;
    }
    rcn_rans_decoder* d = new (std::nothrow) rcn_rans_decoder();
    if (!d) return nullptr;
    d->words.resize((size_t)nbytes / 4);
    memcpy(d->words.data(), stream, (size_t)nbytes);
    d->x = (uint64_t)d->words[0] | ((uint64_t)d->words[1] << 32);
    d->pos = 2;
    return d;
}

extern "C" void rcn_rans_decoder_destroy(rcn_rans_decoder* d) { delete d; }

// The stream and the indexes are untrusted input (a bitstream container can be crafted): every table row index is range-checked,
// the bypass length is bounded by what a 32-bit payload needs, and reading past the end of the stream is an error.
extern "C" int rcn_rans_decode(rcn_rans_decoder* d, const int32_t* indexes, long long n, const int32_t* cdfs, int cdf_stride,
                               int n_rows, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* out) {
    if (!d || !indexes || !cdfs || !cdf_sizes || !offsets || !out || n < 0 || n_rows <= 0 || cdf_stride < 2) {
        rcn::set_error("rcn_rans_decode: bad arguments");
        return RCN_ERR_INVALID;
    }
    const int bad = d->prepare(cdfs, cdf_stride, n_rows, cdf_sizes);
    if (bad >= 0) {
        rcn::set_error("rcn_rans_decode: CDF row %d has size %d (stride %d)", bad, cdf_sizes[bad], cdf_stride);
        return RCN_ERR_INVALID;
    }
    // the state chain runs on local copies (registers): 5.2 M symbols of a 2048^2 tile decode in one call per slice
    const uint16_t* lut = d->lut.data();
    const int32_t* sent = d->sentinel.data();
    const uint32_t* w = d->words.data();
    const size_t nw = d->words.size();
    uint64_t x = d->x;
    size_t pos = d->pos;
#define RCN_REFILL()                                                      \
    do {                                                                  \
        if (x < kLow) { x = (x << 32) | (pos < nw ? w[pos] : 0u); ++pos; } \
    } while (0)
#define RCN_NIBBLE(v)                           \
    do {                                        \
        (v) = (uint32_t)(x & kNibbleMax);       \
        x >>= kNibbleBits;                      \
        RCN_REFILL();                           \
    } while (0)
    for (long long i = 0; i < n; ++i) {
        const int ci = indexes[i];
        if ((unsigned)ci >= (unsigned)n_rows) {
            d->x = x; d->pos = pos;
            rcn::set_error("rcn_rans_decode: index %d at position %lld is outside the %d-row CDF table", ci, i, n_rows);
            return RCN_ERR_INVALID;
        }
        const int32_t* row = cdfs + (long long)ci * cdf_stride;
        const uint32_t cum = (uint32_t)(x & 0xFFFFu);
        // symbol s with row[s] <= cum < row[s+1] (rows are strictly increasing): start from the bucket table, walk forward
        int s = lut[(size_t)ci * rcn_rans_decoder::kLutSize + (cum >> rcn_rans_decoder::kLutShift)];
        while (row[s + 1] <= (int32_t)cum) ++s;
        const uint32_t start = (uint32_t)row[s], freq = (uint32_t)(row[s + 1] - row[s]);
        x = (uint64_t)freq * (x >> kProbBits) + cum - start;
        RCN_REFILL();
        int v = s;
        const int sentinel = sent[ci];
        if (s == sentinel) {
            uint32_t digit;
            RCN_NIBBLE(digit);
            int nb = (int)digit;
            while (digit == kNibbleMax && nb <= 8) { RCN_NIBBLE(digit); nb += (int)digit; }
            if (nb > 8) {   // a 32-bit payload needs at most 8 nibbles
                d->x = x; d->pos = pos;
                rcn::set_error("rcn_rans_decode: corrupt stream (bypass length %d nibbles at position %lld)", nb, i);
                return RCN_ERR_INVALID;
            }
            uint32_t raw = 0;
            for (int j = 0; j < nb; ++j) { uint32_t t; RCN_NIBBLE(t); raw |= t << (j * kNibbleBits); }
            v = (int)(raw >> 1);
            v = (raw & 1u) ? -v - 1 : v + sentinel;
        }
        out[i] = v + offsets[ci];
    }
#undef RCN_NIBBLE
#undef RCN_REFILL
    d->x = x;
    d->pos = pos;
    if (d->pos > d->words.size()) {
        rcn::set_error("rcn_rans_decode: truncated or corrupt stream (read %zu words past its end)", d->pos - d->words.size());
        return RCN_ERR_INVALID;
    }
    return RCN_OK;
}

// pmf (float32) -> strictly increasing 16-bit CDF of length n+1 (CompressAI _CXX.pmf_to_quantized_cdf semantics,
// used by EntropyBottleneck.update / GaussianConditional.update behind raw2bit.py:1759-1764).
extern "C" int rcn_pmf_to_quantized_cdf(const float* pmf, int n, int precision, int32_t* cdf) {
    if (!pmf || !cdf || n <= 0 || precision < 1 || precision > 16) {
        rcn::set_error("rcn_pmf_to_quantized_cdf: bad arguments");
        return RCN_ERR_INVALID;
    }
    const uint32_t one = 1u << precision;
    std::vector<uint32_t> c((size_t)n + 1, 0u);
    uint32_t total = 0;
    for (int i = 0; i < n; ++i) {
        const float p = pmf[i];
        if (!(p >= 0.f) || !std::isfinite(p)) {
            rcn::set_error("rcn_pmf_to_quantized_cdf: pmf[%d] is negative or not finite", i);
            return RCN_ERR_INVALID;
        }
        c[(size_t)i + 1] = (uint32_t)std::round(p * (float)one);
        total += c[(size_t)i + 1];
    }
    if (total == 0) {
        rcn::set_error("rcn_pmf_to_quantized_cdf: pmf sums to zero");
        return RCN_ERR_INVALID;
    }
    uint32_t run = 0;
    for (size_t i = 0; i < c.size(); ++i) {
        run += (uint32_t)(((uint64_t)one * c[i]) / total);
        c[i] = run;
    }
    c.back() = one;
    const int m = n + 1;
    for (int i = 0; i + 1 < m; ++i) {
        if (c[i] != c[i + 1]) continue;
        uint32_t narrowest = ~0u;
        int donor = -1;
        for (int j = 0; j + 1 < m; ++j) {
            const uint32_t f = c[j + 1] - c[j];
            if (f > 1 && f < narrowest) { narrowest = f; donor = j; }
        }
        if (donor < 0) {
            rcn::set_error("rcn_pmf_to_quantized_cdf: cannot make the CDF strictly increasing");
            return RCN_ERR_INVALID;
        }
        if (donor < i) for (int j = donor + 1; j <= i; ++j) c[j]--;
        else for (int j = i + 1; j <= donor; ++j) c[j]++;
    }
    for (int i = 0; i < m; ++i) cdf[i] = (int32_t)c[i];
    return RCN_OK;
}
