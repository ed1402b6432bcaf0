/* Plain-C restatement of the rANS64 range coder used by the reference through
 * CompressAI (TEST ORACLE ONLY -- never linked into the product library).
 *
 * Same published algorithm as oracle/rans.py (see that header for the call sites
 * models/raw2bit.py:1921,1956-1957,1996-1997,2013 and the "parity unpinned" note);
 * this copy exists so multi-million-symbol parity checks finish in seconds.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/librans_oracle.so oracle/rans_c.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PREC 16
#define ORC_BYP 4
#define ORC_BYP_MAX 15
#define ORC_L (1ull << 31)

typedef struct { uint16_t start; uint16_t freq; uint8_t bypass; } orc_sym;

/* returns number of bytes written to out (capacity out_cap bytes), or -1 */
long long orc_rans_encode(const int32_t *symbols, const int32_t *indexes, long long n,
                          const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                          const int32_t *offsets, uint8_t *out, long long out_cap)
{
    long long cap = n + 16, cnt = 0;
    orc_sym *buf = (orc_sym *)malloc((size_t)cap * sizeof(orc_sym));
    if (!buf) return -1;
#define PUSH(S, F, B)                                                              \
    do {                                                                           \
        if (cnt == cap) {                                                          \
            cap = cap * 2;                                                         \
            buf = (orc_sym *)realloc(buf, (size_t)cap * sizeof(orc_sym));          \
            if (!buf) return -1;                                                   \
        }                                                                          \
        buf[cnt].start = (uint16_t)(S); buf[cnt].freq = (uint16_t)(F);             \
        buf[cnt].bypass = (B); cnt++;                                              \
    } while (0)
    for (long long i = 0; i < n; ++i) {
        int ci = indexes[i];
        const int32_t *cdf = cdfs + (long long)ci * cdf_stride;
        int max_value = cdf_sizes[ci] - 2;
        int v = symbols[i] - offsets[ci];
        uint32_t raw = 0;
        if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = max_value; }
        else if (v >= max_value) { raw = (uint32_t)(2 * (v - max_value)); v = max_value; }
        PUSH(cdf[v], cdf[v + 1] - cdf[v], 0);
        if (v == max_value) {
            int nb = 0;
            while (nb < 8 && (raw >> (nb * ORC_BYP)) != 0) ++nb;
            int val = nb;
            while (val >= ORC_BYP_MAX) { PUSH(ORC_BYP_MAX, 0, 1); val -= ORC_BYP_MAX; }
            PUSH(val, 0, 1);
            for (int j = 0; j < nb; ++j) PUSH((raw >> (j * ORC_BYP)) & ORC_BYP_MAX, 0, 1);
        }
    }
#undef PUSH
    uint32_t *words = (uint32_t *)malloc((size_t)(cnt + 2) * sizeof(uint32_t));
    if (!words) { free(buf); return -1; }
    uint32_t *ptr = words + cnt + 2;
    uint64_t x = ORC_L;
    for (long long i = cnt - 1; i >= 0; --i) {
        if (buf[i].bypass) {
            uint64_t x_max = ((ORC_L >> 16) << 32) * (uint64_t)(1u << (16 - ORC_BYP));
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = (x << ORC_BYP) | buf[i].start;
        } else {
            uint32_t f = buf[i].freq;
            if (f == 0) f = 65536; /* a full-range bin wraps in uint16 */
            uint64_t x_max = ((ORC_L >> ORC_PREC) << 32) * (uint64_t)f;
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = ((x / f) << ORC_PREC) + (x % f) + buf[i].start;
        }
    }
    ptr -= 2;
    ptr[0] = (uint32_t)x;
    ptr[1] = (uint32_t)(x >> 32);
    long long nbytes = (long long)((words + cnt + 2) - ptr) * 4;
    long long ret = -1;
    if (nbytes <= out_cap) { memcpy(out, ptr, (size_t)nbytes); ret = nbytes; }
    free(words); free(buf);
    return ret;
}

typedef struct { const uint32_t *w; long long p; uint64_t x; } orc_dec;

static inline uint32_t orc_bits(orc_dec *d, int nb)
{
    uint32_t val = (uint32_t)(d->x & ((1u << nb) - 1));
    d->x >>= nb;
    if (d->x < ORC_L) { d->x = (d->x << 32) | d->w[d->p++]; }
    return val;
}

/* decodes n symbols; state (word position, x) is carried in/out through st[3]
 * so the slice loop can call it repeatedly on one stream (set st[0] = -1 to init) */
int orc_rans_decode(const uint8_t *stream, long long nbytes, const int32_t *indexes, long long n,
                    const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                    const int32_t *offsets, int32_t *out, long long *st)
{
    orc_dec d;
    d.w = (const uint32_t *)stream;
    (void)nbytes;
    if (st[0] < 0) { d.x = (uint64_t)d.w[0] | ((uint64_t)d.w[1] << 32); d.p = 2; }
    else { d.p = st[0]; d.x = ((uint64_t)(uint32_t)st[1]) | ((uint64_t)(uint32_t)st[2] << 32); }
    for (long long i = 0; i < n; ++i) {
        int ci = indexes[i];
        const int32_t *cdf = cdfs + (long long)ci * cdf_stride;
        int size = cdf_sizes[ci], max_value = size - 2;
        uint32_t cum = (uint32_t)(d.x & 0xFFFF);
        int s = 0;
        while (s + 1 < size && (uint32_t)cdf[s + 1] <= cum) ++s;
        uint32_t start = (uint32_t)cdf[s], f = (uint32_t)(cdf[s + 1] - cdf[s]);
        d.x = (uint64_t)f * (d.x >> ORC_PREC) + cum - start;
        if (d.x < ORC_L) { d.x = (d.x << 32) | d.w[d.p++]; }
        int v = s;
        if (v == max_value) {
            int val = (int)orc_bits(&d, ORC_BYP), nb = val;
            while (val == ORC_BYP_MAX) { val = (int)orc_bits(&d, ORC_BYP); nb += val; }
            uint32_t raw = 0;
            for (int j = 0; j < nb; ++j) raw |= orc_bits(&d, ORC_BYP) << (j * ORC_BYP);
            v = (int)(raw >> 1);
            v = (raw & 1) ? -v - 1 : v + max_value;
        }
        out[i] = v + offsets[ci];
    }
    st[0] = d.p; st[1] = (long long)(uint32_t)d.x; st[2] = (long long)(uint32_t)(d.x >> 32);
    return 0;
}
