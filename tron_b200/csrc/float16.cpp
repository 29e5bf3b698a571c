/*
 * float16.cpp -- host-side binary16 conversions (see include/float16.h).
 *
 * Semantics follow /root/reference/src/float16.cu:76-324 bit for bit, including
 * the subnormal rounding quirk (float16.cu:112-126 / 201-216): the significand
 * is shifted right before the ties-to-even test, so bits below the shifted-out
 * position never act as sticky bits.  The implementation is one template over
 * the source format instead of two hand-unrolled copies.
 */
#include "../../include/float16.h"

#include <string.h>

namespace {

/* SRC_MANT = mantissa bits of the source (23 or 52), SRC_BIAS = its exponent bias */
template <typename U, int SRC_MANT, int SRC_BIAS, int SRC_EXPBITS>
inline uint16_t to_half_bits(U v)
{
    const int total = 1 + SRC_EXPBITS + SRC_MANT;
    const U one = 1;
    const uint16_t sign = (uint16_t)((v >> (total - 16)) & 0x8000u);
    const int e = (int)((v >> SRC_MANT) & ((one << SRC_EXPBITS) - 1));      /* biased exponent */
    U m = v & ((one << SRC_MANT) - 1);
    const int drop = SRC_MANT - 10;                                          /* bits rounded away */
    const U halfway = one << (drop - 1), low = (one << (drop + 1)) - 1;

    if (e >= SRC_BIAS + 16) {                                                /* >= 2^16: inf, or nan */
        if (e == (1 << SRC_EXPBITS) - 1 && m != 0) {
            uint16_t q = (uint16_t)(0x7c00u + (uint16_t)(m >> drop));
            if (q == 0x7c00u) ++q;
            return (uint16_t)(sign + q);
        }
        return (uint16_t)(sign + 0x7c00u);
    }
    if (e <= SRC_BIAS - 15) {                                                /* zero or subnormal half */
        if (e < SRC_BIAS - 25) return sign;
        U s = ((one << SRC_MANT) + m) >> (SRC_BIAS - 14 - e);
        if ((s & low) != halfway) s += halfway;
        return (uint16_t)(sign + (uint16_t)(s >> drop));
    }
    const uint16_t he = (uint16_t)((e - (SRC_BIAS - 15)) << 10);
    if ((m & low) != halfway) m += halfway;
    return (uint16_t)(sign + he + (uint16_t)(m >> drop));                    /* mantissa carry bumps he */
}

template <typename U, int DST_MANT, int DST_BIAS, int DST_EXPBITS>
inline U from_half_bits(uint16_t h)
{
    const int total = 1 + DST_EXPBITS + DST_MANT;
    const U sign = (U)(h & 0x8000u) << (total - 16);
    int e = (h >> 10) & 0x1f;
    U m = h & 0x3ffu;
    if (e == 0x1f) return sign + ((((U)1 << DST_EXPBITS) - 1) << DST_MANT) + (m << (DST_MANT - 10));
    if (e == 0) {
        if (m == 0) return sign;
        e = 1;
        while (!(m & 0x400u)) { m <<= 1; --e; }                             /* normalise */
        m &= 0x3ffu;
    }
    return sign + ((U)(e + DST_BIAS - 15) << DST_MANT) + (m << (DST_MANT - 10));
}

} // namespace

uint16_t floatbits_to_halfbits(uint32_t f) { return to_half_bits<uint32_t, 23, 127, 8>(f); }
uint16_t doublebits_to_halfbits(uint64_t d) { return to_half_bits<uint64_t, 52, 1023, 11>(d); }
uint32_t float16bits_to_floatbits(uint16_t h) { return from_half_bits<uint32_t, 23, 127, 8>(h); }
uint64_t float16bits_to_doublebits(uint16_t h) { return from_half_bits<uint64_t, 52, 1023, 11>(h); }

float float16_to_float(float16 h)
{
    uint32_t b = float16bits_to_floatbits(h); float f; memcpy(&f, &b, sizeof f); return f;
}
double float16_to_double(float16 h)
{
    uint64_t b = float16bits_to_doublebits(h); double d; memcpy(&d, &b, sizeof d); return d;
}
float16 float_to_float16(float f)
{
    uint32_t b; memcpy(&b, &f, sizeof b); return floatbits_to_halfbits(b);
}
float16 double_to_float16(double d)
{
    uint64_t b; memcpy(&b, &d, sizeof b); return doublebits_to_halfbits(b);
}

extern "C" {
uint16_t tron_floatbits_to_halfbits(uint32_t f) { return floatbits_to_halfbits(f); }
uint16_t tron_doublebits_to_halfbits(uint64_t d) { return doublebits_to_halfbits(d); }
uint32_t tron_halfbits_to_floatbits(uint16_t h) { return float16bits_to_floatbits(h); }
uint64_t tron_halfbits_to_doublebits(uint16_t h) { return float16bits_to_doublebits(h); }

void tron_float_to_half_array(uint16_t *dst, const float *src, size_t n)
{
    for (size_t i = 0; i < n; ++i) { uint32_t b; memcpy(&b, src + i, 4); dst[i] = floatbits_to_halfbits(b); }
}
void tron_half_to_float_array(float *dst, const uint16_t *src, size_t n)
{
    for (size_t i = 0; i < n; ++i) { uint32_t b = float16bits_to_floatbits(src[i]); memcpy(dst + i, &b, 4); }
}
}
