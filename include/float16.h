/*
 * float16.h -- IEEE binary16 <-> binary32/64 conversions of libtron_b200.
 *
 * Replaces /root/reference/src/float16.h:15-25 (declarations) and
 * /root/reference/src/float16.cu:42-324 (NumPy-derived host routines).
 *
 * The reference declares these with C++ linkage (no extern "C"); a C++ caller
 * that includes this header links against the same mangled names.  FFI users
 * get the unmangled tron_* entry points below.
 *
 * Rounding is round-to-nearest-even as in the reference, INCLUDING its
 * subnormal quirk (float16.cu:112-126: the significand is shifted before the
 * tie test, so sticky bits are dropped for 2^-25 < |f| < 2^-14).  The device
 * fp16 storage path (half2 loads/stores in the CUDA kernels) uses the hardware
 * cvt.rn.f16.f32, which is IEEE RNE; the two agree except on those subnormal
 * inputs (<= 1 subnormal ulp), see tests/test_float16.py.
 */
#ifndef TRON_B200_FLOAT16_H
#define TRON_B200_FLOAT16_H

#include <stddef.h>
#include <stdint.h>

typedef uint16_t float16;

#ifdef __cplusplus
/* same names and linkage as the reference */
float    float16_to_float(float16 h);
double   float16_to_double(float16 h);
float16  float_to_float16(float f);
float16  double_to_float16(double d);
uint16_t floatbits_to_halfbits(uint32_t f);
uint16_t doublebits_to_halfbits(uint64_t d);
uint32_t float16bits_to_floatbits(uint16_t h);
uint64_t float16bits_to_doublebits(uint16_t h);

extern "C" {
#endif

uint16_t tron_floatbits_to_halfbits(uint32_t f);
uint16_t tron_doublebits_to_halfbits(uint64_t d);
uint32_t tron_halfbits_to_floatbits(uint16_t h);
uint64_t tron_halfbits_to_doublebits(uint16_t h);
/* bulk converters used by ra_convert and the CLI's -H path */
void tron_float_to_half_array(uint16_t *dst, const float *src, size_t n);
void tron_half_to_float_array(float *dst, const uint16_t *src, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* TRON_B200_FLOAT16_H */
