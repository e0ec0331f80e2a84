/* divcheck.c — proof by enumeration for `div_small` / `div_small2` (csrc/reduce.cuh), the division of a
 * float by a small integer m used by K4's running means:
 *     y = RN(1/m);  q0 = RN(a*y);  r = fma(-q0, m, a);  q = fma(r, y, q0)
 * is compared with the IEEE quotient a / m for EVERY float a inside the range the kernel uses it on
 * (2^-100 <= |a| <= FLT_MAX; everything else takes the generic division there) and every m in lo..hi.
 *     gcc -O3 -march=native -mfma -fopenmp -ffp-contract=off tools/divcheck.c -o /tmp/divcheck -lm
 *     /tmp/divcheck 1 64            all 2^32 operands per divisor (about 15 s per divisor on 8 cores)
 *     /tmp/divcheck 1 64 4099       every 4099th operand plus the range boundaries (tests/: seconds)
 * Exit status 0 and "failures 0" when the sequence is exact.  Outside the range it is NOT exact
 * (+-inf -> NaN, -0 -> +0, and for even m that are not powers of two the operands below 2^-122):
 * the second column counts those, to show that the guard is needed and where. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t ubits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
int main(int argc, char** argv) {
  int lo = 1, hi = 64;
  long long stride = 1;
  if (argc > 2) { lo = atoi(argv[1]); hi = atoi(argv[2]); }
  if (argc > 3) stride = atoll(argv[3]);
  const float tiny = 0x1p-100f, big = 3.402823466e+38f;
  unsigned long long failures = 0;
  for (int m = lo; m <= hi; ++m) {
    const float fm = (float)m;
    const volatile float one = 1.0f;
    const float y = one / fm;
    unsigned long long bad_in = 0, bad_out = 0, fast = 0;
#pragma omp parallel for reduction(+:bad_in,bad_out,fast) schedule(static)
    for (long long i = 0; i < (1ll << 32); i += stride) {
      /* with a stride, also visit the operands next to the range boundaries and to the powers of two */
      for (int k = 0; k < (stride > 1 ? 5 : 1); ++k) {
        uint32_t u = (uint32_t)i;
        if (k == 1) u = (ubits(tiny) + (uint32_t)(i % 64)) | ((uint32_t)i & 0x80000000u);
        if (k == 2) u = (ubits(big) - (uint32_t)(i % 64)) | ((uint32_t)i & 0x80000000u);
        if (k == 3) u = ((uint32_t)i & 0xff800000u) + (uint32_t)((i / 7) % 3);             /* 1.0, 1.0+ulp, 1.0+2ulp x 2^e */
        if (k == 4) u = ((uint32_t)i | 0x007fffffu) - (uint32_t)((i / 7) % 3);             /* just below a power of two */
        const float a = bits(u);
        const float want = a / fm;
        const float q0 = a * y;
        const float r = fmaf(-q0, fm, a);
        const float q = fmaf(r, y, q0);
        const int same = (ubits(q) == ubits(want)) || (want != want && q != q);
        const int in_range = fabsf(a) >= tiny && fabsf(a) <= big;
        if (in_range) { fast++; if (!same) bad_in++; }
        else if (!same) bad_out++;
      }
    }
    printf("m=%2d  operands on the fast path %llu  wrong there %llu   wrong outside the range (guarded) %llu\n", m, fast, bad_in, bad_out);
    fflush(stdout);
    failures += bad_in;
  }
  printf("failures %llu\n", failures);
  return failures != 0;
}
