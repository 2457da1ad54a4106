// acc (+)= w, c times, each addition rounded to nearest-even exactly as the reference's loop
// `hist[bin] += weight` does for the c samples of one row that fall into the same age bin
// (coal.cpp:2269, 2291-2292) -- but in O(1) instead of c dependent additions.
//
// Inside one binade every rounded addition of the same w moves acc by the same amount
// d = RN(acc + w) - acc (acc is a multiple of its ulp U, w = qU + r, and unless r == U/2 the
// rounding of r does not depend on acc), so acc + c*w "as the reference rounds it" is
// s + (c-1)*d with s = RN(acc + w), provided all partial sums stay in acc's binade, acc >= w
// (so that d and the rounding error are exact) and the sum is not a tie.  Anything else falls
// back to single rounded steps.  Validated against the plain loop by tests/test_host.py
// (host) and the stage-i parity tests (device).
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define ES_HD __host__ __device__ __forceinline__
#else
#define ES_HD inline
#endif

namespace exsum {

#if defined(__CUDA_ARCH__)
ES_HD double add_(double a, double b) { return __dadd_rn(a, b); }
ES_HD double sub_(double a, double b) { return __dsub_rn(a, b); }
ES_HD double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
ES_HD int hi32(double x) { return __double2hiint(x); }
ES_HD double from_hi(int hi) { return __hiloint2double(hi, 0); }
#else
}  // namespace exsum
#include <cmath>
namespace exsum {
ES_HD double add_(double a, double b) { volatile double r = a + b; return r; }
ES_HD double sub_(double a, double b) { volatile double r = a - b; return r; }
ES_HD double fma_(double a, double b, double c) { return std::fma(a, b, c); }
ES_HD int hi32(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
ES_HD double from_hi(int hi) { uint64_t u = (uint64_t)(uint32_t)hi << 32; double x; memcpy(&x, &u, 8); return x; }
#endif

ES_HD double add_repeated(double acc, double w, int c)
{
  while (c > 0) {
    const double s = add_(acc, w);
    if (c == 1) return s;
    const int ea = (hi32(acc) >> 20) & 0x7ff;  // biased exponent of acc (acc >= 0 on this path)
    if (acc >= w && w > 0.0 && ea > 54 && ea < 0x7fe) {
      const double d = sub_(s, acc);                  // exact (Fast2Sum, |acc| >= |w|)
      const double err = sub_(w, d);                  // exact rounding error of acc + w
      const double half_ulp = from_hi((ea - 53) << 20);
      const double t = fma_((double)(c - 1), d, s);   // exact while it stays in the binade
      const int et = (hi32(t) >> 20) & 0x7ff;
      const bool tie = (err == half_ulp) || (err == -half_ulp);
      if (et == ea && !tie) return t;
    }
    acc = s;
    c--;
  }
  return acc;
}

}  // namespace exsum
