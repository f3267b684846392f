// Special functions of the N-independent part (small.cuh), host + device so that tests/test_oracle.py can pin them on the
// CPU against scipy (compiled there with g++ and DPGP_HD defined empty).
#pragma once
#include <math.h>
#ifndef DPGP_HD
#define DPGP_HD __host__ __device__
#endif

namespace dpgp {

// digamma / trigamma for x > 0: upward recurrence to x >= 10, then the asymptotic series (next term < 1e-17 relative).
DPGP_HD inline double digamma_pos(double x) {
  double acc = 0.0;
  while (x < 10.0) { acc -= 1.0 / x; x += 1.0; }
  const double i = 1.0 / x, i2 = i * i;
  double s = 1.0 / 12.0;                                   // B14/14 .. B2/2 (Horner in 1/x^2)
  s = fma(-i2, s, 691.0 / 32760.0);
  s = fma(-i2, s, 1.0 / 132.0);
  s = fma(-i2, s, 1.0 / 240.0);
  s = fma(-i2, s, 1.0 / 252.0);
  s = fma(-i2, s, 1.0 / 120.0);
  s = fma(-i2, s, 1.0 / 12.0);
  return acc + log(x) - 0.5 * i - i2 * s;
}
DPGP_HD inline double trigamma_pos(double x) {
  double acc = 0.0;
  while (x < 10.0) { acc += 1.0 / (x * x); x += 1.0; }
  const double i = 1.0 / x, i2 = i * i;
  double s = 7.0 / 6.0;                                    // B14 .. B2
  s = fma(i2, s, -691.0 / 2730.0);
  s = fma(i2, s, 5.0 / 66.0);
  s = fma(i2, s, -1.0 / 30.0);
  s = fma(i2, s, 1.0 / 42.0);
  s = fma(i2, s, -1.0 / 30.0);
  s = fma(i2, s, 1.0 / 6.0);
  return acc + i + 0.5 * i2 + i * i2 * s;
}
DPGP_HD inline double softplus_d(double x) { return x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x)); }
DPGP_HD inline double sigmoid_d(double x) { return x >= 0.0 ? 1.0 / (1.0 + exp(-x)) : exp(x) / (1.0 + exp(x)); }

}  // namespace dpgp
