// halfband_design.h — host-side design of the 8-coefficient polyphase-IIR
// half-band pair used by every oversampled nonlinearity.
//
// The reference delegates this to the third-party crate `halfband 0.2.0`
// (Cargo.toml:38; call sites src/utils/oversampler.rs:1,39-40,76-79), whose
// source is not part of the reference tree.  The crate is understood to port
// Laurent de Soras' HIIR: coefficients from the elliptic design for a given
// transition bandwidth, here the one that yields the "94 dB" stop-band the
// reference quotes (oversampler.rs:37).  DESIGN.md records this as a
// reconstruction ("parity unpinned").
#pragma once
#include <cmath>

namespace gd {

inline double hb_ipowp(double x, long n) {
  double r = 1.0;
  while (n > 0) { if (n & 1) r *= x; x *= x; n >>= 1; }
  return r;
}

inline void design_halfband8(float out[8], double transition = 0.0343747) {
  const int n = 8, order = 2 * n + 1;
  const double pi = 3.14159265358979323846;
  double k = std::tan((1.0 - transition * 2.0) * pi / 4.0);
  k *= k;
  const double kk = std::pow(1.0 - k * k, 0.25);
  const double e = 0.5 * (1.0 - kk) / (1.0 + kk);
  const double e4 = (e * e) * (e * e);
  const double q = e * (1.0 + e4 * (2.0 + e4 * (15.0 + 150.0 * e4)));
  for (int idx = 0; idx < n; idx++) {
    const int c = idx + 1;
    double num = 0.0, den = 0.0, term;
    int sign = 1;
    for (int i = 0;; i++) {
      term = hb_ipowp(q, (long)i * (i + 1)) * std::sin((i * 2 + 1) * c * pi / order) * sign;
      num += term;
      sign = -sign;
      if (std::fabs(term) <= 1e-100) break;
    }
    num *= std::pow(q, 0.25);
    sign = -1;
    for (int i = 1;; i++) {
      term = hb_ipowp(q, (long)i * i) * std::cos(i * 2 * c * pi / order) * sign;
      den += term;
      sign = -sign;
      if (std::fabs(term) <= 1e-100) break;
    }
    den += 0.5;
    double ww = num / den;
    ww *= ww;
    const double x = std::sqrt((1.0 - ww * k) * (1.0 - ww / k)) / (1.0 + ww);
    out[idx] = (float)((1.0 - x) / (1.0 + x));
  }
}

}  // namespace gd
