// mrh_fmt.h — "%g" for the ASCII PLY writer (geowrapper.cpp:194-229 streams doubles with
// `ostream << double`, i.e. printf("%g"): 6 significant digits, trailing zeros stripped).
// fmt_g6 produces byte-identical text for the values a mesh holds (floats widened to double,
// 1e-4 <= |v| < 1e6 or 0) about 10x faster than sprintf, and hands everything else - exponent
// notation, non-finite values, decimal ties too close to call in double arithmetic - to sprintf.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace mrh {

  inline size_t fmt_uint(uint32_t v, char* dst) {
    char tmp[10];
    int n = 0;
    do {
      tmp[n++] = (char) ('0' + v % 10u);
      v /= 10u;
    } while (v);
    for (int i = 0; i < n; ++i)
      dst[i] = tmp[n - 1 - i];
    return (size_t) n;
  }
  inline size_t fmt_int(int32_t v, char* dst) {
    if (v < 0) {
      *dst = '-';
      return 1 + fmt_uint((uint32_t) (-(int64_t) v), dst + 1);
    }
    return fmt_uint((uint32_t) v, dst);
  }

  inline size_t fmt_g6(double v, char* dst) {
    static const double kPow10[10] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9};
    const double a                 = std::fabs(v);
    if (a == 0.0) {
      if (std::signbit(v)) {
        dst[0] = '-', dst[1] = '0';
        return 2;
      }
      dst[0] = '0';
      return 1;
    }
    if (!(a >= 1e-4 && a < 999999.0))
      return (size_t) sprintf(dst, "%g", v);
    // decimal exponent X of a: 10^X <= a < 10^(X+1)
    int X = a >= 1.0 ? (a >= 1e3 ? (a >= 1e5 ? 5 : (a >= 1e4 ? 4 : 3)) : (a >= 1e2 ? 2 : (a >= 1e1 ? 1 : 0))) : (a >= 1e-2 ? (a >= 1e-1 ? -1 : -2) : (a >= 1e-3 ? -3 : -4));
    const double scaled = a * kPow10[5 - X]; // in [1e5, 1e6), relative error 2^-53
    if (!(scaled >= 1e5 && scaled < 1e6))
      return (size_t) sprintf(dst, "%g", v);
    const double fl   = std::floor(scaled);
    const double frac = scaled - fl;
    if (std::fabs(frac - 0.5) < 1e-6)
      return (size_t) sprintf(dst, "%g", v);
    uint32_t digits = (uint32_t) fl + (frac > 0.5 ? 1u : 0u);
    if (digits >= 1000000u) {
      digits = 100000u;
      if (++X > 5)
        return (size_t) sprintf(dst, "%g", v);
    }
    char d[6];
    for (int i = 5; i >= 0; --i) {
      d[i] = (char) ('0' + digits % 10u);
      digits /= 10u;
    }
    int last = 5; // index of the last non-zero digit
    while (last > 0 && d[last] == '0')
      --last;
    char* p = dst;
    if (std::signbit(v))
      *p++ = '-';
    if (X >= 0) {
      for (int i = 0; i <= X; ++i)
        *p++ = d[i];
      if (last > X) {
        *p++ = '.';
        for (int i = X + 1; i <= last; ++i)
          *p++ = d[i];
      }
    } else {
      *p++ = '0', *p++ = '.';
      for (int i = 0; i < -X - 1; ++i)
        *p++ = '0';
      for (int i = 0; i <= last; ++i)
        *p++ = d[i];
    }
    return (size_t) (p - dst);
  }

} // namespace mrh
