// Java's Double.toString for the output files of both tools (host side only).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <string>

// Double.toString: shortest digit string that round-trips (JDK 19+; older JDKs print a longer string in rare cases),
// decimal notation for 1e-3 <= |x| < 1e7, computerised scientific notation otherwise
// writes at most 32 bytes at `out`, returns the end
inline char* java_double_to(char* out, double x) {
  auto lit = [&](const char* t) { while (*t) *out++ = *t++; return out; };
  if (x != x) return lit("NaN");
  if (x == 1.0 / 0.0) return lit("Infinity");
  if (x == -1.0 / 0.0) return lit("-Infinity");
  if (x == 0.0) return lit(std::signbit(x) ? "-0.0" : "0.0");
  const double a = x < 0 ? -x : x;
  char buf[40];
  auto r = std::to_chars(buf, buf + sizeof buf, a, std::chars_format::scientific);   // d[.ddddd]e[+-]xx (shortest)
  char digits[24];
  int nd = 0;
  const char* q = buf;
  for (; q < r.ptr && *q != 'e'; ++q)
    if (*q != '.') digits[nd++] = *q;
  int exp10 = 0;                              // value = d.ddd * 10^exp10
  {
    const char* z = q + 1;
    const bool neg = *z == '-';
    if (*z == '-' || *z == '+') ++z;
    for (; z < r.ptr; ++z) exp10 = exp10 * 10 + (*z - '0');
    if (neg) exp10 = -exp10;
  }
  while (nd > 1 && digits[nd - 1] == '0') --nd;
  if (x < 0) *out++ = '-';
  if (a >= 1e-3 && a < 1e7) {
    const int point = exp10 + 1;              // digits before the decimal point
    if (point <= 0) {
      *out++ = '0'; *out++ = '.';
      for (int k = 0; k < -point; ++k) *out++ = '0';
      for (int k = 0; k < nd; ++k) *out++ = digits[k];
    } else if (point >= nd) {
      for (int k = 0; k < nd; ++k) *out++ = digits[k];
      for (int k = nd; k < point; ++k) *out++ = '0';
      *out++ = '.'; *out++ = '0';
    } else {
      for (int k = 0; k < point; ++k) *out++ = digits[k];
      *out++ = '.';
      for (int k = point; k < nd; ++k) *out++ = digits[k];
    }
  } else {
    *out++ = digits[0]; *out++ = '.';
    if (nd > 1) for (int k = 1; k < nd; ++k) *out++ = digits[k];
    else *out++ = '0';
    *out++ = 'E';
    out = std::to_chars(out, out + 8, exp10).ptr;
  }
  return out;
}

inline std::string java_double(double x) {
  char b[40];
  return std::string(b, java_double_to(b, x));
}
