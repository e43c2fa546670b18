// Java's Double.toString for the output files of both tools (host side only).
#pragma once
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <string>

// Double.toString: shortest digit string that round-trips (JDK 19+; older JDKs print a longer string in rare cases),
// decimal notation for 1e-3 <= |x| < 1e7, computerised scientific notation otherwise
inline std::string java_double(double x) {
  if (x != x) return "NaN";
  if (x == 1.0 / 0.0) return "Infinity";
  if (x == -1.0 / 0.0) return "-Infinity";
  if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf, x < 0 ? -x : x, std::chars_format::scientific);
  std::string s(buf, r.ptr);                  // d.ddddde[+-]xx (shortest)
  const size_t e = s.find('e');
  std::string digits;
  for (size_t k = 0; k < e; ++k)
    if (s[k] != '.') digits.push_back(s[k]);
  const int exp10 = atoi(s.c_str() + e + 1);  // value = d.ddd * 10^exp10
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const double a = x < 0 ? -x : x;
  std::string out = x < 0 ? "-" : "";
  if (a >= 1e-3 && a < 1e7) {
    const int point = exp10 + 1;              // digits before the decimal point
    if (point <= 0) out += "0." + std::string((size_t)(-point), '0') + digits;
    else if ((size_t)point >= digits.size()) out += digits + std::string((size_t)point - digits.size(), '0') + ".0";
    else out += digits.substr(0, (size_t)point) + "." + digits.substr((size_t)point);
  } else {
    out += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : std::string("0")) + "E" + std::to_string(exp10);
  }
  return out;
}

