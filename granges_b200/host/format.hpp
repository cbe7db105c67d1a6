// format.hpp -- text rendering of results exactly as the reference prints them.
//
// DatumType::Float64(v) is serialised with Rust's `f64::to_string()` (src/data/mod.rs:49-65): the
// shortest decimal string that round-trips, in positional notation, never an exponent ("5" for
// 5.0, "0.30000000000000004", "0.0000001", "NaN", "inf", "-inf", "-0").  DatumType::NoValue
// prints BED_TSV.no_value_string = "." (src/io/tsv.rs:5-12).  C++17 std::to_chars(fixed) is the
// same shortest-round-trip algorithm; only the spelling of NaN differs.
#pragma once
#include <charconv>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace grh {

// Writes v at p (at least 400 bytes free: the longest double in fixed notation has 309 integer digits
// or 324 decimals + the digits) and returns one past the last character written.
inline char *format_f64(char *p, double v) {
    if (std::isnan(v)) {
        memcpy(p, "NaN", 3);
        return p + 3;
    }
    if (std::isinf(v)) {
        if (v < 0) *p++ = '-';
        memcpy(p, "inf", 3);
        return p + 3;
    }
    // shortest round-trip digits and decimal exponent (d.ddd e+-XX), then positional expansion with zero
    // padding -- std::to_chars(fixed) would print the exact integer value of large doubles instead
    char tmp[40];
    auto r = std::to_chars(tmp, tmp + sizeof(tmp), v, std::chars_format::scientific);
    const char *t = tmp;
    if (*t == '-') {
        *p++ = '-';
        t++;
    }
    char digits[24];
    int nd = 0;
    const char *e = t;
    for (; e < r.ptr && *e != 'e'; e++)
        if (*e != '.') digits[nd++] = *e;
    int exp10 = 0;
    if (e < r.ptr) {
        bool neg = e[1] == '-';
        for (const char *q = e + 2; q < r.ptr; q++) exp10 = exp10 * 10 + (*q - '0');
        if (neg) exp10 = -exp10;
    }
    while (nd > 1 && digits[nd - 1] == '0') nd--;  // "5.0e0" style trailing zeros
    if (nd == 1 && digits[0] == '0') {
        *p++ = '0';
        return p;
    }
    const int point = exp10 + 1;  // number of digits before the decimal point
    if (point <= 0) {
        *p++ = '0';
        *p++ = '.';
        for (int i = 0; i < -point; i++) *p++ = '0';
        for (int i = 0; i < nd; i++) *p++ = digits[i];
    } else if (point >= nd) {
        for (int i = 0; i < nd; i++) *p++ = digits[i];
        for (int i = nd; i < point; i++) *p++ = '0';
    } else {
        for (int i = 0; i < point; i++) *p++ = digits[i];
        *p++ = '.';
        for (int i = point; i < nd; i++) *p++ = digits[i];
    }
    return p;
}

inline char *format_u32(char *p, uint32_t v) { return std::to_chars(p, p + 12, v).ptr; }

}  // namespace grh
