// fixed_point.hpp -- host stand-in for the Vitis type the reference hosts use for values and
// queries: real_type_inout = ap_ufixed<32, 1, AP_TRN_ZERO> (src/fpga/src/ip/fpga_types.hpp:16-23).
// Unsigned, 1 integer bit, 31 fractional bits, truncation toward zero, wrap on overflow.
// Only what the host surface needs: construction from double/float, to_float(), ordering,
// the product/sum of the software reference, and the narrowing used by the packet builder.
#pragma once

#include <cmath>
#include <cstdint>
#include <iostream>

struct ufixed32 {
    uint32_t raw = 0;

    ufixed32() = default;
    ufixed32(double v) : raw(from_double(v)) {}
    ufixed32(float v) : raw(from_double((double)v)) {}
    ufixed32(int v) : raw(from_double((double)v)) {}
    static ufixed32 from_raw(uint32_t r) { ufixed32 f; f.raw = r; return f; }

    // (T) value  -- utils.hpp:401, :242
    static uint32_t from_double(double v) {
        if (!(v > 0.0)) return 0u;
        double s = std::floor(v * 2147483648.0);
        s = std::fmod(s, 4294967296.0);
        return (uint32_t)s;
    }
    // ap_fixed_base::to_float(): nearest-even to 24 significant bits
    float to_float() const { return (float)((double)raw / 2147483648.0); }
    double to_double() const { return (double)raw / 2147483648.0; }
    explicit operator float() const { return to_float(); }

    // (real_type) x.to_float() with real_type = ap_ufixed<W,1,AP_TRN_ZERO>  -- fpga_utils.hpp:336-338
    uint32_t narrow_via_float(int W) const {
        double s = std::floor((double)to_float() * (double)(1ull << (W - 1)));
        uint64_t m = (W == 32) ? 0xFFFFFFFFull : ((1ull << W) - 1ull);
        return (uint32_t)(((uint64_t)s) & m);
    }

    // arithmetic of gold_algorithms.hpp:188-246 when V = real_type_inout
    friend ufixed32 operator*(ufixed32 a, ufixed32 b) {
        return from_raw((uint32_t)(((uint64_t)a.raw * (uint64_t)b.raw) >> 31));
    }
    ufixed32 &operator+=(ufixed32 o) { raw += o.raw; return *this; }
    friend bool operator<(ufixed32 a, ufixed32 b) { return a.raw < b.raw; }
    friend bool operator>(ufixed32 a, ufixed32 b) { return a.raw > b.raw; }
    friend bool operator>=(ufixed32 a, ufixed32 b) { return a.raw >= b.raw; }
    friend bool operator==(ufixed32 a, ufixed32 b) { return a.raw == b.raw; }
    friend bool operator!=(ufixed32 a, ufixed32 b) { return a.raw != b.raw; }
    friend ufixed32 operator-(ufixed32 a, ufixed32 b) { return from_raw(a.raw - b.raw); }
    friend std::ostream &operator<<(std::ostream &os, ufixed32 f) { return os << f.to_double(); }
};
