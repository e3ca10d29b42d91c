// ap_int.h -- minimal stand-in for the Xilinx arbitrary-precision integer header (ap_uint<N> only).
//
// TEST INFRASTRUCTURE ONLY.  The Vitis HLS headers are absent from this image, so the
// reference's HLS kernel and FPGA host cannot be compiled as shipped.  This shim implements just the
// subset of the documented ap_uint<N> behaviour those sources use -- N-bit unsigned wrap-around
// arithmetic, .range(hi, lo) and .bit(i) proxies, implicit conversion to a built-in integer -- so that
// oracle/ref_fpga.cpp can compile the reference sources where they lie (oracle/Makefile, target `ref`).
// Storage matches the vendor layout where the reference relies on it (pointer casts between unsigned
// int, ap_uint<N> and ap_ufixed<W,I>): the raw value is the first and only member, held in the smallest
// of 8/16/32/64 bits that fits N, wider types as little-endian 64-bit words.
#pragma once

#include <cstddef>
#include <cstdint>
#include <type_traits>

namespace apshim {

template <int N>
struct raw_of {
    typedef typename std::conditional<
        (N <= 8), uint8_t,
        typename std::conditional<(N <= 16), uint16_t,
                                  typename std::conditional<(N <= 32), uint32_t, uint64_t>::type>::type>::type type;
};

inline uint64_t low_mask(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

template <typename Owner>
struct range_ref {
    Owner *o;
    int hi, lo;
    operator unsigned long long() const { return o->get_bits(lo, hi - lo + 1); }
    range_ref &operator=(unsigned long long v) {
        o->set_bits(lo, hi - lo + 1, v);
        return *this;
    }
    range_ref &operator=(const range_ref &r) { return *this = (unsigned long long)r; }
};

template <typename Owner>
struct bit_ref {
    Owner *o;
    int i;
    operator bool() const { return o->get_bits(i, 1) != 0; }
    bit_ref &operator=(unsigned long long v) {
        o->set_bits(i, 1, v != 0);
        return *this;
    }
    bit_ref &operator=(const bit_ref &r) { return *this = (unsigned long long)(bool)r; }
};

}  // namespace apshim

template <int N, bool WIDE = (N > 64)>
struct ap_uint;

// ---- N <= 64: one machine word ------------------------------------------------------------------
template <int N>
struct ap_uint<N, false> {
    typedef typename apshim::raw_of<N>::type raw_t;
    raw_t V;

    ap_uint() : V(0) {}
    template <typename T, typename = typename std::enable_if<std::is_integral<T>::value>::type>
    ap_uint(T v) : V((raw_t)((unsigned long long)v & apshim::low_mask(N))) {}

    operator unsigned long long() const { return V; }

    ap_uint &operator++() { V = (raw_t)((V + 1ull) & apshim::low_mask(N)); return *this; }
    ap_uint operator++(int) { ap_uint t = *this; ++*this; return t; }
    ap_uint &operator--() { V = (raw_t)((V - 1ull) & apshim::low_mask(N)); return *this; }
    ap_uint operator--(int) { ap_uint t = *this; --*this; return t; }
    template <typename T> ap_uint &operator+=(T v) { V = (raw_t)((V + (unsigned long long)v) & apshim::low_mask(N)); return *this; }
    template <typename T> ap_uint &operator-=(T v) { V = (raw_t)((V - (unsigned long long)v) & apshim::low_mask(N)); return *this; }

    unsigned long long get_bits(int lo, int len) const { return ((unsigned long long)V >> lo) & apshim::low_mask(len); }
    void set_bits(int lo, int len, unsigned long long v) {
        const unsigned long long m = apshim::low_mask(len) << lo;
        V = (raw_t)((((unsigned long long)V & ~m) | ((v << lo) & m)) & apshim::low_mask(N));
    }
    apshim::range_ref<ap_uint> range(int hi, int lo) { return apshim::range_ref<ap_uint>{this, hi, lo}; }
    unsigned long long range(int hi, int lo) const { return get_bits(lo, hi - lo + 1); }
    apshim::bit_ref<ap_uint> bit(int i) { return apshim::bit_ref<ap_uint>{this, i}; }
    bool bit(int i) const { return get_bits(i, 1) != 0; }
};

// ---- N > 64: little-endian 64-bit words (bit 0 = LSB of word 0 = LSB of byte 0) --------------------
template <int N>
struct ap_uint<N, true> {
    enum { WORDS = (N + 63) / 64 };
    uint64_t w[WORDS];

    ap_uint() { for (int i = 0; i < WORDS; i++) w[i] = 0; }
    template <typename T, typename = typename std::enable_if<std::is_integral<T>::value>::type>
    ap_uint(T v) { for (int i = 0; i < WORDS; i++) w[i] = 0; w[0] = (uint64_t)v; }

    unsigned long long get_bits(int lo, int len) const {   // len <= 64
        const int q = lo >> 6, s = lo & 63;
        unsigned long long v = w[q] >> s;
        if (s != 0 && q + 1 < WORDS) v |= w[q + 1] << (64 - s);
        return v & apshim::low_mask(len);
    }
    void set_bits(int lo, int len, unsigned long long v) {   // len <= 64
        for (int b = 0; b < len; b++) {
            const int p = lo + b;
            if (p >= N) break;
            const uint64_t m = 1ull << (p & 63);
            if ((v >> b) & 1ull) w[p >> 6] |= m; else w[p >> 6] &= ~m;
        }
    }
    apshim::range_ref<ap_uint> range(int hi, int lo) { return apshim::range_ref<ap_uint>{this, hi, lo}; }
    unsigned long long range(int hi, int lo) const { return get_bits(lo, hi - lo + 1); }
    apshim::bit_ref<ap_uint> bit(int i) { return apshim::bit_ref<ap_uint>{this, i}; }
    bool bit(int i) const { return get_bits(i, 1) != 0; }
};
