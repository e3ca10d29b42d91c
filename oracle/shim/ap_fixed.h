// ap_fixed.h -- minimal stand-in for the Xilinx fixed-point header (ap_ufixed<W,I,Q,O> only).
//
// TEST INFRASTRUCTURE ONLY (see ap_int.h).  Implements the documented semantics of the one
// configuration the reference uses, ap_ufixed<W, 1, AP_TRN_ZERO> with the default AP_WRAP overflow
// (src/fpga/src/ip/fpga_types.hpp:20-22): an unsigned W-bit word with W-I fraction bits;
//   * a * b is exact (ap_ufixed<W1+W2, I1+I2>), a + b is exact (one more integer bit);
//   * assignment / cast to a narrower type drops low fraction bits (truncation: AP_TRN and
//     AP_TRN_ZERO coincide for unsigned values) and high integer bits (wrap);
//   * construction from double/float truncates toward zero; to_float() rounds to nearest-even.
// Widths up to 128 bits (storage: 8/16/32/64-bit word, or unsigned __int128 above 64), which covers every
// type the reference instantiates (32 x 32 -> 64 in the kernel; 65 bits in spmv_coo_gold4's reduction).
#pragma once

#include <cmath>
#include <cstdint>
#include <ostream>
#include <type_traits>

#include "ap_int.h"

enum ap_q_mode { AP_RND, AP_RND_ZERO, AP_RND_MIN_INF, AP_RND_INF, AP_RND_CONV, AP_TRN, AP_TRN_ZERO };
enum ap_o_mode { AP_SAT, AP_SAT_ZERO, AP_SAT_SYM, AP_WRAP, AP_WRAP_SM };

template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP, int NB = 0>
struct ap_ufixed {
    static_assert(W >= 1 && W <= 128, "shim: widths up to 128 bits");
    static_assert(O == AP_WRAP && (Q == AP_TRN || Q == AP_TRN_ZERO), "shim: truncate + wrap only");
    typedef typename std::conditional<(W <= 64), typename apshim::raw_of<(W <= 64 ? W : 64)>::type, unsigned __int128>::type raw_t;
    enum { width = W, iwidth = I, F = W - I };
    raw_t V;

    ap_ufixed() : V(0) {}
    template <typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
    ap_ufixed(T v) : V((raw_t)(shl((unsigned __int128)(unsigned long long)v, F) & mask())) {}
    template <typename T, typename std::enable_if<std::is_floating_point<T>::value, long>::type = 0>
    ap_ufixed(T d) {
        const double s = std::floor(std::ldexp((double)d, F));   // exact scaling, then truncation
        V = (raw_t)((unsigned __int128)(s < 0 ? (unsigned long long)(long long)s : (unsigned long long)s) & mask());
    }
    template <int W2, int I2, ap_q_mode Q2, ap_o_mode O2, int N2>
    ap_ufixed(const ap_ufixed<W2, I2, Q2, O2, N2> &o) {
        const int F2 = W2 - I2;
        const unsigned __int128 v = (unsigned __int128)o.V;
        V = (raw_t)((F2 >= F ? (v >> (F2 - F)) : shl(v, F - F2)) & mask());
    }

    static unsigned __int128 mask() { return W >= 128 ? ~(unsigned __int128)0 : ((((unsigned __int128)1) << (W & 127)) - 1); }
    static unsigned __int128 shl(unsigned __int128 v, int s) { return s >= 128 ? 0 : (s <= 0 ? v : v << s); }

    double to_double() const { return (double)std::ldexp((long double)V, -F); }
    float to_float() const { return (float)std::ldexp((long double)V, -F); }   // round to nearest-even
    operator double() const { return to_double(); }   // implicit, like the vendor type (std::pair<I, V> -> pair<I, double>)
    explicit operator float() const { return to_float(); }
    explicit operator bool() const { return V != 0; }

    template <int W2, int I2, ap_q_mode Q2, ap_o_mode O2, int N2>
    ap_ufixed &operator+=(const ap_ufixed<W2, I2, Q2, O2, N2> &o) { *this = *this + o; return *this; }
    template <int W2, int I2, ap_q_mode Q2, ap_o_mode O2, int N2>
    ap_ufixed &operator*=(const ap_ufixed<W2, I2, Q2, O2, N2> &o) { *this = *this * o; return *this; }
};

namespace apshim {
template <int A, int B> struct max_of { enum { v = A > B ? A : B }; };
template <int W1, int I1, ap_q_mode Q1, ap_o_mode O1, int N1, int W2, int I2, ap_q_mode Q2, ap_o_mode O2, int N2>
inline int fx_cmp(const ap_ufixed<W1, I1, Q1, O1, N1> &a, const ap_ufixed<W2, I2, Q2, O2, N2> &b) {
    const int F1 = W1 - I1, F2 = W2 - I2, FM = F1 > F2 ? F1 : F2;
    const unsigned __int128 x = (unsigned __int128)a.V << (FM - F1), y = (unsigned __int128)b.V << (FM - F2);
    return x < y ? -1 : (x > y ? 1 : 0);
}
}  // namespace apshim

template <int W1, int I1, ap_q_mode Q1, ap_o_mode O1, int N1, int W2, int I2, ap_q_mode Q2, ap_o_mode O2, int N2>
inline ap_ufixed<W1 + W2, I1 + I2, AP_TRN, AP_WRAP, 0> operator*(const ap_ufixed<W1, I1, Q1, O1, N1> &a,
                                                                 const ap_ufixed<W2, I2, Q2, O2, N2> &b) {
    static_assert(W1 + W2 <= 128, "shim: product wider than 128 bits");
    ap_ufixed<W1 + W2, I1 + I2, AP_TRN, AP_WRAP, 0> r;
    r.V = (typename ap_ufixed<W1 + W2, I1 + I2, AP_TRN, AP_WRAP, 0>::raw_t)((unsigned __int128)a.V * (unsigned __int128)b.V);
    return r;
}

template <int W1, int I1, ap_q_mode Q1, ap_o_mode O1, int N1, int W2, int I2, ap_q_mode Q2, ap_o_mode O2, int N2>
inline ap_ufixed<apshim::max_of<I1, I2>::v + 1 + apshim::max_of<W1 - I1, W2 - I2>::v, apshim::max_of<I1, I2>::v + 1, AP_TRN, AP_WRAP, 0>
operator+(const ap_ufixed<W1, I1, Q1, O1, N1> &a, const ap_ufixed<W2, I2, Q2, O2, N2> &b) {
    enum { IM = apshim::max_of<I1, I2>::v + 1, FM = apshim::max_of<W1 - I1, W2 - I2>::v };
    static_assert(IM + FM <= 128, "shim: sum wider than 128 bits");
    ap_ufixed<IM + FM, IM, AP_TRN, AP_WRAP, 0> r;
    r.V = (typename ap_ufixed<IM + FM, IM, AP_TRN, AP_WRAP, 0>::raw_t)(((unsigned __int128)a.V << (FM - (W1 - I1))) +
                                                                        ((unsigned __int128)b.V << (FM - (W2 - I2))));
    return r;
}

// a - b, wrapping (the reference only subtracts the smaller from the larger, utils.hpp:208)
template <int W1, int I1, ap_q_mode Q1, ap_o_mode O1, int N1, int W2, int I2, ap_q_mode Q2, ap_o_mode O2, int N2>
inline ap_ufixed<apshim::max_of<I1, I2>::v + 1 + apshim::max_of<W1 - I1, W2 - I2>::v, apshim::max_of<I1, I2>::v + 1, AP_TRN, AP_WRAP, 0>
operator-(const ap_ufixed<W1, I1, Q1, O1, N1> &a, const ap_ufixed<W2, I2, Q2, O2, N2> &b) {
    enum { IM = apshim::max_of<I1, I2>::v + 1, FM = apshim::max_of<W1 - I1, W2 - I2>::v };
    static_assert(IM + FM <= 128, "shim: difference wider than 128 bits");
    ap_ufixed<IM + FM, IM, AP_TRN, AP_WRAP, 0> r;
    r.V = (typename ap_ufixed<IM + FM, IM, AP_TRN, AP_WRAP, 0>::raw_t)(
        ((((unsigned __int128)a.V << (FM - (W1 - I1))) - ((unsigned __int128)b.V << (FM - (W2 - I2))))) & ap_ufixed<IM + FM, IM, AP_TRN, AP_WRAP, 0>::mask());
    return r;
}

// fixed * built-in integer: the integer is an exact ap_ufixed<32,32>
template <int W1, int I1, ap_q_mode Q1, ap_o_mode O1, int N1, typename T, typename std::enable_if<std::is_integral<T>::value, int>::type = 0>
inline ap_ufixed<W1 + 32, I1 + 32, AP_TRN, AP_WRAP, 0> operator*(const ap_ufixed<W1, I1, Q1, O1, N1> &a, T b) {
    return a * ap_ufixed<32, 32, AP_TRN, AP_WRAP, 0>((unsigned)b);
}

#define APSHIM_CMP(op)                                                                                                   \
    template <int W1, int I1, ap_q_mode Q1, ap_o_mode O1, int N1, int W2, int I2, ap_q_mode Q2, ap_o_mode O2, int N2>    \
    inline bool operator op(const ap_ufixed<W1, I1, Q1, O1, N1> &a, const ap_ufixed<W2, I2, Q2, O2, N2> &b) {            \
        return apshim::fx_cmp(a, b) op 0;                                                                                \
    }                                                                                                                    \
    template <int W1, int I1, ap_q_mode Q1, ap_o_mode O1, int N1, typename T,                                            \
              typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>                                      \
    inline bool operator op(const ap_ufixed<W1, I1, Q1, O1, N1> &a, T b) {                                               \
        return a.to_double() op (double)b;                                                                               \
    }                                                                                                                    \
    template <int W1, int I1, ap_q_mode Q1, ap_o_mode O1, int N1, typename T,                                            \
              typename std::enable_if<std::is_arithmetic<T>::value, int>::type = 0>                                      \
    inline bool operator op(T a, const ap_ufixed<W1, I1, Q1, O1, N1> &b) {                                               \
        return (double)a op b.to_double();                                                                               \
    }
APSHIM_CMP(<)
APSHIM_CMP(<=)
APSHIM_CMP(>)
APSHIM_CMP(>=)
APSHIM_CMP(==)
APSHIM_CMP(!=)
#undef APSHIM_CMP

template <int W, int I, ap_q_mode Q, ap_o_mode O, int N>
inline std::ostream &operator<<(std::ostream &os, const ap_ufixed<W, I, Q, O, N> &v) { return os << v.to_double(); }
