// TEST INFRASTRUCTURE: the Rust std semantics that tools/transliterate_gen.py's output relies on, restated for C++17.
// Every helper cites the Rust definition it mirrors (library/core, library/std of Rust 1.85, the reference's MSRV).  Floating-point
// helpers are single IEEE-754 operations or glibc libm calls -- what Rust's std compiles to on x86-64 Linux.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <tuple>
#include <type_traits>

namespace rs {

// `[None; N]` / defaulted aggregates
struct Default {
    template <class T>
    operator T() const { return T{}; }
};
struct NoneT {};
static constexpr NoneT none{};
template <class T>
struct Some_ { T v; };
template <class T>
Some_<T> some(T v) { return {v}; }
template <class T>
struct Option {  // core::option::Option<T> for Copy payloads
    bool is_some = false;
    T value{};
    Option() {}
    Option(NoneT) {}
    Option(Some_<T> s) : is_some(true), value(s.v) {}
};

// array expressions `[a, b, c]` and `[x; N]`
template <class... T>
auto arr(T... t) -> std::array<std::common_type_t<T...>, sizeof...(T)> { return {{t...}}; }
template <class E, class... T>
std::array<E, sizeof...(T)> arr_of(T... t) { return {{(E)t...}}; }
template <size_t N, class T>
std::array<T, N> fill(T v) {
    std::array<T, N> a;
    a.fill(v);
    return a;
}

// `expr as T` (Rust reference "Numeric cast"): float -> int saturates and maps NaN to 0; everything else is a plain conversion
template <class To, class From>
To as_(From x) {
    if constexpr (std::is_floating_point_v<From> && std::is_integral_v<To>) {
        if (x != x) return (To)0;
        if (x <= (From)std::numeric_limits<To>::min()) return std::numeric_limits<To>::min();
        if (x >= (From)std::numeric_limits<To>::max()) return std::numeric_limits<To>::max();
        return (To)x;
    } else {
        return static_cast<To>(x);
    }
}

// ---- f64 methods (library/core/src/num/f64.rs, library/std/src/f64.rs) ----
inline double m_abs(double x) { return std::fabs(x); }
inline double m_max(double a, double b) { return std::fmax(a, b); }  // IEEE maxNum
inline double m_min(double a, double b) { return std::fmin(a, b); }
inline double m_clamp(double x, double lo, double hi) {  // f64::clamp: NaN stays NaN
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}
inline bool m_is_finite(double x) { return std::isfinite(x); }
inline bool m_is_nan(double x) { return x != x; }
inline bool m_is_sign_negative(double x) { return std::signbit(x); }
inline double m_mul_add(double x, double a, double b) { return std::fma(x, a, b); }
inline uint64_t m_to_bits(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
inline double m_sqrt(double x) { return std::sqrt(x); }
inline double m_exp(double x) { return std::exp(x); }
inline double m_ln(double x) { return std::log(x); }
inline double m_log10(double x) { return std::log10(x); }
inline double m_powf(double x, double y) { return std::pow(x, y); }
inline double m_powi(double a, int b) {  // llvm.powi -> compiler-rt __powidf2
    const bool recip = b < 0;
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}
inline double m_tanh(double x) { return std::tanh(x); }
inline double m_sin(double x) { return std::sin(x); }
inline double m_cos(double x) { return std::cos(x); }
inline double m_floor(double x) { return std::floor(x); }
inline double m_ceil(double x) { return std::ceil(x); }
inline double m_round(double x) { return std::round(x); }
inline double m_recip(double x) { return 1.0 / x; }
inline double m_signum(double x) { return x != x ? x : (std::signbit(x) ? -1.0 : 1.0); }
inline double m_copysign(double x, double s) { return std::copysign(x, s); }

// ---- integer methods ----
template <class T, class U, class = std::enable_if_t<std::is_integral_v<T>>>
T m_saturating_add(T a, U b) { T r; return __builtin_add_overflow(a, (T)b, &r) ? std::numeric_limits<T>::max() : r; }
template <class T, class U, class = std::enable_if_t<std::is_integral_v<T>>>
T m_wrapping_add(T a, U b) { return (T)((std::make_unsigned_t<T>)a + (std::make_unsigned_t<T>)b); }
template <class T, class U, class = std::enable_if_t<std::is_integral_v<T>>>
T m_wrapping_mul(T a, U b) { return (T)((std::make_unsigned_t<T>)a * (std::make_unsigned_t<T>)b); }
inline uint64_t m_rotate_left(uint64_t x, unsigned n) { return (x << (n & 63)) | (x >> ((64 - n) & 63)); }
template <class T, class = std::enable_if_t<std::is_integral_v<T>>>
T m_max(T a, T b) { return a > b ? a : b; }
template <class T, class = std::enable_if_t<std::is_integral_v<T>>>
T m_min(T a, T b) { return a < b ? a : b; }

// ---- slices ----
template <class T, size_t N>
void m_swap(std::array<T, N>& a, size_t i, size_t j) { std::swap(a[i], a[j]); }  // <[T]>::swap
template <class T, size_t N>
size_t m_len(const std::array<T, N>&) { return N; }
template <class T>
struct Iter { const T* p; size_t n; };
template <class T, size_t N>
Iter<T> iter(const std::array<T, N>& a) { return {a.data(), N}; }
template <class T>
Iter<T> m_take(Iter<T> it, size_t n) { return {it.p, n < it.n ? n : it.n}; }
template <class T, class F>
bool m_any(Iter<T> it, F f) { for (size_t i = 0; i < it.n; i++) if (f(it.p[i])) return true; return false; }
template <class T, class F>
bool m_all(Iter<T> it, F f) { for (size_t i = 0; i < it.n; i++) if (!f(it.p[i])) return false; return true; }

template <class T>
void mem_swap(T& a, T& b) { std::swap(a, b); }  // core::mem::swap

namespace f64_ {
static constexpr double EPSILON = std::numeric_limits<double>::epsilon();
static constexpr double MAX = std::numeric_limits<double>::max();
static constexpr double MIN = std::numeric_limits<double>::lowest();
static constexpr double MIN_POSITIVE = std::numeric_limits<double>::min();
static constexpr double INFINITY_ = std::numeric_limits<double>::infinity();
static constexpr double NAN_ = std::numeric_limits<double>::quiet_NaN();
inline double from_bits(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
}  // namespace f64_
namespace u64_ { static constexpr uint64_t MAX = std::numeric_limits<uint64_t>::max(); }
namespace u32_ { static constexpr uint32_t MAX = std::numeric_limits<uint32_t>::max(); }
namespace usize_ { static constexpr size_t MAX = std::numeric_limits<size_t>::max(); }

// std::f64::consts (library/core/src/num/f64.rs)
static constexpr double f64_consts_LOG2_E = 1.44269504088896340735992468100189214;
static constexpr double f64_consts_LN_2 = 0.693147180559945309417232121458176568;
static constexpr double f64_consts_LN_10 = 2.30258509299404568401799145468436421;
static constexpr double f64_consts_LOG10_E = 0.434294481903251827651128918916605082;
static constexpr double f64_consts_PI = 3.14159265358979323846264338327950288;
static constexpr double f64_consts_TAU = 6.28318530717958647692528676655900577;
static constexpr double f64_consts_E = 2.71828182845904523536028747135266250;
static constexpr double f64_consts_SQRT_2 = 1.41421356237309504880168872420969808;
static constexpr double f64_consts_FRAC_PI_2 = 1.57079632679489661923132169163975144;

}  // namespace rs
