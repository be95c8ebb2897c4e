// prologue.h — lets the reference's ISPC kernel source (backends/embree/render_embree.ispc and the
// .ih headers it includes) compile as SCALAR C++: one program instance, "varying" = one lane.
// The Makefile pipes  prologue.h + (render_embree.ispc with its two `foreach` headers rewritten as
// `for` by sed) + epilogue.h  into g++ -x c++ -fsingle-precision-constant (ISPC reads an unsuffixed
// floating literal as float) -ffp-contract=off; nothing of the reference is written to disk.
// ISPC's standard library functions used by those files are supplied below on top of libm, so
// transcendental results are libm's, not the ISPC built-in math library's (last-ulp differences).
// TEST INFRASTRUCTURE (oracle/_ref/libcrt_embree.so).
#include <cmath>
#include <cstdint>
#include <cstring>

namespace ispc {

// ---- ISPC type qualifiers and statements that have no meaning for a single program instance ----
#define uniform
#define varying
#define unmasked
#define cif if
#define cwhile while
#define cfor for
#define export
// ISPC spells sized integers int8/int16/int64; util.ih builds <cstdint> names from them
#define int8 char
#define int16 short
#define int64 long
typedef unsigned int uint32;
typedef unsigned int uint;

// Path depth: the reference fixes MAX_PATH_DEPTH = 5 at compile time (util.ih:10). The Makefile's sed
// points the one use in render_embree.ispc:336 at this variable instead (default: util.ih's value, set in
// epilogue.h), so that BASELINE config 2 ("max depth 8") can be run through the reference's kernel too.
extern int crt_ref_max_path_depth;

// ---- the ISPC standard library subset the reference uses ----
inline float sqrt(float x) { return ::sqrtf(x); }
inline float rsqrt(float x) { return 1.f / ::sqrtf(x); }
inline float rcp(float x) { return 1.f / x; }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float tan(float x) { return ::tanf(x); }
inline float asin(float x) { return ::asinf(x); }
inline float acos(float x) { return ::acosf(x); }
inline float atan(float x) { return ::atanf(x); }
inline float atan2(float y, float x) { return ::atan2f(y, x); }
inline float exp(float x) { return ::expf(x); }
inline float log(float x) { return ::logf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float floor(float x) { return ::floorf(x); }
inline float ceil(float x) { return ::ceilf(x); }
inline float round(float x) { return ::nearbyintf(x); }
inline float ldexp(float x, int e) { return ::ldexpf(x, e); }
inline float abs(float x) { return ::fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float min(float a, float b) { return a < b ? a : b; }
inline float max(float a, float b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline unsigned int intbits(float x)
{
    unsigned int u;
    std::memcpy(&u, &x, 4);
    return u;
}
inline float floatbits(unsigned int u)
{
    float x;
    std::memcpy(&x, &u, 4);
    return x;
}
inline float select(bool c, float a, float b) { return c ? a : b; }
inline bool isnan(float x) { return x != x; }
// ISPC stdlib float_to_srgb8 is a table-driven conversion (the ISPC distribution is absent here); it is
// supplied as round-to-nearest of the exact sRGB curve, like the oracle: the 8-bit image is compared
// within +-1 LSB, parity proper is judged on the float framebuffer.
inline unsigned char float_to_srgb8(float x)
{
    if (!(x > 0.f)) {
        return 0;
    }
    if (x >= 1.f) {
        return 255;
    }
    const float s = x <= 0.0031308f ? 12.92f * x : 1.055f * ::powf(x, 1.f / 2.4f) - 0.055f;
    return (unsigned char)(s * 255.f + 0.5f);
}

// ---- argument evaluation order ----
// ISPC evaluates function arguments left to right; C++ leaves the order unspecified and g++ goes right
// to left. The only calls in these sources whose arguments have side effects are the two
// make_float2(lcg_randomf(rng), lcg_randomf(rng)) (disney_bsdf.ih:377, render_embree.ispc:134). float3.ih
// is pulled in here (it is #pragma once), then make_float2 calls are routed through a braced-init-list,
// whose elements C++ evaluates strictly in order.
#include "float3.ih"
struct OrderedFloat2Args {
    float a, b;
    OrderedFloat2Args(float x) : a(x), b(x) {}
    OrderedFloat2Args(float x, float y) : a(x), b(y) {}
};
inline float2 make_float2_ordered(const OrderedFloat2Args &v) { return make_float2(v.a, v.b); }
#define make_float2(...) make_float2_ordered(OrderedFloat2Args{__VA_ARGS__})
