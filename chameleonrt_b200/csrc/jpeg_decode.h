// jpeg_decode.h — JPEG textures for the native scene loader (scene_io.cpp), host code only.
//
// The reference decodes textures with stb_image (util/stb_image.h, pulled in by util/util.cpp:17): stbi_load / stbi_load_from_memory
// with four requested components. For a valid JPEG stream the entropy decoding is fixed by the standard (ITU T.81): every
// conforming decoder recovers the same quantised coefficients, so that part is written here from the standard (sequential
// and progressive Huffman modes, restart intervals, 8-bit samples, 8- and 16-bit quantisation tables). What is NOT fixed by
// the standard — and what decides the bytes of the texture — is restated from stb_image, with its integer arithmetic:
//   * the inverse DCT               stbi__idct_block            (stb_image.h:2271-2332; jidctint-style, 12-bit constants,
//                                                                 columns kept at 2 extra bits, +128 level shift folded in)
//   * chroma upsampling             stbi__resample_row_v_2 / _h_2 / _hv_2 / _generic, driven as load_jpeg_image drives them
//                                   (:3240-3300, :3421-3432, :3640-3700: "near" and "far" rows, 3:1 weights, nearest for
//                                   other factors)
//   * YCbCr -> RGB                  stbi__YCbCr_to_RGB_row      (:3435-3459: 20-bit fixed point, the Cb term of green
//                                                                 masked to 16 bits)
//   * which streams are RGB already load_jpeg_image's is_rgb    (:3655: component ids 'R','G','B', or an Adobe marker with
//                                                                 transform 0 and no JFIF marker)
// stb_image's SSE2 kernels produce the same bytes as these scalar ones (its own comments, :2334, :3433), so the build of the
// reference does not matter. Grey and three-component streams are read; four-component (CMYK / YCCK) streams, 12-bit
// samples and arithmetic coding are errors (the latter two are errors in stb_image as well).
#pragma once

#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic ignored "-Wpsabi"  // (256-bit vector values through inline helpers when the build has no -mavx: no ABI crosses a library boundary)
#endif

#ifdef __AVX2__
#include <immintrin.h>
#endif

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace crt_jpeg {

// ---- stb_image's numeric kernels ----
constexpr int f2f(float x)
{
    return (int)(x * 4096 + 0.5);
}
inline uint8_t clamp_u8(int x)
{
    return (unsigned)x > 255u ? (x < 0 ? 0 : 255) : (uint8_t)x;
}

// Wrap-around 32-bit arithmetic (what stb_image's int arithmetic does in practice): a corrupt stream can carry coefficients
// large enough to overflow, which must give garbage pixels, not undefined behaviour.
inline int wadd(int a, int b)
{
    return (int)((unsigned)a + (unsigned)b);
}
inline int wsub(int a, int b)
{
    return (int)((unsigned)a - (unsigned)b);
}
inline int wmul(int a, int b)
{
    return (int)((unsigned)a * (unsigned)b);
}
#if defined(__GNUC__) && !defined(CRT_JPEG_NO_VECTOR)
typedef int32_t v8i __attribute__((vector_size(32)));
typedef uint32_t v8u __attribute__((vector_size(32)));
typedef int16_t v8s __attribute__((vector_size(16)));
typedef uint8_t v8b __attribute__((vector_size(8)));
inline v8i wadd(v8i a, v8i b)
{
    return (v8i)((v8u)a + (v8u)b);
}
inline v8i wadd(v8i a, int b)
{
    return (v8i)((v8u)a + (unsigned)b);
}
inline v8i wsub(v8i a, v8i b)
{
    return (v8i)((v8u)a - (v8u)b);
}
inline v8i wmul(v8i a, int b)
{
    return (v8i)((v8u)a * (unsigned)b);
}
#endif

// one pass of the separable transform over eight values (T: int, or a vector of ints — one transform per lane); the results
// come back still scaled by 4096
template <typename T>
struct Idct1d {
    T x0, x1, x2, x3, t0, t1, t2, t3;
    Idct1d(T s0, T s1, T s2, T s3, T s4, T s5, T s6, T s7)
    {
        T p1, p2, p3, p4, p5;
        p2 = s2;
        p3 = s6;
        p1 = wmul(wadd(p2, p3), f2f(0.5411961f));
        t2 = wadd(p1, wmul(p3, f2f(-1.847759065f)));
        t3 = wadd(p1, wmul(p2, f2f(0.765366865f)));
        p2 = s0;
        p3 = s4;
        t0 = wmul(wadd(p2, p3), 4096);
        t1 = wmul(wsub(p2, p3), 4096);
        x0 = wadd(t0, t3);
        x3 = wsub(t0, t3);
        x1 = wadd(t1, t2);
        x2 = wsub(t1, t2);
        t0 = s7;
        t1 = s5;
        t2 = s3;
        t3 = s1;
        p3 = wadd(t0, t2);
        p4 = wadd(t1, t3);
        p1 = wadd(t0, t3);
        p2 = wadd(t1, t2);
        p5 = wmul(wadd(p3, p4), f2f(1.175875602f));
        t0 = wmul(t0, f2f(0.298631336f));
        t1 = wmul(t1, f2f(2.053119869f));
        t2 = wmul(t2, f2f(3.072711026f));
        t3 = wmul(t3, f2f(1.501321110f));
        p1 = wadd(p5, wmul(p1, f2f(-0.899976223f)));
        p2 = wadd(p5, wmul(p2, f2f(-2.562915447f)));
        p3 = wmul(p3, f2f(-1.961570560f));
        p4 = wmul(p4, f2f(-0.390180644f));
        t3 = wadd(t3, wadd(p1, p4));
        t2 = wadd(t2, wadd(p2, p3));
        t1 = wadd(t1, wadd(p2, p4));
        t0 = wadd(t0, wadd(p1, p3));
    }
    // the eight outputs: (x_k +- t_k + rounding) >> shift, in the order 0, 7, 1, 6, 2, 5, 3, 4
    void finish(T out[8], int rounding, int shift)
    {
        x0 = wadd(x0, rounding), x1 = wadd(x1, rounding), x2 = wadd(x2, rounding), x3 = wadd(x3, rounding);
        out[0] = wadd(x0, t3) >> shift;
        out[7] = wsub(x0, t3) >> shift;
        out[1] = wadd(x1, t2) >> shift;
        out[6] = wsub(x1, t2) >> shift;
        out[2] = wadd(x2, t1) >> shift;
        out[5] = wsub(x2, t1) >> shift;
        out[3] = wadd(x3, t0) >> shift;
        out[4] = wsub(x3, t0) >> shift;
    }
};

inline void idct_block_scalar(uint8_t *out, size_t out_stride, const short d[64])
{
    int val[64];
    for (int i = 0; i < 8; ++i) {  // columns; the constants scaled things by 1 << 12: back down, keeping two extra bits
        const short *c = d + i;
        int v[8];
        if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
            v[0] = v[1] = v[2] = v[3] = v[4] = v[5] = v[6] = v[7] = c[0] * 4;
        } else {
            Idct1d<int>(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]).finish(v, 512, 10);
        }
        for (int r = 0; r < 8; ++r) {
            val[8 * r + i] = v[r];
        }
    }
    for (int i = 0; i < 8; ++i) {  // rows: 17 bits to remove, rounded, with the +128 level shift added before the shift
        const int *v = val + 8 * i;
        int o[8];
        Idct1d<int>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]).finish(o, 65536 + (128 << 17), 17);
        for (int c = 0; c < 8; ++c) {
            out[out_stride * i + c] = clamp_u8(o[c]);
        }
    }
}

// A block whose only non-zero coefficient is the DC term: both passes reduce to one value for all 64 samples (column 0
// becomes d0 * 4 in every row, each row then has x0 = x1 = x2 = x3 = s0 * 4096 + bias and t0..t3 = 0).
inline void idct_dc_only(uint8_t *out, size_t out_stride, short dc)
{
    const uint8_t v = clamp_u8(wadd(wmul(dc * 4, 4096), 65536 + (128 << 17)) >> 17);
    for (int r = 0; r < 8; ++r) {
        std::memset(out + out_stride * r, v, 8);
    }
}

#if defined(__GNUC__) && !defined(CRT_JPEG_NO_VECTOR)
// The same arithmetic on eight lanes (GCC / Clang vector extensions; AVX2 or 2 x SSE2 underneath): the column pass with one
// lane per column, an 8 x 8 transpose, the row pass with one lane per row, a transpose back. Integer adds, multiplies and
// arithmetic shifts lane by lane — the values are the scalar version's (the all-zero-column shortcut of the scalar code
// computes the same numbers as the full pass: (d0 * 4096 + 512) >> 10 == d0 * 4).

#ifdef __AVX2__
inline void transpose8(v8i r[8])  // (written with intrinsics: the compiler turns the generic shuffles below into ~5x as many instructions)
{
    __m256i t[8], u[8];
    for (int k = 0; k < 4; ++k) {
        t[2 * k] = _mm256_unpacklo_epi32((__m256i)r[2 * k], (__m256i)r[2 * k + 1]);
        t[2 * k + 1] = _mm256_unpackhi_epi32((__m256i)r[2 * k], (__m256i)r[2 * k + 1]);
    }
    for (int k = 0; k < 2; ++k) {
        u[4 * k] = _mm256_unpacklo_epi64(t[4 * k], t[4 * k + 2]);
        u[4 * k + 1] = _mm256_unpackhi_epi64(t[4 * k], t[4 * k + 2]);
        u[4 * k + 2] = _mm256_unpacklo_epi64(t[4 * k + 1], t[4 * k + 3]);
        u[4 * k + 3] = _mm256_unpackhi_epi64(t[4 * k + 1], t[4 * k + 3]);
    }
    for (int k = 0; k < 4; ++k) {
        r[k] = (v8i)_mm256_permute2x128_si256(u[k], u[k + 4], 0x20);
        r[k + 4] = (v8i)_mm256_permute2x128_si256(u[k], u[k + 4], 0x31);
    }
}
#else
inline void transpose8(v8i r[8])
{
    const v8i lo32 = {0, 8, 1, 9, 4, 12, 5, 13}, hi32 = {2, 10, 3, 11, 6, 14, 7, 15};
    const v8i lo64 = {0, 1, 8, 9, 4, 5, 12, 13}, hi64 = {2, 3, 10, 11, 6, 7, 14, 15};
    const v8i lo128 = {0, 1, 2, 3, 8, 9, 10, 11}, hi128 = {4, 5, 6, 7, 12, 13, 14, 15};
    v8i t[8], u[8];
    for (int k = 0; k < 4; ++k) {
        t[2 * k] = __builtin_shuffle(r[2 * k], r[2 * k + 1], lo32);
        t[2 * k + 1] = __builtin_shuffle(r[2 * k], r[2 * k + 1], hi32);
    }
    for (int k = 0; k < 2; ++k) {
        u[4 * k] = __builtin_shuffle(t[4 * k], t[4 * k + 2], lo64);
        u[4 * k + 1] = __builtin_shuffle(t[4 * k], t[4 * k + 2], hi64);
        u[4 * k + 2] = __builtin_shuffle(t[4 * k + 1], t[4 * k + 3], lo64);
        u[4 * k + 3] = __builtin_shuffle(t[4 * k + 1], t[4 * k + 3], hi64);
    }
    for (int k = 0; k < 4; ++k) {
        r[k] = __builtin_shuffle(u[k], u[k + 4], lo128);
        r[k + 4] = __builtin_shuffle(u[k], u[k + 4], hi128);
    }
}
#endif

inline void idct_block(uint8_t *out, size_t out_stride, const short d[64])
{
    v8i v[8];
    for (int r = 0; r < 8; ++r) {
        v8s row;
        std::memcpy(&row, d + 8 * r, sizeof(row));
        v[r] = __builtin_convertvector(row, v8i);
    }
    Idct1d<v8i>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]).finish(v, 512, 10);
    transpose8(v);  // v[c]: column c of the intermediate, one lane per row
    Idct1d<v8i>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]).finish(v, 65536 + (128 << 17), 17);
    transpose8(v);  // back to rows
#ifdef __AVX2__
    for (int r = 0; r < 8; r += 2) {  // two rows at a time; the saturating packs are the clamp (|x| < 2^14 after the shift by 17)
        const __m256i words = _mm256_permute4x64_epi64(_mm256_packs_epi32((__m256i)v[r], (__m256i)v[r + 1]), 0xD8);  // a0-7 | b0-7
        const __m256i bytes = _mm256_packus_epi16(words, words);                                                     // a0-7 a0-7 | b0-7 b0-7
        _mm_storel_epi64(reinterpret_cast<__m128i *>(out + out_stride * r), _mm256_castsi256_si128(bytes));
        _mm_storel_epi64(reinterpret_cast<__m128i *>(out + out_stride * (r + 1)), _mm256_extracti128_si256(bytes, 1));
    }
#else
    for (int r = 0; r < 8; ++r) {
        for (int c = 0; c < 8; ++c) {
            out[out_stride * r + c] = clamp_u8(v[r][c]);
        }
    }
#endif
}
#else
inline void idct_block(uint8_t *out, size_t out_stride, const short d[64])
{
    idct_block_scalar(out, out_stride, d);
}
#endif

// one output row of a component from its two nearest stored rows; returns where the row is (in `out` or `near` itself)
inline const uint8_t *resample_row(uint8_t *out, const uint8_t *near, const uint8_t *far, int w, int hs, int vs)
{
    if (hs == 1 && vs == 1) {
        return near;
    }
    if (hs == 1 && vs == 2) {
        for (int i = 0; i < w; ++i) {
            out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
        }
        return out;
    }
    if (hs == 2 && vs == 1) {
        if (w == 1) {
            out[0] = out[1] = near[0];
            return out;
        }
        out[0] = near[0];
        out[1] = (uint8_t)((near[0] * 3 + near[1] + 2) >> 2);
        int i;
        for (i = 1; i < w - 1; ++i) {
            const int n = 3 * near[i] + 2;
            out[i * 2 + 0] = (uint8_t)((n + near[i - 1]) >> 2);
            out[i * 2 + 1] = (uint8_t)((n + near[i + 1]) >> 2);
        }
        out[i * 2 + 0] = (uint8_t)((near[w - 2] * 3 + near[w - 1] + 2) >> 2);
        out[i * 2 + 1] = near[w - 1];
        return out;
    }
    if (hs == 2 && vs == 2) {
        if (w == 1) {
            out[0] = out[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2);
            return out;
        }
        // t[i] = 3 * near[i] + far[i]; out[2i - 1] and out[2i] blend t[i - 1] and t[i] 3:1 / 1:3. In chunks, so that both loops are
        // free of loop-carried values (the compiler vectorises them)
        out[0] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2);
        for (int base = 1; base < w; base += 256) {
            const int n = std::min(256, w - base);
            uint16_t t[257];
            for (int k = 0; k <= n; ++k) {
                t[k] = (uint16_t)(3 * near[base - 1 + k] + far[base - 1 + k]);
            }
            for (int k = 0; k < n; ++k) {
                out[(base + k) * 2 - 1] = (uint8_t)((3 * t[k] + t[k + 1] + 8) >> 4);
                out[(base + k) * 2] = (uint8_t)((3 * t[k + 1] + t[k] + 8) >> 4);
            }
        }
        const int t1 = 3 * near[w - 1] + far[w - 1];
        out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
        return out;
    }
    for (int i = 0; i < w; ++i) {  // any other factor: nearest
        for (int j = 0; j < hs; ++j) {
            out[i * hs + j] = near[i];
        }
    }
    return out;
}

constexpr int float2fixed(float x)
{
    return ((int)(x * 4096.0f + 0.5f)) << 8;
}
inline void ycbcr_to_rgba_row(uint8_t *out, const uint8_t *y, const uint8_t *pcb, const uint8_t *pcr, int count)
{
    int i = 0;
#if defined(__AVX2__) && !defined(CRT_JPEG_NO_VECTOR)
    // the arithmetic below on eight pixels (32-bit lanes wrap like the unsigned scalar expression; the saturating packs are the
    // clamp: after the shift by 20 the values lie within +-2048)
    const __m256i c_r = _mm256_set1_epi32(float2fixed(1.40200f)), c_g1 = _mm256_set1_epi32(-float2fixed(0.71414f));
    const __m256i c_g2 = _mm256_set1_epi32(-float2fixed(0.34414f)), c_b = _mm256_set1_epi32(float2fixed(1.77200f));
    const __m256i half = _mm256_set1_epi32(1 << 19), mid = _mm256_set1_epi32(128), mask = _mm256_set1_epi32((int)0xffff0000u);
    const __m128i opaque = _mm_set1_epi8((char)0xff);
    for (; i + 8 <= count; i += 8) {
        const __m256i yy = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(y + i)));
        const __m256i cb = _mm256_sub_epi32(_mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(pcb + i))), mid);
        const __m256i cr = _mm256_sub_epi32(_mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(pcr + i))), mid);
        const __m256i y_fixed = _mm256_add_epi32(_mm256_slli_epi32(yy, 20), half);
        const __m256i r = _mm256_srai_epi32(_mm256_add_epi32(y_fixed, _mm256_mullo_epi32(cr, c_r)), 20);
        const __m256i g = _mm256_srai_epi32(_mm256_add_epi32(_mm256_add_epi32(y_fixed, _mm256_mullo_epi32(cr, c_g1)),
                                                             _mm256_and_si256(_mm256_mullo_epi32(cb, c_g2), mask)), 20);
        const __m256i b = _mm256_srai_epi32(_mm256_add_epi32(y_fixed, _mm256_mullo_epi32(cb, c_b)), 20);
        const __m256i rg16 = _mm256_permute4x64_epi64(_mm256_packs_epi32(r, g), 0xD8);  // r0-7 | g0-7 as 16-bit
        const __m256i b16 = _mm256_permute4x64_epi64(_mm256_packs_epi32(b, b), 0xD8);   // b0-7 | b0-7
        const __m128i r8 = _mm_packus_epi16(_mm256_castsi256_si128(rg16), _mm256_castsi256_si128(rg16));
        const __m128i g8 = _mm_packus_epi16(_mm256_extracti128_si256(rg16, 1), _mm256_extracti128_si256(rg16, 1));
        const __m128i b8 = _mm_packus_epi16(_mm256_castsi256_si128(b16), _mm256_castsi256_si128(b16));
        const __m128i rg = _mm_unpacklo_epi8(r8, g8), ba = _mm_unpacklo_epi8(b8, opaque);
        _mm_storeu_si128(reinterpret_cast<__m128i *>(out + 4 * i), _mm_unpacklo_epi16(rg, ba));
        _mm_storeu_si128(reinterpret_cast<__m128i *>(out + 4 * i + 16), _mm_unpackhi_epi16(rg, ba));
    }
#endif
    for (; i < count; ++i) {
        const int y_fixed = (y[i] << 20) + (1 << 19);
        const int cr = pcr[i] - 128, cb = pcb[i] - 128;
        int r = y_fixed + cr * float2fixed(1.40200f);
        int g = (int)((unsigned)(y_fixed + (cr * -float2fixed(0.71414f))) + ((unsigned)(cb * -float2fixed(0.34414f)) & 0xffff0000u));
        int b = y_fixed + cb * float2fixed(1.77200f);
        r >>= 20;
        g >>= 20;
        b >>= 20;
        out[4 * i] = clamp_u8(r);
        out[4 * i + 1] = clamp_u8(g);
        out[4 * i + 2] = clamp_u8(b);
        out[4 * i + 3] = 255;
    }
}

// ---- the stream (ITU T.81) ----
struct HuffmanTable {
    bool defined = false;
    uint8_t values[256];
    int mincode[17], maxcode[18], valptr[17];  // per code length (annex F.2.2.3)
    uint16_t fast[512];                        // 9 leading bits -> (length << 8) | symbol, 0 = longer code

    void build(const uint8_t counts[16], const uint8_t *symbols, int n)
    {
        std::memcpy(values, symbols, (size_t)n);
        std::memset(fast, 0, sizeof(fast));
        int code = 0, k = 0;
        for (int len = 1; len <= 16; ++len) {
            valptr[len] = k;
            mincode[len] = code;
            if (code + counts[len - 1] > (1 << len)) {  // more codes of this length than the code space has left
                throw std::runtime_error("bad code lengths");
            }
            for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
                if (len <= 9) {
                    const int first = code << (9 - len);
                    for (int f = 0; f < (1 << (9 - len)); ++f) {
                        fast[first + f] = (uint16_t)((len << 8) | values[k]);
                    }
                }
            }
            maxcode[len] = counts[len - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        defined = true;
    }
};

class Decoder {
public:
    Decoder(const uint8_t *data, size_t size) : p(data), end(data + size) {}

    // -> RGBA, `flip`: rows bottom-up (stbi_set_flip_vertically_on_load)
    void decode(std::vector<uint8_t> &out, int &width, int &height, bool flip)
    {
        if (next_marker() != 0xD8) {
            throw std::runtime_error("no SOI");
        }
        for (;;) {
            const int m = next_marker();
            if (m == 0xD9) {
                break;
            }
            if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
                if (frame_seen) {
                    throw std::runtime_error("second frame header");
                }
                frame_header(m == 0xC2);
            } else if (m == 0xDA) {
                if (!frame_seen) {
                    throw std::runtime_error("scan before the frame header");
                }
                scan();
            } else if (m == 0xDC) {  // DNL
                const int len = get16(), lines = get16();
                if (len != 4 || lines != img_y) {
                    throw std::runtime_error("bad DNL");
                }
            } else if (m < 0) {
                throw std::runtime_error("expected marker (the data ends without an EOI)");  // stb_image fails here as well
            } else {
                table_or_misc(m);
            }
        }
        if (!frame_seen || !scans) {
            throw std::runtime_error("no image data");
        }
        if (progressive) {
            for (Component &c : comps) {
                const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
                for (int j = 0; j < h; ++j) {
                    for (int i = 0; i < w; ++i) {
                        short *block = &c.coeff[64 * ((size_t)i + (size_t)j * c.blocks_w)];
                        for (int k = 0; k < 64; ++k) {
                            block[k] = (short)(block[k] * quant[c.tq][k]);
                        }
                        idct_block(&c.data[(size_t)c.w2 * j * 8 + i * 8], (size_t)c.w2, block);
                    }
                }
            }
        }
        assemble(out, flip);
        width = img_x;
        height = img_y;
    }

private:
    struct Component {
        int id = 0, h = 1, v = 1, tq = 0, hd = 0, ha = 0, dc_pred = 0;
        int x = 0, y = 0, w2 = 0, h2 = 0, blocks_w = 0;
        std::vector<uint8_t> data;
        std::vector<short> coeff;  // progressive only
    };
    const uint8_t *p, *end;
    bool frame_seen = false, progressive = false, jfif = false;
    int adobe_transform = -1, scans = 0;
    int img_x = 0, img_y = 0, h_max = 1, v_max = 1, mcu_x = 0, mcu_y = 0, restart_interval = 0;
    std::vector<Component> comps;
    uint16_t quant[4][64] = {};
    HuffmanTable dc_tables[4], ac_tables[4];
    // entropy-coded segment
    uint32_t bit_buffer = 0;
    int bit_count = 0;
    int pending_marker = -1;  // a marker met while filling the bit buffer
    int eob_run = 0;

    int get8()
    {
        return p < end ? *p++ : 0;
    }
    int get16()
    {
        const int hi = get8();
        return (hi << 8) | get8();
    }
    // the next marker code; -1 at the end of the data. Bytes that are no marker are skipped (padding, stray data)
    int next_marker()
    {
        if (pending_marker >= 0) {
            const int m = pending_marker;
            pending_marker = -1;
            return m;
        }
        while (p < end) {
            if (*p++ != 0xFF) {
                continue;
            }
            while (p < end && *p == 0xFF) {
                ++p;
            }
            if (p < end && *p != 0) {
                return *p++;
            }
        }
        return -1;
    }
    static const uint8_t *zigzag()
    {
        static const uint8_t z[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                      41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                      30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
        return z;
    }

    void table_or_misc(int m)
    {
        if (m == 0xDD) {  // DRI
            if (get16() != 4) {
                throw std::runtime_error("bad DRI");
            }
            restart_interval = get16();
            return;
        }
        if (m == 0xDB) {  // DQT
            int len = get16() - 2;
            while (len > 0) {
                const int q = get8(), precision = q >> 4, t = q & 15;
                if (precision > 1 || t > 3) {
                    throw std::runtime_error("bad DQT");
                }
                for (int i = 0; i < 64; ++i) {
                    quant[t][zigzag()[i]] = (uint16_t)(precision ? get16() : get8());
                }
                len -= precision ? 129 : 65;
            }
            if (len != 0) {
                throw std::runtime_error("bad DQT length");
            }
            return;
        }
        if (m == 0xC4) {  // DHT
            int len = get16() - 2;
            while (len > 0) {
                const int q = get8(), tc = q >> 4, th = q & 15;
                if (tc > 1 || th > 3) {
                    throw std::runtime_error("bad DHT");
                }
                uint8_t counts[16], symbols[256];
                int n = 0;
                for (int i = 0; i < 16; ++i) {
                    counts[i] = (uint8_t)get8();
                    n += counts[i];
                }
                if (n > 256) {
                    throw std::runtime_error("bad DHT");
                }
                for (int i = 0; i < n; ++i) {
                    symbols[i] = (uint8_t)get8();
                }
                (tc ? ac_tables : dc_tables)[th].build(counts, symbols, n);
                len -= 17 + n;
            }
            if (len != 0) {
                throw std::runtime_error("bad DHT length");
            }
            return;
        }
        if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) {  // APPn, COM
            int len = get16();
            if (len < 2) {
                throw std::runtime_error("bad segment length");
            }
            len -= 2;
            if (m == 0xE0 && len >= 5) {
                jfif = jfif || (end - p >= 5 && std::memcmp(p, "JFIF\0", 5) == 0);
            } else if (m == 0xEE && len >= 12 && end - p >= 12 && std::memcmp(p, "Adobe\0", 6) == 0) {
                adobe_transform = p[11];
            }
            p = (size_t)(end - p) < (size_t)len ? end : p + len;
            return;
        }
        if (m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            throw std::runtime_error("lossless, hierarchical or arithmetic-coded JPEG");
        }
        throw std::runtime_error("unknown marker");
    }

    void frame_header(bool is_progressive)
    {
        progressive = is_progressive;
        const int len = get16(), precision = get8();
        img_y = get16();
        img_x = get16();
        const int n = get8();
        if (precision != 8) {
            throw std::runtime_error("only 8-bit samples");
        }
        if (img_x == 0 || img_y == 0) {
            throw std::runtime_error("no image size in the frame header");
        }
        if (n == 4) {
            throw std::runtime_error("four-component (CMYK) JPEG");
        }
        if ((n != 1 && n != 3) || len != 8 + 3 * n) {
            throw std::runtime_error("bad frame header");
        }
        if ((uint64_t)img_x * (uint64_t)img_y > ((uint64_t)1 << 28)) {
            throw std::runtime_error("image too large");
        }
        comps.resize((size_t)n);
        for (Component &c : comps) {
            c.id = get8();
            const int q = get8();
            c.h = q >> 4;
            c.v = q & 15;
            c.tq = get8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) {
                throw std::runtime_error("bad component");
            }
            h_max = std::max(h_max, c.h);
            v_max = std::max(v_max, c.v);
        }
        mcu_x = (img_x + h_max * 8 - 1) / (h_max * 8);
        mcu_y = (img_y + v_max * 8 - 1) / (v_max * 8);
        for (Component &c : comps) {
            c.x = (img_x * c.h + h_max - 1) / h_max;
            c.y = (img_y * c.v + v_max - 1) / v_max;
            c.w2 = mcu_x * c.h * 8;
            c.h2 = mcu_y * c.v * 8;
            c.blocks_w = c.w2 / 8;
            c.data.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) {
                c.coeff.assign((size_t)c.w2 * c.h2, 0);
            }
        }
        frame_seen = true;
    }

    // ---- bits ----
    void fill()
    {
        while (bit_count <= 24) {
            uint32_t b = 0;
            if (pending_marker < 0 && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    int c = get8();
                    while (c == 0xFF) {
                        c = get8();
                    }
                    if (c != 0) {  // a marker ends the entropy-coded data; zeros from here on
                        pending_marker = c;
                        b = 0;
                    }
                }
            }
            bit_buffer |= b << (24 - bit_count);
            bit_count += 8;
        }
    }
    int get_bits(int n)
    {
        if (n == 0) {
            return 0;
        }
        if (bit_count < n) {
            fill();
        }
        const int v = (int)(bit_buffer >> (32 - n));
        bit_buffer <<= n;
        bit_count -= n;
        return v;
    }
    int receive_extend(int n)  // F.2.2.1
    {
        if (n == 0) {
            return 0;
        }
        const int v = get_bits(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
    int huffman(const HuffmanTable &t)
    {
        if (bit_count < 16) {
            fill();
        }
        const uint16_t f = t.fast[bit_buffer >> 23];
        if (f) {
            const int len = f >> 8;
            bit_buffer <<= len;
            bit_count -= len;
            return f & 255;
        }
        int code = (int)(bit_buffer >> 22), len = 10;  // codes of 10 to 16 bits
        for (; len <= 16; ++len, code = (int)(bit_buffer >> (32 - len))) {
            if (t.maxcode[len] >= 0 && code <= t.maxcode[len] && code >= t.mincode[len]) {
                bit_buffer <<= len;
                bit_count -= len;
                return t.values[t.valptr[len] + code - t.mincode[len]];
            }
        }
        throw std::runtime_error("bad huffman code");
    }
    void restart()
    {
        bit_buffer = 0;
        bit_count = 0;
        eob_run = 0;
        for (Component &c : comps) {
            c.dc_pred = 0;
        }
    }

    // ---- blocks ----
    // returns whether any AC coefficient was decoded
    bool sequential_block(Component &c, short block[64])
    {
        bool any_ac = false;
        std::memset(block, 0, 64 * sizeof(short));
        const int t = huffman(dc_tables[c.hd]);
        if (t > 15) {
            throw std::runtime_error("bad DC code");
        }
        c.dc_pred = wadd(c.dc_pred, receive_extend(t));
        block[0] = (short)wmul(c.dc_pred, quant[c.tq][0]);
        const HuffmanTable &ac = ac_tables[c.ha];
        for (int k = 1; k < 64;) {
            const int rs = huffman(ac), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (rs != 0xF0) {
                    break;
                }
                k += 16;
                continue;
            }
            k += r;
            if (k > 63) {
                throw std::runtime_error("coefficient index past the block");
            }
            const int z = zigzag()[k++];
            block[z] = (short)(receive_extend(s) * quant[c.tq][z]);
            any_ac = true;
        }
        return any_ac;
    }
    void progressive_dc(Component &c, short block[64], int succ_high, int succ_low)
    {
        if (succ_high == 0) {
            const int t = huffman(dc_tables[c.hd]);
            if (t > 15) {
                throw std::runtime_error("bad DC code");
            }
            c.dc_pred = wadd(c.dc_pred, receive_extend(t));
            block[0] = (short)wmul(c.dc_pred, 1 << succ_low);
        } else if (get_bits(1)) {
            block[0] = (short)(block[0] + (1 << succ_low));
        }
    }
    void progressive_ac(Component &c, short block[64], int spec_start, int spec_end, int succ_high, int succ_low)
    {
        const HuffmanTable &ac = ac_tables[c.ha];
        if (succ_high == 0) {  // first pass over this band (G.1.2.2)
            if (eob_run) {
                --eob_run;
                return;
            }
            for (int k = spec_start; k <= spec_end;) {
                const int rs = huffman(ac), r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = (1 << r) - 1 + get_bits(r);
                        break;
                    }
                    k += 16;
                    continue;
                }
                k += r;
                if (k > 63) {
                    throw std::runtime_error("coefficient index past the block");
                }
                block[zigzag()[k++]] = (short)(receive_extend(s) * (1 << succ_low));
            }
            return;
        }
        // refinement (G.1.2.3): one more bit for every coefficient that is already non-zero, new +-1 coefficients in between
        const short bit = (short)(1 << succ_low);
        const auto refine = [&](short &v) {
            if (get_bits(1) && (v & bit) == 0) {
                v = (short)(v > 0 ? v + bit : v - bit);
            }
        };
        if (eob_run) {
            --eob_run;
            for (int k = spec_start; k <= spec_end; ++k) {
                short &v = block[zigzag()[k]];
                if (v != 0) {
                    refine(v);
                }
            }
            return;
        }
        for (int k = spec_start; k <= spec_end;) {
            const int rs = huffman(ac);
            int r = rs >> 4, s = rs & 15;
            short value = 0;
            if (s == 0) {
                if (r < 15) {
                    eob_run = (1 << r) - 1 + get_bits(r);
                    r = 64;  // to the end of the band: only refinement bits follow for this block
                }
            } else {
                if (s != 1) {
                    throw std::runtime_error("bad refinement code");
                }
                value = get_bits(1) ? bit : (short)-bit;
            }
            while (k <= spec_end) {
                short &v = block[zigzag()[k++]];
                if (v != 0) {
                    refine(v);
                } else {
                    if (r == 0) {
                        v = value;
                        break;
                    }
                    --r;
                }
            }
        }
    }

    void scan()
    {
        const int len = get16(), n = get8();
        if (n < 1 || n > (int)comps.size() || len != 6 + 2 * n) {
            throw std::runtime_error("bad scan header");
        }
        Component *order[4];
        for (int i = 0; i < n; ++i) {
            const int id = get8(), q = get8();
            order[i] = nullptr;
            for (Component &c : comps) {
                if (c.id == id) {
                    order[i] = &c;
                }
            }
            if (!order[i] || (q >> 4) > 3 || (q & 15) > 3) {
                throw std::runtime_error("bad scan component");
            }
            order[i]->hd = q >> 4;
            order[i]->ha = q & 15;
        }
        int spec_start = get8(), spec_end = get8();
        const int aa = get8(), succ_high = aa >> 4, succ_low = aa & 15;
        if (progressive) {
            if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13 ||
                (spec_start == 0 && spec_end != 0) || (spec_start != 0 && n != 1)) {
                throw std::runtime_error("bad progressive scan");
            }
        } else {
            if (spec_start != 0 || succ_high != 0 || succ_low != 0) {
                throw std::runtime_error("bad scan parameters");
            }
            spec_end = 63;
        }
        for (int i = 0; i < n; ++i) {
            const bool need_dc = spec_start == 0 && succ_high == 0, need_ac = !progressive || spec_start != 0;
            if ((need_dc && !dc_tables[order[i]->hd].defined) || (need_ac && !ac_tables[order[i]->ha].defined)) {
                throw std::runtime_error("scan without its huffman table");
            }
        }
        ++scans;
        restart();
        pending_marker = -1;
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        // after every restart interval: the next marker must be RSTn; if it is not, the scan ends with what it has (stb_image)
        const auto interval_done = [&]() {
            if (--todo > 0) {
                return false;
            }
            if (bit_count < 24) {
                fill();
            }
            if (pending_marker < 0xD0 || pending_marker > 0xD7) {
                return true;
            }
            pending_marker = -1;
            restart();
            todo = restart_interval ? restart_interval : 0x7fffffff;
            return false;
        };
        short block[64];
        const auto one_block = [&](Component &c, int bx, int by) {
            if (!progressive) {
                uint8_t *dst = &c.data[(size_t)c.w2 * by * 8 + bx * 8];
                if (sequential_block(c, block)) {
                    idct_block(dst, (size_t)c.w2, block);
                } else {
                    idct_dc_only(dst, (size_t)c.w2, block[0]);
                }
                return;
            }
            short *coeff = &c.coeff[64 * ((size_t)bx + (size_t)by * c.blocks_w)];
            if (spec_start == 0) {
                progressive_dc(c, coeff, succ_high, succ_low);
            } else {
                progressive_ac(c, coeff, spec_start, spec_end, succ_high, succ_low);
            }
        };
        if (n == 1) {  // not interleaved: the component's own blocks, row by row
            Component &c = *order[0];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h; ++j) {
                for (int i = 0; i < w; ++i) {
                    one_block(c, i, j);
                    if (interval_done()) {
                        return;
                    }
                }
            }
            return;
        }
        for (int j = 0; j < mcu_y; ++j) {
            for (int i = 0; i < mcu_x; ++i) {
                for (int k = 0; k < n; ++k) {
                    Component &c = *order[k];
                    for (int y = 0; y < c.v; ++y) {
                        for (int x = 0; x < c.h; ++x) {
                            one_block(c, i * c.h + x, j * c.v + y);
                        }
                    }
                }
                if (interval_done()) {
                    return;
                }
            }
        }
    }

    // load_jpeg_image (stb_image.h:3640-3700) for four requested components
    void assemble(std::vector<uint8_t> &out, bool flip)
    {
        const size_t n = comps.size();
        int rgb_ids = 0;
        for (size_t k = 0; k < n; ++k) {
            rgb_ids += (n == 3 && comps[k].id == "RGB"[k]) ? 1 : 0;
        }
        const bool is_rgb = n == 3 && (rgb_ids == 3 || (adobe_transform == 0 && !jfif));
        struct Resample {
            int hs, vs, ystep, w_lores, ypos;
            const uint8_t *line0, *line1;
            std::vector<uint8_t> buffer;
        };
        std::vector<Resample> res(n);
        for (size_t k = 0; k < n; ++k) {
            Resample &r = res[k];
            r.hs = h_max / comps[k].h;
            r.vs = v_max / comps[k].v;
            r.ystep = r.vs >> 1;
            r.w_lores = (img_x + r.hs - 1) / r.hs;
            r.ypos = 0;
            r.line0 = r.line1 = comps[k].data.data();
            r.buffer.resize((size_t)img_x + 3 + 8);
        }
        out.resize((size_t)img_x * img_y * 4);
        const uint8_t *rows[3] = {nullptr, nullptr, nullptr};
        for (int j = 0; j < img_y; ++j) {
            uint8_t *dst = &out[(size_t)img_x * 4 * (flip ? img_y - 1 - j : j)];
            for (size_t k = 0; k < n; ++k) {
                Resample &r = res[k];
                const bool y_bot = r.ystep >= (r.vs >> 1);
                rows[k] = resample_row(r.buffer.data(), y_bot ? r.line1 : r.line0, y_bot ? r.line0 : r.line1, r.w_lores, r.hs, r.vs);
                if (++r.ystep >= r.vs) {
                    r.ystep = 0;
                    r.line0 = r.line1;
                    if (++r.ypos < comps[k].y) {
                        r.line1 += comps[k].w2;
                    }
                }
            }
            if (n == 3 && !is_rgb) {
                ycbcr_to_rgba_row(dst, rows[0], rows[1], rows[2], img_x);
            } else {
                for (int i = 0; i < img_x; ++i) {
                    dst[4 * i] = rows[0][i];
                    dst[4 * i + 1] = n == 3 ? rows[1][i] : rows[0][i];
                    dst[4 * i + 2] = n == 3 ? rows[2][i] : rows[0][i];
                    dst[4 * i + 3] = 255;
                }
            }
        }
    }
};

inline bool is_jpeg(const uint8_t *data, size_t size)
{
    return size >= 2 && data[0] == 0xFF && data[1] == 0xD8;
}

// RGBA pixels of a JPEG file as stbi_load_from_memory(..., 4) returns them; `name` for messages
inline void decode_rgba(const uint8_t *data, size_t size, const std::string &name, std::vector<uint8_t> &out, int &width, int &height,
                        bool flip)
{
    try {
        Decoder(data, size).decode(out, width, height, flip);
    } catch (const std::exception &e) {
        throw std::runtime_error("JPEG " + name + ": " + e.what());
    }
}

}  // namespace crt_jpeg
