// image_decode.h — texture files to 8-bit RGBA for the native scene loader (scene_io.cpp), host code only: PNG and TGA here,
// JPEG in jpeg_decode.h; the bytes are the ones stb_image returns for stbi_load(..., 4) (the reference decodes every texture
// with stb_image: util/material.cpp:5-17, util/scene.cpp:488-511, tinygltf's LoadImageData).
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "jpeg_decode.h"

#if defined(__SSE4_1__)
#include <smmintrin.h>
#endif

namespace crt_image {

inline uint32_t be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

// `file`: the bytes of a PNG file; `path`: its name for messages. Every bit depth (1, 2, 4, 8, 16) and colour type, Adam7
// interlacing, tRNS transparency; to 8-bit RGBA the way stbi_load(..., 4) gets there (stb_image.h: stbi__parse_png_file):
// grey samples of fewer than 8 bits are scaled to the full range (x 0xff / 0x55 / 0x11), palette indices are looked up, a
// 16-bit sample keeps its high byte, and the one transparent colour of a grey / RGB image (tRNS) is compared at the file's
// own sample width.
inline void decode_png_rgba(const uint8_t *file_data, size_t file_size, const std::string &path, std::vector<uint8_t> &out, int &width,
                     int &height, bool flip)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file_size < 8 + 25 || std::memcmp(file_data, sig, 8) != 0) {
        throw std::runtime_error("not a PNG, JPEG, TGA or BMP file (the formats this loader reads): " + path);
    }
    // the chunks, with the structural checks stb_image makes (stbi__parse_png_file: a file it refuses is refused here)
    size_t pos = 8;
    int bit_depth = 0, color_type = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool have_header = false, have_idat = false, ended = false;
    width = height = 0;
    const auto corrupt = [&](const char *what) { return std::runtime_error(std::string("corrupt PNG (") + what + "): " + path); };
    while (!ended && pos + 8 <= file_size) {
        const uint32_t len = be32(file_data + pos);
        const char *type = reinterpret_cast<const char *>(file_data + pos + 4);
        const uint8_t *body = file_data + pos + 8;
        if ((size_t)len > file_size - (pos + 8)) {
            throw std::runtime_error("truncated PNG: " + path);
        }
        const auto is = [&](const char *name) { return std::memcmp(type, name, 4) == 0; };
        if (!have_header && !is("IHDR") && !is("CgBI")) {
            throw corrupt("first not IHDR");
        }
        if (is("IHDR")) {
            if (have_header || len != 13) {
                throw corrupt("IHDR");
            }
            have_header = true;
            width = (int)be32(body);
            height = (int)be32(body + 4);
            bit_depth = body[8];
            color_type = body[9];
            interlace = body[12];
            if (be32(body) > (1u << 24) || be32(body + 4) > (1u << 24) || body[10] != 0 || body[11] != 0) {
                throw corrupt("size, compression or filter method");
            }
        } else if (is("PLTE")) {
            if (len > 256 * 3 || len % 3 != 0) {
                throw corrupt("invalid PLTE");
            }
            palette.assign(body, body + len);
        } else if (is("tRNS")) {
            if (have_idat) {
                throw corrupt("tRNS after IDAT");
            }
            if (color_type == 3) {
                if (palette.empty() || len > palette.size() / 3) {
                    throw corrupt("tRNS and PLTE");
                }
            } else if ((color_type & 4) || len != (uint32_t)((color_type & 2) ? 6 : 2)) {
                throw corrupt("tRNS of an image with alpha, or of the wrong length");
            }
            trns.assign(body, body + len);
        } else if (is("IDAT")) {
            if (color_type == 3 && palette.empty()) {
                throw corrupt("no PLTE");
            }
            have_idat = true;
            idat.insert(idat.end(), body, body + len);
        } else if (is("IEND")) {
            ended = true;
        } else if ((type[0] & 0x20) == 0 && !is("CgBI")) {
            throw std::runtime_error("PNG not supported: unknown critical chunk: " + path);
        }
        pos += 8 + (size_t)len + 4;  // (the CRC is not checked, as in stb_image)
    }
    if (!have_header || !have_idat) {
        throw corrupt("no IHDR or no IDAT");
    }
    int ch;
    switch (color_type) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 3: ch = 1; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: throw std::runtime_error("unsupported PNG colour type: " + path);
    }
    const bool depth_ok = bit_depth == 8 || (bit_depth == 16 && color_type != 3) ||
                          ((bit_depth == 1 || bit_depth == 2 || bit_depth == 4) && (color_type == 0 || color_type == 3));
    if (width <= 0 || height <= 0 || (uint64_t)width * (uint64_t)height > ((uint64_t)1 << 28) || !depth_ok || interlace > 1) {
        throw std::runtime_error("unsupported or corrupt PNG header: " + path);
    }
    // ---- the passes of the image (one, or the seven of Adam7) and their sizes in the inflated stream ----
    struct Pass {
        int x0, y0, dx, dy, w, h;
        size_t row_bytes;
    };
    std::vector<Pass> passes;
    const auto add_pass = [&](int x0, int y0, int dx, int dy) {
        const int w = (width - x0 + dx - 1) / dx, h = (height - y0 + dy - 1) / dy;
        if (w > 0 && h > 0) {
            passes.push_back(Pass{x0, y0, dx, dy, w, h, ((size_t)w * ch * bit_depth + 7) / 8});
        }
    };
    if (interlace) {
        static const int x0[7] = {0, 4, 0, 2, 0, 1, 0}, y0[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
        for (int k = 0; k < 7; ++k) {
            add_pass(x0[k], y0[k], dx[k], dy[k]);
        }
    } else {
        add_pass(0, 0, 1, 1);
    }
    size_t raw_size = 0;
    for (const Pass &ps : passes) {
        raw_size += (ps.row_bytes + 1) * (size_t)ps.h;
    }
    std::vector<uint8_t> raw(raw_size + 16);  // (+ room for the 4-byte loads of the last 3-byte pixel of the last row)
    {
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit(&zs) != Z_OK) {
            throw std::runtime_error("zlib: " + path);
        }
        zs.next_in = idat.data();
        zs.avail_in = (uInt)idat.size();
        zs.next_out = raw.data();
        zs.avail_out = (uInt)raw_size;
        const int rc = inflate(&zs, Z_FINISH);
        const size_t got = raw_size - zs.avail_out;
        inflateEnd(&zs);
        if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || got != raw_size) {  // (more data than the image needs is ignored)
            throw std::runtime_error("corrupt PNG data: " + path);
        }
    }
    // ---- per pass: undo the scanline filters in place (PNG specification, section 9), then place the samples ----
    const int sample_bytes = bit_depth == 16 ? 2 : 1;
    const size_t filter_bpp = std::max<size_t>(1, (size_t)ch * bit_depth / 8);
    const bool rows_in_place = !interlace && bit_depth >= 8;  // the defiltered scanlines already are the rows of the image
    std::vector<uint8_t> samples;  // else: one byte (two for 16 bits) per sample, unscaled, in image order
    if (!rows_in_place) {
        samples.resize((size_t)width * height * ch * sample_bytes);
    }
    size_t offset = 0;
    size_t longest_row = 0;
    for (const Pass &ps : passes) {
        longest_row = std::max(longest_row, ps.row_bytes);
    }
    const std::vector<uint8_t> zero_row(longest_row + 16, 0);  // the row "above" the first one of a pass (+ room as above)
    for (const Pass &ps : passes) {
        const size_t n = ps.row_bytes, bpp = std::min(filter_bpp, ps.row_bytes);
        const uint8_t *prev = zero_row.data();
        for (int y = 0; y < ps.h; ++y) {
            const uint8_t filter = raw[offset];
            uint8_t *cur = &raw[offset + 1];
            offset += n + 1;
            switch (filter) {
            case 0: break;
            case 1:
                for (size_t x = bpp; x < n; ++x) {
                    cur[x] = (uint8_t)(cur[x] + cur[x - bpp]);
                }
                break;
            case 2:
                for (size_t x = 0; x < n; ++x) {
                    cur[x] = (uint8_t)(cur[x] + prev[x]);
                }
                break;
            case 3:
                for (size_t x = 0; x < bpp; ++x) {
                    cur[x] = (uint8_t)(cur[x] + (prev[x] >> 1));
                }
                for (size_t x = bpp; x < n; ++x) {
                    cur[x] = (uint8_t)(cur[x] + ((cur[x - bpp] + prev[x]) >> 1));
                }
                break;
            case 4:
                for (size_t x = 0; x < bpp; ++x) {
                    cur[x] = (uint8_t)(cur[x] + prev[x]);  // (a = c = 0: the predictor is b)
                }
                {
                    size_t x = bpp;
#if defined(__SSE4_1__)
                    if ((bpp == 3 || bpp == 4) && n >= 2 * bpp) {
                        // one pixel per step, its channels in 16-bit lanes: pa = |b - c|, pb = |a - c|, pc = |a + b - 2c|; the
                        // predictor is a, b or c by the smallest of them, ties in that order (the scalar rule below)
                        const auto load = [](const uint8_t *p) {
                            int32_t v;
                            std::memcpy(&v, p, 4);  // (4 bytes also for 3-byte pixels: the buffer is padded)
                            return _mm_cvtepu8_epi16(_mm_cvtsi32_si128(v));
                        };
                        __m128i a = load(cur), c = load(prev);  // the first pixel is done: it is the left neighbour of the next
                        for (; x + bpp <= n; x += bpp) {
                            const __m128i b = load(prev + x);
                            const __m128i d_bc = _mm_sub_epi16(b, c), d_ac = _mm_sub_epi16(a, c);
                            const __m128i pa = _mm_abs_epi16(d_bc), pb = _mm_abs_epi16(d_ac), pc = _mm_abs_epi16(_mm_add_epi16(d_bc, d_ac));
                            const __m128i smallest = _mm_min_epi16(pc, _mm_min_epi16(pa, pb));
                            const __m128i nearest = _mm_blendv_epi8(_mm_blendv_epi8(c, b, _mm_cmpeq_epi16(smallest, pb)), a, _mm_cmpeq_epi16(smallest, pa));
                            a = _mm_and_si128(_mm_add_epi16(load(cur + x), nearest), _mm_set1_epi16(0xff));
                            c = b;
                            const int32_t px = _mm_cvtsi128_si32(_mm_packus_epi16(a, a));
                            std::memcpy(cur + x, &px, bpp);
                        }
                    }
#endif
                    for (; x < n; ++x) {
                        const int a = cur[x - bpp], b = prev[x], c = prev[x - bpp];
                        const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                        cur[x] = (uint8_t)(cur[x] + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)));
                    }
                }
                break;
            default: throw std::runtime_error("corrupt PNG filter: " + path);
            }
            prev = cur;
            if (rows_in_place) {
                continue;
            }
            uint8_t *dst_row = &samples[(size_t)(ps.y0 + y * ps.dy) * width * ch * sample_bytes];
            for (int x = 0; x < ps.w; ++x) {
                uint8_t *dst = dst_row + (size_t)(ps.x0 + x * ps.dx) * ch * sample_bytes;
                if (bit_depth >= 8) {
                    std::memcpy(dst, &cur[(size_t)x * ch * sample_bytes], (size_t)ch * sample_bytes);
                } else {  // one channel, several samples per byte, the leftmost in the high bits
                    const int per_byte = 8 / bit_depth, shift = (per_byte - 1 - x % per_byte) * bit_depth;
                    dst[0] = (uint8_t)((cur[(size_t)x / per_byte] >> shift) & ((1 << bit_depth) - 1));
                }
            }
        }
    }
    const size_t sample_stride = rows_in_place ? passes[0].row_bytes + 1 : (size_t)width * ch * sample_bytes;
    const uint8_t *sample_base = rows_in_place ? raw.data() + 1 : samples.data();
    // ---- to RGBA as stbi_load(..., 4) does; flip: rows bottom-up (stbi_set_flip_vertically_on_load(1), util/material.cpp:8) ----
    static const int depth_scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    const int scale = color_type == 0 && bit_depth < 8 ? depth_scale[bit_depth] : 1;
    const auto sample16 = [](const uint8_t *p) { return (uint32_t)((p[0] << 8) | p[1]); };
    uint32_t key[3] = {0, 0, 0};  // the transparent colour of a grey / RGB image, at the sample width the comparison uses
    const bool has_key = (color_type == 0 && trns.size() >= 2) || (color_type == 2 && trns.size() >= 6);
    for (int k = 0; has_key && k < (color_type == 0 ? 1 : 3); ++k) {
        const uint32_t v = sample16(&trns[2 * (size_t)k]);
        key[k] = bit_depth == 16 ? v : (uint32_t)(uint8_t)((v & 255) * (uint32_t)scale);
    }
    out.resize((size_t)width * height * 4);
    for (int y = 0; y < height; ++y) {
        const uint8_t *src = sample_base + (size_t)(flip ? height - 1 - y : y) * sample_stride;
        uint8_t *dst = &out[(size_t)width * 4 * y];
        if (bit_depth == 8 && !has_key && color_type == 6) {  // the common cases without per-pixel decisions
            std::memcpy(dst, src, (size_t)width * 4);
            continue;
        }
        if (bit_depth == 8 && !has_key && color_type == 2) {
            int x = 0;
#if defined(__SSE4_1__)
            const __m128i spread = _mm_setr_epi8(0, 1, 2, -1, 3, 4, 5, -1, 6, 7, 8, -1, 9, 10, 11, -1), alpha = _mm_set1_epi32((int)0xff000000u);
            for (; x + 6 <= width; x += 4) {  // four pixels per step; the 16-byte load reads into the fifth and sixth
                const __m128i rgb = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 3 * x));
                _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + 4 * x), _mm_or_si128(_mm_shuffle_epi8(rgb, spread), alpha));
            }
#endif
            for (; x < width; ++x) {
                dst[4 * x] = src[3 * x], dst[4 * x + 1] = src[3 * x + 1], dst[4 * x + 2] = src[3 * x + 2], dst[4 * x + 3] = 255;
            }
            continue;
        }
        for (int x = 0; x < width; ++x) {
            const uint8_t *px = src + (size_t)x * ch * sample_bytes;
            uint8_t r, g, b, a = 255;
            if (color_type == 3) {
                const size_t i = px[0];
                if (3 * i + 2 >= palette.size()) {
                    throw std::runtime_error("corrupt PNG palette: " + path);
                }
                r = palette[3 * i], g = palette[3 * i + 1], b = palette[3 * i + 2];
                a = i < trns.size() ? trns[i] : 255;
            } else if (bit_depth == 16) {
                uint32_t v[4] = {0, 0, 0, 0xffff};
                for (int k = 0; k < ch; ++k) {
                    v[k] = sample16(px + 2 * k);
                }
                switch (color_type) {
                case 0: r = g = b = (uint8_t)(v[0] >> 8), a = (has_key && v[0] == key[0]) ? 0 : 255; break;
                case 2:
                    r = (uint8_t)(v[0] >> 8), g = (uint8_t)(v[1] >> 8), b = (uint8_t)(v[2] >> 8);
                    a = (has_key && v[0] == key[0] && v[1] == key[1] && v[2] == key[2]) ? 0 : 255;
                    break;
                case 4: r = g = b = (uint8_t)(v[0] >> 8), a = (uint8_t)(v[1] >> 8); break;
                default: r = (uint8_t)(v[0] >> 8), g = (uint8_t)(v[1] >> 8), b = (uint8_t)(v[2] >> 8), a = (uint8_t)(v[3] >> 8); break;
                }
            } else {
                switch (color_type) {
                case 0:
                    r = g = b = (uint8_t)(px[0] * scale);
                    a = (has_key && r == key[0]) ? 0 : 255;
                    break;
                case 2:
                    r = px[0], g = px[1], b = px[2];
                    a = (has_key && r == key[0] && g == key[1] && b == key[2]) ? 0 : 255;
                    break;
                case 4: r = g = b = px[0], a = px[1]; break;
                default: r = px[0], g = px[1], b = px[2], a = px[3]; break;
                }
            }
            dst[4 * x] = r, dst[4 * x + 1] = g, dst[4 * x + 2] = b, dst[4 * x + 3] = a;
        }
    }
}

// ---- TGA (stb_image.h: stbi__tga_test, stbi__tga_load): true-colour, grey and colour-mapped images, raw or run-length
// encoded, 8 / 15 / 16 / 24 / 32 bits; stb_image's reading of the format: 15- and 16-bit pixels are 5-5-5 RGB without alpha
// ((c * 255) / 31), a 16-bit grey image is grey + alpha, the colour map starts `first entry index` BYTES into its data, an
// index past the map reads entry 0, bytes missing at the end read as 0. `path`: its name for messages.
inline bool looks_like_tga(const uint8_t *d, size_t n)
{
    if (n < 18 || d[1] > 1) {
        return false;
    }
    const int type = d[2], bpp = d[16];
    if (d[1] == 1) {
        const int pal_bits = d[7];
        if ((type != 1 && type != 9) || (pal_bits != 8 && pal_bits != 15 && pal_bits != 16 && pal_bits != 24 && pal_bits != 32) ||
            (bpp != 8 && bpp != 16)) {
            return false;
        }
    } else if (type != 2 && type != 3 && type != 10 && type != 11) {
        return false;
    }
    const int w = d[12] | (d[13] << 8), h = d[14] | (d[15] << 8);
    return w >= 1 && h >= 1 && (bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32);
}

inline void decode_tga_rgba(const uint8_t *data, size_t size, const std::string &path, std::vector<uint8_t> &out, int &width, int &height, bool flip)
{
    if (!looks_like_tga(data, size)) {
        throw std::runtime_error("not a PNG, JPEG, TGA or BMP file (the formats this loader reads): " + path);
    }
    size_t pos = 0;
    const auto get8 = [&]() -> int { return pos < size ? data[pos++] : (++pos, 0); };
    const auto get16 = [&]() {
        const int lo = get8();
        return lo | (get8() << 8);
    };
    const int id_length = get8(), indexed = get8();
    int type = get8();
    const bool rle = type >= 8;
    type -= rle ? 8 : 0;
    const int pal_start = get16(), pal_len = get16(), pal_bits = get8();
    get16();
    get16();
    width = get16();
    height = get16();
    const int bpp = get8(), descriptor = get8();
    const bool bottom_up = ((descriptor >> 5) & 1) == 0;
    bool rgb16 = false;
    const auto components = [&](int bits, bool grey) {
        switch (bits) {
        case 8: return 1;
        case 16:
            if (grey) {
                return 2;
            }
            rgb16 = true;
            return 3;
        case 15: rgb16 = true; return 3;
        case 24: return 3;
        case 32: return 4;
        default: return 0;
        }
    };
    const int comp = indexed ? components(pal_bits, false) : components(bpp, type == 3);
    if (!comp) {
        throw std::runtime_error("unsupported TGA pixel format: " + path);
    }
    const auto read_rgb16 = [&](uint8_t *o) {
        const int px = get16();
        o[0] = (uint8_t)((((px >> 10) & 31) * 255) / 31);
        o[1] = (uint8_t)((((px >> 5) & 31) * 255) / 31);
        o[2] = (uint8_t)(((px & 31) * 255) / 31);
    };
    pos += (size_t)id_length;
    std::vector<uint8_t> map;
    if (indexed) {
        pos += (size_t)pal_start;
        map.resize((size_t)pal_len * comp);
        for (int i = 0; i < pal_len; ++i) {
            if (rgb16) {
                read_rgb16(&map[(size_t)i * comp]);
            } else {
                for (int j = 0; j < comp; ++j) {
                    map[(size_t)i * comp + j] = (uint8_t)get8();
                }
            }
        }
    }
    const size_t count = (size_t)width * height;
    std::vector<uint8_t> px(count * comp);
    uint8_t current[4] = {0, 0, 0, 0};
    int run = 0;
    bool repeating = false;
    for (size_t i = 0; i < count; ++i) {
        bool read = true;
        if (rle) {
            if (run == 0) {
                const int cmd = get8();
                run = 1 + (cmd & 127);
                repeating = (cmd >> 7) != 0;
            } else {
                read = !repeating;
            }
        }
        if (read) {
            if (indexed) {
                int idx = bpp == 8 ? get8() : get16();
                idx = idx >= pal_len ? 0 : idx;
                for (int j = 0; j < comp; ++j) {
                    current[j] = (size_t)idx * comp + j < map.size() ? map[(size_t)idx * comp + j] : 0;
                }
            } else if (rgb16) {
                read_rgb16(current);
            } else {
                for (int j = 0; j < comp; ++j) {
                    current[j] = (uint8_t)get8();
                }
            }
        }
        std::memcpy(&px[i * comp], current, (size_t)comp);
        --run;
    }
    out.resize(count * 4);
    for (int y = 0; y < height; ++y) {
        // the file's rows are bottom-up unless the descriptor says otherwise; stbi returns top-down rows, `flip` turns them again
        const bool reverse = bottom_up != flip;
        const uint8_t *src = &px[(size_t)(reverse ? height - 1 - y : y) * width * comp];
        uint8_t *dst = &out[(size_t)y * width * 4];
        for (int x = 0; x < width; ++x, src += comp, dst += 4) {
            switch (comp) {
            case 1: dst[0] = dst[1] = dst[2] = src[0], dst[3] = 255; break;
            case 2: dst[0] = dst[1] = dst[2] = src[0], dst[3] = src[1]; break;
            case 3:
                if (rgb16) {
                    dst[0] = src[0], dst[1] = src[1], dst[2] = src[2];
                } else {
                    dst[0] = src[2], dst[1] = src[1], dst[2] = src[0];  // stored blue first
                }
                dst[3] = 255;
                break;
            default: dst[0] = src[2], dst[1] = src[1], dst[2] = src[0], dst[3] = src[3]; break;
            }
        }
    }
}

// ---- BMP (stb_image.h: stbi__bmp_parse_header, stbi__bmp_load): core (12-byte), info (40 / 56-byte) and V4 / V5 headers;
// 1 / 4 / 8-bit palette images, 16-bit (5-5-5 or bit fields), 24-bit, 32-bit (BGRA or bit fields); either row order; no RLE.
// stb_image's reading of the format: a 32-bit BI_RGB image whose alpha bytes are all 0 is opaque; channel masks of any width
// are expanded to 8 bits by bit replication (stbi__shiftsigned); the palette of a core-header file is counted from offset - 38.
inline bool looks_like_bmp(const uint8_t *d, size_t n)
{
    if (n < 18 || d[0] != 'B' || d[1] != 'M') {
        return false;
    }
    const uint32_t hsz = (uint32_t)d[14] | ((uint32_t)d[15] << 8) | ((uint32_t)d[16] << 16) | ((uint32_t)d[17] << 24);
    return hsz == 12 || hsz == 40 || hsz == 56 || hsz == 108 || hsz == 124;
}

inline void decode_bmp_rgba(const uint8_t *data, size_t size, const std::string &path, std::vector<uint8_t> &out, int &width, int &height, bool flip)
{
    size_t pos = 2;
    const auto get8 = [&]() -> uint32_t { return pos < size ? data[pos++] : (++pos, 0u); };  // (bytes past the end read as 0)
    const auto get16 = [&]() {
        const uint32_t lo = get8();
        return lo | (get8() << 8);
    };
    const auto get32 = [&]() {
        const uint32_t lo = get16();
        return lo | (get16() << 16);
    };
    const auto bad = [&](const char *what) { return std::runtime_error(std::string("unsupported or corrupt BMP (") + what + "): " + path); };
    get32();
    get16();
    get16();
    const int64_t offset = (int32_t)get32();
    const uint32_t hsz = get32();
    int32_t img_x, img_y;
    if (hsz == 12) {
        img_x = (int32_t)get16();
        img_y = (int32_t)get16();
    } else {
        img_x = (int32_t)get32();
        img_y = (int32_t)get32();
    }
    if (get16() != 1) {
        throw bad("planes");
    }
    const int bpp = (int)get16();
    uint32_t mr = 0, mg = 0, mb = 0, ma = 0, all_a = 255;
    if (hsz != 12) {
        const uint32_t compress = get32();
        if (compress == 1 || compress == 2) {
            throw bad("run-length encoded");
        }
        for (int k = 0; k < 5; ++k) {
            get32();  // image size, resolutions, colours used / important
        }
        if (hsz == 40 || hsz == 56) {
            for (int k = 0; hsz == 56 && k < 4; ++k) {
                get32();
            }
            if (bpp == 16 || bpp == 32) {
                if (compress == 0) {
                    if (bpp == 32) {
                        mr = 0xffu << 16, mg = 0xffu << 8, mb = 0xffu, ma = 0xffu << 24;
                        all_a = 0;  // an alpha channel that turns out to be all 0 is replaced by 255
                    } else {
                        mr = 31u << 10, mg = 31u << 5, mb = 31u;
                    }
                } else if (compress == 3) {
                    mr = get32(), mg = get32(), mb = get32();
                    if (mr == mg && mg == mb) {
                        throw bad("masks");
                    }
                } else {
                    throw bad("compression");
                }
            }
        } else {
            mr = get32(), mg = get32(), mb = get32(), ma = get32();
            for (int k = 0; k < 13 + (hsz == 124 ? 4 : 0); ++k) {
                get32();  // colour space, its parameters, (V5) intent and profile
            }
        }
    }
    const bool bottom_up = img_y > 0;
    if (img_y < 0) {
        img_y = img_y == INT32_MIN ? 0 : -img_y;
    }
    if (img_x <= 0 || img_y <= 0 || (uint64_t)img_x * (uint64_t)img_y > ((uint64_t)1 << 28)) {
        throw bad("size");
    }
    width = img_x;
    height = img_y;
    int64_t psize = 0;
    if (hsz == 12) {
        psize = bpp < 24 ? (offset - 14 - 24) / 3 : 0;
    } else if (bpp < 16) {
        psize = (offset - 14 - (int64_t)hsz) >> 2;
    }
    std::vector<uint8_t> px((size_t)img_x * img_y * 4);  // rows in file order
    size_t z = 0;
    const auto skip = [&](int64_t n) {
        pos = n < 0 ? std::max(pos, size) : pos + (size_t)n;  // (stbi__skip: a negative count goes to the end of the data)
    };
    if (bpp < 16) {
        if (psize <= 0 || psize > 256) {
            throw bad("palette size");
        }
        uint8_t pal[256][3] = {};
        for (int64_t i = 0; i < psize; ++i) {
            pal[i][2] = (uint8_t)get8();
            pal[i][1] = (uint8_t)get8();
            pal[i][0] = (uint8_t)get8();
            if (hsz != 12) {
                get8();
            }
        }
        skip(offset - 14 - (int64_t)hsz - psize * (hsz == 12 ? 3 : 4));
        int row_bytes;
        if (bpp == 1) {
            row_bytes = (img_x + 7) >> 3;
        } else if (bpp == 4) {
            row_bytes = (img_x + 1) >> 1;
        } else if (bpp == 8) {
            row_bytes = img_x;
        } else {
            throw bad("bits per pixel");
        }
        const int pad = (-row_bytes) & 3;
        for (int j = 0; j < img_y; ++j) {
            uint32_t v = 0;
            for (int i = 0; i < img_x; ++i) {
                uint32_t index;
                if (bpp == 1) {
                    v = (i & 7) == 0 ? get8() : v;
                    index = (v >> (7 - (i & 7))) & 1u;
                } else if (bpp == 4) {
                    v = (i & 1) == 0 ? get8() : v;
                    index = (i & 1) == 0 ? (v >> 4) : (v & 15u);
                } else {
                    index = get8();
                }
                px[z++] = pal[index][0], px[z++] = pal[index][1], px[z++] = pal[index][2], px[z++] = 255;
            }
            skip(pad);
        }
    } else {
        skip(offset - 14 - (int64_t)hsz);
        const int pad = bpp == 24 ? (-(3 * img_x)) & 3 : (bpp == 16 ? (-(2 * img_x)) & 3 : 0);
        const int easy = bpp == 24 ? 1 : ((bpp == 32 && mb == 0xffu && mg == 0xff00u && mr == 0x00ff0000u && ma == 0xff000000u) ? 2 : 0);
        const auto high_bit = [](uint32_t v) {
            int n = -1;
            for (; v; v >>= 1) {
                ++n;
            }
            return n;
        };
        const auto bit_count = [](uint32_t v) {
            int n = 0;
            for (; v; v &= v - 1) {
                ++n;
            }
            return n;
        };
        // an arbitrarily placed field of `bits` bits -> 8 bits, the pattern repeated to fill them (stbi__shiftsigned)
        const auto expand = [](uint32_t v, int shift, int bits) -> uint8_t {
            static const uint32_t mul_table[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01};
            static const uint32_t shift_table[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
            v = shift < 0 ? v << -shift : v >> shift;
            v >>= (8 - bits);
            return (uint8_t)((int)(v * mul_table[bits]) >> shift_table[bits]);
        };
        int rshift = 0, gshift = 0, bshift = 0, ashift = 0, rcount = 0, gcount = 0, bcount = 0, acount = 0;
        if (!easy) {
            if (!mr || !mg || !mb) {
                throw bad("masks");
            }
            rshift = high_bit(mr) - 7, rcount = bit_count(mr);
            gshift = high_bit(mg) - 7, gcount = bit_count(mg);
            bshift = high_bit(mb) - 7, bcount = bit_count(mb);
            ashift = high_bit(ma) - 7, acount = bit_count(ma);
            if (rcount > 8 || gcount > 8 || bcount > 8 || acount > 8) {
                throw bad("masks wider than 8 bits");  // (stb_image indexes past its tables here)
            }
        }
        for (int j = 0; j < img_y; ++j) {
            for (int i = 0; i < img_x; ++i) {
                uint32_t a;
                if (easy) {
                    px[z + 2] = (uint8_t)get8();
                    px[z + 1] = (uint8_t)get8();
                    px[z + 0] = (uint8_t)get8();
                    a = easy == 2 ? get8() : 255u;
                    z += 3;
                } else {
                    const uint32_t v = bpp == 16 ? get16() : get32();
                    px[z++] = expand(v & mr, rshift, rcount);
                    px[z++] = expand(v & mg, gshift, gcount);
                    px[z++] = expand(v & mb, bshift, bcount);
                    a = ma ? expand(v & ma, ashift, acount) : 255u;
                }
                all_a |= a;
                px[z++] = (uint8_t)a;
            }
            skip(pad);
        }
    }
    if (all_a == 0) {
        for (size_t i = 3; i < px.size(); i += 4) {
            px[i] = 255;
        }
    }
    out.resize(px.size());
    const size_t row = (size_t)img_x * 4;
    for (int y = 0; y < img_y; ++y) {  // stbi returns top-down rows; `flip` turns them again
        const bool reverse = bottom_up != flip;
        std::memcpy(&out[row * (size_t)y], &px[row * (size_t)(reverse ? img_y - 1 - y : y)], row);
    }
}

// stbi_load_from_memory(..., 4) for the formats textures come in: PNG, TGA and BMP (above), JPEG (jpeg_decode.h)
inline void decode_image_rgba(const uint8_t *data, size_t size, const std::string &name, std::vector<uint8_t> &out, int &width, int &height,
                       bool flip)
{
    static const uint8_t png_sig[4] = {0x89, 'P', 'N', 'G'};
    if (crt_jpeg::is_jpeg(data, size)) {
        crt_jpeg::decode_rgba(data, size, name, out, width, height, flip);
    } else if (size >= 4 && std::memcmp(data, png_sig, 4) == 0) {
        decode_png_rgba(data, size, name, out, width, height, flip);
    } else if (looks_like_bmp(data, size)) {
        decode_bmp_rgba(data, size, name, out, width, height, flip);
    } else {  // (TGA has no signature: stb_image tries it last, too)
        decode_tga_rgba(data, size, name, out, width, height, flip);
    }
}

}  // namespace crt_image
