// jpeg_encode.cpp — frame bytes -> JPEG file (mororo18/draw src/app/mod.rs:316-379, ImgFileFormat::Jpeg).
//
// The reference hands the byte-swapped frame (R,G,B,pad per pixel) to stb_image_write's stbi_write_jpg with
// components = 4 and, in the quality slot, the value it passes as the PNG row stride: width * 4 (app/mod.rs:371-377).
// stbi_write_jpg clamps quality to 1..100, so every export is a quality-100 file: all quantisers 1, no chroma
// subsampling (the writer subsamples only at quality <= 90), the alpha byte ignored.  draw_image_write_jpg restates
// that writer's published scheme — JFIF header, the two Annex K quantisation tables scaled and clamped, one baseline
// frame (SOF0) with the Annex K Huffman tables, RGB -> YCbCr in float, an 8x8 AAN float DCT per block with edge
// pixels replicated, round-half-away quantisation, 1-bit fill at the end — and takes the same quality argument.
//
// Parity: UNPINNED at the byte level (stb_image_write is not in this image, and a float DCT's last bit depends on how
// its compiler contracts multiplies and adds); pinned at the pixel level: tests/test_jpeg_cpu.py decodes the file
// with libjpeg and with the library's own decoder and bounds the difference to the frame.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/draw_b200.h"

namespace drawb200 {
int loader_fail(int code, const char *msg); // scene.cpp: sets draw_last_error
}

namespace {

using drawb200::loader_fail;

const uint8_t ZIGZAG[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                            41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                            46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63}; // natural -> zigzag position

// ITU-T T.81 Annex K.1 quantisation tables (natural order) and K.3 Huffman tables
const int Q_LUMA[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,
                        69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55, 64,
                        81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95,  98,  112, 100, 103, 99};
const int Q_CHROMA[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                          99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                          99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const uint8_t DC_LUMA_COUNTS[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t DC_CHROMA_COUNTS[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t DC_VALUES[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t AC_LUMA_COUNTS[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t AC_LUMA_VALUES[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
    0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18,
    0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
    0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5,
    0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t AC_CHROMA_COUNTS[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t AC_CHROMA_VALUES[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
    0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25,
    0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
    0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
    0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
    0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4,
    0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

struct Code {
    uint16_t bits, len;
};
void build_codes(const uint8_t counts[16], const uint8_t *values, Code out[256]) { // T.81 Annex C
    std::memset(out, 0, 256 * sizeof(Code));
    int code = 0, k = 0;
    for (int len = 1; len <= 16; len++) {
        for (int i = 0; i < counts[len - 1]; i++) out[values[k++]] = {(uint16_t)code++, (uint16_t)len};
        code <<= 1;
    }
}

struct BitWriter {
    std::vector<uint8_t> &out;
    uint32_t buf = 0;
    int cnt = 0;
    void put(uint32_t bits, int len) {
        cnt += len;
        buf |= bits << (24 - cnt);
        while (cnt >= 8) {
            const uint8_t c = (uint8_t)(buf >> 16);
            out.push_back(c);
            if (c == 0xff) out.push_back(0); // byte stuffing
            buf <<= 8;
            cnt -= 8;
        }
    }
};

// One pass of the AAN forward DCT over eight values with stride s (Arai, Agui, Nakajima; the scale factors go into the
// quantisation reciprocals), in float as stb_image_write does it.
void dct8(float *d, int s) {
    const float d0 = d[0], d1 = d[s], d2 = d[2 * s], d3 = d[3 * s], d4 = d[4 * s], d5 = d[5 * s], d6 = d[6 * s], d7 = d[7 * s];
    const float t0 = d0 + d7, t7 = d0 - d7, t1 = d1 + d6, t6 = d1 - d6, t2 = d2 + d5, t5 = d2 - d5, t3 = d3 + d4, t4 = d3 - d4;
    float t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    d[0] = t10 + t11;
    d[4 * s] = t10 - t11;
    const float z1 = (t12 + t13) * 0.707106781f;
    d[2 * s] = t13 + z1;
    d[6 * s] = t13 - z1;
    t10 = t4 + t5;
    t11 = t5 + t6;
    t12 = t6 + t7;
    const float z5 = (t10 - t12) * 0.382683433f;
    const float z2 = t10 * 0.541196100f + z5, z4 = t12 * 1.306562965f + z5, z3 = t11 * 0.707106781f;
    const float z11 = t7 + z3, z13 = t7 - z3;
    d[5 * s] = z13 + z2;
    d[3 * s] = z13 - z2;
    d[s] = z11 + z4;
    d[7 * s] = z11 - z4;
}

int magnitude_bits(int v, uint32_t *bits) { // T.81 F.1.2.1: size category and the low bits of v (v - 1 when negative)
    int a = v < 0 ? -v : v, n = 0;
    while (a) {
        ++n;
        a >>= 1;
    }
    *bits = (uint32_t)(v < 0 ? v - 1 : v) & ((1u << n) - 1u);
    return n;
}

int encode_block(BitWriter &bw, float *block, const float *recip, int dc_pred, const Code *dc, const Code *ac) {
    for (int r = 0; r < 8; r++) dct8(block + 8 * r, 1);
    for (int c = 0; c < 8; c++) dct8(block + c, 8);
    int zz[64];
    for (int i = 0; i < 64; i++) {
        const float v = block[i] * recip[i];
        int q = (int)(v < 0 ? v - 0.5f : v + 0.5f);
        if (i) q = q < -1023 ? -1023 : q > 1023 ? 1023 : q; // the largest AC size category of an 8-bit stream is 10
        zz[ZIGZAG[i]] = q;
    }
    uint32_t bits;
    const int diff = zz[0] - dc_pred;
    int n = magnitude_bits(diff, &bits);
    bw.put(dc[n].bits, dc[n].len);
    if (n) bw.put(bits, n);
    int last = 63;
    while (last > 0 && zz[last] == 0) --last;
    for (int i = 1; i <= last; i++) {
        int run = 0;
        while (zz[i] == 0) {
            ++run;
            ++i;
        }
        for (; run >= 16; run -= 16) bw.put(ac[0xf0].bits, ac[0xf0].len);
        n = magnitude_bits(zz[i], &bits);
        bw.put(ac[run << 4 | n].bits, ac[run << 4 | n].len);
        bw.put(bits, n);
    }
    if (last != 63) bw.put(ac[0].bits, ac[0].len); // end of block
    return zz[0];
}

} // namespace

extern "C" int draw_image_write_jpg(const char *path, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t components,
                                    int quality) {
    try {
        if (!path || !pixels || !width || !height || width > 65535 || height > 65535 || (components != 3 && components != 4))
            return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "draw_image_write_jpg: bad argument (components 3 or 4, sides up to 65535)");
        // stbi_write_jpg's quality handling: 0 -> 90; subsampling at <= 90 only; clamp to 1..100; the IJG scale
        quality = quality ? quality : 90;
        const bool subsample = quality <= 90;
        quality = quality < 1 ? 1 : quality > 100 ? 100 : quality;
        const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
        if (subsample)
            return loader_fail(DRAW_ERR_INVALID_ARGUMENT,
                               "draw_image_write_jpg: quality <= 90 selects 4:2:0, which the reference's export never reaches (it passes width * 4)");
        uint8_t qt[2][64]; // natural order
        float recip[2][64];
        static const float AASF[8] = {1.0f * 2.828427125f,         1.387039845f * 2.828427125f, 1.306562965f * 2.828427125f,
                                      1.175875602f * 2.828427125f, 1.0f * 2.828427125f,         0.785694958f * 2.828427125f,
                                      0.541196100f * 2.828427125f, 0.275899379f * 2.828427125f};
        for (int i = 0; i < 64; i++) {
            const int y = (Q_LUMA[i] * scale + 50) / 100, c = (Q_CHROMA[i] * scale + 50) / 100;
            qt[0][i] = (uint8_t)(y < 1 ? 1 : y > 255 ? 255 : y);
            qt[1][i] = (uint8_t)(c < 1 ? 1 : c > 255 ? 255 : c);
            recip[0][i] = 1.0f / (qt[0][i] * AASF[i >> 3] * AASF[i & 7]);
            recip[1][i] = 1.0f / (qt[1][i] * AASF[i >> 3] * AASF[i & 7]);
        }
        std::vector<uint8_t> out;
        out.reserve((size_t)width * height);
        auto put = [&](std::initializer_list<int> bytes) {
            for (int b : bytes) out.push_back((uint8_t)b);
        };
        put({0xff, 0xd8, 0xff, 0xe0, 0, 0x10, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
        put({0xff, 0xdb, 0, 0x84, 0});
        uint8_t zz[64];
        for (int i = 0; i < 64; i++) zz[ZIGZAG[i]] = qt[0][i];
        out.insert(out.end(), zz, zz + 64);
        out.push_back(1);
        for (int i = 0; i < 64; i++) zz[ZIGZAG[i]] = qt[1][i];
        out.insert(out.end(), zz, zz + 64);
        put({0xff, 0xc0, 0, 0x11, 8, (int)(height >> 8), (int)(height & 255), (int)(width >> 8), (int)(width & 255), 3, 1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1});
        put({0xff, 0xc4, 0x01, 0xa2, 0x00});
        out.insert(out.end(), DC_LUMA_COUNTS, DC_LUMA_COUNTS + 16);
        out.insert(out.end(), DC_VALUES, DC_VALUES + 12);
        out.push_back(0x10);
        out.insert(out.end(), AC_LUMA_COUNTS, AC_LUMA_COUNTS + 16);
        out.insert(out.end(), AC_LUMA_VALUES, AC_LUMA_VALUES + 162);
        out.push_back(0x01);
        out.insert(out.end(), DC_CHROMA_COUNTS, DC_CHROMA_COUNTS + 16);
        out.insert(out.end(), DC_VALUES, DC_VALUES + 12);
        out.push_back(0x11);
        out.insert(out.end(), AC_CHROMA_COUNTS, AC_CHROMA_COUNTS + 16);
        out.insert(out.end(), AC_CHROMA_VALUES, AC_CHROMA_VALUES + 162);
        put({0xff, 0xda, 0, 0x0c, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 0x3f, 0});

        Code dc_y[256], ac_y[256], dc_c[256], ac_c[256];
        build_codes(DC_LUMA_COUNTS, DC_VALUES, dc_y);
        build_codes(AC_LUMA_COUNTS, AC_LUMA_VALUES, ac_y);
        build_codes(DC_CHROMA_COUNTS, DC_VALUES, dc_c);
        build_codes(AC_CHROMA_COUNTS, AC_CHROMA_VALUES, ac_c);
        BitWriter bw{out};
        int pred_y = 0, pred_u = 0, pred_v = 0;
        float Y[64], U[64], V[64];
        for (uint32_t by = 0; by < height; by += 8)
            for (uint32_t bx = 0; bx < width; bx += 8) {
                for (uint32_t r = 0; r < 8; r++) {
                    const uint32_t yy = by + r < height ? by + r : height - 1; // edge pixels replicated
                    for (uint32_t c = 0; c < 8; c++) {
                        const uint32_t xx = bx + c < width ? bx + c : width - 1;
                        const uint8_t *px = pixels + ((size_t)yy * width + xx) * components;
                        const float R = px[0], G = px[1], B = px[2];
                        Y[r * 8 + c] = 0.29900f * R + 0.58700f * G + 0.11400f * B - 128.0f;
                        U[r * 8 + c] = -0.16874f * R - 0.33126f * G + 0.50000f * B;
                        V[r * 8 + c] = 0.50000f * R - 0.41869f * G - 0.08131f * B;
                    }
                }
                pred_y = encode_block(bw, Y, recip[0], pred_y, dc_y, ac_y);
                pred_u = encode_block(bw, U, recip[1], pred_u, dc_c, ac_c);
                pred_v = encode_block(bw, V, recip[1], pred_v, dc_c, ac_c);
            }
        bw.put(0x7f, 7); // fill the last byte with ones
        put({0xff, 0xd9});
        FILE *f = std::fopen(path, "wb");
        if (!f) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, (std::string("cannot write ") + path).c_str());
        bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
        ok = (std::fclose(f) == 0) && ok;
        return ok ? DRAW_OK : loader_fail(DRAW_ERR_INTERNAL, "JPEG: short write");
    } catch (const std::bad_alloc &) {
        return loader_fail(DRAW_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (...) {
        return loader_fail(DRAW_ERR_INTERNAL, "internal error");
    }
}
