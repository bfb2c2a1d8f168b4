// image_decode.cpp — texture file -> TextureMap bytes on the host (mororo18/draw scene/mod.rs:174-202).
//
// The reference decodes texture files with stb_image (`stbi_info_from_reader` + `stbi_load_from_reader`
// with the file's own channel count, which must be 3 or 4, :185-189) and keeps the bytes as they come:
// width * height * components, row 0 = top of the image.  This file is the library's own decoder for the
// lossless format among the reference's assets, PNG (models/lemur/lemurT.png): any correct PNG decoder
// produces the same bytes as stb_image, so parity does not depend on whose it is.  JPEG is lossy and
// decoders differ in their IDCT / upsampling arithmetic: jpeg_decode.cpp restates stb_image's.
//
// PNG (ISO/IEC 15948): signature, IHDR, [PLTE], [tRNS], IDAT..., IEND; the concatenated IDAT payload is a
// zlib stream (inflated with zlib) of filtered scanlines; filters None / Sub / Up / Average / Paeth.
// Supported: colour types 2 (RGB), 6 (RGBA) at 8 or 16 bits per sample (16 -> the high byte, like
// stb_image's 8-bit API), colour type 3 (palette, 1/2/4/8 bits; with tRNS -> 4 components), non-interlaced.
// Greyscale files (types 0, 4) have 1 or 2 native components, which the reference rejects
// (`unreachable!`, :187), and so does this decoder.  CRCs are not checked (stb_image does not either).
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/draw_b200.h"

namespace drawb200 {
int loader_fail(int code, const char *msg); // scene.cpp: sets draw_last_error
int decode_jpeg(const std::vector<uint8_t> &file, uint8_t **out_pixels, uint32_t *out_w, uint32_t *out_h, uint32_t *out_comp); // jpeg_decode.cpp
}

namespace {

using drawb200::loader_fail;

uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool read_file(const char *path, std::vector<uint8_t> &out) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) {
        std::fclose(f);
        return false;
    }
    out.resize((size_t)n);
    const size_t got = n ? std::fread(out.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

int decode_png(const std::vector<uint8_t> &file, uint8_t **out_pixels, uint32_t *out_w, uint32_t *out_h, uint32_t *out_comp) {
    static const uint8_t SIG[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 + 25 || std::memcmp(file.data(), SIG, 8) != 0) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "not a PNG file");
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    size_t pos = 8;
    bool seen_end = false;
    while (!seen_end && pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t *type = &file[pos + 4], *data = &file[pos + 8];
        if (len > file.size() - pos - 12) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: truncated chunk");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: bad IHDR");
            w = be32(data); h = be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            if (data[10] != 0 || data[11] != 0) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: unknown compression or filter method");
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            seen_end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (ctype < 0 || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24)) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: missing or bad IHDR");
    if (interlace) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: interlaced files are not supported");
    int samples = 0; // per pixel in the file
    if (ctype == 2 && (depth == 8 || depth == 16)) samples = 3;
    else if (ctype == 6 && (depth == 8 || depth == 16)) samples = 4;
    else if (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) samples = 1;
    else if (ctype == 0 || ctype == 4)
        return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: greyscale image (the reference accepts 3 or 4 components only, scene/mod.rs:185-189)");
    else return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: unsupported colour type / bit depth");
    if (ctype == 3 && (palette.empty() || palette.size() % 3)) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: palette image without PLTE");

    const size_t bits_pp = (size_t)samples * depth, bpp = (bits_pp + 7) / 8; // filter unit: whole bytes per pixel, at least 1
    const size_t stride = ((size_t)w * bits_pp + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    {
        z_stream zs{};
        if (inflateInit(&zs) != Z_OK) return loader_fail(DRAW_ERR_INTERNAL, "PNG: inflateInit failed");
        zs.next_in = idat.data();
        zs.avail_in = (uInt)idat.size();
        zs.next_out = raw.data();
        zs.avail_out = (uInt)raw.size();
        const int rc = inflate(&zs, Z_FINISH);
        const size_t produced = raw.size() - zs.avail_out;
        inflateEnd(&zs);
        if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || produced != raw.size())
            return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: corrupt or short image data");
    }
    // unfilter in place (row r: filter byte, then `stride` bytes)
    for (uint32_t y = 0; y < h; y++) {
        uint8_t *row = &raw[(stride + 1) * (size_t)y + 1];
        const uint8_t *up = y ? row - (stride + 1) : nullptr;
        const int filter = row[-1];
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? row[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int add;
            switch (filter) {
                case 0: add = 0; break;
                case 1: add = a; break;
                case 2: add = b; break;
                case 3: add = (a + b) >> 1; break;
                case 4: add = paeth(a, b, c); break;
                default: return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "PNG: unknown scanline filter");
            }
            row[i] = (uint8_t)(row[i] + add);
        }
    }
    // colour type 2 with a tRNS colour key: stb_image reports 4 components, alpha 0 where the pixel equals the key
    const bool color_key = ctype == 2 && trns.size() >= 6;
    const uint32_t comp = ctype == 3 ? (trns.empty() ? 3u : 4u) : (color_key ? 4u : (uint32_t)samples);
    uint8_t *px = static_cast<uint8_t *>(std::malloc((size_t)w * h * comp));
    if (!px) return loader_fail(DRAW_ERR_OUT_OF_MEMORY, "PNG: out of memory");
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *row = &raw[(stride + 1) * (size_t)y + 1];
        uint8_t *dst = px + (size_t)y * w * comp;
        if (ctype == 3) {
            for (uint32_t x = 0; x < w; x++) {
                const size_t bit = (size_t)x * depth;
                const uint32_t idx = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
                const bool known = (size_t)idx * 3 + 2 < palette.size();
                for (int c = 0; c < 3; c++) dst[x * comp + c] = known ? palette[idx * 3 + c] : 0;
                if (comp == 4) dst[x * 4 + 3] = idx < trns.size() ? trns[idx] : 255;
            }
        } else if (color_key) {
            const size_t bytes = depth / 8; // per sample; the key holds 16-bit values, 8-bit files use their low byte
            for (uint32_t x = 0; x < w; x++) {
                bool match = true;
                for (int c = 0; c < 3; c++) {
                    const uint8_t *sp = row + ((size_t)x * 3 + c) * bytes;
                    const uint32_t v = bytes == 2 ? (uint32_t)sp[0] << 8 | sp[1] : sp[0];
                    const uint32_t k = bytes == 2 ? (uint32_t)trns[2 * c] << 8 | trns[2 * c + 1] : trns[2 * c + 1];
                    match = match && v == k;
                    dst[x * 4 + c] = sp[0];
                }
                dst[x * 4 + 3] = match ? 0 : 255;
            }
        } else if (depth == 8) {
            std::memcpy(dst, row, (size_t)w * comp);
        } else { // 16 bits per sample, big-endian: keep the high byte
            for (size_t i = 0; i < (size_t)w * comp; i++) dst[i] = row[2 * i];
        }
    }
    *out_pixels = px;
    *out_w = w;
    *out_h = h;
    *out_comp = comp;
    return DRAW_OK;
}

} // namespace

extern "C" {

int draw_image_load(const char *path, uint8_t **out_pixels, uint32_t *out_w, uint32_t *out_h, uint32_t *out_components) {
    try {
        if (!path || !out_pixels || !out_w || !out_h || !out_components) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "NULL argument");
        *out_pixels = nullptr;
        std::vector<uint8_t> file;
        if (!read_file(path, file)) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, (std::string("cannot read ") + path).c_str());
        if (file.size() >= 8 && file[0] == 0x89 && file[1] == 'P') return decode_png(file, out_pixels, out_w, out_h, out_components);
        if (file.size() >= 4 && file[0] == 0xff && file[1] == 0xd8) return drawb200::decode_jpeg(file, out_pixels, out_w, out_h, out_components);
        return loader_fail(DRAW_ERR_INVALID_ARGUMENT,
                           (std::string(path) + ": only PNG and JPEG are decoded by the library; pass a draw_image_loader for other formats").c_str());
    } catch (const std::bad_alloc &) {
        return loader_fail(DRAW_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (...) {
        return loader_fail(DRAW_ERR_INTERNAL, "internal error");
    }
}

void draw_image_free(uint8_t *pixels) { std::free(pixels); }

// stbi_write_png's job in the reference (app/mod.rs:362-378): width*height*components bytes -> a PNG file.
// Lossless, so any conforming encoder yields a file that decodes to the same pixels; this one uses zlib's
// deflate and the Up filter on every row but the first (cheap, and effective on rendered frames whose
// background is constant).  components: 3 (RGB) or 4 (RGBA), 8 bits per sample.
int draw_image_write_png(const char *path, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t components) {
    try {
        if (!path || !pixels || !width || !height || (components != 3 && components != 4))
            return loader_fail(DRAW_ERR_INVALID_ARGUMENT, "draw_image_write_png: bad argument (components must be 3 or 4)");
        const size_t stride = (size_t)width * components;
        std::vector<uint8_t> raw((stride + 1) * (size_t)height);
        for (uint32_t y = 0; y < height; y++) {
            uint8_t *row = &raw[(stride + 1) * (size_t)y];
            const uint8_t *src = pixels + stride * y, *up = y ? src - stride : nullptr;
            row[0] = up ? 2 : 0;
            for (size_t i = 0; i < stride; i++) row[1 + i] = up ? (uint8_t)(src[i] - up[i]) : src[i];
        }
        uLongf zlen = compressBound((uLong)raw.size());
        std::vector<uint8_t> z(zlen);
        if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return loader_fail(DRAW_ERR_INTERNAL, "PNG: deflate failed");
        FILE *f = std::fopen(path, "wb");
        if (!f) return loader_fail(DRAW_ERR_INVALID_ARGUMENT, (std::string("cannot write ") + path).c_str());
        auto put_chunk = [&](const char *type, const uint8_t *data, size_t len) {
            uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                              (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
            uLong crc = crc32(0L, hdr + 4, 4);
            if (len) crc = crc32(crc, data, (uInt)len);
            const uint8_t tail[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
            return std::fwrite(hdr, 1, 8, f) == 8 && (!len || std::fwrite(data, 1, len, f) == len) && std::fwrite(tail, 1, 4, f) == 4;
        };
        static const uint8_t SIG[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        const uint8_t ihdr[13] = {(uint8_t)(width >> 24), (uint8_t)(width >> 16), (uint8_t)(width >> 8), (uint8_t)width,
                                  (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height,
                                  8, (uint8_t)(components == 3 ? 2 : 6), 0, 0, 0};
        bool ok = std::fwrite(SIG, 1, 8, f) == 8 && put_chunk("IHDR", ihdr, 13) && put_chunk("IDAT", z.data(), zlen) && put_chunk("IEND", nullptr, 0);
        ok = (std::fclose(f) == 0) && ok;
        return ok ? DRAW_OK : loader_fail(DRAW_ERR_INTERNAL, "PNG: short write");
    } catch (const std::bad_alloc &) {
        return loader_fail(DRAW_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (...) {
        return loader_fail(DRAW_ERR_INTERNAL, "internal error");
    }
}

int draw_image_loader_builtin(const char *path, void *, uint8_t **out_pixels, uint32_t *out_w, uint32_t *out_h,
                              uint32_t *out_components) {
    return draw_image_load(path, out_pixels, out_w, out_h, out_components) == DRAW_OK ? 0 : 1;
}

} // extern "C"
