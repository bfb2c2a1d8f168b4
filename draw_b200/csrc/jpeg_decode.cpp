// jpeg_decode.cpp — JPEG texture file -> TextureMap bytes on the host (mororo18/draw scene/mod.rs:174-202).
//
// The reference decodes textures with stb_image (crate `stb`, git branch fix-ubuntu-24.10-build of mororo18/stb,
// Cargo.toml:16 — a third-party dependency that is not in the reference tree).  The entropy-coded part of a JPEG
// stream (ITU-T T.81: Huffman, baseline or progressive) decodes to the same quantised coefficients in every
// conforming decoder; what differs between decoders, and so what fixes the texel bytes the renderer sees, is the
// arithmetic after it.  This decoder restates stb_image's published arithmetic for those three steps:
//
//   dequantisation   coefficient * table entry, truncated to 16 bits
//   inverse DCT      the 12-bit fixed-point two-pass IDCT (column pass to >> 10 with + 512, row pass to >> 17 with
//                    + 65536 + (128 << 17), constants round(x * 4096)); stb_image's SSE2 / NEON variants are
//                    documented to give bit-identical results to this scalar form
//   upsampling       h2v1: (3 * near + far + 2) >> 2 along the row; h1v2: the same between rows; h2v2: the
//                    separable 3:1 filter to >> 4 with + 8; any other ratio: pixel replication; the near / far
//                    row pairing steps through the rows the way stb_image's load loop does
//   colour           YCbCr -> RGB in 20-bit fixed point (constants round(x * 4096) << 8, the Cb term of green masked
//                    to its high 16 bits), Y offset (1 << 19) for rounding; RGB files (component ids 'R','G','B',
//                    or an Adobe APP14 marker with transform 0 and no JFIF marker) are copied
//
// Parity: UNPINNED against stb_image itself — it is not in this image.  tests/test_jpeg_cpu.py checks the decoder
// against libjpeg (through Pillow) on the reference's one JPEG asset and on files Pillow writes in every mode the
// decoder supports: the two IDCTs round differently, so the bound there is a few grey levels per texel, not equality.
//
// Supported: 8-bit baseline / extended sequential (SOF0, SOF1) and progressive (SOF2) Huffman streams, 3 components
// with any sampling factors up to 4, restart intervals.  1-component (greyscale) files have one native channel, which
// the reference rejects (`unreachable!`, scene/mod.rs:187), and so does this decoder; 4-component (CMYK) and
// arithmetic-coded files are rejected as stb_image rejects or mis-handles them.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/draw_b200.h"

namespace drawb200 {
int loader_fail(int code, const char *msg); // scene.cpp: sets draw_last_error

namespace {

struct JpegError {
    const char *what;
};
[[noreturn]] void bad(const char *what) { throw JpegError{what}; }

const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huffman { // T.81 Annex C / F.2.2.3: canonical codes by length
    bool defined = false;
    int32_t mincode[17], maxcode[18], valptr[17];
    uint8_t values[256];
    void build(const uint8_t counts[16], const uint8_t *vals, int n) {
        std::memcpy(values, vals, n);
        int code = 0, k = 0;
        for (int len = 1; len <= 16; len++) {
            valptr[len] = k;
            mincode[len] = code;
            code += counts[len - 1];
            k += counts[len - 1];
            maxcode[len] = counts[len - 1] ? code - 1 : -1;
            if (code > (1 << len)) bad("bad code lengths");
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        defined = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int hd = 0, ha = 0;        // DC / AC table of the current scan
    int dc_pred = 0;
    int x = 0, y = 0;          // size in samples
    int w2 = 0, h2 = 0;        // size padded to whole MCUs
    int bw = 0, bh = 0;        // blocks per row / column of the padded plane
    std::vector<int16_t> coeff; // progressive: bw * bh blocks of 64
    std::vector<uint8_t> plane; // w2 * h2 samples
};

struct Decoder {
    const uint8_t *p, *end;
    uint16_t dequant[4][64];
    Huffman hdc[4], hac[4];
    Component comp[3];
    int img_x = 0, img_y = 0, n_comp = 0;
    int h_max = 1, v_max = 1, mcus_x = 0, mcus_y = 0;
    bool progressive = false, jfif = false;
    int app14_transform = -1, rgb_ids = 0;
    int restart_interval = 0;
    // scan state
    int scan_n = 0, order[3];
    int spec_start = 0, spec_end = 63, succ_high = 0, succ_low = 0, eob_run = 0;
    // bit reader
    uint32_t bitbuf = 0;
    int bitcnt = 0;
    bool hit_marker = false;
    uint8_t marker = 0;

    int u8() { return p < end ? *p++ : 0; }
    int u16() {
        const int a = u8();
        return a << 8 | u8();
    }

    // ---- entropy-coded segment --------------------------------------------------------------
    void fill() {
        while (bitcnt <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p++;
                if (b == 0xff) {
                    int c = p < end ? *p++ : 0xd9;
                    while (c == 0xff) c = p < end ? *p++ : 0xd9; // fill bytes
                    if (c != 0) {                                    // a marker ends the segment: zeros from here on
                        marker = (uint8_t)c;
                        hit_marker = true;
                        b = 0;
                    }
                }
            }
            bitbuf |= (uint32_t)b << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int bits(int n) {
        if (n == 0) return 0;
        if (bitcnt < n) fill();
        const int v = (int)(bitbuf >> (32 - n));
        bitbuf <<= n;
        bitcnt -= n;
        return v;
    }
    int bit() { return bits(1); }
    int decode(const Huffman &h) {
        if (!h.defined) bad("missing Huffman table");
        if (bitcnt < 16) fill();
        int code = 0;
        for (int len = 1; len <= 16; len++) {
            code = (int)(bitbuf >> (32 - len));
            if (code <= h.maxcode[len] && h.maxcode[len] >= 0) {
                bitbuf <<= len;
                bitcnt -= len;
                return h.values[h.valptr[len] + code - h.mincode[len]];
            }
        }
        bad("bad Huffman code");
    }
    int receive_extend(int n) { // T.81 F.2.2.1
        if (n == 0) return 0;
        const int v = bits(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
    void reset_entropy() {
        bitbuf = 0;
        bitcnt = 0;
        hit_marker = false;
        marker = 0;
        for (auto &c : comp) c.dc_pred = 0;
        eob_run = 0;
    }

    // ---- blocks --------------------------------------------------------------------------------
    void block_baseline(Component &c, int16_t data[64]) {
        std::memset(data, 0, 64 * sizeof(int16_t));
        const int t = decode(hdc[c.hd]);
        if (t > 16) bad("bad DC size");
        c.dc_pred += receive_extend(t);
        data[0] = (int16_t)(c.dc_pred * dequant[c.tq][0]);
        for (int k = 1; k < 64;) {
            const int rs = decode(hac[c.ha]), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (rs != 0xf0) break;
                k += 16;
            } else {
                k += r;
                if (k > 63) bad("bad AC run");
                const int z = ZIGZAG[k++];
                data[z] = (int16_t)(receive_extend(s) * dequant[c.tq][z]);
            }
        }
    }
    void block_prog_dc(Component &c, int16_t *data) { // T.81 G.1.2.1
        if (spec_end != 0) bad("DC scan with AC coefficients");
        if (succ_high == 0) {
            std::memset(data, 0, 64 * sizeof(int16_t));
            const int t = decode(hdc[c.hd]);
            if (t > 16) bad("bad DC size");
            c.dc_pred += receive_extend(t);
            data[0] = (int16_t)(c.dc_pred * (1 << succ_low));
        } else if (bit()) {
            data[0] = (int16_t)(data[0] + (1 << succ_low));
        }
    }
    void block_prog_ac(Component &c, int16_t *data) { // T.81 G.1.2.2 / G.1.2.3
        if (spec_start == 0) bad("AC scan starting at the DC coefficient");
        const Huffman &h = hac[c.ha];
        if (succ_high == 0) {
            if (eob_run) {
                --eob_run;
                return;
            }
            for (int k = spec_start; k <= spec_end;) {
                const int rs = decode(h), r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = (1 << r) - 1;
                        if (r) eob_run += bits(r);
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    if (k > 63) bad("bad AC run");
                    data[ZIGZAG[k++]] = (int16_t)(receive_extend(s) * (1 << succ_low));
                }
            }
            return;
        }
        const int16_t plus = (int16_t)(1 << succ_low);
        auto refine = [&](int16_t &v) {
            if (bit() && (v & plus) == 0) v = (int16_t)(v > 0 ? v + plus : v - plus);
        };
        int k = spec_start;
        if (eob_run > 0) { // inside an end-of-band run: only correction bits for the coefficients already non-zero
            --eob_run;
        } else {
            while (k <= spec_end) {
                const int rs = decode(h), s = rs & 15;
                int r = rs >> 4, value = 0;
                if (s == 0) {
                    if (r < 15) { // end of band for this block and eob_run more; the rest of this block is refined below
                        eob_run = (1 << r) - 1;
                        if (r) eob_run += bits(r);
                        break;
                    }
                } else {
                    if (s != 1) bad("bad refinement code");
                    value = bit() ? plus : -plus;
                }
                while (k <= spec_end) { // skip r zero coefficients (refining the non-zero ones on the way), then place the value
                    int16_t &v = data[ZIGZAG[k++]];
                    if (v != 0) {
                        refine(v);
                    } else {
                        if (r == 0) {
                            if (value) v = (int16_t)value;
                            break;
                        }
                        --r;
                    }
                }
            }
        }
        for (; k <= spec_end; k++) {
            int16_t &v = data[ZIGZAG[k]];
            if (v != 0) refine(v);
        }
    }

    // ---- inverse DCT, stb_image's fixed-point form ------------------------------------------------------
    static int f2f(float x) { return (int)(x * 4096 + 0.5); }
    struct Idct1D {
        int t0, t1, t2, t3, x0, x1, x2, x3;
    };
    static Idct1D idct_1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
        static const int C0 = f2f(0.5411961f), C1 = f2f(-1.847759065f), C2 = f2f(0.765366865f), C3 = f2f(1.175875602f),
                         C4 = f2f(0.298631336f), C5 = f2f(2.053119869f), C6 = f2f(3.072711026f), C7 = f2f(1.501321110f),
                         C8 = f2f(-0.899976223f), C9 = f2f(-2.562915447f), C10 = f2f(-1.961570560f), C11 = f2f(-0.390180644f);
        Idct1D o;
        int p1, p2, p3, p4, p5;
        p2 = s2;
        p3 = s6;
        p1 = (p2 + p3) * C0;
        o.t2 = p1 + p3 * C1;
        o.t3 = p1 + p2 * C2;
        p2 = s0;
        p3 = s4;
        o.t0 = (p2 + p3) * 4096;
        o.t1 = (p2 - p3) * 4096;
        o.x0 = o.t0 + o.t3;
        o.x3 = o.t0 - o.t3;
        o.x1 = o.t1 + o.t2;
        o.x2 = o.t1 - o.t2;
        o.t0 = s7;
        o.t1 = s5;
        o.t2 = s3;
        o.t3 = s1;
        p3 = o.t0 + o.t2;
        p4 = o.t1 + o.t3;
        p1 = o.t0 + o.t3;
        p2 = o.t1 + o.t2;
        p5 = (p3 + p4) * C3;
        o.t0 = o.t0 * C4;
        o.t1 = o.t1 * C5;
        o.t2 = o.t2 * C6;
        o.t3 = o.t3 * C7;
        p1 = p5 + p1 * C8;
        p2 = p5 + p2 * C9;
        p3 = p3 * C10;
        p4 = p4 * C11;
        o.t3 += p1 + p4;
        o.t2 += p2 + p3;
        o.t1 += p2 + p4;
        o.t0 += p1 + p3;
        return o;
    }
    static uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : x > 255 ? 255 : x); }
    static void idct_block(uint8_t *out, int stride, const int16_t d[64]) {
        int val[64];
        for (int i = 0; i < 8; i++) { // columns
            int *v = val + i;
            const int16_t *c = d + i;
            if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
                const int dc = c[0] * 4;
                v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
            } else {
                Idct1D r = idct_1d(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]);
                r.x0 += 512;
                r.x1 += 512;
                r.x2 += 512;
                r.x3 += 512;
                v[0] = (r.x0 + r.t3) >> 10;
                v[56] = (r.x0 - r.t3) >> 10;
                v[8] = (r.x1 + r.t2) >> 10;
                v[48] = (r.x1 - r.t2) >> 10;
                v[16] = (r.x2 + r.t1) >> 10;
                v[40] = (r.x2 - r.t1) >> 10;
                v[24] = (r.x3 + r.t0) >> 10;
                v[32] = (r.x3 - r.t0) >> 10;
            }
        }
        for (int i = 0; i < 8; i++) { // rows
            const int *v = val + 8 * i;
            uint8_t *o = out + (size_t)i * stride;
            Idct1D r = idct_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            const int bias = 65536 + (128 << 17);
            r.x0 += bias;
            r.x1 += bias;
            r.x2 += bias;
            r.x3 += bias;
            o[0] = clamp8((r.x0 + r.t3) >> 17);
            o[7] = clamp8((r.x0 - r.t3) >> 17);
            o[1] = clamp8((r.x1 + r.t2) >> 17);
            o[6] = clamp8((r.x1 - r.t2) >> 17);
            o[2] = clamp8((r.x2 + r.t1) >> 17);
            o[5] = clamp8((r.x2 - r.t1) >> 17);
            o[3] = clamp8((r.x3 + r.t0) >> 17);
            o[4] = clamp8((r.x3 - r.t0) >> 17);
        }
    }

    // ---- scans ---------------------------------------------------------------------------------
    void restart_if_due(int &todo) {
        if (restart_interval == 0 || --todo > 0) return;
        if (bitcnt < 24) fill();
        if (!hit_marker || marker < 0xd0 || marker > 0xd7) {
            todo = 0x7fffffff; // no restart marker where one is due: carry on without (stb_image stops the scan; a valid file never gets here)
            return;
        }
        reset_entropy();
        todo = restart_interval;
    }
    void one_block(Component &c, int bx, int by) {
        if (bx >= c.bw || by >= c.bh) bad("block outside the image");
        if (progressive) {
            int16_t *data = c.coeff.data() + 64 * ((size_t)by * c.bw + bx);
            if (spec_start == 0)
                block_prog_dc(c, data);
            else
                block_prog_ac(c, data);
        } else {
            int16_t data[64];
            block_baseline(c, data);
            idct_block(c.plane.data() + (size_t)by * 8 * c.w2 + (size_t)bx * 8, c.w2, data);
        }
    }
    void decode_scan() {
        reset_entropy();
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        if (scan_n == 1) { // non-interleaved: the component's own blocks, row by row
            Component &c = comp[order[0]];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h; j++)
                for (int i = 0; i < w; i++) {
                    one_block(c, i, j);
                    restart_if_due(todo);
                }
        } else {
            for (int j = 0; j < mcus_y; j++)
                for (int i = 0; i < mcus_x; i++) {
                    for (int k = 0; k < scan_n; k++) {
                        Component &c = comp[order[k]];
                        for (int y = 0; y < c.v; y++)
                            for (int x = 0; x < c.h; x++) one_block(c, i * c.h + x, j * c.v + y);
                    }
                    restart_if_due(todo);
                }
        }
        // the segment ends at the next marker: hand it back to the marker loop
        if (!hit_marker) {
            while (p < end) {
                if (*p++ != 0xff) continue;
                while (p < end && *p == 0xff) ++p;
                if (p < end && *p != 0) {
                    marker = *p++;
                    hit_marker = true;
                    break;
                }
            }
        }
    }
    void finish_progressive() {
        for (int n = 0; n < n_comp; n++) {
            Component &c = comp[n];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h; j++)
                for (int i = 0; i < w; i++) {
                    int16_t *data = c.coeff.data() + 64 * ((size_t)j * c.bw + i);
                    for (int k = 0; k < 64; k++) data[k] = (int16_t)(data[k] * dequant[c.tq][k]);
                    idct_block(c.plane.data() + (size_t)j * 8 * c.w2 + (size_t)i * 8, c.w2, data);
                }
        }
    }

    // ---- headers -------------------------------------------------------------------------------
    void read_dqt(int len) {
        while (len > 0) {
            const int q = u8(), wide = q >> 4, t = q & 15;
            if (wide > 1 || t > 3) bad("bad DQT");
            for (int i = 0; i < 64; i++) dequant[t][ZIGZAG[i]] = (uint16_t)(wide ? u16() : u8());
            len -= wide ? 129 : 65;
        }
    }
    void read_dht(int len) {
        while (len > 0) {
            const int q = u8(), tc = q >> 4, th = q & 15;
            if (tc > 1 || th > 3) bad("bad DHT");
            uint8_t counts[16], vals[256];
            int n = 0;
            for (int i = 0; i < 16; i++) n += counts[i] = (uint8_t)u8();
            if (n > 256) bad("bad DHT");
            for (int i = 0; i < n; i++) vals[i] = (uint8_t)u8();
            (tc ? hac : hdc)[th].build(counts, vals, n);
            len -= 17 + n;
        }
    }
    void read_sof(int len) {
        if (len < 9) bad("bad SOF");
        if (u8() != 8) bad("only 8-bit samples");
        img_y = u16();
        img_x = u16();
        n_comp = u8();
        if (img_x == 0 || img_y == 0) bad("empty image");
        if ((unsigned long long)img_x * (unsigned long long)img_y > (1ull << 28)) bad("image too large (more than 2^28 pixels)");
        if (n_comp == 1) bad("greyscale JPEG: one native channel (the reference accepts 3 or 4, scene/mod.rs:185-189)");
        if (n_comp != 3) bad("only three-component JPEG files are decoded");
        if (len != 6 + 3 * n_comp) bad("bad SOF length");
        static const char RGB[3] = {'R', 'G', 'B'};
        for (int i = 0; i < n_comp; i++) {
            Component &c = comp[i];
            c.id = u8();
            if (c.id == RGB[i]) ++rgb_ids;
            const int q = u8();
            c.h = q >> 4;
            c.v = q & 15;
            c.tq = u8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) bad("bad component");
            h_max = c.h > h_max ? c.h : h_max;
            v_max = c.v > v_max ? c.v : v_max;
        }
        for (int i = 0; i < n_comp; i++)
            if (h_max % comp[i].h || v_max % comp[i].v) bad("fractional sampling ratio");
        mcus_x = (img_x + 8 * h_max - 1) / (8 * h_max);
        mcus_y = (img_y + 8 * v_max - 1) / (8 * v_max);
        for (int i = 0; i < n_comp; i++) {
            Component &c = comp[i];
            c.x = (img_x * c.h + h_max - 1) / h_max;
            c.y = (img_y * c.v + v_max - 1) / v_max;
            c.w2 = mcus_x * c.h * 8;
            c.h2 = mcus_y * c.v * 8;
            c.bw = c.w2 / 8;
            c.bh = c.h2 / 8;
            c.plane.assign((size_t)c.w2 * c.h2, 0);
            if (progressive) c.coeff.assign((size_t)c.bw * c.bh * 64, 0);
        }
    }
    void read_sos(int len) {
        scan_n = u8();
        if (scan_n < 1 || scan_n > n_comp || len != 4 + 2 * scan_n) bad("bad SOS");
        for (int k = 0; k < scan_n; k++) {
            const int id = u8(), q = u8();
            int which = -1;
            for (int i = 0; i < n_comp; i++)
                if (comp[i].id == id) which = i;
            if (which < 0) bad("SOS names an unknown component");
            comp[which].hd = q >> 4;
            comp[which].ha = q & 15;
            if (comp[which].hd > 3 || comp[which].ha > 3) bad("bad table selector");
            order[k] = which;
        }
        spec_start = u8();
        spec_end = u8();
        const int a = u8();
        succ_high = a >> 4;
        succ_low = a & 15;
        if (progressive) {
            if (spec_start > 63 || spec_end > 63 || spec_start > spec_end || succ_high > 13 || succ_low > 13) bad("bad SOS");
        } else {
            if (spec_start != 0 || succ_high != 0 || succ_low != 0) bad("bad SOS");
            spec_end = 63;
        }
    }

    // ---- upsampling + colour, stb_image's arithmetic ----------------------------------------------------
    static void up_h2(uint8_t *out, const uint8_t *in, int w) {
        if (w == 1) {
            out[0] = out[1] = in[0];
            return;
        }
        out[0] = in[0];
        out[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
        int i;
        for (i = 1; i < w - 1; i++) {
            const int n = 3 * in[i] + 2;
            out[i * 2 + 0] = (uint8_t)((n + in[i - 1]) >> 2);
            out[i * 2 + 1] = (uint8_t)((n + in[i + 1]) >> 2);
        }
        out[i * 2 + 0] = (uint8_t)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
        out[i * 2 + 1] = in[w - 1];
    }
    static void up_v2(uint8_t *out, const uint8_t *near, const uint8_t *far, int w) {
        for (int i = 0; i < w; i++) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
    }
    static void up_hv2(uint8_t *out, const uint8_t *near, const uint8_t *far, int w) {
        if (w == 1) {
            out[0] = out[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2);
            return;
        }
        int t0, t1 = 3 * near[0] + far[0];
        out[0] = (uint8_t)((t1 + 2) >> 2);
        for (int i = 1; i < w; i++) {
            t0 = t1;
            t1 = 3 * near[i] + far[i];
            out[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
            out[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
        }
        out[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
    }
    static int f2fixed(float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; }

    void to_rgb(uint8_t *out) {
        struct Resample {
            const uint8_t *line0, *line1;
            int hs, vs, w_lores, ystep, ypos;
            std::vector<uint8_t> buf;
        } rs[3];
        for (int k = 0; k < 3; k++) {
            Resample &r = rs[k];
            r.hs = h_max / comp[k].h;
            r.vs = v_max / comp[k].v;
            r.ystep = r.vs >> 1;
            r.w_lores = (img_x + r.hs - 1) / r.hs;
            r.ypos = 0;
            r.line0 = r.line1 = comp[k].plane.data();
            r.buf.assign((size_t)img_x + 8, 0);
        }
        const bool is_rgb = rgb_ids == 3 || (app14_transform == 0 && !jfif);
        const int CR_R = f2fixed(1.40200f), CR_G = -f2fixed(0.71414f), CB_G = -f2fixed(0.34414f), CB_B = f2fixed(1.77200f);
        for (int j = 0; j < img_y; j++) {
            const uint8_t *row[3];
            for (int k = 0; k < 3; k++) {
                Resample &r = rs[k];
                const bool bot = r.ystep >= (r.vs >> 1);
                const uint8_t *near = bot ? r.line1 : r.line0, *far = bot ? r.line0 : r.line1;
                if (r.hs == 1 && r.vs == 1) {
                    row[k] = near;
                } else {
                    if (r.hs == 1 && r.vs == 2)
                        up_v2(r.buf.data(), near, far, r.w_lores);
                    else if (r.hs == 2 && r.vs == 1)
                        up_h2(r.buf.data(), near, r.w_lores);
                    else if (r.hs == 2 && r.vs == 2)
                        up_hv2(r.buf.data(), near, far, r.w_lores);
                    else
                        for (int i = 0; i < r.w_lores; i++)
                            for (int q = 0; q < r.hs; q++)
                                if ((size_t)(i * r.hs + q) < r.buf.size()) r.buf[i * r.hs + q] = near[i];
                    row[k] = r.buf.data();
                }
                if (++r.ystep >= r.vs) {
                    r.ystep = 0;
                    r.line0 = r.line1;
                    if (++r.ypos < comp[k].y) r.line1 += comp[k].w2;
                }
            }
            uint8_t *o = out + (size_t)j * img_x * 3;
            if (is_rgb) {
                for (int i = 0; i < img_x; i++, o += 3) {
                    o[0] = row[0][i];
                    o[1] = row[1][i];
                    o[2] = row[2][i];
                }
            } else {
                for (int i = 0; i < img_x; i++, o += 3) {
                    const int y_fixed = (row[0][i] << 20) + (1 << 19);
                    const int cb = row[1][i] - 128, cr = row[2][i] - 128;
                    int r = y_fixed + cr * CR_R;
                    int g = y_fixed + cr * CR_G + (int)((uint32_t)(cb * CB_G) & 0xffff0000u);
                    int b = y_fixed + cb * CB_B;
                    r >>= 20;
                    g >>= 20;
                    b >>= 20;
                    o[0] = clamp8(r);
                    o[1] = clamp8(g);
                    o[2] = clamp8(b);
                }
            }
        }
    }

    void run(uint8_t **out_pixels, uint32_t *out_w, uint32_t *out_h, uint32_t *out_comp) {
        std::memset(dequant, 0, sizeof dequant);
        if (u8() != 0xff || u8() != 0xd8) bad("not a JPEG file");
        bool have_sof = false, done = false;
        while (!done) {
            int m;
            if (hit_marker) {
                m = marker;
                hit_marker = false;
            } else {
                if (p >= end) bad("truncated file");
                if (u8() != 0xff) continue;
                m = u8();
                while (m == 0xff) m = u8();
                if (m == 0) continue;
            }
            if (m == 0xd9) break;
            if (m >= 0xd0 && m <= 0xd7) continue; // stray restart marker
            if (m == 0x01) continue;
            const int len = u16() - 2;
            if (len < 0 || p + len > end) bad("bad segment length");
            const uint8_t *next = p + len;
            switch (m) {
            case 0xc0: case 0xc1: case 0xc2:
                if (have_sof) bad("two frame headers");
                progressive = m == 0xc2;
                read_sof(len + 2 - 2);
                have_sof = true;
                break;
            case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
                bad("lossless, hierarchical and arithmetic-coded JPEG files are not decoded");
            case 0xc4: read_dht(len); break;
            case 0xdb: read_dqt(len); break;
            case 0xdd: restart_interval = u16(); break;
            case 0xe0:
                if (len >= 5 && !std::memcmp(p, "JFIF\0", 5)) jfif = true;
                break;
            case 0xee:
                if (len >= 12 && !std::memcmp(p, "Adobe\0", 6)) app14_transform = p[11];
                break;
            case 0xda:
                if (!have_sof) bad("scan before the frame header");
                read_sos(len);
                p = next;
                decode_scan();
                next = p;
                break;
            default: break;
            }
            p = next;
        }
        if (!have_sof) bad("no frame header");
        if (progressive) finish_progressive();
        uint8_t *px = static_cast<uint8_t *>(std::malloc((size_t)img_x * img_y * 3));
        if (!px) throw std::bad_alloc();
        to_rgb(px);
        *out_pixels = px;
        *out_w = (uint32_t)img_x;
        *out_h = (uint32_t)img_y;
        *out_comp = 3;
    }
};

} // namespace

// Called by draw_image_load (image_decode.cpp) for files that start with the SOI marker.
int decode_jpeg(const std::vector<uint8_t> &file, uint8_t **out_pixels, uint32_t *out_w, uint32_t *out_h, uint32_t *out_comp) {
    Decoder d;
    d.p = file.data();
    d.end = file.data() + file.size();
    try {
        d.run(out_pixels, out_w, out_h, out_comp);
    } catch (const JpegError &e) {
        return loader_fail(DRAW_ERR_INVALID_ARGUMENT, (std::string("JPEG: ") + e.what).c_str());
    }
    return DRAW_OK;
}

} // namespace drawb200
