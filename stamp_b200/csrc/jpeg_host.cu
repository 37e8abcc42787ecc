// Host half of the tile-cache JPEG decode: marker parsing and Huffman (entropy) decoding of baseline JPEG into
// quantised DCT coefficients.  Plain C++ (no CUDA calls): this is the inherently sequential part of the format; the
// arithmetic (dequantisation, inverse DCT, chroma up-sampling, colour conversion) runs on the GPU in jpeg.cu.
//
// reference call site: _tiles_from_cache_file, src/stamp/preprocessing/tiling.py:380-406 -- ``Image.open(tile_fp)``
// + ``img.load()`` for every cached tile (written by Pillow: baseline sequential DCT, one interleaved scan, 4:2:0
// or 4:4:4).  Format: ITU-T T.81 (Annex B markers, F.2.2 Huffman decoding); the coefficient order handed to the GPU is
// the natural (row-major) order, de-zigzagged here.
#include <cstdint>
#include <cstring>

#include "stamp_b200.h"

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int LOOK = 9;   // bits of the first-level lookup

struct HuffTable {
    bool present = false;
    uint16_t fast[1 << LOOK];      // (length << 8) | symbol for codes of <= LOOK bits, 0 = longer code
    int32_t maxcode[18];           // largest code of each length (-1: none), T.81 F.2.2.3
    int32_t valoff[17];            // symbol index = code + valoff[length]
    uint8_t symbols[256];
};

bool build_table(const uint8_t* counts, const uint8_t* symbols, int n, HuffTable* t) {
    std::memset(t->fast, 0, sizeof(t->fast));
    std::memcpy(t->symbols, symbols, n);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        t->valoff[len] = k - code;
        const int c = counts[len - 1];
        if (c) {
            if (code + c > (1 << len)) return false;
            if (len <= LOOK) {
                for (int i = 0; i < c; ++i) {
                    const int first = (code + i) << (LOOK - len);
                    for (int j = 0; j < (1 << (LOOK - len)); ++j)
                        t->fast[first + j] = static_cast<uint16_t>((len << 8) | symbols[k + i]);
                }
            }
            code += c;
            k += c;
            t->maxcode[len] = code - 1;
        } else {
            t->maxcode[len] = -1;
        }
        code <<= 1;
    }
    t->maxcode[17] = 0x7fffffff;
    t->present = true;
    return true;
}

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;
    int n = 0;
    bool marker = false;   // a marker was met: zeros are fed from here on (T.81 F.2.2.5)
    bool eof = false;      // ... or the data ended without one (truncated file)

    inline void fill() {
        while (n <= 56) {
            uint32_t b = 0;
            if (!marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    if (p < end && *p == 0) {
                        ++p;
                    } else {
                        --p;
                        marker = true;
                        b = 0;
                    }
                }
            } else {
                if (!marker) eof = true;
                marker = true;
            }
            acc = (acc << 8) | b;
            n += 8;
        }
    }
    inline uint32_t peek(int k) { return static_cast<uint32_t>((acc >> (n - k)) & ((1u << k) - 1)); }
    inline void skip(int k) { n -= k; }
    inline uint32_t get(int k) {
        const uint32_t v = peek(k);
        n -= k;
        return v;
    }
};

inline int decode_symbol(BitReader& br, const HuffTable& t) {
    if (br.n < 16) br.fill();
    const uint16_t f = t.fast[br.peek(LOOK)];
    if (f) {
        br.skip(f >> 8);
        return f & 0xFF;
    }
    int len = LOOK + 1;
    int32_t code = static_cast<int32_t>(br.peek(len));
    while (code > t.maxcode[len]) {
        if (++len > 16) return -1;   // no code of any length matches: corrupt data
        code = static_cast<int32_t>(br.peek(len));
    }
    br.skip(len);
    return t.symbols[(code + t.valoff[len]) & 0xFF];
}

inline int extend(uint32_t v, int s) { return v < (1u << (s - 1)) ? static_cast<int>(v) - (1 << s) + 1 : static_cast<int>(v); }

struct Parsed {
    StampJpegInfo info;
    HuffTable dc[4], ac[4];
    uint16_t qt[4][64];   // natural order
    bool qt_present[4] = {false, false, false, false};
    int tq[3], td[3], ta[3];
    int restart_interval = 0;
    const uint8_t* scan = nullptr;
};

inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

int parse_headers(const uint8_t* data, size_t n, Parsed* P) {
    if (data == nullptr || n < 4 || data[0] != 0xFF || data[1] != 0xD8) return STAMP_ERR_BAD_ARG;
    size_t p = 2;
    bool have_frame = false;
    int comp_id[3] = {0, 0, 0};
    std::memset(&P->info, 0, sizeof(P->info));
    while (p + 4 <= n) {
        if (data[p] != 0xFF) return STAMP_ERR_BAD_ARG;
        const int marker = data[p + 1];
        p += 2;
        if (marker == 0xFF) { --p; continue; }   // fill byte
        if (marker == 0xD8 || marker == 0x01 || (marker >= 0xD0 && marker <= 0xD7)) continue;
        if (marker == 0xD9) return STAMP_ERR_BAD_ARG;
        const int len = be16(data + p);
        if (len < 2 || p + len > n) return STAMP_ERR_BAD_ARG;
        const uint8_t* seg = data + p + 2;
        const int slen = len - 2;
        switch (marker) {
        case 0xDB: {
            int q = 0;
            while (q < slen) {
                const int pq = seg[q] >> 4, t = seg[q] & 15;
                if (t > 3 || q + 1 + (pq ? 128 : 64) > slen) return STAMP_ERR_BAD_ARG;
                for (int i = 0; i < 64; ++i)
                    P->qt[t][kZigzag[i]] = pq ? static_cast<uint16_t>(be16(seg + q + 1 + 2 * i)) : seg[q + 1 + i];
                P->qt_present[t] = true;
                q += 1 + (pq ? 128 : 64);
            }
            break;
        }
        case 0xC0:
        case 0xC1: {
            if (slen < 6 || seg[0] != 8) return STAMP_ERR_UNSUPPORTED;   // 8-bit samples only
            P->info.height = be16(seg + 1);
            P->info.width = be16(seg + 3);
            P->info.n_comp = seg[5];
            if (P->info.n_comp != 3 || slen < 6 + 9 || P->info.height <= 0 || P->info.width <= 0) return STAMP_ERR_UNSUPPORTED;
            for (int i = 0; i < 3; ++i) {
                comp_id[i] = seg[6 + 3 * i];
                P->info.h[i] = seg[7 + 3 * i] >> 4;
                P->info.v[i] = seg[7 + 3 * i] & 15;
                P->tq[i] = seg[8 + 3 * i];
                if (P->tq[i] > 3) return STAMP_ERR_BAD_ARG;
            }
            const bool s420 = P->info.h[0] == 2 && P->info.v[0] == 2;
            const bool s444 = P->info.h[0] == 1 && P->info.v[0] == 1;
            if (!(s420 || s444) || P->info.h[1] != 1 || P->info.v[1] != 1 || P->info.h[2] != 1 || P->info.v[2] != 1)
                return STAMP_ERR_UNSUPPORTED;   // 4:2:0 and 4:4:4 (what Pillow writes for quality <= 100 / subsampling=0)
            const int mx = 8 * P->info.h[0], my = 8 * P->info.v[0];
            P->info.mcus_x = (P->info.width + mx - 1) / mx;
            P->info.mcus_y = (P->info.height + my - 1) / my;
            have_frame = true;
            break;
        }
        case 0xC2: case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE:
        case 0xCF:
            return STAMP_ERR_UNSUPPORTED;   // progressive / lossless / arithmetic coding
        case 0xC4: {
            int q = 0;
            while (q < slen) {
                if (q + 17 > slen) return STAMP_ERR_BAD_ARG;
                const int tc = seg[q] >> 4, th = seg[q] & 15;
                int cnt = 0;
                for (int i = 0; i < 16; ++i) cnt += seg[q + 1 + i];
                if (tc > 1 || th > 3 || cnt > 256 || q + 17 + cnt > slen) return STAMP_ERR_BAD_ARG;
                if (!build_table(seg + q + 1, seg + q + 17, cnt, tc ? &P->ac[th] : &P->dc[th])) return STAMP_ERR_BAD_ARG;
                q += 17 + cnt;
            }
            break;
        }
        case 0xDD:
            if (slen < 2) return STAMP_ERR_BAD_ARG;
            P->restart_interval = be16(seg);
            break;
        case 0xDA: {
            if (!have_frame || slen < 1 + 2 * 3 + 3 || seg[0] != 3) return STAMP_ERR_UNSUPPORTED;   // one interleaved scan
            for (int i = 0; i < 3; ++i) {
                if (seg[1 + 2 * i] != comp_id[i]) return STAMP_ERR_UNSUPPORTED;
                P->td[i] = seg[2 + 2 * i] >> 4;
                P->ta[i] = seg[2 + 2 * i] & 15;
                if (P->td[i] > 3 || P->ta[i] > 3 || !P->dc[P->td[i]].present || !P->ac[P->ta[i]].present ||
                    !P->qt_present[P->tq[i]])
                    return STAMP_ERR_BAD_ARG;
            }
            P->scan = data + p + len;
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 64; ++k) P->info.quant[i][k] = P->qt[P->tq[i]][k];
            return STAMP_OK;
        }
        default:
            break;   // APPn, COM, ...
        }
        p += len;
    }
    return STAMP_ERR_BAD_ARG;
}

}  // namespace

extern "C" int stamp_jpeg_read_header(const uint8_t* data, size_t n, StampJpegInfo* info) {
    if (info == nullptr) return STAMP_ERR_BAD_ARG;
    Parsed P;
    const int rc = parse_headers(data, n, &P);
    if (rc == STAMP_OK) *info = P.info;
    return rc;
}

extern "C" size_t stamp_jpeg_coef_count(const StampJpegInfo* info) {
    if (info == nullptr || info->mcus_x <= 0 || info->mcus_y <= 0) return 0;
    size_t blocks = 0;
    for (int i = 0; i < 3; ++i) blocks += static_cast<size_t>(info->mcus_x) * info->h[i] * info->mcus_y * info->v[i];
    return blocks * 64;
}

extern "C" int stamp_jpeg_entropy_decode(const uint8_t* data, size_t n, const StampJpegInfo* expect, int16_t* coef,
                                         uint16_t* quant) {
    if (coef == nullptr || quant == nullptr) return STAMP_ERR_BAD_ARG;
    Parsed P;
    int rc = parse_headers(data, n, &P);
    if (rc != STAMP_OK) return rc;
    const StampJpegInfo& I = P.info;
    if (expect != nullptr && (expect->width != I.width || expect->height != I.height || expect->h[0] != I.h[0] ||
                              expect->v[0] != I.v[0]))
        return STAMP_ERR_UNSUPPORTED;   // every tile of a batch shares one geometry
    std::memcpy(quant, I.quant, sizeof(I.quant));
    int16_t* plane[3];
    int bx[3];
    size_t off = 0;
    for (int i = 0; i < 3; ++i) {
        plane[i] = coef + off;
        bx[i] = I.mcus_x * I.h[i];
        off += static_cast<size_t>(bx[i]) * I.mcus_y * I.v[i] * 64;
    }
    BitReader br{P.scan, data + n};
    bool truncated = false;
    int pred[3] = {0, 0, 0};
    const int n_mcu = I.mcus_x * I.mcus_y;
    int next_rst = 0;
    for (int m = 0; m < n_mcu; ++m) {
        if (P.restart_interval && m && m % P.restart_interval == 0) {
            // byte-align, expect RSTn, reset predictions
            const uint8_t* q = br.p;
            while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) ++q;
            if (q + 1 >= br.end || q[1] != 0xD0 + next_rst) return STAMP_ERR_BAD_ARG;
            next_rst = (next_rst + 1) & 7;
            truncated |= br.eof;
            br = BitReader{q + 2, data + n};
            pred[0] = pred[1] = pred[2] = 0;
        }
        const int my = m / I.mcus_x, mx = m - my * I.mcus_x;
        for (int c = 0; c < 3; ++c) {
            const HuffTable& dct = P.dc[P.td[c]];
            const HuffTable& act = P.ac[P.ta[c]];
            for (int by = 0; by < I.v[c]; ++by)
                for (int bxi = 0; bxi < I.h[c]; ++bxi) {
                    int16_t blk[64];
                    std::memset(blk, 0, sizeof(blk));
                    int s = decode_symbol(br, dct);
                    if (s < 0 || s > 11) return STAMP_ERR_BAD_ARG;
                    if (s) {
                        if (br.n < s) br.fill();
                        pred[c] += extend(br.get(s), s);
                    }
                    blk[0] = static_cast<int16_t>(pred[c]);
                    for (int k = 1; k < 64;) {
                        const int rs = decode_symbol(br, act);
                        if (rs < 0) return STAMP_ERR_BAD_ARG;
                        const int r = rs >> 4;
                        s = rs & 15;
                        if (s == 0) {
                            if (r != 15) break;
                            k += 16;
                            continue;
                        }
                        k += r;
                        if (k > 63) return STAMP_ERR_BAD_ARG;
                        if (br.n < s) br.fill();
                        blk[kZigzag[k]] = static_cast<int16_t>(extend(br.get(s), s));
                        ++k;
                    }
                    const size_t b = static_cast<size_t>(my * I.v[c] + by) * bx[c] + mx * I.h[c] + bxi;
                    std::memcpy(plane[c] + b * 64, blk, sizeof(blk));
                }
        }
    }
    return (truncated || br.eof) ? STAMP_ERR_BAD_ARG : STAMP_OK;   // ran off the end of the data: truncated file
}
