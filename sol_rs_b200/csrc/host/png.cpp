// png.cpp — sol::image::decode_png: 8- and 16-bit greyscale / RGB / palette / grey-alpha / RGBA, non-interlaced -> rgba8.
//
// The reference decodes images with the `image` crate (Texture2d::new, src/texture.rs:488-494: open -> flipv -> to_rgba8);
// nothing on its ray-tracing path samples one (texture_offset is reserved, src/ray/mod.rs:20).  This decoder exists so that
// the C++ host mirror can hand base-colour textures to solb_scene_set_textures without a dependency: inflate
// (RFC 1951: stored, fixed and dynamic Huffman blocks, decoded bit by bit with canonical-code counting), the zlib wrapper
// (RFC 1950, Adler-32 checked), the five PNG row filters.  16-bit samples become 8-bit by (v + 128) / 257, the crate's rule.
#include <cstring>

#include "sol.hpp"

namespace sol {
namespace image {

namespace {

struct Bits {
    const uint8_t *p, *end;
    uint32_t buf = 0;
    int cnt = 0;
    uint32_t get(int n) {  // n <= 16, LSB first
        while (cnt < n) {
            if (p >= end) throw Error(SOLB_ERR_INVALID, "png: deflate stream ends early");
            buf |= (uint32_t)*p++ << cnt;
            cnt += 8;
        }
        const uint32_t v = buf & ((1u << n) - 1u);
        buf >>= n;
        cnt -= n;
        return v;
    }
    void align() { buf = 0; cnt = 0; }
};

struct Huffman {
    uint16_t count[16];
    uint16_t symbol[288];
    void build(const uint8_t *len, int n) {
        std::memset(count, 0, sizeof(count));
        for (int i = 0; i < n; i++) count[len[i]]++;
        count[0] = 0;
        uint16_t offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
        for (int i = 0; i < n; i++)
            if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    }
    int decode(Bits &b) const {  // canonical codes: walk the lengths, first code of each length known from the counts
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; l++) {
            code |= (int)b.get(1);
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        throw Error(SOLB_ERR_INVALID, "png: bad Huffman code");
    }
};

const uint16_t LEN_BASE[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
const uint8_t LEN_EXTRA[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
const uint16_t DIST_BASE[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
                                 6145, 8193, 12289, 16385, 24577 };
const uint8_t DIST_EXTRA[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

void inflate_codes(Bits &b, const Huffman &lit, const Huffman &dist, std::vector<uint8_t> &out) {
    for (;;) {
        const int s = lit.decode(b);
        if (s < 256) out.push_back((uint8_t)s);
        else if (s == 256) return;
        else {
            if (s > 285) throw Error(SOLB_ERR_INVALID, "png: bad length symbol");
            const size_t len = LEN_BASE[s - 257] + b.get(LEN_EXTRA[s - 257]);
            const int ds = dist.decode(b);
            if (ds > 29) throw Error(SOLB_ERR_INVALID, "png: bad distance symbol");
            const size_t d = DIST_BASE[ds] + b.get(DIST_EXTRA[ds]);
            if (d > out.size()) throw Error(SOLB_ERR_INVALID, "png: distance beyond the output");
            const size_t from = out.size() - d;
            for (size_t i = 0; i < len; i++) out.push_back(out[from + i]);  // may overlap: byte by byte
        }
    }
}

std::vector<uint8_t> zlib_inflate(const uint8_t *data, size_t n, size_t expect) {
    if (n < 6 || (data[0] & 0x0f) != 8 || ((data[0] << 8) | data[1]) % 31 != 0 || (data[1] & 0x20))
        throw Error(SOLB_ERR_INVALID, "png: not a zlib stream");
    Bits b{ data + 2, data + n };
    std::vector<uint8_t> out;
    out.reserve(expect);
    for (bool last = false; !last;) {
        last = b.get(1) != 0;
        const uint32_t type = b.get(2);
        if (type == 0) {
            b.align();
            if (b.end - b.p < 4) throw Error(SOLB_ERR_INVALID, "png: stored block header missing");
            const uint32_t len = b.p[0] | (b.p[1] << 8), nlen = b.p[2] | (b.p[3] << 8);
            b.p += 4;
            if ((len ^ 0xffffu) != nlen || (size_t)(b.end - b.p) < len) throw Error(SOLB_ERR_INVALID, "png: bad stored block");
            out.insert(out.end(), b.p, b.p + len);
            b.p += len;
        } else if (type == 1) {
            uint8_t l[288 + 30];
            for (int i = 0; i < 144; i++) l[i] = 8;
            for (int i = 144; i < 256; i++) l[i] = 9;
            for (int i = 256; i < 280; i++) l[i] = 7;
            for (int i = 280; i < 288; i++) l[i] = 8;
            for (int i = 0; i < 30; i++) l[288 + i] = 5;
            Huffman lit, dist;
            lit.build(l, 288);
            dist.build(l + 288, 30);
            inflate_codes(b, lit, dist, out);
        } else if (type == 2) {
            const int nlen = (int)b.get(5) + 257, ndist = (int)b.get(5) + 1, ncode = (int)b.get(4) + 4;
            if (nlen > 286 || ndist > 30) throw Error(SOLB_ERR_INVALID, "png: too many codes");
            static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
            uint8_t l[320];
            std::memset(l, 0, sizeof(l));
            for (int i = 0; i < ncode; i++) l[order[i]] = (uint8_t)b.get(3);
            Huffman code;
            code.build(l, 19);
            uint8_t lens[320];
            for (int i = 0; i < nlen + ndist;) {
                const int s = code.decode(b);
                if (s < 16) lens[i++] = (uint8_t)s;
                else {
                    uint8_t v = 0;
                    int rep;
                    if (s == 16) {
                        if (i == 0) throw Error(SOLB_ERR_INVALID, "png: repeat without a length");
                        v = lens[i - 1];
                        rep = 3 + (int)b.get(2);
                    } else if (s == 17) rep = 3 + (int)b.get(3);
                    else rep = 11 + (int)b.get(7);
                    if (i + rep > nlen + ndist) throw Error(SOLB_ERR_INVALID, "png: code lengths overflow");
                    while (rep--) lens[i++] = v;
                }
            }
            Huffman lit, dist;
            lit.build(lens, nlen);
            dist.build(lens + nlen, ndist);
            inflate_codes(b, lit, dist, out);
        } else throw Error(SOLB_ERR_INVALID, "png: reserved block type");
    }
    b.align();
    if (b.end - b.p >= 4) {
        uint32_t a = 1, s2 = 0;
        for (uint8_t v : out) { a = (a + v) % 65521u; s2 = (s2 + a) % 65521u; }
        const uint32_t want = ((uint32_t)b.p[0] << 24) | ((uint32_t)b.p[1] << 16) | ((uint32_t)b.p[2] << 8) | b.p[3];
        if (((s2 << 16) | a) != want) throw Error(SOLB_ERR_INVALID, "png: Adler-32 mismatch");
    }
    return out;
}

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

}  // namespace

bool is_png(const uint8_t *data, size_t n) {
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
    return n >= 8 && std::memcmp(data, sig, 8) == 0;
}

Rgba8 decode_png(const uint8_t *data, size_t n) {
    if (!is_png(data, n)) throw Error(SOLB_ERR_INVALID, "png: bad signature");
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1;
    std::vector<uint8_t> idat, plte, trns;
    for (size_t p = 8; p + 12 <= n;) {
        const uint32_t len = be32(data + p);
        if (p + 12 + (size_t)len > n) throw Error(SOLB_ERR_INVALID, "png: chunk beyond the file");
        const uint8_t *tag = data + p + 4, *body = data + p + 8;
        if (!std::memcmp(tag, "IHDR", 4)) {
            if (len < 13) throw Error(SOLB_ERR_INVALID, "png: short IHDR");
            w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9];
            if (body[10] != 0 || body[11] != 0) throw Error(SOLB_ERR_UNSUPPORTED, "png: unknown compression / filter method");
            if (body[12] != 0) throw Error(SOLB_ERR_UNSUPPORTED, "png: interlaced images are not supported");
        } else if (!std::memcmp(tag, "PLTE", 4)) plte.assign(body, body + len);
        else if (!std::memcmp(tag, "tRNS", 4)) trns.assign(body, body + len);
        else if (!std::memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(tag, "IEND", 4)) break;
        p += 12 + (size_t)len;
    }
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: throw Error(SOLB_ERR_INVALID, "png: bad colour type");
    }
    if (w == 0 || h == 0 || w > 32768u || h > 32768u) throw Error(SOLB_ERR_INVALID, "png: bad size");
    if (!(depth == 8 || (depth == 16 && ctype != 3))) throw Error(SOLB_ERR_UNSUPPORTED, "png: only 8- and 16-bit samples are supported");
    const size_t bpp = (size_t)channels * depth / 8, stride = (size_t)w * bpp;
    std::vector<uint8_t> raw = zlib_inflate(idat.data(), idat.size(), (stride + 1) * h);
    if (raw.size() < (stride + 1) * h) throw Error(SOLB_ERR_INVALID, "png: image data too short");
    std::vector<uint8_t> prev(stride, 0);
    Rgba8 out;
    out.width = w;
    out.height = h;
    out.pixels.resize((size_t)w * h * 4);
    for (uint32_t y = 0; y < h; y++) {
        uint8_t *row = raw.data() + (size_t)y * (stride + 1) + 1;
        const int filter = row[-1];
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? row[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v = row[i];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: throw Error(SOLB_ERR_INVALID, "png: bad row filter");
            }
            row[i] = (uint8_t)v;
        }
        std::memcpy(prev.data(), row, stride);
        uint8_t *dst = out.pixels.data() + (size_t)y * w * 4;
        for (uint32_t x = 0; x < w; x++) {
            uint8_t s[4] = { 0, 0, 0, 255 };
            for (int k = 0; k < channels; k++) {
                if (depth == 8) s[k] = row[x * bpp + k];
                else s[k] = (uint8_t)(((((uint32_t)row[x * bpp + 2 * k] << 8) | row[x * bpp + 2 * k + 1]) + 128u) / 257u);
            }
            uint8_t *o = dst + 4 * x;
            if (ctype == 0) { o[0] = o[1] = o[2] = s[0]; o[3] = 255; }
            else if (ctype == 2) { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = 255; }
            else if (ctype == 4) { o[0] = o[1] = o[2] = s[0]; o[3] = s[1]; }
            else if (ctype == 6) { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3]; }
            else {
                const size_t idx = s[0];
                if (3 * idx + 2 >= plte.size()) throw Error(SOLB_ERR_INVALID, "png: palette index out of range");
                o[0] = plte[3 * idx]; o[1] = plte[3 * idx + 1]; o[2] = plte[3 * idx + 2];
                o[3] = idx < trns.size() ? trns[idx] : 255;
            }
        }
    }
    return out;
}

}  // namespace image
}  // namespace sol
