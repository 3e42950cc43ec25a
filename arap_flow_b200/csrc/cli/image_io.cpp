// image_io.cpp -- see image_io.h
#include "image_io.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

namespace arapcli {
namespace {

bool read_file(const std::string& path, std::vector<uint8_t>& buf)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    bool ok = n >= 0 && fread(buf.data(), 1, buf.size(), f) == buf.size();
    fclose(f);
    return ok;
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(std::vector<uint8_t>& v, uint32_t x)
{
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

void append_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n)
{
    put_be32(out, (uint32_t)n);
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    put_be32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4)));
}

} // namespace

bool load_png_rgb(const std::string& path, ImageRGB& out)
{
    std::vector<uint8_t> f;
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (!read_file(path, f) || f.size() < 8 || memcmp(f.data(), sig, 8) != 0) {
        fprintf(stderr, "load_png: cannot read '%s' as PNG\n", path.c_str());
        return false;
    }
    int W = 0, H = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    size_t pos = 8;
    while (pos + 12 <= f.size()) {
        const uint32_t len = be32(&f[pos]);
        const char* type = (const char*)&f[pos + 4];
        if (pos + 12 + len > f.size()) break;
        const uint8_t* d = &f[pos + 8];
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            W = (int)be32(d); H = (int)be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(d, d + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    const int chan = (ctype == 0) ? 1 : (ctype == 2) ? 3 : (ctype == 3) ? 1 : (ctype == 4) ? 2 : (ctype == 6) ? 4 : 0;
    if (W <= 0 || H <= 0 || !chan || interlace || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) {
        fprintf(stderr, "load_png: '%s': unsupported PNG (size %dx%d, depth %d, colour type %d, interlace %d)\n",
                path.c_str(), W, H, depth, ctype, interlace);
        return false;
    }
    const size_t bpp_bits = (size_t)chan * depth, stride = ((size_t)W * bpp_bits + 7) / 8, bpp = (bpp_bits + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * (size_t)H);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) {
        fprintf(stderr, "load_png: '%s': bad image data\n", path.c_str());
        return false;
    }
    // un-filter in place
    std::vector<uint8_t> prev(stride, 0);
    for (int y = 0; y < H; ++y) {
        uint8_t* row = &raw[(stride + 1) * (size_t)y];
        const int ft = row[0];
        uint8_t* p = row + 1;
        for (size_t i = 0; i < stride; ++i) {
            const int a = (i >= bpp) ? p[i - bpp] : 0, b = prev[i], c = (i >= bpp) ? prev[i - bpp] : 0;
            int v = p[i];
            switch (ft) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: fprintf(stderr, "load_png: '%s': bad filter\n", path.c_str()); return false;
            }
            p[i] = (uint8_t)v;
        }
        memcpy(prev.data(), p, stride);
    }
    out.W = W; out.H = H;
    out.px.assign((size_t)W * H * 3, 0);
    const int maxv = (1 << (depth < 8 ? depth : 8)) - 1;
    for (int y = 0; y < H; ++y) {
        const uint8_t* p = &raw[(stride + 1) * (size_t)y + 1];
        for (int x = 0; x < W; ++x) {
            int s[4] = {0, 0, 0, 0};
            for (int c = 0; c < chan; ++c) {
                if (depth == 8) s[c] = p[(size_t)x * chan + c];
                else if (depth == 16) s[c] = p[((size_t)x * chan + c) * 2]; // most significant byte, like LodePNG
                else {
                    const size_t bit = (size_t)x * depth; // chan == 1 for sub-byte depths
                    s[c] = (p[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
                }
            }
            uint8_t* o = &out.px[((size_t)y * W + x) * 3];
            if (ctype == 3) {
                const size_t idx = (size_t)s[0] * 3;
                if (idx + 2 < plte.size()) { o[0] = plte[idx]; o[1] = plte[idx + 1]; o[2] = plte[idx + 2]; }
            } else if (ctype == 0 || ctype == 4) {
                const int g = (depth < 8) ? (s[0] * 255) / maxv : s[0];
                o[0] = o[1] = o[2] = (uint8_t)g;
            } else {
                o[0] = (uint8_t)s[0]; o[1] = (uint8_t)s[1]; o[2] = (uint8_t)s[2];
            }
        }
    }
    return true;
}

bool save_png_rgb(const std::string& path, int W, int H, const uint8_t* rgb)
{
    const size_t stride = (size_t)W * 3;
    std::vector<uint8_t> raw((stride + 1) * (size_t)H);
    for (int y = 0; y < H; ++y) {
        raw[(stride + 1) * (size_t)y] = 0; // filter type "none"
        memcpy(&raw[(stride + 1) * (size_t)y + 1], rgb + stride * (size_t)y, stride);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    std::vector<uint8_t> out = {137, 80, 78, 71, 13, 10, 26, 10};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)W);
    put_be32(ihdr, (uint32_t)H);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    append_chunk(out, "IHDR", ihdr.data(), ihdr.size());
    append_chunk(out, "IDAT", comp.data(), clen);
    append_chunk(out, "IEND", nullptr, 0);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) {
        fprintf(stderr, "save_png: cannot write '%s'\n", path.c_str());
        return false;
    }
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

bool read_flo(const std::string& path, int& W, int& H, std::vector<float>& uv)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        fprintf(stderr, "ReadFlowFile: could not open %s\n", path.c_str());
        return false;
    }
    float tag = 0.f;
    int w = 0, h = 0;
    bool ok = fread(&tag, sizeof(float), 1, f) == 1 && fread(&w, sizeof(int), 1, f) == 1 && fread(&h, sizeof(int), 1, f) == 1;
    if (!ok || tag != 202021.25f) { // main.h:7 TAG_FLOAT
        fprintf(stderr, "ReadFlowFile(%s): wrong tag (possibly due to big-endian machine?)\n", path.c_str());
        fclose(f);
        return false;
    }
    if (w < 1 || w > 99999 || h < 1 || h > 99999) {
        fprintf(stderr, "ReadFlowFile(%s): illegal size %d x %d\n", path.c_str(), w, h);
        fclose(f);
        return false;
    }
    uv.resize((size_t)2 * w * h);
    ok = fread(uv.data(), sizeof(float), uv.size(), f) == uv.size();
    if (!ok) fprintf(stderr, "ReadFlowFile(%s): file is too short\n", path.c_str());
    else if (fgetc(f) != EOF) fprintf(stderr, "ReadFlowFile(%s): file is too long\n", path.c_str());
    fclose(f);
    W = w; H = h;
    return ok;
}

bool write_flo(const std::string& path, int W, int H, const float* uv)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) {
        fprintf(stderr, "WriteFlowFile(%s): cannot open\n", path.c_str());
        return false;
    }
    bool ok = fwrite("PIEH", 1, 4, f) == 4 && fwrite(&W, sizeof(int), 1, f) == 1 && fwrite(&H, sizeof(int), 1, f) == 1;
    ok = ok && fwrite(uv, sizeof(float), (size_t)2 * W * H, f) == (size_t)2 * W * H;
    if (!ok) fprintf(stderr, "WriteFlowFile(%s): problem writing data\n", path.c_str());
    fclose(f);
    return ok;
}

bool read_constraints(const std::string& path, std::vector<int32_t>& xyxy)
{
    std::ifstream in(path);
    if (!in.good()) {
        fprintf(stderr, "Could not open marker file %s\n", path.c_str());
        return false;
    }
    unsigned n = 0;
    in >> n;
    xyxy.assign((size_t)4 * n, 0);
    for (size_t i = 0; i < xyxy.size(); ++i) in >> xyxy[i];
    return true;
}

} // namespace arapcli
