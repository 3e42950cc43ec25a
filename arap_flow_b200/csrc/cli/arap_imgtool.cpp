// arap_imgtool -- exercises the CLI codecs without a GPU (used by the CPU test-suite):
//   arap_imgtool png2raw in.png out.raw      (out.raw = int32 W, int32 H, then W*H*3 bytes RGB)
//   arap_imgtool raw2png in.raw out.png
//   arap_imgtool flocopy in.flo out.flo      (read + re-write)
//   arap_imgtool cstr in.txt                 (prints n and the sum of all integers)
#include "image_io.h"

#include <cstdio>
#include <cstring>

using namespace arapcli;

int main(int argc, const char* argv[])
{
    if (argc >= 4 && !strcmp(argv[1], "png2raw")) {
        ImageRGB im;
        if (!load_png_rgb(argv[2], im)) return 1;
        FILE* f = fopen(argv[3], "wb");
        if (!f) return 1;
        fwrite(&im.W, 4, 1, f); fwrite(&im.H, 4, 1, f);
        fwrite(im.px.data(), 1, im.px.size(), f);
        fclose(f);
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "raw2png")) {
        FILE* f = fopen(argv[2], "rb");
        if (!f) return 1;
        int W = 0, H = 0;
        if (fread(&W, 4, 1, f) != 1 || fread(&H, 4, 1, f) != 1) return 1;
        std::vector<uint8_t> px((size_t)W * H * 3);
        if (fread(px.data(), 1, px.size(), f) != px.size()) return 1;
        fclose(f);
        return save_png_rgb(argv[3], W, H, px.data()) ? 0 : 1;
    }
    if (argc >= 4 && !strcmp(argv[1], "flocopy")) {
        int W, H;
        std::vector<float> uv;
        if (!read_flo(argv[2], W, H, uv)) return 1;
        return write_flo(argv[3], W, H, uv.data()) ? 0 : 1;
    }
    if (argc >= 3 && !strcmp(argv[1], "cstr")) {
        std::vector<int32_t> c;
        if (!read_constraints(argv[2], c)) return 1;
        long long s = 0;
        for (int v : c) s += v;
        printf("%zu %lld\n", c.size() / 4, s);
        return 0;
    }
    fprintf(stderr, "usage: arap_imgtool png2raw|raw2png|flocopy|cstr ...\n");
    return 2;
}
