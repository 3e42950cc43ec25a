// image_io.h -- the file formats on either side of the hot path, for the command-line tools:
//   PNG  : decoded to 8-bit RGB (the reference goes through LodePNG into ColorImageR8G8B8:
//          ARAP/external/mLib/src/ext-lodepng/imageLoaderLodePNG.cpp:5-51); written as 8-bit RGB.
//          Only decoded pixel values are contractual.  Minimal codec over zlib: all colour
//          types and bit depths, no interlacing.
//   .flo : "PIEH", int32 W, int32 H, rows of interleaved (u, v) float32 (ARAP/deformation/src/main.cpp:53-75,
//          reader ARAP/warping/src/main.cpp:228-274)
//   cstr : n, then n x (x1 y1 x2 y2) integers (ARAP/deformation/src/main.cpp:26-50)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace arapcli {

struct ImageRGB {
    int W = 0, H = 0;
    std::vector<uint8_t> px; // W*H*3
};

// false + message on stderr when the file cannot be read / is not a PNG this codec understands
bool load_png_rgb(const std::string& path, ImageRGB& out);
bool save_png_rgb(const std::string& path, int W, int H, const uint8_t* rgb);

bool read_flo(const std::string& path, int& W, int& H, std::vector<float>& uv);
bool write_flo(const std::string& path, int W, int H, const float* uv);

bool read_constraints(const std::string& path, std::vector<int32_t>& xyxy);

} // namespace arapcli
