// warp_image -- command-line front end with the argv contract of the reference's
// ARAP/warping/src/main.cpp:302-336 (driven by run_warp.py:8-19):
//   warp_image image mask flow warped_image warped_mask
#include "../../../include/arapb200.h"
#include "image_io.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace arapcli;

static void usage()
{
    puts("Usage:");
    puts("./warp_image image mask flow warped_image warped_mask");
    puts("Forward-warp the object of an image (and its mask) with the given optical flow field.");
    puts("\timage: path to image with png extension");
    puts("\tmask: path to mask image with png extension, red channel 0 for object, non-zero for background");
    puts("\tflo: path to optical flow with flo extension");
    puts("\twarped_image: path to output warped image (.png), all intermediate directories must exist");
    puts("\twarped_mask: path to output warped mask (.png), all intermediate directories must exist");
}

int main(int argc, const char* argv[])
{
    if (argc != 6) {
        printf("Invalid Input! ");
        usage();
        return 1;
    }
    ImageRGB rgb, mask;
    int W = 0, H = 0;
    std::vector<float> flow;
    if (!load_png_rgb(argv[1], rgb) || !load_png_rgb(argv[2], mask)) return 1;
    const char* dot = strrchr(argv[3], '.');
    if (!dot || strcmp(dot, ".flo") != 0) printf("ReadFlowFile (%s): extension .flo expected", argv[3]);
    if (!read_flo(argv[3], W, H, flow)) return 1;
    if (rgb.W != W || rgb.H != H || mask.W != W || mask.H != H) {
        fprintf(stderr, "warp_image: image (%dx%d), mask (%dx%d) and flow (%dx%d) differ in size\n", rgb.W, rgb.H, mask.W,
                mask.H, W, H);
        return 1;
    }
    std::vector<uint8_t> red((size_t)W * H), wrgb((size_t)3 * W * H), wmask((size_t)W * H), m3((size_t)3 * W * H);
    for (size_t k = 0; k < red.size(); ++k) red[k] = mask.px[3 * k];
    if (int rc = arapb200_warp_flow(W, H, flow.data(), rgb.px.data(), red.data(), wrgb.data(), wmask.data(), NULL)) {
        fprintf(stderr, "warp_image: warp failed (%d)\n", rc);
        return rc;
    }
    for (size_t k = 0; k < wmask.size(); ++k) m3[3 * k] = m3[3 * k + 1] = m3[3 * k + 2] = wmask[k];
    if (!save_png_rgb(argv[4], W, H, wrgb.data()) || !save_png_rgb(argv[5], W, H, m3.data())) return 1;
    printf("Saved\n");
    return 0;
}
