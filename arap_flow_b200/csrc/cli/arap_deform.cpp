// arap_deform -- command-line front end with the argv / list-file / environment contract of the reference's
// ARAP/deformation/src/main.cpp:162-241, as driven by para_gen.py:178-214, run_arap.py:10-15 and generate.py:
//   arap_deform RGB MASK CSTR FLO_out WRGB_out WMASK_out      or      arap_deform LISTFILE
// Differences from the reference are on the inside only: consecutive list entries of one image size are
// solved together (several problems per cooperative launch), nothing is interpreted from $ARAP_PLAN.
#include "../../../include/Opt.h"
#include "../../../include/arapb200.h"
#include "image_io.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

using namespace arapcli;

struct InputPaths {
    std::string rgb, mask, cstr, flo, wrgb, wmask;
};

static void usage()
{
    puts("Usage:\n");
    puts("./arap_deform RGB Mask Constraint Flow warped_RGB warped_Mask\n");
    puts("Deform the object of an image as rigidly as possible so that the given point matches are met,");
    puts("and write the dense flow, the warped image and the warped mask.\n");
    puts("RGB \t\t [input]  path to an input RGB image (.png only)");
    puts("Mask\t\t [input]  path to an input mask image (.png only), red channel 0 for object, non-zero for background");
    puts("Constraint \t [input]  path to list of constraints, text file: n, then n lines x1 y1 x2 y2");
    puts("Flow \t\t [output] path to optical flow (.flo only)");
    puts("warped_RGB \t [output] path to output warped image (.png), all intermediate directories must exist");
    puts("warped_Mask \t [output] path to output warped mask (.png), all intermediate directories must exist");
    puts("\n./arap_deform LISTFILE   (one such 6-tuple per line)");
    puts("Environment: ARAP_PLAN = path of the ARAP energy file (default ./arap_plan.t); CUDA_VISIBLE_DEVICES selects the GPU;");
    puts("             ARAP_BATCH = problems solved together (default 8)");
}

struct Loaded {
    ImageRGB rgb;
    std::vector<uint8_t> mask_red;
    std::vector<int32_t> cstr;
    std::vector<float> flow;
    std::vector<uint8_t> wrgb, wmask;
};

int main(int argc, const char* argv[])
{
    std::vector<InputPaths> lines;
    if (argc == 7) {
        lines.push_back({argv[1], argv[2], argv[3], argv[4], argv[5], argv[6]});
    } else if (argc == 2) { // a list file: main.cpp:182-193
        std::ifstream infile(argv[1]);
        std::string line;
        while (getline(infile, line)) {
            std::stringstream s(line);
            InputPaths p;
            if (s >> p.rgb >> p.mask >> p.cstr >> p.flo >> p.wrgb >> p.wmask) lines.push_back(p);
        }
    } else {
        printf("Invalid Input!\n");
        usage();
        return 1;
    }
    if (lines.empty()) {
        printf("No file to be processed");
        return 1;
    }
    const char* planPath = getenv("ARAP_PLAN") == NULL ? "arap_plan.t" : getenv("ARAP_PLAN");
    printf("Optimization plan at %s\n", planPath);
    {
        Opt_InitializationParameters ip = {0, 0, 0, 0};
        Opt_State* st = Opt_NewState(ip);
        Opt_Problem* pr = st ? Opt_ProblemDefine(st, planPath, "gaussNewtonGPU") : NULL;
        if (!pr) {
            printf(" Not found! Please run export ARAP_PLAN=/path/to/plan.t or copy "
                   "the file to the running folder with name arap_plan.t");
            return 1;
        }
        Opt_ProblemDelete(st, pr);
    }
    // the solver budget is a compile-time constant of the reference: main.cpp:215-221
    const int nCont = 19, nGN = 8, nPCG = 400;
    int batch = getenv("ARAP_BATCH") ? atoi(getenv("ARAP_BATCH")) : 8;
    if (batch < 1) batch = 1;

    arapb200_batch* ctx = NULL;
    int ctxW = 0, ctxH = 0;
    size_t i = 0;
    while (i < lines.size()) {
        // gather up to `batch` consecutive entries of one image size
        std::vector<Loaded> group;
        size_t j = i;
        int W = 0, H = 0;
        for (; j < lines.size() && (int)group.size() < batch; ++j) {
            Loaded L;
            ImageRGB m;
            if (!read_constraints(lines[j].cstr, L.cstr) || !load_png_rgb(lines[j].rgb, L.rgb) || !load_png_rgb(lines[j].mask, m))
                return 1;
            if (m.W != L.rgb.W || m.H != L.rgb.H) {
                fprintf(stderr, "mask %s and image %s differ in size\n", lines[j].mask.c_str(), lines[j].rgb.c_str());
                return 1;
            }
            if (group.empty()) { W = L.rgb.W; H = L.rgb.H; }
            else if (L.rgb.W != W || L.rgb.H != H) break; // next group
            L.mask_red.resize((size_t)W * H);
            for (size_t k = 0; k < L.mask_red.size(); ++k) L.mask_red[k] = m.px[3 * k]; // red channel only
            L.flow.resize((size_t)2 * W * H);
            L.wrgb.resize((size_t)3 * W * H);
            L.wmask.resize((size_t)W * H);
            group.push_back(std::move(L));
        }
        if (!ctx || W != ctxW || H != ctxH) {
            if (ctx) {
                printf("Warning: Input image has different size to one in the prebuilt plan.\n"
                       "To avoid re-building the plan and to save time, put images of the "
                       "same size in the same list.\nStarting to re-build plan...\n");
                arapb200_batch_destroy(ctx);
            }
            ctx = arapb200_batch_create(W, H, batch, nCont, nGN, nPCG, ARAPB200_BACKEND_AUTO);
            if (!ctx) return 1;
            ctxW = W; ctxH = H;
        }
        for (size_t g = 0; g < group.size(); ++g) {
            Loaded& L = group[g];
            if (arapb200_batch_submit(ctx, (int)g, W, H, L.rgb.px.data(), L.mask_red.data(), L.cstr.data(),
                                      (int)(L.cstr.size() / 4), L.flow.data(), L.wrgb.data(), L.wmask.data(), NULL))
                return 1;
        }
        if (int rc = arapb200_batch_run(ctx)) {
            fprintf(stderr, "arap_deform: solver failed (%d)\n", rc);
            return rc;
        }
        for (size_t g = 0; g < group.size(); ++g) {
            Loaded& L = group[g];
            const InputPaths& p = lines[i + g];
            std::vector<uint8_t> m3((size_t)3 * W * H); // warped mask as an RGB image, 255 = object (README.md:25-32)
            for (size_t k = 0; k < L.wmask.size(); ++k) m3[3 * k] = m3[3 * k + 1] = m3[3 * k + 2] = L.wmask[k];
            if (!save_png_rgb(p.wrgb, W, H, L.wrgb.data()) || !save_png_rgb(p.wmask, W, H, m3.data()) ||
                !write_flo(p.flo, W, H, L.flow.data()))
                return 1;
            printf("Saved\n");
        }
        i += group.size();
    }
    if (ctx) arapb200_batch_destroy(ctx);
    return 0;
}
