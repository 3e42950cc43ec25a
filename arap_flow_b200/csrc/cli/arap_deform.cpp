// arap_deform -- command-line front end with the argv / list-file / environment contract of the reference's
// ARAP/deformation/src/main.cpp:162-241, as driven by para_gen.py:178-214, run_arap.py:10-15 and generate.py:
//   arap_deform RGB MASK CSTR FLO_out WRGB_out WMASK_out      or      arap_deform LISTFILE
// Differences from the reference are on the inside only: consecutive list entries of one image size are
// solved together (several problems per cooperative launch), nothing is interpreted from $ARAP_PLAN.
#include "../../../include/Opt.h"
#include "../../../include/arapb200.h"
#include "image_io.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <deque>
#include <future>
#include <memory>
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
    puts("             ARAP_BATCH = problems solved together (default 9: three cooperative launches of three)");
    puts("             ARAP_PCG_RTOL = opt-in relative PCG tolerance, e.g. 1e-3 (default 0: fixed 400 iterations)");
    puts("             ARAP_GN_RTOL = opt-in relative cost-decrease tolerance of the Gauss-Newton steps (default 0: fixed 8 steps)");
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
    int batch = getenv("ARAP_BATCH") ? atoi(getenv("ARAP_BATCH")) : 9;
    if (batch < 1) batch = 1;
    // opt-in, off by default: convergence-aware PCG loops (changes results; include/arapb200.h)
    const double pcg_rtol = getenv("ARAP_PCG_RTOL") ? atof(getenv("ARAP_PCG_RTOL")) : 0.0;
    const double gn_rtol = getenv("ARAP_GN_RTOL") ? atof(getenv("ARAP_GN_RTOL")) : 0.0;
    // ARAP_TIMING=1: per-stage wall times on stderr
    const bool timing = getenv("ARAP_TIMING") != NULL;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    double t_decode = 0, t_gpu_wait = 0, t_encode = 0, t_create = 0;

    // Three overlapped stages: the main thread decodes group k+1 and encodes group k-1 while
    // a worker thread runs group k on the GPU.  Output order and the "Saved" lines stay in list order.
    struct Group {
        size_t first = 0;
        int W = 0, H = 0;
        std::vector<Loaded> items;
    };
    arapb200_batch* ctx = NULL;
    int ctxW = 0, ctxH = 0;
    auto gpu_stage = [&](std::shared_ptr<Group> g) -> int {
        if (!ctx || g->W != ctxW || g->H != ctxH) {
            if (ctx) {
                printf("Warning: Input image has different size to one in the prebuilt plan.\n"
                       "To avoid re-building the plan and to save time, put images of the "
                       "same size in the same list.\nStarting to re-build plan...\n");
                arapb200_batch_destroy(ctx);
            }
            const auto t0 = std::chrono::steady_clock::now();
            ctx = arapb200_batch_create(g->W, g->H, batch, nCont, nGN, nPCG, ARAPB200_BACKEND_AUTO);
            t_create += since(t0);
            if (!ctx) return 1;
            if (pcg_rtol > 0.0 && arapb200_batch_set_option(ctx, "pcg_rtol", pcg_rtol)) {
                fprintf(stderr, "ARAP_PCG_RTOL must be in [0, 1)\n");
                return 1;
            }
            if (gn_rtol > 0.0 && arapb200_batch_set_option(ctx, "gn_rtol", gn_rtol)) {
                fprintf(stderr, "ARAP_GN_RTOL must be in [0, 1)\n");
                return 1;
            }
            ctxW = g->W; ctxH = g->H;
        }
        for (size_t k = 0; k < g->items.size(); ++k) {
            Loaded& L = g->items[k];
            if (arapb200_batch_submit(ctx, (int)k, g->W, g->H, L.rgb.px.data(), L.mask_red.data(), L.cstr.data(),
                                      (int)(L.cstr.size() / 4), L.flow.data(), L.wrgb.data(), L.wmask.data(), NULL))
                return 1;
        }
        return arapb200_batch_run(ctx);
    };
    // decode / encode run one task per list entry (zlib dominates: ~0.1 s of CPU per pair at 854x480)
    auto save_one = [&](const Group& g, size_t k) -> bool {
        const Loaded& L = g.items[k];
        const InputPaths& p = lines[g.first + k];
        std::vector<uint8_t> m3((size_t)3 * g.W * g.H); // warped mask as an RGB image, 255 = object (README.md:25-32)
        for (size_t q = 0; q < L.wmask.size(); ++q) m3[3 * q] = m3[3 * q + 1] = m3[3 * q + 2] = L.wmask[q];
        return save_png_rgb(p.wrgb, g.W, g.H, L.wrgb.data()) && save_png_rgb(p.wmask, g.W, g.H, m3.data()) &&
               write_flo(p.flo, g.W, g.H, L.flow.data());
    };
    auto save_stage = [&](const Group& g) -> bool {
        std::vector<std::future<bool>> jobs;
        for (size_t k = 0; k < g.items.size(); ++k) jobs.push_back(std::async(std::launch::async, save_one, std::cref(g), k));
        bool ok = true;
        for (auto& j : jobs) {
            if (j.get()) printf("Saved\n"); // list order
            else ok = false;
        }
        return ok;
    };
    auto load_one = [&](size_t idx, Loaded* out) -> bool {
        Loaded& L = *out;
        ImageRGB m;
        if (!read_constraints(lines[idx].cstr, L.cstr) || !load_png_rgb(lines[idx].rgb, L.rgb) || !load_png_rgb(lines[idx].mask, m))
            return false;
        if (m.W != L.rgb.W || m.H != L.rgb.H) {
            fprintf(stderr, "mask %s and image %s differ in size\n", lines[idx].mask.c_str(), lines[idx].rgb.c_str());
            return false;
        }
        const size_t N = (size_t)m.W * m.H;
        L.mask_red.resize(N);
        for (size_t q = 0; q < N; ++q) L.mask_red[q] = m.px[3 * q]; // red channel only
        L.flow.resize(2 * N);
        L.wrgb.resize(3 * N);
        L.wmask.resize(N);
        return true;
    };

    std::future<int> running;
    std::shared_ptr<Group> in_flight, done;
    std::deque<Loaded> ready; // decoded, not yet grouped (a size change leaves the tail for the next group)
    size_t decoded = 0, grouped = 0;
    while (grouped < lines.size() || in_flight) {
        // ---- decode ahead: keep up to `batch` entries ready, all of them in parallel ----
        std::shared_ptr<Group> next;
        if (grouped < lines.size()) {
            const size_t want = std::min(lines.size() - grouped, (size_t)batch);
            if (ready.size() < want) {
                const auto t0 = std::chrono::steady_clock::now();
                const size_t n_new = want - ready.size();
                std::vector<Loaded> fresh(n_new);
                std::vector<std::future<bool>> jobs;
                for (size_t k = 0; k < n_new; ++k) jobs.push_back(std::async(std::launch::async, load_one, decoded + k, &fresh[k]));
                bool ok = true;
                for (auto& j : jobs) ok = j.get() && ok;
                if (!ok) return 1;
                for (auto& L : fresh) ready.push_back(std::move(L));
                decoded += n_new;
                t_decode += since(t0);
            }
            // ---- the next group: consecutive entries of one image size ----
            next = std::make_shared<Group>();
            next->first = grouped;
            next->W = ready.front().rgb.W;
            next->H = ready.front().rgb.H;
            while (!ready.empty() && (int)next->items.size() < batch && ready.front().rgb.W == next->W &&
                   ready.front().rgb.H == next->H) {
                next->items.push_back(std::move(ready.front()));
                ready.pop_front();
                ++grouped;
            }
        }
        // ---- wait for the group on the GPU, start the next one, then encode the finished one ----
        if (in_flight) {
            const auto t0 = std::chrono::steady_clock::now();
            const int rc = running.get();
            t_gpu_wait += since(t0);
            if (rc) {
                fprintf(stderr, "arap_deform: solver failed (%d)\n", rc);
                return rc;
            }
            done = in_flight;
            in_flight.reset();
        }
        if (next) {
            in_flight = next;
            running = std::async(std::launch::async, gpu_stage, next);
        }
        if (done) {
            const auto t0 = std::chrono::steady_clock::now();
            if (!save_stage(*done)) return 1;
            t_encode += since(t0);
            done.reset();
        }
    }
    if (timing)
        fprintf(stderr, "arap_deform timing: total %.3f s | decode %.3f | waiting for the GPU %.3f (context + buffers %.3f) | "
                        "encode %.3f\n", since(t_begin), t_decode, t_gpu_wait, t_create, t_encode);
    if (ctx) arapb200_batch_destroy(ctx);
    return 0;
}
