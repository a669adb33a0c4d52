// render_cli.cpp -- the reference's intended render command, `render <spp> <res_y> <scene-id|index|file>`
// (dead code in the reference: src/cmd_render.rs:17-44, .vscode/launch.json args "500 300 mesh"), driving the
// B200 backend through the C ABI only.  Output follows render() (src/render/mod.rs:1031-1088): a P3 PPM named
// out/<timestamp>-scene-<id>-spp<N>-res<H>-.ppm plus a `latest.ppm` symlink.
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ptb.h"

static const char *kSceneIds[] = {"cornell", "mesh", "single-sphere", "two-spheres", "three-spheres", "cartesian"};

static void usage(const char *argv0) {
    std::fprintf(stderr,
                 "Run with:\n  %s <samplesPerPixel = 100> <y-resolution = 300> <scene = 'mesh'> [--seed N] [--width W] [--gpu K | --gpus N] [--out FILE]\n\nScenes:",
                 argv0);
    for (size_t i = 0; i < sizeof kSceneIds / sizeof *kSceneIds; ++i) std::fprintf(stderr, " %zu: %s,", i, kSceneIds[i]);
    std::fprintf(stderr, " or a path to a scene .json\n");
}

int main(int argc, char **argv) {
    // defaults of the GUI that is the reference's only caller: spp 100, res_y 300, scene "mesh" (main.rs:79,91-92)
    unsigned long long spp = 100, seed = 0;
    int res_y = 300, width = 0, gpu = 0, gpus = 1;  // --gpus N: devices 0..N-1 in one context (ptb_create_multi)
    std::string scene = "mesh", out_path;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { usage(argv[0]); std::exit(1); } return argv[++i]; };
        if (a == "--seed") seed = std::strtoull(next(), nullptr, 10);
        else if (a == "--width") width = std::atoi(next());
        else if (a == "--gpu") gpu = std::atoi(next());
        else if (a == "--gpus") gpus = std::atoi(next());
        else if (a == "--out") out_path = next();
        else if (a == "-h" || a == "--help") { usage(argv[0]); return 0; }
        else pos.push_back(a);
    }
    if (pos.size() != 0 && pos.size() != 3) { usage(argv[0]); return 1; }
    if (pos.size() == 3) {
        spp = std::strtoull(pos[0].c_str(), nullptr, 10);
        res_y = std::atoi(pos[1].c_str());
        scene = pos[2];
        char *end = nullptr;
        unsigned long idx = std::strtoul(scene.c_str(), &end, 10);  // SceneId::Int, cmd_render.rs:20-24
        if (end && *end == 0 && !scene.empty()) {
            if (idx >= sizeof kSceneIds / sizeof *kSceneIds) { usage(argv[0]); return 1; }
            scene = kSceneIds[idx];
        }
    }
    if (spp == 0 || res_y <= 0) { usage(argv[0]); return 1; }
    if (width <= 0) width = res_y * 3 / 2;  // main.rs:176
    std::string json = scene.find(".json") != std::string::npos ? scene : "scenes/" + scene + ".json";

    char err[512];
    ptb_scene *sc = nullptr;
    if (ptb_scene_load_json(json.c_str(), ".", &sc, err, sizeof err) != PTB_OK) { std::fprintf(stderr, "error: %s\n", err); return 1; }
    const ptb_scene_desc *desc = ptb_scene_get_desc(sc);
    ptb_ctx *ctx = nullptr;
    if (gpus < 1 || gpus > ptb_device_count()) { std::fprintf(stderr, "error: --gpus %d, but %d CUDA device(s) are visible\n", gpus, ptb_device_count()); return 2; }
    std::vector<int> ids;
    for (int g = 0; g < gpus; ++g) ids.push_back(gpus == 1 ? gpu : g);
    if (ptb_create_multi(ids.data(), gpus, &ctx) != PTB_OK) { std::fprintf(stderr, "error: %s\n", ptb_last_error(nullptr)); return 2; }
    if (ptb_upload_scene(ctx, desc) != PTB_OK) { std::fprintf(stderr, "error: %s\n", ptb_last_error(ctx)); return 2; }
    std::printf("Rendering scene %s (%llu objects), %llu samples per pixel, %dx%d resolution, %d GPU(s)\n", ptb_scene_id(sc),
                (unsigned long long)desc->n_objects, spp, width, res_y, gpus);
    std::vector<float> img((size_t)width * res_y * 3);
    auto t0 = std::chrono::steady_clock::now();
    // progress line with elapsed / estimated total time, like the reference's print_progress (cmd_render.rs:54-80)
    volatile uint64_t samples_done = 0;
    std::atomic<bool> finished{false};
    const double total = (double)width * res_y * (double)spp;
    auto fmt = [](double s, char *buf, size_t n) {
        unsigned long long t = (unsigned long long)s, h = t / 3600, m = (t / 60) % 60, sec = t % 60;
        if (h == 0) std::snprintf(buf, n, "%llum:%02llus", m, sec);
        else std::snprintf(buf, n, "%llu:%02llu:%02llu", h, m, sec);
    };
    std::thread progress([&]() {
        while (!finished.load()) {
            std::this_thread::sleep_for(std::chrono::milliseconds(500));
            const double frac = (double)samples_done / total;
            if (frac <= 0.0 || finished.load()) continue;
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            char a[32], b[32];
            fmt(el, a, sizeof a); fmt(el / frac, b, sizeof b);
            std::printf("\rRendering ... %5.1f%% (%s / %s)", 100.0 * frac, a, b);
            std::fflush(stdout);
        }
    });
    int rc = ptb_render(ctx, width, res_y, 0, spp, seed, PTB_OUT_MEAN, img.data(), nullptr, &samples_done);
    finished.store(true);
    progress.join();
    if (samples_done) std::printf("\n");
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc < 0) { std::fprintf(stderr, "error: %s\n", ptb_last_error(ctx)); return 2; }
    ptb_stats st;
    ptb_get_stats(ctx, &st);
    std::printf("Rendering complete: %.3f s (kernels %.3f s), %.2f Mpaths/s, %.2f Mray-segments/s, hash %016llx\n", sec,
                st.render_ms * 1e-3, (double)st.samples / sec * 1e-6, (double)st.segments / sec * 1e-6,
                (unsigned long long)ptb_hash_pixels(img.data(), (uint64_t)width * res_y));

    if (out_path.empty()) {
        mkdir("out", 0755);
        char ts[64];
        std::time_t now = std::time(nullptr);
        std::strftime(ts, sizeof ts, "%Y-%m-%d_%H:%M:%S", std::localtime(&now));
        out_path = std::string("out/") + ts + "-scene-" + ptb_scene_id(sc) + "-spp" + std::to_string(spp) + "-res" +
                   std::to_string(res_y) + "-.ppm";
    }
    if (ptb_write_ppm(out_path.c_str(), img.data(), width, res_y, spp, ptb_scene_id(sc), (uint64_t)sec) != PTB_OK) {
        std::fprintf(stderr, "error: %s\n", ptb_last_error(nullptr));
        return 3;
    }
    unlink("latest.ppm");
    if (symlink(out_path.c_str(), "latest.ppm") != 0)
        std::printf("Could not create symlink to latest image. You can find it at %s\n", out_path.c_str());
    ptb_destroy(ctx);
    ptb_scene_free(sc);
    return 0;
}
