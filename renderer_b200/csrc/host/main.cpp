// main.cpp — the host program: same command line and main loop as the reference's renderer
// (reference src/renderer.cc:168-642), with the render call going through the C-ABI to the B200 kernels.
//
//   b200renderer [-h] [-r] [-b] [-n N] [-w] [-m mode] [extra long options] FILENAME
//
// Kept from the reference: -h help, -r fps report every 5 s, -b benchmark N frames (default 100) along the
// deterministic orbit, -n N, -w second light, -m <mode> (1..9, 0 = anti-aliased ray tracing), default mode 8,
// the final "Rendering N frames in S seconds. (F fps)" line.  There is no window here (no SDL, no display):
// without -b the program renders the same orbit until -n frames are done (default 100).
// New (the reference's compile-time #defines made runtime):
//   --width W --height H      (reference: WIDTH/HEIGHT in src/Defines.h:26-27, default 800x600)
//   --no-reflections          (REFLECTIONS, src/Raytracer.cc:67)      --no-shadows (USE_SHADOWS :63)
//   --ao N                    (AMBIENT_OCCLUSION + AMBIENT_SAMPLES, src/Raytracer.cc:77-79)
//   --mlaa                    (--enable-mlaa build + Screen::ShowScreen hook, src/Screen.h:130-137)
//   --dump PREFIX --frames a,b,c   write PREFIX_<frame>.xrgb (raw 0x00RRGGBB words) for the listed frames
//   --device D
//   --gpus N [--assemble push|nccl]   ray tracing on N GPUs of this node: one thread + one context per GPU, rows dealt
//                             round-robin, frames assembled on every GPU (b200r_pipeline_*)
#include <getopt.h>

#include <atomic>
#include <thread>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "../../../include/b200render.h"

static void usage()
{
    fprintf(stderr, "%s\n", b200r_version());
    fprintf(stderr, "Usage: b200renderer [OPTIONS] [FILENAME]\n\n"
                    "  -h         this help\n"
                    "  -r         print FPS reports to stdout (every 5 seconds)\n"
                    "  -b         benchmark rendering of N frames (default: 100)\n"
                    "  -n N       set number of benchmarking frames\n"
                    "  -w         use two lights\n"
                    "  -m <mode>  rendering mode:\n"
                    "       1 : point mode\n"
                    "       2 : points based on triangles (culling,color)\n"
                    "       3 : triangles, wireframe anti-aliased\n"
                    "       4 : triangles, ambient colors\n"
                    "       5 : triangles, Gouraud shading, ZBuffer\n"
                    "       6 : triangles, per-pixel Phong, ZBuffer\n"
                    "       7 : triangles, per-pixel Phong, ZBuffer, Shadowmaps\n"
                    "       8 : triangles, per-pixel Phong, ZBuffer, Soft shadowmaps\n"
                    "       9 : raytracing, with shadows and reflections\n"
                    "       0 : raytracing, with shadows, reflections and anti-aliasing\n"
                    "  --width W --height H --no-reflections --no-shadows --ao N --mlaa\n"
                    "  --dump PREFIX --frames a,b,c --device D --host-bvh --frames-in-flight N (1..8, default 2)\n"
                    "  --gpus N [--assemble push|nccl]   ray tracing (-m 9 / -m 0) on N GPUs of this node\n");
    exit(0);
}

static double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char* argv[])
{
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);      // one hardware queue per stream of the frame pipeline (before CUDA starts)
    unsigned mode = B200R_MODE_PHONG_SOFTSHADOWMAPS;     // reference default (renderer.cc:177)
    bool doReports = false, doBenchmark = false, useTwoLights = false;
    unsigned benchmarkFrames = 100;
    unsigned W = 800, H = 600, flags = B200R_F_DEFAULT, ao = 0;
    int device = 0;
    unsigned inFlight = 4;               // frames rendering concurrently (b200r_set_pipeline_depth; measured: 2: 3950 fps, 3: 4410, 4: 4550 on C2)
    bool hostBvh = false;
    unsigned gpus = 1, assemble = B200R_ASSEMBLE_PUSH;
    std::string dumpPrefix;
    std::set<unsigned> dumpFrames;

    static option longopts[] = {{"width", required_argument, 0, 1000}, {"height", required_argument, 0, 1001},
                                {"no-reflections", no_argument, 0, 1002}, {"no-shadows", no_argument, 0, 1003},
                                {"ao", required_argument, 0, 1004}, {"mlaa", no_argument, 0, 1005},
                                {"dump", required_argument, 0, 1006}, {"frames", required_argument, 0, 1007},
                                {"device", required_argument, 0, 1008}, {"host-bvh", no_argument, 0, 1009},
                                {"frames-in-flight", required_argument, 0, 1010}, {"gpus", required_argument, 0, 1011},
                                {"assemble", required_argument, 0, 1012}, {0, 0, 0, 0}};
    int c;
    opterr = 0;
    while ((c = getopt_long(argc, argv, "hbrwn:m:", longopts, nullptr)) != -1) switch (c) {
        case 'h': usage(); break;
        case 'm':
            if (atoi(optarg) == 0) mode = B200R_MODE_RAYTRACE_AA; else mode = (unsigned)atoi(optarg);
            if (atoi(optarg) < 0 || atoi(optarg) > 9) usage();        // the reference's switch knows modes 0..9 (renderer.cc:196-206)
            break;
        case 'b': doBenchmark = true; break;
        case 'r': doReports = true; break;
        case 'w': useTwoLights = true; break;
        case 'n': benchmarkFrames = (unsigned)atoi(optarg); break;
        case 1000: W = (unsigned)atoi(optarg); break;
        case 1001: H = (unsigned)atoi(optarg); break;
        case 1002: flags &= ~B200R_F_REFLECTIONS; break;
        case 1003: flags &= ~B200R_F_SHADOWS; break;
        case 1004: ao = (unsigned)atoi(optarg); if (ao) flags |= B200R_F_AO; break;
        case 1005: flags |= B200R_F_MLAA; break;
        case 1006: dumpPrefix = optarg; break;
        case 1007: { char* p = optarg; while (*p) { dumpFrames.insert((unsigned)strtoul(p, &p, 10)); if (*p == ',') p++; else break; } } break;
        case 1008: device = atoi(optarg); break;
        case 1009: hostBvh = true; break;
        case 1010: inFlight = (unsigned)atoi(optarg); if (inFlight < 1 || inFlight > B200R_MAX_FRAMES_IN_FLIGHT) usage(); break;
        case 1011: gpus = (unsigned)atoi(optarg); if (gpus < 1 || gpus > 16) usage(); break;
        case 1012: assemble = !strcmp(optarg, "nccl") ? B200R_ASSEMBLE_NCCL : B200R_ASSEMBLE_PUSH; break;
        case '?': fprintf(stderr, "No such option (%c)\n", (char)optopt); usage(); break;
        default: break;
    }
    if (optind == argc) usage();
    const char* fname = argv[optind];

    b200r_scene* scene = nullptr;
    if (b200r_scene_load(fname, &scene)) { fprintf(stderr, "%s\n", b200r_last_error(nullptr)); return 1; }
    uint32_t nv = 0, nt = 0;
    b200r_scene_vertices(scene, &nv); b200r_scene_tris(scene, &nt);
    printf("Vertexes: %u Triangles: %u\n", nv, nt);
    const bool raytrace = (mode == B200R_MODE_RAYTRACE || mode == B200R_MODE_RAYTRACE_AA);
    b200r_ctx* ctx = nullptr;
    if (b200r_init(device, &ctx)) { fprintf(stderr, "%s\n", b200r_last_error(nullptr)); return 1; }
    if (raytrace) {
        puts("Creating BVH... please wait...");
        const std::string cache = std::string(fname) + ".bvh";      // same cache file as the reference
        const double t0 = now_ms();
        // the SAH build runs on the device (same tree, same cache file); --host-bvh keeps it on the host
        const int rc = hostBvh ? b200r_scene_build_bvh(scene, cache.c_str(), 0) : b200r_scene_build_bvh_device(scene, ctx, cache.c_str(), 0);
        if (rc) { fprintf(stderr, "%s\n", b200r_last_error(hostBvh ? nullptr : ctx)); return 1; }
        printf("BVH ready in %.2f seconds (depth %d)\n", (now_ms() - t0) / 1000., b200r_scene_bvh_depth(scene));
    }

    if (b200r_upload_scene_handle(ctx, scene)) { fprintf(stderr, "%s\n", b200r_last_error(ctx)); return 1; }

    const unsigned nLights = useTwoLights ? 2 : 1;
    if (gpus > 1) {
        // ---- several GPUs: one thread per GPU, each with its own context + pipeline; rank 0 receives the assembled frames
        if (!raytrace) { fprintf(stderr, "--gpus needs a ray-tracing mode (-m 9 or -m 0)\n"); return 1; }
        char uid[128];
        if (b200r_dist_unique_id(uid)) { fprintf(stderr, "%s\n", b200r_last_error(nullptr)); return 1; }
        std::vector<b200r_ctx*> ctxs(gpus, nullptr);
        std::vector<b200r_pipeline*> pipes(gpus, nullptr);
        ctxs[0] = ctx;
        std::atomic<int> failed{0}, ready{0}, done{0};
        std::vector<uint32_t*> host(inFlight, nullptr);
        for (unsigned d = 0; d < inFlight; d++) if (b200r_host_alloc((uint64_t)W * H * 4, (void**)&host[d])) { fprintf(stderr, "%s\n", b200r_last_error(nullptr)); return 1; }
        // the orbit is a float recurrence (renderer.cc:486-493): iterate it once, hand every rank the same frames
        std::vector<b200r_frame> frames(benchmarkFrames);
        { b200r_orbit orbit; b200r_orbit_init(&orbit);
          for (unsigned k = 0; k < benchmarkFrames; k++) {
              float eye[3], mv[9];
              b200r_orbit_step(&orbit, eye, mv);
              b200r_frame_defaults(&frames[k], mode, W, H, eye, mv, nLights);
              frames[k].flags = flags; if (ao) frames[k].ao_samples = ao; frames[k].frame_index = k;
          } }
        double t_start = 0, t_end = 0;
        auto worker = [&](unsigned r) {
            auto die = [&](const char* what) { fprintf(stderr, "rank %u: %s\n", r, what); failed = 1; };
            if (r > 0) {
                if (b200r_init(device + (int)r, &ctxs[r])) { die(b200r_last_error(nullptr)); }
                else if (b200r_upload_scene_handle(ctxs[r], scene)) { die(b200r_last_error(ctxs[r])); }
            }
            if (!failed && b200r_pipeline_create(ctxs[r], W, H, inFlight, r, gpus, uid, assemble, &pipes[r])) die(b200r_last_error(nullptr));
            ready++;
            while (ready < (int)gpus) std::this_thread::yield();
            if (failed) return;
            if (r == 0) t_start = now_ms();
            for (unsigned k = 0; k < benchmarkFrames && !failed; k++)
                if (b200r_pipeline_submit(pipes[r], &frames[k], r == 0 ? host[k % inFlight] : nullptr)) die(b200r_pipeline_last_error(pipes[r]));
            if (pipes[r] && b200r_pipeline_drain(pipes[r])) die(b200r_pipeline_last_error(pipes[r]));
            done++;
            while (done < (int)gpus) std::this_thread::yield();          // nobody tears down before everybody has drained
            if (r == 0) t_end = now_ms();
        };
        std::vector<std::thread> th;
        for (unsigned r = 0; r < gpus; r++) th.emplace_back(worker, r);
        for (auto& t : th) t.join();
        if (failed) return 1;
        if (!dumpPrefix.empty()) {           // the last frame of the run, as rank 0 received it
            const std::string name = dumpPrefix + "_" + std::to_string(benchmarkFrames - 1) + ".xrgb";
            FILE* fp = fopen(name.c_str(), "wb");
            if (!fp) { perror(name.c_str()); return 2; }
            fwrite(host[(benchmarkFrames - 1) % inFlight], 4, (size_t)W * H, fp);
            fclose(fp);
        }
        printf("Rendering %u frames in %g seconds. (%g fps) on %u GPUs\n", benchmarkFrames, (t_end - t_start) / 1000.0,
               benchmarkFrames / ((t_end - t_start) / 1000.0), gpus);
        for (unsigned r = 0; r < gpus; r++) b200r_pipeline_destroy(pipes[r]);
        for (unsigned r = 0; r < gpus; r++) b200r_destroy(ctxs[r]);
        for (auto p : host) b200r_host_free(p);
        b200r_scene_free(scene);
        return 0;
    }
    if (mode == B200R_MODE_PHONG_SHADOWMAPS || mode == B200R_MODE_PHONG_SOFTSHADOWMAPS) {
        // pLight->RenderSceneIntoShadowBuffer(scene) before the loop (renderer.cc:320,325); the light never moves here
        for (unsigned i = 0; i < nLights; i++) {
            float lp[3], w2l[9];
            b200r_default_light_pos((int)i, lp);
            b200r_light_world_to_light(lp, w2l);
            if (b200r_render_shadowmap(ctx, (int)i, lp, w2l)) { fprintf(stderr, "%s\n", b200r_last_error(ctx)); return 1; }
        }
    }

    // Frames of the orbit do not depend on each other, so the loop is pipelined: b200r_render_async returns when frame i is
    // enqueued, its copy-out and its tail overlap the following frames (inFlight + 1 host frames rotate). Dumped frames use the
    // blocking call.
    if (b200r_set_pipeline_depth(ctx, inFlight)) { fprintf(stderr, "%s\n", b200r_last_error(ctx)); return 1; }
    // page-locked host frames: the copy-out is a DMA straight into them (pageable memory would go through a staging copy)
    std::vector<uint32_t*> fb(inFlight + 1, nullptr);
    for (auto& p : fb) if (b200r_host_alloc((uint64_t)W * H * 4, (void**)&p)) { fprintf(stderr, "%s\n", b200r_last_error(nullptr)); return 1; }
    b200r_orbit orbit; b200r_orbit_init(&orbit);
    unsigned framesDrawn = 0;
    double msSpentDrawing = 0, lastReport = now_ms();
    (void)doBenchmark;      // with no window/keyboard, every run follows the benchmark orbit
    while (framesDrawn != benchmarkFrames) {
        float eye[3], mv[9];
        b200r_orbit_step(&orbit, eye, mv);
        b200r_frame f;
        b200r_frame_defaults(&f, mode, W, H, eye, mv, nLights);
        f.flags = flags; if (ao) f.ao_samples = ao; f.frame_index = framesDrawn;
        const bool dump = !dumpPrefix.empty() && (dumpFrames.empty() || dumpFrames.count(framesDrawn));
        uint32_t* out = fb[framesDrawn % (inFlight + 1)];
        const double t0 = now_ms();
        const int rc = dump ? b200r_render(ctx, &f, out) : b200r_render_async(ctx, &f, out);
        if (rc) { fprintf(stderr, "%s\n", b200r_last_error(ctx)); return 1; }
        msSpentDrawing += now_ms() - t0;
        if (dump) {
            const std::string name = dumpPrefix + "_" + std::to_string(framesDrawn) + ".xrgb";
            FILE* fp = fopen(name.c_str(), "wb");
            if (!fp) { perror(name.c_str()); return 2; }
            fwrite(out, 4, (size_t)W * H, fp);
            fclose(fp);
        }
        framesDrawn++;
        if (doReports && now_ms() - lastReport > 5000) {
            lastReport = now_ms();
            if (msSpentDrawing > 0) printf("FPS: %g\n", framesDrawn / (msSpentDrawing / 1000.0));
        }
    }
    {
        const double t0 = now_ms();
        if (b200r_wait(ctx)) { fprintf(stderr, "%s\n", b200r_last_error(ctx)); return 1; }
        msSpentDrawing += now_ms() - t0;
    }
    if (msSpentDrawing > 0)
        printf("Rendering %u frames in %g seconds. (%g fps)\n", framesDrawn, msSpentDrawing / 1000.0,
               framesDrawn / (msSpentDrawing / 1000.0));
    for (auto p : fb) b200r_host_free(p);
    b200r_destroy(ctx);
    b200r_scene_free(scene);
    return 0;
}
