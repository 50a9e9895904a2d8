// cpu_bench.cpp -- frame-parallel CPU driver for the timed CPU baseline (TEST / BENCH INFRASTRUCTURE, never part of the
// product).  Mirrors how the reference's own tool runs the detector over many frames: one detector state per thread, frames
// handed out round-robin (src/app/acf/acf.cpp:443-455), per-stage wall times in the style of ScopeTimeLogger
// (src/app/common/ScopeTimeLogger.h:21-59).  No Python anywhere in the timed loop.
//
// Built by oracle/Makefile three ways from the same sources as the oracle libraries:
//   _ref/cpu_bench_native      reference toolbox objects as shipped (SSE rcpps / rsqrtps), -O2, baseline x86-64 (the reference's
//                              own build flags)  -> the CPU baseline bench.py reports
//   _ref/cpu_bench_native_o3   the same objects with -O3 -mavx2 (courtesy row; -march=native is not portable between the build
//                              container and the GPU box's host CPU)
//   cpu_bench_port             this repo's restatement (only used when /root/reference was absent at build time)
//
// usage: cpu_bench <blob> <threads> <repeats> [frames_for_1thread_row]
// blob (written by bench.py): "ACFB" u32 version=1 | oracle_opts | nTrees nTreeNodes treeDepth (i32) | fids u32[] thrs f32[] child
//                             u32[] hs f32[] | nFrames rows cols (i32) | frames u8 RGB HWC
// prints one JSON object.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <malloc.h>
#include "acf_oracle.h"

namespace
{
enum Stage { S_RGB = 0, S_TRI1, S_GRADMAG, S_TRI5, S_NORM, S_HIST, S_RESAMPLE, S_COUNT };
const char* kStageName[S_COUNT] = { "rgbConvert", "convTri1 (image + final channel smoothing)", "gradMag", "convTri r=5 (normalisation triangle)",
                                    "gradMagNorm", "gradHist", "imResample" };
thread_local double tl_stage[S_COUNT];

struct Scope
{
    Stage s;
    std::chrono::steady_clock::time_point t0;
    explicit Scope(Stage st) : s(st), t0(std::chrono::steady_clock::now()) {}
    ~Scope() { tl_stage[s] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

const OracleL1& base() { return oracle_l1(); }
void t_rgb(float* I, float* J, int n, int d, int flag, float nrm) { Scope sc(S_RGB); base().rgbConvert(I, J, n, d, flag, nrm); }
void t_tri1(float* I, float* O, int h, int w, int d, float p, int s) { Scope sc(S_TRI1); base().convTri1(I, O, h, w, d, p, s); }
void t_tri(float* I, float* O, int h, int w, int d, int r, int s) { Scope sc(S_TRI5); base().convTri(I, O, h, w, d, r, s); }
void t_gm(float* I, float* M, float* O, int h, int w, int d, bool full) { Scope sc(S_GRADMAG); base().gradMag(I, M, O, h, w, d, full); }
void t_gmn(float* M, float* S, int h, int w, float norm) { Scope sc(S_NORM); base().gradMagNorm(M, S, h, w, norm); }
void t_gh(float* M, float* O, float* H, int h, int w, int bin, int nO, int sb, bool full) { Scope sc(S_HIST); base().gradHist(M, O, H, h, w, bin, nO, sb, full); }
void t_rs(float* A, float* B, int ha, int hb, int wa, int wb, int d, float r) { Scope sc(S_RESAMPLE); base().resample(A, B, ha, hb, wa, wb, d, r); }
} // namespace

// acf_oracle.cpp is compiled with -Doracle_l1=oracle_l1_timed for this binary, so the orchestration calls through these wrappers
const OracleL1& oracle_l1_timed()
{
    static const OracleL1 t = { base().kind, t_rgb, t_tri1, t_tri, t_gm, t_gmn, t_gh, t_rs };
    return t;
}

struct Blob
{
    oracle_opts opts;
    oracle_clf clf;
    std::vector<uint32_t> fids, child;
    std::vector<float> thrs, hs;
    int nFrames = 0, rows = 0, cols = 0;
    std::vector<uint8_t> frames;
};

static bool readBlob(const char* path, Blob& b)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char magic[4]; uint32_t ver = 0;
    bool ok = fread(magic, 1, 4, f) == 4 && !memcmp(magic, "ACFB", 4) && fread(&ver, 4, 1, f) == 1 && ver == 1;
    ok = ok && fread(&b.opts, sizeof(b.opts), 1, f) == 1;
    int hdr[3];
    ok = ok && fread(hdr, 4, 3, f) == 3;
    if (ok)
    {
        const size_t n = (size_t)hdr[0] * hdr[1];
        b.fids.resize(n); b.thrs.resize(n); b.child.resize(n); b.hs.resize(n);
        ok = fread(b.fids.data(), 4, n, f) == n && fread(b.thrs.data(), 4, n, f) == n && fread(b.child.data(), 4, n, f) == n && fread(b.hs.data(), 4, n, f) == n;
        b.clf = oracle_clf{ hdr[0], hdr[1], hdr[2], b.fids.data(), b.thrs.data(), b.child.data(), b.hs.data() };
    }
    int fr[3];
    ok = ok && fread(fr, 4, 3, f) == 3;
    if (ok)
    {
        b.nFrames = fr[0]; b.rows = fr[1]; b.cols = fr[2];
        b.frames.resize((size_t)b.nFrames * b.rows * b.cols * 3);
        ok = fread(b.frames.data(), 1, b.frames.size(), f) == b.frames.size();
    }
    fclose(f);
    return ok;
}

struct RunResult { double seconds = 0, pyramid = 0, cascade = 0, stage[S_COUNT] = {}; long long hits = 0, trees = 0; int frames = 0; };

// every frame [0, nFrames) `repeats` times over `threads` threads
static RunResult run(const Blob& b, int nFrames, int threads, int repeats)
{
    std::vector<RunResult> per(threads);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> ts;
    const size_t frameBytes = (size_t)b.rows * b.cols * 3;
    for (int t = 0; t < threads; t++)
        ts.emplace_back([&, t] {
            for (int s = 0; s < S_COUNT; s++) tl_stage[s] = 0;
            RunResult& r = per[t];
            std::vector<oracle_det> out(1 << 16);
            for (int rep = 0; rep < repeats; rep++)
                for (int i = t; i < nFrames; i += threads)
                {
                    const auto a = std::chrono::steady_clock::now();
                    void* P = oracle_pyramid_create(&b.opts, b.frames.data() + (size_t)i * frameBytes, b.rows, b.cols, 0, nullptr, nullptr);
                    const auto c = std::chrono::steady_clock::now();
                    if (!P) { fprintf(stderr, "cpu_bench: %s\n", oracle_last_error()); exit(2); }
                    int total = 0; uint64_t ne = 0;
                    oracle_detect(P, &b.opts, &b.clf, out.data(), (int)out.size(), &total, nullptr, nullptr, nullptr, &ne);
                    oracle_pyramid_destroy(P);
                    const auto d = std::chrono::steady_clock::now();
                    r.pyramid += std::chrono::duration<double>(c - a).count();
                    r.cascade += std::chrono::duration<double>(d - c).count();
                    r.hits += total; r.trees += (long long)ne; r.frames++;
                }
            for (int s = 0; s < S_COUNT; s++) r.stage[s] = tl_stage[s];
        });
    for (auto& t : ts) t.join();
    RunResult R;
    R.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& r : per)
    {
        R.pyramid += r.pyramid; R.cascade += r.cascade; R.hits += r.hits; R.trees += r.trees; R.frames += r.frames;
        for (int s = 0; s < S_COUNT; s++) R.stage[s] += r.stage[s];
    }
    return R;
}

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: cpu_bench <blob> <threads> <repeats> [frames_for_1thread_row]\n"); return 1; }
    // keep freed plane buffers in the process instead of returning them to the kernel after every frame (glibc would mmap / munmap
    // each 8 MB plane and page-fault it in again): the allocator setting a production deployment of the CPU path would use
    mallopt(M_MMAP_THRESHOLD, 1 << 30); mallopt(M_TRIM_THRESHOLD, 1 << 30);
    Blob b;
    if (!readBlob(argv[1], b)) { fprintf(stderr, "cpu_bench: cannot read %s\n", argv[1]); return 1; }
    const int threads = std::max(1, atoi(argv[2])), repeats = std::max(1, atoi(argv[3]));
    const int n1 = argc > 4 ? std::max(0, std::min(b.nFrames, atoi(argv[4]))) : 0;
    run(b, std::min(b.nFrames, threads), threads, 1); // warm-up: page in, spin up
    RunResult one;
    if (n1 > 0) one = run(b, n1, 1, 1);
    const RunResult R = run(b, b.nFrames, threads, repeats);
    printf("{\"kind\": \"%s\", \"threads\": %d, \"frames\": %d, \"seconds\": %.4f, \"fps\": %.4f, \"fps_per_thread\": %.4f, \"hits\": %lld, \"trees\": %lld",
           oracle_kind(), threads, R.frames, R.seconds, R.frames / R.seconds, R.frames / R.seconds / threads, R.hits, R.trees);
    if (n1 > 0) printf(", \"fps_1thread\": %.4f, \"frames_1thread\": %d", one.frames / one.seconds, one.frames);
    // per-frame stage times: thread-seconds summed over all threads / frames (what one core spends on one frame)
    printf(", \"stage_ms\": {");
    double l1 = 0;
    for (int s = 0; s < S_COUNT; s++) { printf("\"%s\": %.3f, ", kStageName[s], 1e3 * R.stage[s] / R.frames); l1 += R.stage[s]; }
    printf("\"other pyramid work (transpose, u8->f32, padding, copies)\": %.3f, \"acfDetect (cascade)\": %.3f, \"frame total\": %.3f}}\n",
           1e3 * (R.pyramid - l1) / R.frames, 1e3 * R.cascade / R.frames, 1e3 * (R.pyramid + R.cascade) / R.frames);
    return 0;
}
