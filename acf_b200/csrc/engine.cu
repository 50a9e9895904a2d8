// engine.cu -- device engine + the C ABI of include/acf_b200.h.
//
// One engine = one CUDA device, one stream, persistent buffers sized at first use for a frame size
// and batch (no per-frame allocation in steady state).  Host work kept from the reference because
// it is scalar fp64 bookkeeping: scale schedule (plan.cpp), hit ordering, box rescale
// (ACF.cpp:302-311), bbNms / prune (bbNms.cpp:111-304, ObjectDetector.cpp:28-44).
// No CPU fallback exists: without a CUDA device acfb_engine_create fails.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "kernels.cuh"
#include "model.h"
#include "plan.h"
#include "dist.h"

namespace acfb
{

static thread_local std::string g_err;

#define CUDA_OK(x)                                                                                         \
    do                                                                                                     \
    {                                                                                                      \
        cudaError_t e_ = (x);                                                                              \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #x); \
    } while (0)

#define NCCL_OK(x)                                                                                         \
    do                                                                                                     \
    {                                                                                                      \
        int r_ = (x);                                                                                      \
        if (r_ != 0) throw std::runtime_error(std::string("NCCL: ") + NcclApi::get().GetErrorString(r_) + " at " #x); \
    } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
static PFN_cuTensorMapEncodeTiled_v12000 tensorMapEncoder()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) throw std::runtime_error("engine: cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// 4-D map of one scale's channel block: dims y (contiguous) | x | channel | frame, box = one cascade tile (all channels)
static CUtensorMap channelTileMap(const float* base, int H, int W, int nChns, int nFrames, int64_t pitch, int64_t planeStride, int64_t frameStride,
                                  const CascTileGeom& g)
{
    CUtensorMap m;
    const cuuint64_t dims[4] = { (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)nChns, (cuuint64_t)std::max(1, nFrames) };
    const cuuint64_t strides[3] = { (cuuint64_t)pitch * 4, (cuuint64_t)planeStride * 4, (cuuint64_t)std::max<int64_t>(frameStride, 4) * 4 };
    const cuuint32_t box[4] = { (cuuint32_t)g.BY, (cuuint32_t)g.BX, (cuuint32_t)nChns, 1u };
    const cuuint32_t es[4] = { 1u, 1u, 1u, 1u };
    const CUresult r = tensorMapEncoder()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("engine: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

// 3-D map of a full-resolution float plane set [frame][x][y] (y contiguous): box = 32 rows x 32 columns of one frame, 128-byte
// swizzle (k_triyhist_tma).  false: the layout does not meet the copy engine's 16-byte stride / address rules.
static bool planeChunkMap(const float* base, int H, int W, int nFrames, int64_t frameStride, CUtensorMap& m)
{
    if ((H & 3) || (frameStride & 3) || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    const cuuint64_t dims[3] = { (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)std::max(1, nFrames) };
    const cuuint64_t strides[2] = { (cuuint64_t)H * 4, (cuuint64_t)std::max<int64_t>(frameStride, 4) * 4 };
    const cuuint32_t box[3] = { 32u, 32u, 1u };
    const cuuint32_t es[3] = { 1u, 1u, 1u };
    return tensorMapEncoder()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    void ensure(size_t count)
    {
        if (count <= n) return;
        release();
        CUDA_OK(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

struct AxisUpload
{
    DevBuf<int> ints;    // start | cnt
    DevBuf<float> wts;
    AxisDev dev{};
    void upload(const AxisCoef& c, cudaStream_t s)
    {
        if (c.nOut == 0) { dev = AxisDev{}; return; }
        ints.ensure(2 * (size_t)c.nOut);
        wts.ensure(c.wt.size());
        CUDA_OK(cudaMemcpyAsync(ints.p, c.start.data(), c.nOut * sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaMemcpyAsync(ints.p + c.nOut, c.cnt.data(), c.nOut * sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaMemcpyAsync(wts.p, c.wt.data(), c.wt.size() * sizeof(float), cudaMemcpyHostToDevice, s));
        CUDA_OK(cudaStreamSynchronize(s));
        dev.start = ints.p; dev.cnt = ints.p + c.nOut; dev.wt = wts.p;
        dev.nOut = c.nOut; dev.nIn = c.nIn; dev.mode = c.mode; dev.ymul = c.ymul;
    }
};

// everything that depends on the frame size
struct SizeState
{
    Plan plan;
    std::vector<std::unique_ptr<AxisUpload>> realAx;  // 2 per real scale (x, y)
    std::vector<std::unique_ptr<AxisUpload>> scaleAx; // 2 per scale
    DevBuf<AxisDev> axes;                              // [2*nScales]
    std::vector<ChanJob> chanJobsHost;
    DevBuf<ChanJob> chanJobs;
    std::vector<PadJob> padJobsHost;
    DevBuf<PadJob> padJobs;
    int64_t padTotal = 0;
    std::vector<CascScale> cascHost;
    DevBuf<CascScale> casc;
    int cascBlocksPerFrame = 0;
    std::vector<CascTileScale> ctHost;  // k_cascade_tile: per-scale tile grids (tile0 counts from the scale's octave group)
    DevBuf<CascTileScale> ct;
    DevBuf<PostScale> postScales;       // scale, scaleshw per pyramid scale (k_post)
    DevBuf<CUtensorMap> tmaps;          // one 4-D tensor map per scale over the resident pyramid (re-encoded when it is re-allocated)
    // per octave group (= real scale): scale range, k_chan job range, k_pad job range, cascade task count
    // jobBeg..jobEnd: planes of at most 128 rows (independent warps); mJobBeg..mJobEnd: taller planes, mWarps jobs per plane
    // mRanges: the taller planes, one job range per strip count (a block = the strips of one plane, so planes of two strips run as
    // 64-thread blocks and planes of three as 96-thread blocks: no idle padding warp holds registers and a warp slot)
    struct MRange { int beg, end, warps; };
    struct Group { int sBeg = 0, sEnd = 0, jobBeg = 0, jobEnd = 0, mJobBeg = 0, mJobEnd = 0, mWarps = 0, padBeg = 0, padEnd = 0, cascTasks = 0, cascTiles = 0; int64_t padTotal = 0; std::vector<MRange> mRanges; };
    std::vector<Group> groups;
    uint64_t windowsPerFrame = 0;
    std::vector<int64_t> realOff; // float offset of each real scale's channel block inside a frame's R block
    std::vector<int64_t> moOff;   // float offset of each real scale's full-resolution plane inside a frame's M / O / U block
    int64_t moFloatsPerFrame = 0;
    DevBuf<float> gM, gU;         // raw gradient magnitude, x pass of the normalisation triangle
    DevBuf<uint16_t> gO;          // orientation as acos-table index (GradArgs::outO)
    int64_t rFloatsPerFrame = 0;
    int lastN = 0, lastLanes = 0; // frame -> lane mapping of the previous submitted batch (see Engine::submitAll)
    // batch-sized buffers
    int batchCap = 0;
    DevBuf<float> resT;           // scratch of the two-pass real-scale resample: per frame max over real scales of planes * w * srcH
    int64_t resTFloatsPerFrame = 0;
    DevBuf<float> I0;
    std::vector<std::unique_ptr<DevBuf<float>>> In, C; // per real scale
    DevBuf<float> R, pyr;
};

struct Engine
{
    Model model;
    acfb_options opt{};
    int device = 0;
    int maxRows = 0, maxCols = 0, maxBatch = 0;
    cudaStream_t stream = nullptr;
    DevBuf<float> lut, acosTab;
    DevBuf<float> opA, opB, opC, opS;  // scratch planes of the stand-alone operators (acfb_op_*)
    DevBuf<uint16_t> opO;
    DevBuf<uint32_t> cascTab, cascTabU8;
    int recWords = 0;
    // k_cascade_tile (depth-2 float models whose window fits a shared-memory tile): geometry + tile-local tree table
    bool tileCascade = false;
    CascTileGeom tileGeom{};
    DevBuf<uint32_t> cascTabTile;
    DevBuf<uint32_t> cascTabSoA;  // the same records as three arrays with window-local offsets (k_cascade_tail_win)
    CascTileGeom winGeom{};       // BY x BX = one window's footprint: the box of the window tensor maps
    bool tailOnWindows = false;   // the hand-over is finished by k_cascade_tail_win (else k_cascade_tail)
    std::vector<CascHeadRec> tileHead; // records of the first kCascHeadTrees trees (kernel parameters), empty for shorter models
    DevBuf<CUtensorMap> scratchMaps;
    DevBuf<CascTileScale> scratchTileScale;
    cudaStream_t copyStream = nullptr;
    // multi-GPU: frames shard by batch over ranks; the only exchange is the gather of the boxes k_post leaves
    NcclComm comm = nullptr;
    ShmExchange* exch = nullptr;       // single-node exchange through shared memory (the default); comm stays null then
    bool distEarly = false;            // ACFB_DIST_EARLY=1: enqueue the gather at submit time (behind k_post) instead of at collect time: the
                                       // collective's kernel then spins on an SM until the slowest rank reaches the same batch
    bool ownsComm = true;              // false: communicator and communication stream belong to the engine's first pipeline
    int distRank = 0, distWorld = 1;
    cudaStream_t commStream = nullptr; // the gather of batch k runs here while the kernels of batch k+1 run on the engine's streams
    cudaStream_t finStream = nullptr; // joins the lanes of a submitted batch, reads its counters back and signals Slot::done
    // A lane is a pair of compute streams: `a` runs colour + real-scale kernels, `b` runs the final channels + cascade of
    // octave group k as soon as real scale k is done (overlapping the real-scale kernels of group k+1).  A batch is split
    // over nLanes lanes (lane 0's `a` is the engine's main stream) so that independent kernels fill idle issue slots.
    struct Lane
    {
        cudaStream_t a = nullptr, b = nullptr, c = nullptr; // a: colour + gradient chain, b: final channels + cascade, c: image resample + smoothing chain
        std::vector<cudaEvent_t> evReal, evSmooth, evChan;
        cudaEvent_t evColor = nullptr;
        cudaEvent_t evB = nullptr, evStart = nullptr, evEnd = nullptr;
    };
    static constexpr int kMaxLanes = 4;
    Lane lanes[kMaxLanes];
    // Measured in round 2 (profiles/r2_overlap.md): every kernel of the path fills the SMs it runs on, so kernels of different
    // streams do not co-reside and the lane / stream split only adds scheduling gaps (22.3 ms per step serial, 23.9 with three
    // streams, 24.2 with two lanes x three streams).  Default: one stream; ACFB_OVERLAP=1 / ACFB_LANES=n re-enable the split.
    int nLanes = 1;
    bool overlap = false;
    int pixfmt = 0; // acfb_set_input_format: 0 RGB24, 1 BGR24, 2 RGBA32, 3 BGRA32, 4 GRAY8, 5 RGB32F, 6 PLANAR32F, 7 NV12
    int triyBlocksPerSm = 3; // ACFB_TRIY_BPS (k_triyhist_tma: at most 3; k_triyhist: 2 fit)
    bool triyTma = true;     // ACFB_TRIY_TMA=0: k_triyhist (register-staged loads) instead of k_triyhist_tma
    // L2 prefetch distance (columns past the register banks) of the marching kernels.  Measured (256 frames in flight):
    // k_smooth needs it (its eight-step banks do not cover the loaded DRAM latency: 1.47 ms without, 1.05 ms with), but far
    // ahead the lines are evicted again before use (64 columns: DRAM reads 2x the plane); k_trix is faster without (1.13 -> 0.82 ms)
    int marchPrefetch = 16;  // ACFB_MARCH_PF
    int trixPrefetch = 0;    // ACFB_TRIX_PF
    int triyFastScan = 1;    // ACFB_TRIY_FAST
    int gradCols = 8;        // ACFB_GRAD_COLS
    bool twoPassResample = true; // ACFB_RESAMPLE_2PASS
    bool devicePost = true;  // ACFB_DEVICE_POST=0: ordering, rescale, bbNms and prune on the host for every batch
    int postOut = 16;        // records per frame k_post writes (min(maxDet, 64) rounded up)
    bool useFront = true;    // ACFB_FRONT=0: k_smooth + k_gradmag + k_trix instead of the fused march k_front
    bool keepC = false;      // acfb_set_debug_taps: keep every real scale's smoothed image for acfb_tap("C")
    bool fuseDown2 = true;   // ACFB_FUSE_DOWN2: k_smooth also writes the half-resolution image of the next octave
    int cascBlocksPerSm = 0; // ACFB_CASC_BPS
    int cascSparseMax = 16;     // ACFB_CASC_SPARSE: see CascTileArgs::sparseMax (only reached when the hand-over list is full)
    int cascExportMax = 48;     // ACFB_CASC_EXPORT: see CascTileArgs::exportMax
    int cascHeadLevels = 3;     // ACFB_CASC_HEAD_LEVELS: 5 or 3, see CascTileArgs::headLevels
    bool cascTailOnWin = true;  // ACFB_CASC_TAIL_WIN=0: finish the hand-over with global gathers (k_cascade_tail) instead of TMA-staged window footprints
    int cascTailCap = 1 << 18;  // hand-over entries per cascade launch (4 MB)
    bool useTileCascade = true; // ACFB_CASC_TILE=0: every model through the global-gather kernel (k_cascade)
    int cascPrefetch = 1; // ACFB_CASC_PF: 0 none, 1 L2 (default, -2 % cascade time), 2 L1 (see CascArgs::prefetch)
    bool isTranspose = false, isLuv = false; // Detector::setIsTranspose / setIsLuv (ACF.h:560-576)
    int bpp() const { return pixfmt <= 1 ? 3 : pixfmt <= 3 ? 4 : pixfmt == 4 ? 1 : pixfmt == 7 ? 1 : 12; }
    // bytes of one frame in the current input format (NV12: a luma byte per pixel + a (U, V) byte pair per 2 x 2 pixels)
    size_t frameBytes(int rows, int cols) const { return pixfmt == 7 ? (size_t)rows * cols * 3 / 2 : (size_t)rows * cols * bpp(); }
    // rgbConvert's dispatch (rgbConvert.cpp:102-170, chnsPyramid.cpp:231-261) for the current input format
    int colorMode() const
    {
        const int cs = opt.color_space;
        if (pixfmt == 4 && cs != 0 && cs != 4) throw std::runtime_error("engine: single-channel frames need colorSpace gray or orig (rgbConvert.cpp:139-147)");
        if (isLuv)
        {
            if (cs != 2) throw std::runtime_error("engine: setIsLuv needs a luv model (rgbConvert.cpp:150-155)");
            if (pixfmt == 4) throw std::runtime_error("engine: setIsLuv needs three input planes");
            return 1;
        }
        return cs == 0 ? 0 : cs == 2 ? 2 : cs == 3 ? 3 : 1;
    }
    std::map<std::pair<int, int>, std::unique_ptr<SizeState>> sizes;
    SizeState* cur = nullptr;
    int curN = 0;
    // hits
    int hitCap = 4096;
    // Up to three batches may be in flight (submit k+1, k+2 while batch k computes): everything a batch owns until it is
    // collected lives in a slot -- the H2D staging buffer, hit counters / records and their pinned host mirrors.
    struct Slot
    {
        DevBuf<uint8_t> frames;
        DevBuf<int> hitCount;
        DevBuf<int4> hits;
        DevBuf<unsigned long long> stats;
        int* hCount = nullptr;              // pinned
        unsigned long long* hStats = nullptr; // pinned
        int hCountCap = 0;
        cudaEvent_t copied = nullptr, done = nullptr;
        SizeState* st = nullptr;
        int n = 0;
        int nextCounter = 0; // task counters handed to the cascade launches of this batch
        size_t statsWords = 64;
        // k_post output: one packed buffer [int32 detCount[n] | int32 fallback, pad | PostDet dets[n][postOut]] and its pinned mirror
        DevBuf<uint8_t> post;
        uint8_t* hPost = nullptr;
        size_t postBytes = 0, hPostCap = 0;
        bool posted = false;     // this batch ran k_post
        // multi-GPU: every rank's `post` buffer, all-gathered on commStream, and its pinned mirror
        DevBuf<uint8_t> gath;
        uint8_t* hGath = nullptr;
        size_t hGathCap = 0;
        cudaEvent_t evPost = nullptr, gathDone = nullptr;
        bool gathered = false;   // the all-gather of this batch has been enqueued
        DevBuf<int4> tail;       // k_cascade_tile -> k_cascade_tail hand-over lists, tailCap entries per cascade launch
        DevBuf<int> tailCount;   // one per cascade launch
        bool pending = false;
    };
    static constexpr int kSlots = 3; // batches that may be in flight: one computing, one copying in, one being collected
    Slot slots[kSlots];
    int subSlot = 0, colSlot = 0;
    bool anyPending() const { for (auto& s : slots) if (s.pending) return true; return false; }
    cudaStream_t d2hStream = nullptr;
    DevBuf<int> scratchCount;
    DevBuf<unsigned long long> scratchStats;
    std::vector<int> hHitCount;
    std::vector<int4> hHits;
    unsigned long long hStats[2] = { 0, 0 };
    std::vector<acfb_hit> lastHits;
    std::vector<unsigned char> distBuf; // this rank's record for the shared-memory exchange
    bool lastHitsValid = true;    // false: the last batch was finished by k_post, only the raw hit count is known
    long long lastHitTotal = 0;
    // options of ObjectDetector
    bool doNms = false;
    int maxDet = 10;
    double pruneRatio = 0.0;
    uint64_t launches = 0;
    double collectWaitMs = 0, collectTailMs = 0; // host wall time of the last collect (acfb_collect_times)
    // stage timing
    bool timing = false;
    std::vector<cudaEvent_t> evs;
    std::vector<const char*> evNames;
    std::vector<float> stageMs;
    std::vector<const char*> stageNames;
    std::vector<double> lambdasFromImage;
    // scratch for acfb_acf_detect1
    DevBuf<float> scratch;
    DevBuf<CascScale> scratchScale;
    DevBuf<int4> scratchHits, scratchTail;

    // k_post serves bbNms "max" / "maxg" with 1..64 reported boxes, at most 64 scales and window grids below 8192 (its sort key
    // packs scale | c | r into 32 bits); everything else -- and raw-hit output (NMS off) -- stays on the host
    bool postOnDevice(const SizeState& st) const
    {
        if (!devicePost || !doNms || maxDet < 1 || maxDet > 64 || timing) return false;
        const std::string type = opt.nms_type;
        if (type != "max" && type != "maxg") return false;
        if (st.plan.geom.size() > 64) return false;
        for (const CascScale& c : st.cascHost) if (c.width1 >= 8192 || c.height1 >= 8192) return false;
        return true;
    }
    static size_t postHeaderBytes(int n) { return ((size_t)(n + 1) * sizeof(int) + 15) & ~(size_t)15; }
    void launchPostKernel(SizeState& st, Slot& S, int f0, int n, cudaStream_t s)
    {
        PostArgs p{};
        p.hits = S.hits.p + (size_t)f0 * hitCap; p.hitCount = S.hitCount.p + f0; p.hitCap = hitCap; p.n = n; p.frame0 = f0;
        int cap2 = 1;
        while (cap2 < hitCap && cap2 < 4096) cap2 <<= 1;
        p.cap2 = cap2;
        p.scales = st.postScales.p; p.stride = opt.stride; p.modelDs_w = opt.modelDs_w; p.modelDs_h = opt.modelDs_h;
        p.shift_w = (opt.modelDsPad_w - opt.modelDs_w) / 2 - opt.pad_w; p.shift_h = (opt.modelDsPad_h - opt.modelDs_h) / 2 - opt.pad_h;
        p.greedy = std::string(opt.nms_type) == "maxg"; p.ovrUnion = std::string(opt.nms_ovrDnm) != "min";
        p.maxDet = maxDet; p.maxOut = postOut; p.overlap = opt.nms_overlap; p.pruneRatio = pruneRatio;
        int* hdr = reinterpret_cast<int*>(S.post.p);
        p.detCount = hdr + f0; p.fallback = hdr + S.n;
        p.dets = reinterpret_cast<PostDet*>(S.post.p + postHeaderBytes(S.n)) + (size_t)f0 * postOut;
        launchPost(p, s); launches++;
    }

    // every stream of the engine (submitted batches run on the lanes' streams and finish on finStream)
    void syncAll()
    {
        for (int l = 0; l < kMaxLanes; l++)
            for (cudaStream_t q : { lanes[l].a, lanes[l].b, lanes[l].c })
                if (q) CUDA_OK(cudaStreamSynchronize(q));
        for (cudaStream_t q : { copyStream, finStream, d2hStream })
            if (q) CUDA_OK(cudaStreamSynchronize(q));
    }

    ~Engine()
    {
        cudaSetDevice(device);
        try { syncAll(); } catch (...) {}
        for (auto e : evs) cudaEventDestroy(e);
        for (auto& s : slots)
        {
            if (s.copied) cudaEventDestroy(s.copied);
            if (s.done) cudaEventDestroy(s.done);
            if (s.hCount) cudaFreeHost(s.hCount);
            if (s.hStats) cudaFreeHost(s.hStats);
            if (s.hPost) cudaFreeHost(s.hPost);
            if (s.hGath) cudaFreeHost(s.hGath);
            if (s.evPost) cudaEventDestroy(s.evPost);
            if (s.gathDone) cudaEventDestroy(s.gathDone);
        }
        if (exch && ownsComm) delete exch;
        if (comm && ownsComm) { try { NcclApi::get().CommDestroy(comm); } catch (...) {} }
        if (commStream && ownsComm) cudaStreamDestroy(commStream);
        if (d2hStream) cudaStreamDestroy(d2hStream);
        if (copyStream) cudaStreamDestroy(copyStream);
        if (finStream) cudaStreamDestroy(finStream);
        for (int l = 0; l < kMaxLanes; l++)
        {
            Lane& L = lanes[l];
            for (auto ev : L.evReal) cudaEventDestroy(ev);
            for (auto ev : L.evSmooth) cudaEventDestroy(ev);
            for (auto ev : L.evChan) cudaEventDestroy(ev);
            if (L.evColor) cudaEventDestroy(L.evColor);
            if (L.c) cudaStreamDestroy(L.c);
            if (L.evB) cudaEventDestroy(L.evB);
            if (L.evStart) cudaEventDestroy(L.evStart);
            if (L.evEnd) cudaEventDestroy(L.evEnd);
            if (L.b) cudaStreamDestroy(L.b);
            if (l > 0 && L.a) cudaStreamDestroy(L.a);
        }
        if (stream) cudaStreamDestroy(stream);
    }

    void mark(const char* name)
    {
        if (!timing) return;
        cudaEvent_t e;
        CUDA_OK(cudaEventCreate(&e));
        CUDA_OK(cudaEventRecord(e, stream));
        evs.push_back(e);
        evNames.push_back(name);
    }

    void init()
    {
        CUDA_OK(cudaSetDevice(device));
        CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&finStream, cudaStreamNonBlocking));
        if (const char* ov = getenv("ACFB_OVERLAP")) overlap = atoi(ov) != 0;
        if (const char* nl = getenv("ACFB_LANES")) nLanes = std::max(1, std::min(kMaxLanes, atoi(nl)));
        if (const char* mp = getenv("ACFB_MARCH_PF")) marchPrefetch = std::max(0, std::min(256, atoi(mp)));
        if (const char* r2 = getenv("ACFB_RESAMPLE_2PASS")) twoPassResample = atoi(r2) != 0;
        if (const char* gc = getenv("ACFB_GRAD_COLS")) gradCols = atoi(gc);
        if (const char* tf = getenv("ACFB_TRIY_FAST")) triyFastScan = atoi(tf) != 0;
        if (const char* tp = getenv("ACFB_TRIX_PF")) trixPrefetch = std::max(0, std::min(256, atoi(tp)));
        if (const char* fd = getenv("ACFB_FUSE_DOWN2")) fuseDown2 = atoi(fd) != 0;
        if (const char* fr = getenv("ACFB_FRONT")) useFront = atoi(fr) != 0;
        if (const char* dp = getenv("ACFB_DEVICE_POST")) devicePost = atoi(dp) != 0;
        if (const char* tt = getenv("ACFB_TRIY_TMA")) triyTma = atoi(tt) != 0;
        if (const char* tb = getenv("ACFB_TRIY_BPS")) triyBlocksPerSm = std::max(1, std::min(4, atoi(tb)));
        if (const char* bp = getenv("ACFB_CASC_BPS")) cascBlocksPerSm = std::max(0, std::min(3, atoi(bp)));
        if (const char* pf = getenv("ACFB_CASC_PF")) cascPrefetch = std::max(0, std::min(2, atoi(pf)));
        if (const char* tc = getenv("ACFB_CASC_TILE")) useTileCascade = atoi(tc) != 0;
        if (const char* sp = getenv("ACFB_CASC_SPARSE")) cascSparseMax = std::max(0, atoi(sp));
        if (const char* ex = getenv("ACFB_CASC_EXPORT")) cascExportMax = std::max(0, atoi(ex));
        if (const char* tt = getenv("ACFB_CASC_TAIL_WIN")) cascTailOnWin = atoi(tt) != 0;
        if (const char* de = getenv("ACFB_DIST_EARLY")) distEarly = atoi(de) != 0;
        if (const char* hl = getenv("ACFB_CASC_HEAD_LEVELS")) cascHeadLevels = atoi(hl) == 3 ? 3 : 5;
        for (int l = 0; l < kMaxLanes; l++)
        {
            Lane& L = lanes[l];
            if (l == 0) L.a = stream; else CUDA_OK(cudaStreamCreateWithFlags(&L.a, cudaStreamNonBlocking));
            CUDA_OK(cudaStreamCreateWithFlags(&L.b, cudaStreamNonBlocking));
            CUDA_OK(cudaStreamCreateWithFlags(&L.c, cudaStreamNonBlocking));
            CUDA_OK(cudaEventCreateWithFlags(&L.evColor, cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&L.evB, cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&L.evStart, cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&L.evEnd, cudaEventDisableTiming));
        }
        CUDA_OK(cudaStreamCreateWithFlags(&d2hStream, cudaStreamNonBlocking));
        for (auto& s : slots)
        {
            CUDA_OK(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
            CUDA_OK(cudaMallocHost(&s.hStats, 2 * sizeof(unsigned long long)));
            s.stats.ensure(64);
        }
        // L lookup table, rgbConvertMex.cpp:20-59 (host pow, exactly as the reference builds it)
        {
            std::vector<float> t(1064);
            const float y0 = (float)((6.0 / 29) * (6.0 / 29) * (6.0 / 29)), aa = (float)((29.0 / 3) * (29.0 / 3) * (29.0 / 3));
            const float maxi = (float)1.0 / 270;
            for (int i = 0; i < 1025; i++)
            {
                const float y = (float)(i / 1024.0);
                const float l = y > y0 ? 116 * (float)pow((double)y, 1.0 / 3.0) - 16 : y * aa;
                t[i] = l * maxi;
            }
            for (int i = 1025; i < 1064; i++) t[i] = t[i - 1];
            lut.ensure(t.size());
            CUDA_OK(cudaMemcpy(lut.p, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
        // acos table, gradientMex.cpp:103-165
        {
            const int n = 10000, b = 10;
            std::vector<float> t(2 * (n + b));
            float* a1 = t.data() + n + b;
            const float PI = 3.14159265f;
            for (int i = -n - b; i < -n; i++) a1[i] = PI;
            for (int i = -n; i < n; i++) a1[i] = float(std::acos(i / float(n)));
            for (int i = n; i < n + b; i++) a1[i] = 0;
            for (int i = -n - b; i < n / 10; i++)
                if (a1[i] > PI - 1e-6f) a1[i] = PI - 1e-6f;
            acosTab.ensure(t.size());
            CUDA_OK(cudaMemcpy(acosTab.p, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
        buildCascadeTable();
        scratchCount.ensure(2); // [0] hits, [1] hand-over entries of acfb_acf_detect1
        scratchStats.ensure(4);
    }

    // per tree: (2^D - 1) x {packed (z, c, r), threshold bits}, then 2^D leaf outputs.
    // feature id -> (z, c, r): acfDetect1.cpp:390-406 (r fastest, then c, then z).
    void buildCascadeTable()
    {
        const int D = model.clf.treeDepth;
        const int nT = model.nTrees(), nN = model.nTreeNodes();
        const int mH = opt.modelDsPad_w / opt.shrink; // rows of the window in channel px (orig y)
        const int mW = opt.modelDsPad_h / opt.shrink;
        const uint32_t* fids = model.clf.fids.ptr<uint32_t>();
        const float* thrs = model.clf.thrs.ptr<float>();
        const float* hs = model.clf.hs.ptr<float>();
        const uint32_t* child = model.clf.child.ptr<uint32_t>();
        std::vector<uint32_t> t;
        auto node = [&](uint32_t* dst, size_t idx) {
            const uint32_t fid = fids[idx];
            dst[0] = fid / (mH * mW);      // z
            dst[1] = (fid / mH) % mW;      // c
            dst[2] = fid % mH;             // r
            memcpy(&dst[3], &thrs[idx], 4);
        };
        if (D == 0)
        {   // variable depth: all N nodes {z,c,r,thr}, N outputs, N child links, node count in the last word
            recWords = (6 * nN + 1 + 3) & ~3;
            t.assign((size_t)nT * recWords, 0u);
            for (int i = 0; i < nT; i++)
            {
                uint32_t* rec = &t[(size_t)i * recWords];
                for (int k = 0; k < nN; k++)
                {
                    const size_t idx = (size_t)i * nN + k;
                    if (child[idx] && !(child[idx] > (uint32_t)k + 1 && child[idx] <= (uint32_t)nN - 1)) throw std::runtime_error("model: child link does not point forward inside the tree");
                    if (child[idx]) node(&rec[4 * k], idx);
                    memcpy(&rec[4 * nN + k], &hs[idx], 4);
                    rec[5 * nN + k] = child[idx];
                }
                rec[recWords - 1] = (uint32_t)nN;
            }
        }
        else
        {
            const int nInt = (1 << D) - 1, nLeaf = 1 << D;
            recWords = (4 * nInt + nLeaf + 3) & ~3; // internal nodes {z, c, r, thr} (16 B each), then the leaf outputs
            t.assign((size_t)nT * recWords, 0u);
            for (int i = 0; i < nT; i++)
            {
                uint32_t* rec = &t[(size_t)i * recWords];
                for (int k = 0; k < nInt; k++) node(&rec[4 * k], (size_t)i * nN + k);
                for (int k = 0; k < nLeaf; k++) memcpy(&rec[4 * nInt + k], &hs[(size_t)i * nN + nInt + k], 4);
            }
        }
        cascTab.ensure(t.size());
        CUDA_OK(cudaMemcpy(cascTab.p, t.data(), t.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        // byte-channel variant: thresholds pre-scaled like thrs.convertTo(thrsU8, CV_8UC1, 255.0f) (ACFIOArchive.h:96-99):
        // saturate_cast<uchar>(cvRound(thr * 255)), compared as float(ftr) < float(thrU8) (acfDetect1.cpp:72-82,157-166)
        {
            std::vector<uint32_t> t8 = t;
            const int nThr = (D == 0) ? nN : (1 << D) - 1;
            for (int i = 0; i < nT; i++)
                for (int k = 0; k < nThr; k++)
                {
                    float thr;
                    memcpy(&thr, &t8[(size_t)i * recWords + 4 * k + 3], 4);
                    const float q = std::nearbyint(thr * 255.0f); // float arithmetic + cvRound, as cv::Mat::convertTo does for CV_32F -> CV_8U
                    const float u = q < 0 ? 0.f : q > 255 ? 255.f : q;
                    memcpy(&t8[(size_t)i * recWords + 4 * k + 3], &u, 4);
                }
            cascTabU8.ensure(t8.size());
            CUDA_OK(cudaMemcpy(cascTabU8.p, t8.data(), t8.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
        // k_cascade_tile's table: node -> byte offset inside a shared-memory tile (pitch BY, plane BX * BY), so a gather is
        // base(window) + offset(node).  Only depth-2 trees, windows that fit a tile, strides that are multiples of shrink.
        tileCascade = false;
        if (useTileCascade && D == 2 && opt.stride % opt.shrink == 0 && mH >= 1 && mW >= 1 &&
            cascTileGeometry(mH, mW, (opt.color_enabled ? (opt.color_space == 0 ? 1 : 3) : 0) + 1 + opt.gh_nOrients, opt.stride / opt.shrink, cascBlocksPerSm, tileGeom))
        {
            const int rw = cascTileRecWords();
            std::vector<uint32_t> tt((size_t)nT * rw, 0u);
            for (int i = 0; i < nT; i++)
            {
                uint32_t* rec = &tt[(size_t)i * rw];
                for (int k = 0; k < 3; k++)
                {
                    const size_t idx = (size_t)i * nN + k;
                    const uint32_t fid = fids[idx];
                    const uint32_t z = fid / (mH * mW), c = (fid / mH) % mW, r = fid % mH;
                    if ((int)z >= tileGeom.nChns) throw std::runtime_error("model: feature id outside the channel planes");
                    rec[k] = 4u * ((z * (uint32_t)tileGeom.BX + c) * (uint32_t)tileGeom.BY + r);
                    memcpy(&rec[k == 0 ? 3 : 3 + k], &thrs[idx], 4);
                }
                for (int k = 0; k < 4; k++) memcpy(&rec[6 + k], &hs[(size_t)i * nN + 3 + k], 4);
            }
            cascTabTile.ensure(tt.size());
            CUDA_OK(cudaMemcpy(cascTabTile.p, tt.data(), tt.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            {   // k_cascade_tail_win's table: the same records as three arrays (lanes = trees: a warp reads 32 consecutive trees), offsets
                // local to ONE window's footprint (mHp >= mH + 3 rows x mW columns per channel, the box of the window tensor maps)
                winGeom = tileGeom;
                winGeom.BY = (mH + 3 + 3) & ~3; winGeom.BX = mW; // rows: the copy starts at the 16-byte aligned row at or above the window's first row
                winGeom.boxBytes = winGeom.BY * winGeom.BX * tileGeom.nChns * 4;
                std::vector<uint32_t> soa((size_t)nT * 10, 0u);
                for (int i = 0; i < nT; i++)
                {
                    const uint32_t* rec = &tt[(size_t)i * rw];
                    uint32_t w[12];
                    memcpy(w, rec, sizeof(w));
                    for (int k = 0; k < 3; k++)
                    {
                        const uint32_t fid = fids[(size_t)i * nN + k];
                        const uint32_t z = fid / (mH * mW), c = (fid / mH) % mW, r = fid % mH;
                        w[k] = 4u * ((z * (uint32_t)winGeom.BX + c) * (uint32_t)winGeom.BY + r);
                    }
                    memcpy(&soa[(size_t)i * 4], w, 16);
                    memcpy(&soa[(size_t)nT * 4 + (size_t)i * 4], w + 4, 16);
                    memcpy(&soa[(size_t)nT * 8 + (size_t)i * 2], w + 8, 8);
                }
                cascTabSoA.ensure(soa.size());
                CUDA_OK(cudaMemcpy(cascTabSoA.p, soa.data(), soa.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
                tailOnWindows = cascTailOnWin && cascTailWinSmem(winGeom.boxBytes) <= 220 * 1024;
            }
            tileHead.clear();
            if (nT >= kCascHeadTrees)
            {
                tileHead.resize(kCascHeadTrees);
                for (int i = 0; i < kCascHeadTrees; i++) memcpy(&tileHead[i], &tt[(size_t)i * rw], sizeof(CascHeadRec));
            }
            tileCascade = true;
        }
    }

    SizeState& sizeState(int rows, int cols)
    {
        auto key = std::make_pair(rows, cols);
        auto it = sizes.find(key);
        if (it != sizes.end()) return *it->second;
        if (rows > maxRows || cols > maxCols) throw std::runtime_error("engine: frame larger than the size the engine was created for");
        std::unique_ptr<SizeState> st(new SizeState());
        st->plan = makePlan(opt, rows, cols);
        Plan& P = st->plan;
        // resample tables
        for (auto& r : P.reals)
        {
            st->realAx.emplace_back(new AxisUpload());
            st->realAx.emplace_back(new AxisUpload());
            if (r.mode != RealScale::ALIAS)
            {
                st->realAx[st->realAx.size() - 2]->upload(r.cx, stream);
                st->realAx[st->realAx.size() - 1]->upload(r.cy, stream);
            }
        }
        std::vector<AxisDev> axes(2 * P.geom.size());
        for (size_t i = 0; i < P.geom.size(); i++)
        {
            st->scaleAx.emplace_back(new AxisUpload());
            st->scaleAx.emplace_back(new AxisUpload());
            const ScaleGeom& g = P.geom[i];
            if (!g.isReal)
            {
                for (int v : g.cx.cnt) if (v > 3) throw std::runtime_error("engine: approximated scale needs more than 3 x taps");
                for (int v : g.cy.cnt) if (v > 3) throw std::runtime_error("engine: approximated scale needs more than 3 y taps");
                st->scaleAx[2 * i]->upload(g.cx, stream);
                st->scaleAx[2 * i + 1]->upload(g.cy, stream);
            }
            axes[2 * i] = st->scaleAx[2 * i]->dev;
            axes[2 * i + 1] = st->scaleAx[2 * i + 1]->dev;
        }
        st->axes.ensure(axes.size());
        CUDA_OK(cudaMemcpy(st->axes.p, axes.data(), axes.size() * sizeof(AxisDev), cudaMemcpyHostToDevice));
        // real-scale channel block layout
        int64_t off = 0;
        for (auto& r : P.reals)
        {
            st->realOff.push_back(off);
            off += (int64_t)P.nChns * r.cw * r.cP;
        }
        st->rFloatsPerFrame = (off + 31) / 32 * 32;
        int64_t mo = 0;
        for (auto& r : P.reals) { st->moOff.push_back(mo); mo += (int64_t)r.h * r.w; }
        st->moFloatsPerFrame = mo;
        if (P.reals.size() > 14) throw std::runtime_error("engine: more than 14 octaves are not supported");
        buildJobs(*st);
        // cascade geometry, acfDetect1.cpp:252-259
        const int modelHt = opt.modelDsPad_w, modelWd = opt.modelDsPad_h;
        int blk = 0, blkAll = 0;
        for (size_t si = 0; si < P.geom.size(); si++)
        {
            const ScaleGeom& g = P.geom[si];
            if (si > 0 && g.realK != P.geom[si - 1].realK) { st->groups[P.geom[si - 1].realK].cascTasks = blk; blk = 0; }
            CascScale c{};
            c.scaleIdx = (int)si;
            c.off = g.offset; c.P = g.P; c.planeStride = g.W * g.P;
            c.height1 = (int)ceil(float(g.H * opt.shrink - modelHt + 1) / opt.stride);
            c.width1 = (int)ceil(float(g.W * opt.shrink - modelWd + 1) / opt.stride);
            if (c.height1 < 0) c.height1 = 0;
            if (c.width1 < 0) c.width1 = 0;
            c.blk0 = blk;
            const int64_t nwin = (int64_t)c.height1 * c.width1;
            blk += (int)((nwin + kCascTask - 1) / kCascTask);
            blkAll += (int)((nwin + kCascTask - 1) / kCascTask);
            st->windowsPerFrame += nwin;
            st->cascHost.push_back(c);
        }
        st->groups[P.geom.back().realK].cascTasks = blk;
        st->cascBlocksPerFrame = blkAll;
        st->casc.ensure(st->cascHost.size());
        CUDA_OK(cudaMemcpy(st->casc.p, st->cascHost.data(), st->cascHost.size() * sizeof(CascScale), cudaMemcpyHostToDevice));
        if (tileCascade)
        {   // tile grids of k_cascade_tile, numbered per octave group like the tasks above
            int tile = 0;
            for (size_t si = 0; si < P.geom.size(); si++)
            {
                if (si > 0 && P.geom[si].realK != P.geom[si - 1].realK) { st->groups[P.geom[si - 1].realK].cascTiles = tile; tile = 0; }
                const CascScale& c = st->cascHost[si];
                CascTileScale t{};
                t.tile0 = tile; t.width1 = c.width1; t.height1 = c.height1; t.scaleIdx = c.scaleIdx;
                t.nTx = (c.width1 + tileGeom.Wc - 1) / tileGeom.Wc; t.nTy = (c.height1 + tileGeom.Wr - 1) / tileGeom.Wr;
                if (c.width1 <= 0 || c.height1 <= 0) { t.nTx = 0; t.nTy = 1; }
                tile += t.nTx * t.nTy;
                st->ctHost.push_back(t);
            }
            st->groups[P.geom.back().realK].cascTiles = tile;
            st->ct.ensure(st->ctHost.size());
            CUDA_OK(cudaMemcpy(st->ct.p, st->ctHost.data(), st->ctHost.size() * sizeof(CascTileScale), cudaMemcpyHostToDevice));
        }
        {
            std::vector<PostScale> ps(P.geom.size());
            for (size_t i = 0; i < ps.size(); i++) ps[i] = PostScale{ P.scales[i], P.scaleshw[i].first, P.scaleshw[i].second };
            st->postScales.ensure(ps.size());
            CUDA_OK(cudaMemcpy(st->postScales.p, ps.data(), ps.size() * sizeof(PostScale), cudaMemcpyHostToDevice));
        }
        SizeState& ref = *st;
        sizes[key] = std::move(st);
        return ref;
    }

    void buildJobs(SizeState& st)
    {
        const Plan& P = st.plan;
        st.chanJobsHost.clear();
        st.padJobsHost.clear();
        {
            std::vector<SizeState::Group> old = st.groups; // keep the cascade task counts across a rebuild (image-derived lambdas)
            st.groups.assign(P.reals.size(), SizeState::Group());
            for (size_t k = 0; k < old.size() && k < st.groups.size(); k++) st.groups[k].cascTasks = old[k].cascTasks;
        }
        for (size_t k = 0; k < st.groups.size(); k++) st.groups[k].sBeg = st.groups[k].sEnd = -1;
        int64_t cum = 0;
        std::vector<std::vector<ChanJob>> single(st.groups.size());
        std::vector<std::vector<std::vector<ChanJob>>> multi(st.groups.size()); // [group][plane][strip]
        for (size_t i = 0; i < P.geom.size(); i++)
        {
            const ScaleGeom& g = P.geom[i];
            const RealScale& r = P.reals[g.realK];
            SizeState::Group& G = st.groups[g.realK];
            if (G.sBeg < 0) { G.sBeg = (int)i; G.padBeg = (int)st.padJobsHost.size(); cum = 0; }
            G.sEnd = (int)i + 1;
            const int nStrips = (g.h + kStripRows - 1) / kStripRows;
            if (nStrips > 8) throw std::runtime_error("engine: channel planes taller than 1024 rows are not supported");
            // one kind for all strips of a plane (they share the block's barriers): bilinear only if every strip qualifies
            int kind = 1;
            if (!g.isReal)
            {
                int xTaps = 0;
                for (int v : g.cx.cnt) xTaps = std::max(xTaps, v);
                kind = (g.cy.mode == 2 && xTaps <= 2) ? 2 : 0;
                for (int s = 0; s < nStrips; s++)
                {   // k_chan stages the x pass of all source rows a strip touches in a 192-entry buffer
                    const int ya = std::min(s * kStripRows, g.h - 1), yb = std::min(s * kStripRows + kStripRows - 1, g.h - 1);
                    const int span = g.cy.start[yb] + g.cy.cnt[yb] - g.cy.start[ya];
                    if (span > 192) throw std::runtime_error("engine: approximated scale spans too many source rows per strip");
                    if (span > 128) kind = 0;
                }
            }
            for (int z = 0; z < P.nChns; z++)
            {
                const int type = z < P.typeFirst[1] ? 0 : (z < P.typeFirst[2] ? 1 : 2);
                std::vector<ChanJob> strips;
                for (int s = 0; s < nStrips; s++)
                {
                    ChanJob j{};
                    j.srcOff = st.realOff[g.realK] + (int64_t)z * r.cw * r.cP;
                    j.dstOff = g.offset + (int64_t)z * g.W * g.P;
                    j.srcH = r.ch; j.srcW = r.cw; j.srcP = r.cP;
                    j.h = g.h; j.w = g.w; j.P = g.P; j.padX = P.padX; j.padY = P.padY;
                    j.strip = s; j.kind = kind; j.axis = (int)i;
                    j.r = g.isReal ? 1.0f : g.ratio[type];
                    strips.push_back(j);
                }
                if (nStrips == 1) single[g.realK].push_back(strips[0]);
                else { multi[g.realK].push_back(strips); G.mWarps = std::max(G.mWarps, nStrips); }
            }
            if (P.padX || P.padY)
                for (int type = 0; type < 3; type++)
                {
                    if (!P.typeCount[type]) continue;
                    PadJob pj{};
                    pj.off = g.offset + (int64_t)P.typeFirst[type] * g.W * g.P;
                    pj.h = g.h; pj.w = g.w; pj.P = g.P; pj.W = g.W; pj.H = g.H; pj.padX = P.padX; pj.padY = P.padY;
                    pj.d = P.typeCount[type];
                    pj.cum = cum;
                    cum += (int64_t)pj.d * ((int64_t)g.W * g.H - (int64_t)g.w * g.h); // border elements only
                    st.padJobsHost.push_back(pj);
                }
            G.padEnd = (int)st.padJobsHost.size(); G.padTotal = cum;
        }
        for (size_t k = 0; k < st.groups.size(); k++)
        {
            SizeState::Group& G = st.groups[k];
            G.jobBeg = (int)st.chanJobsHost.size();
            st.chanJobsHost.insert(st.chanJobsHost.end(), single[k].begin(), single[k].end());
            G.jobEnd = G.mJobBeg = (int)st.chanJobsHost.size();
            G.mRanges.clear();
            for (int w = 2; w <= G.mWarps; w++)
            {
                const int beg = (int)st.chanJobsHost.size();
                for (auto& strips : multi[k])
                    if ((int)strips.size() == w) st.chanJobsHost.insert(st.chanJobsHost.end(), strips.begin(), strips.end());
                const int end = (int)st.chanJobsHost.size();
                if (end > beg) G.mRanges.push_back({ beg, end, w });
            }
            G.mJobEnd = (int)st.chanJobsHost.size();
        }
        st.padTotal = cum;
        st.chanJobs.ensure(st.chanJobsHost.size());
        CUDA_OK(cudaMemcpy(st.chanJobs.p, st.chanJobsHost.data(), st.chanJobsHost.size() * sizeof(ChanJob), cudaMemcpyHostToDevice));
        if (!st.padJobsHost.empty())
        {
            st.padJobs.ensure(st.padJobsHost.size());
            CUDA_OK(cudaMemcpy(st.padJobs.p, st.padJobsHost.data(), st.padJobsHost.size() * sizeof(PadJob), cudaMemcpyHostToDevice));
        }
    }

    // frame-0 pointers of real scale k's input image X_k (the frame itself, a smoothed earlier scale, or a resampled copy)
    // and of its smoothed image C_k = what chnsCompute works on after its in-place convTri (chnsCompute.cpp:239)
    const float* imgIn(const SizeState& st, int k) const
    {
        const RealScale& r = st.plan.reals[k];
        if (r.mode != RealScale::ALIAS) return st.In[k]->p;
        return r.srcKind == RealScale::FROM_I0 ? st.I0.p : imgSmooth(st, r.srcReal);
    }
    const float* imgSmooth(const SizeState& st, int k) const { return opt.color_smooth > 0 ? st.C[k]->p : imgIn(st, k); }

    void ensureBatch(SizeState& st, int n, bool needFrames)
    {
        const Plan& P = st.plan;
        const size_t img = (size_t)P.rows * P.cols;
        (void)needFrames; (void)img;
        if (n <= st.batchCap) return;
        st.I0.ensure((size_t)n * P.nImgPlanes * img);
        st.In.resize(P.reals.size());
        st.C.resize(P.reals.size());
        for (size_t k = 0; k < P.reals.size(); k++)
        {
            const RealScale& r = P.reals[k];
            if (!st.In[k]) st.In[k].reset(new DevBuf<float>());
            if (!st.C[k]) st.C[k].reset(new DevBuf<float>());
            if (r.mode != RealScale::ALIAS) st.In[k]->ensure((size_t)n * P.nImgPlanes * r.h * r.w);
            if (opt.color_smooth > 0) st.C[k]->ensure((size_t)n * P.nImgPlanes * r.h * r.w);
        }
        st.resTFloatsPerFrame = 0;
        for (const RealScale& r : P.reals)
            if (r.mode != RealScale::ALIAS) st.resTFloatsPerFrame = std::max<int64_t>(st.resTFloatsPerFrame, ((int64_t)P.nImgPlanes * r.w * r.srcH + 3) / 4 * 4);
        if (st.resTFloatsPerFrame) st.resT.ensure((size_t)n * st.resTFloatsPerFrame);
        st.gM.ensure((size_t)n * st.moFloatsPerFrame);
        st.gO.ensure((size_t)n * st.moFloatsPerFrame);
        if (opt.gm_normRad) st.gU.ensure((size_t)n * st.moFloatsPerFrame);
        st.R.ensure((size_t)n * st.rFloatsPerFrame + 1024); // slack: k_chan loads (never uses) up to 191 rows past a strip's last source row
        st.pyr.ensure((size_t)n * P.floatsPerFrame + 64); // slack: the cascade prefetches 32 elements past the lines it gathers
        // the pitch / alignment padding of the pyramid is never written by the kernels: clear it once
        CUDA_OK(cudaMemsetAsync(st.pyr.p, 0, (size_t)n * P.floatsPerFrame * sizeof(float), stream));
        CUDA_OK(cudaMemsetAsync(st.R.p, 0, (size_t)n * st.rFloatsPerFrame * sizeof(float), stream));
        if (tileCascade)
        {   // the pyramid moved: re-encode the per-scale tensor maps (frame = 4th dimension, so lanes only differ in a coordinate)
            std::vector<CUtensorMap> maps(2 * P.geom.size()); // box = a cascade tile, then box = one window's footprint
            for (size_t i = 0; i < P.geom.size(); i++)
            {
                const ScaleGeom& g = P.geom[i];
                maps[i] = channelTileMap(st.pyr.p + g.offset, g.H, g.W, P.nChns, n, g.P, (int64_t)g.W * g.P, P.floatsPerFrame, tileGeom);
                maps[P.geom.size() + i] = channelTileMap(st.pyr.p + g.offset, g.H, g.W, P.nChns, n, g.P, (int64_t)g.W * g.P, P.floatsPerFrame, winGeom);
            }
            st.tmaps.ensure(maps.size());
            CUDA_OK(cudaMemcpyAsync(st.tmaps.p, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, stream));
            CUDA_OK(cudaStreamSynchronize(stream)); // `maps` is a local
        }
        st.batchCap = n;
    }

    // ------------------------------------------------------------------------------------------
    // validate, select the per-size state, size the buffers, reset per-call instrumentation
    SizeState& beginBatch(const uint8_t* frames, int n, int rows, int cols, bool onDevice)
    {
        if (n < 1 || n > maxBatch) throw std::runtime_error("engine: batch size outside [1, max_batch]");
        if (!frames) throw std::runtime_error("engine: null frame pointer");
        CUDA_OK(cudaSetDevice(device));
        SizeState& st = sizeState(rows, cols);
        ensureBatch(st, n, !onDevice);
        cur = &st; curN = n;
        for (auto e : evs) cudaEventDestroy(e);
        evs.clear(); evNames.clear();
        mark("begin");
        return st;
    }

    void runPyramid(const uint8_t* frames, int n, int rows, int cols, bool onDevice)
    {
        SizeState& st = beginBatch(frames, n, rows, cols, onDevice);
        const uint8_t* dFrames = frames;
        if (!onDevice)
        {
            Slot& S = slots[0];
            if (anyPending()) throw std::runtime_error("engine: collect the submitted batches first");
            S.frames.ensure((size_t)n * frameBytes(rows, cols));
            CUDA_OK(cudaMemcpyAsync(S.frames.p, frames, (size_t)n * frameBytes(rows, cols), cudaMemcpyHostToDevice, stream));
            dFrames = S.frames.p;
            mark("h2d");
        }
        pyramidRange(st, dFrames, 0, n);
    }

    // pyramid + cascade for a batch, asynchronously.  Host frames go through the slot's staging buffer on the copy
    // stream, so the H2D copy of batch k+1 overlaps the kernels of batch k when the caller keeps more than one batch in flight.
    void submitAll(const uint8_t* frames, int n, int rows, int cols, bool onDevice)
    {
        Slot& S = slots[subSlot];
        if (S.pending) throw std::runtime_error("engine: three batches already in flight; call acfb_collect first");
        SizeState& st = beginBatch(frames, n, rows, cols, onDevice);
        const uint8_t* dFrames = frames;
        if (!onDevice)
        {
            const size_t bytes = (size_t)n * frameBytes(rows, cols);
            S.frames.ensure(bytes);
            CUDA_OK(cudaMemcpyAsync(S.frames.p, frames, bytes, cudaMemcpyHostToDevice, copyStream));
            CUDA_OK(cudaEventRecord(S.copied, copyStream));
            CUDA_OK(cudaStreamWaitEvent(stream, S.copied, 0));
            dFrames = S.frames.p;
        }
        resetHits(S, n, st.groups.size() * kMaxLanes);
        S.n = n; S.st = &st;
        S.posted = postOnDevice(st);
        if (S.posted)
        {
            postOut = std::max(1, std::min(maxDet, 64));
            S.postBytes = postHeaderBytes(n) + (size_t)n * postOut * sizeof(PostDet);
            S.post.ensure(S.postBytes);
            if (S.hPostCap < S.postBytes)
            {
                if (S.hPost) cudaFreeHost(S.hPost);
                CUDA_OK(cudaMallocHost(&S.hPost, S.postBytes));
                S.hPostCap = S.postBytes;
            }
            CUDA_OK(cudaMemsetAsync(S.post.p, 0, postHeaderBytes(n), stream));
        }
        const int useLanes = (overlap && !timing && !st.plan.lambdasFromImage && n >= 2 * nLanes) ? nLanes : 1;
        // Batches in flight are only ordered lane by lane (a lane's streams, plus the per-lane R_k event).  That is enough
        // while every frame index stays in its lane; when the batch size or the lane count changes, frame ranges move
        // between lanes, so this batch must first wait for everything the previous batches still run on ANY lane.
        if (st.lastN != 0 && (st.lastN != n || st.lastLanes != useLanes))
            for (int l = 0; l < kMaxLanes; l++) CUDA_OK(cudaStreamWaitEvent(stream, lanes[l].evB, 0));
        st.lastN = n; st.lastLanes = useLanes;
        if (useLanes == 1)
        {
            pyramidRange(st, dFrames, 0, n, &S, 0);
            fetchCounters(S, n, stream);
            S.st = &st; S.n = n; S.pending = true;
            CUDA_OK(cudaEventRecord(S.done, stream));
        }
        else
        {
            // Lanes are not joined back into `stream`: the colour / gradient chain of the NEXT batch starts while the final
            // channels and the cascade of this one are still running on the b streams.  finStream alone waits for them.
            const size_t img = frameBytes(rows, cols);
            const int per = (n + useLanes - 1) / useLanes;
            CUDA_OK(cudaEventRecord(lanes[0].evStart, stream)); // everything queued so far (H2D wait, counter reset)
            for (int l = 0; l < useLanes; l++)
            {
                const int f0 = l * per, nc = std::min(per, n - f0);
                if (nc <= 0) break;
                if (l > 0) CUDA_OK(cudaStreamWaitEvent(lanes[l].a, lanes[0].evStart, 0));
                pyramidRange(st, dFrames + (size_t)f0 * img, f0, nc, &S, l, /*joinB=*/false);
                CUDA_OK(cudaStreamWaitEvent(finStream, lanes[l].evB, 0));
            }
            fetchCounters(S, n, finStream);
            S.st = &st; S.n = n; S.pending = true;
            CUDA_OK(cudaEventRecord(S.done, finStream));
        }
        subSlot = (subSlot + 1) % kSlots;
    }

    // launches every pyramid kernel for frames [f0, f0 + n); dFrames points at frame f0 (device memory)
    // S != nullptr: also run the cascade into slot S.  With `overlap`, the final-channel kernel, border fill and cascade of
    // octave group k run on streamB as soon as real scale k is done, concurrently with the real-scale kernels of group k+1.
    void pyramidRange(SizeState& st, const uint8_t* dFrames, int f0, int n, Slot* S = nullptr, int lane = 0, bool joinB = true)
    {
        Lane& L = lanes[lane];
        const Plan& P = st.plan;
        const int rows = P.rows, cols = P.cols;
        const size_t img = (size_t)rows * cols;
        static const int kOff[8][3] = { { 0, 1, 2 }, { 2, 1, 0 }, { 0, 1, 2 }, { 2, 1, 0 }, { 0, 0, 0 }, { 0, 1, 2 }, { 0, 1, 2 }, { 0, 1, 2 } };
        ColorArgs ca{ dFrames, st.I0.p + (size_t)f0 * P.nImgPlanes * img, lut.p, rows, cols, n, colorMode(),
                      bpp(), kOff[pixfmt][0], kOff[pixfmt][1], kOff[pixfmt][2],
                      pixfmt == 5 ? 1 : pixfmt == 6 ? 2 : pixfmt == 7 ? 3 : 0, isTranspose ? 1 : 0 };
        if (pixfmt == 7 && ((rows | cols) & 1)) throw std::runtime_error("engine: NV12 frames need even rows and cols");
        if (pixfmt == 7 && isTranspose) throw std::runtime_error("engine: NV12 frames cannot be handed over transposed");
        launchColor(ca, L.a); launches++;
        mark("color");
        const double rs = opt.color_smooth;
        const bool ovl = overlap && !P.lambdasFromImage && !timing;
        while (L.evReal.size() < P.reals.size()) { cudaEvent_t ev; CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); L.evReal.push_back(ev); }
        while (L.evChan.size() < P.reals.size()) { cudaEvent_t ev; CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); L.evChan.push_back(ev); }
        while (L.evSmooth.size() < P.reals.size()) { cudaEvent_t ev; CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); L.evSmooth.push_back(ev); }
        // The images of the real scales form their own dependency chain (I0 -> C0 -> resample -> C1 -> ...): with overlap
        // it runs ahead on stream c, and the gradient kernels of scale k (stream a) start when C_k is ready.
        cudaStream_t sImg = ovl ? L.c : L.a;
        if (ovl)
        {
            CUDA_OK(cudaEventRecord(L.evColor, L.a));
            CUDA_OK(cudaStreamWaitEvent(L.c, L.evColor, 0));
        }
        // A real scale that is exactly half of an earlier one's smoothed image (the reference's /2 fast path) is written by
        // that scale's k_smooth while the smoothed columns are still in registers: no separate k_down2 pass over the plane.
        std::vector<int> halfOf(P.reals.size(), -1);
        std::vector<char> fused(P.reals.size(), 0);
        if (fuseDown2 && rs > 0)
            for (size_t k = 0; k < P.reals.size(); k++)
            {
                const RealScale& r = P.reals[k];
                if (r.mode == RealScale::DOWN2 && r.h % 4 == 0 && r.srcKind == RealScale::FROM_C && r.srcReal >= 0 && r.srcReal < (int)k &&
                    halfOf[r.srcReal] < 0 && r.srcH == 2 * r.h && r.srcW == 2 * r.w)
                { halfOf[r.srcReal] = (int)k; fused[k] = 1; }
            }
        for (size_t k = 0; k < P.reals.size(); k++)
        {
            const RealScale& r = P.reals[k];
            const int64_t srcStride = (int64_t)P.nImgPlanes * r.srcH * r.srcW;
            const int64_t ownStride = (int64_t)P.nImgPlanes * r.h * r.w;
            if (r.mode != RealScale::ALIAS && !fused[k])
            {   // I1 = imResample(I, sz1) (chnsPyramid.cpp:303-312), incl. the exact /2 fast path of imResampleMex.cpp:198-203,284-301
                const float* src = ((r.srcKind == RealScale::FROM_I0) ? st.I0.p : imgSmooth(st, r.srcReal)) + (size_t)f0 * srcStride;
                ResampleArgs ra{};
                ra.src = src; ra.dst = st.In[k]->p + (size_t)f0 * ownStride; ra.srcFrameStride = srcStride;
                ra.dstFrameStride = ownStride;
                ra.ha = r.srcH; ra.wa = r.srcW; ra.hb = r.h; ra.wb = r.w; ra.d = P.nImgPlanes; ra.n = n;
                ra.cx = st.realAx[2 * k]->dev; ra.cy = st.realAx[2 * k + 1]->dev; ra.r = r.r;
                if (twoPassResample && st.resT.p) { ra.tmp = st.resT.p + (size_t)f0 * st.resTFloatsPerFrame; ra.tmpFrameStride = st.resTFloatsPerFrame; }
                if (r.mode == RealScale::DOWN2 && r.h % 4 == 0) launchDown2(ra, sImg); else { launchResample(ra, sImg); if (ra.tmp) launches++; }
                launches++;
            }
            float* Mk = st.gM.p + (size_t)f0 * st.moFloatsPerFrame + st.moOff[k];
            uint16_t* Ok = st.gO.p + (size_t)f0 * st.moFloatsPerFrame + st.moOff[k];
            float* Rk = st.R.p + (size_t)f0 * st.rFloatsPerFrame + st.realOff[k];
            float* Uk = opt.gm_normRad ? st.gU.p + (size_t)f0 * st.moFloatsPerFrame + st.moOff[k] : nullptr;
            // k_front: smoothing + gradMag + x pass of the normalisation triangle in one march (needs the smoothing, planes of at
            // most 2304 rows); the smoothed image itself is only written when something still reads it -- a later real scale that
            // resamples from it with a ratio other than the fused 1/2, the colour channels' shrink, or the debug taps
            const bool front = useFront && rs > 0 && r.h % 4 == 0 && r.h >= 16 && r.w >= 16 && r.h <= 2304;
            bool needC = keepC || opt.color_enabled;
            for (size_t j = k + 1; j < P.reals.size(); j++)
                if (P.reals[j].mode != RealScale::ALIAS && !fused[j] && P.reals[j].srcKind == RealScale::FROM_C && P.reals[j].srcReal == (int)k) needC = true;
            const float sp = rs > 0 ? (float)(12.0 / rs / (rs + 2.0) - 2.0) : 0.f; // convTri.cpp:215-218, convConst.cpp:496
            if (front)
            {
                FrontArgs fa{};
                fa.src = imgIn(st, (int)k) + (size_t)f0 * ownStride; fa.dstC = needC ? st.C[k]->p + (size_t)f0 * ownStride : nullptr;
                fa.outM = Mk; fa.outO = Ok; fa.outU = Uk; fa.moFrameStride = st.moFloatsPerFrame;
                fa.H = r.h; fa.W = r.w; fa.nc = P.nImgPlanes; fa.nPlanes = n * P.nImgPlanes; fa.gradPlane = opt.gm_colorChn; fa.full = opt.gm_full;
                fa.p = sp; fa.nrm = 1.0f / ((sp + 2) * (sp + 2));
                if (halfOf[k] >= 0)
                {
                    const RealScale& h = P.reals[halfOf[k]];
                    fa.dst2 = st.In[halfOf[k]]->p + (size_t)f0 * P.nImgPlanes * h.h * h.w; fa.r2 = h.r / 2;
                }
                launchFront(fa, sImg); launches++;
            }
            else if (rs > 0)
            {   // the in-place smoothing of the image planes, bit exact (see k_smooth)
                SmoothArgs sa{};
                sa.src = imgIn(st, (int)k) + (size_t)f0 * ownStride; sa.dst = st.C[k]->p + (size_t)f0 * ownStride;
                sa.H = r.h; sa.W = r.w; sa.nPlanes = n * P.nImgPlanes;
                sa.p = sp; sa.nrm = 1.0f / ((sa.p + 2) * (sa.p + 2));
                sa.pfAhead = marchPrefetch;
                if (halfOf[k] >= 0)
                {
                    const RealScale& h = P.reals[halfOf[k]];
                    sa.dst2 = st.In[halfOf[k]]->p + (size_t)f0 * P.nImgPlanes * h.h * h.w; sa.r2 = h.r / 2;
                }
                launchSmooth(sa, sImg); launches++;
            }
            if (ovl)
            {
                CUDA_OK(cudaEventRecord(L.evSmooth[k], L.c));
                CUDA_OK(cudaStreamWaitEvent(L.a, L.evSmooth[k], 0));
            }
            const float* Ck = imgSmooth(st, (int)k) + (size_t)f0 * ownStride;
            if (!front)
            {   // gradientMag of plane pGradMag.colorChn (chnsCompute.cpp:262-282)
                GradArgs ga{};
                ga.src = Ck + (size_t)opt.gm_colorChn * r.h * r.w; ga.outM = Mk; ga.outO = Ok; ga.acosTab = acosTab.p;
                ga.srcFrameStride = ownStride; ga.moFrameStride = st.moFloatsPerFrame;
                ga.H = r.h; ga.W = r.w; ga.n = n; ga.full = opt.gm_full; ga.colsPerThread = gradCols;
                launchGradMag(ga, L.a); launches++;
            }
            // gradientHist + the shrunk magnitude and colour channels (chnsCompute.cpp:241-258,283-338)
            HistArgs ha{};
            ha.M = Mk; ha.O = Ok; ha.acosTab = acosTab.p; ha.C = Ck; ha.cFrameStride = ownStride; ha.outR = Rk;
            ha.moFrameStride = st.moFloatsPerFrame; ha.rFrameStride = st.rFloatsPerFrame;
            ha.H = r.h; ha.W = r.w; ha.n = n; ha.cP = r.cP; ha.firstPlane = opt.color_enabled ? P.nImgPlanes : 0; ha.nOrients = opt.gh_nOrients;
            {
                const float PI = 3.14159265f;
                ha.oMult = (float)opt.gh_nOrients / (opt.gm_full ? 2 * PI : PI);
                const float sh = (float)opt.shrink; ha.sInv2 = 1 / sh / sh;
                float q = 1.0f; q /= 4; q /= float(1 + 1e-6); ha.shrinkMul = q / 4; // imResampleMex.cpp:153-157, 314
            }
            if (opt.gm_normRad)
            {   // convTri(M, S, normRad) as the reference's two running-sum passes (gradientMag.cpp:125-131); the y pass
                // normalises and bins in the same kernel, so S and the normalised magnitude never reach HBM
                if (!front)
                {
                    TrixArgs xa{ Mk, Uk, st.moFloatsPerFrame, r.h, r.w, n, trixPrefetch };
                    launchTrix(xa, L.a); launches++;
                }
                if (ovl) CUDA_OK(cudaStreamWaitEvent(L.a, L.evChan[k], 0)); // R_k is still read by the previous batch's k_chan (no-op the first time)
                TriyArgs ta{};
                ta.U = Uk; ta.h = ha; ta.frameStride = st.moFloatsPerFrame; ta.H = r.h; ta.W = r.w; ta.n = n;
                ta.normConst = (float)opt.gm_normConst; ta.blocksPerSm = triyBlocksPerSm; ta.fastScan = triyFastScan;
                CUtensorMap mapU, mapM;
                if (triyTma && planeChunkMap(Uk, r.h, r.w, n, st.moFloatsPerFrame, mapU) && planeChunkMap(Mk, r.h, r.w, n, st.moFloatsPerFrame, mapM))
                    launchTriyHistTma(ta, mapU, mapM, L.a);
                else launchTriyHist(ta, L.a);
                launches++;
                ha.doMag = 0;
            }
            else
            {
                if (ovl) CUDA_OK(cudaStreamWaitEvent(L.a, L.evChan[k], 0));
                ha.doMag = 1;
            }
            if (ha.doMag || ha.firstPlane > 0) { launchHist(ha, L.a); launches++; }
            if (ovl)
            {
                CUDA_OK(cudaEventRecord(L.evReal[k], L.a));
                CUDA_OK(cudaStreamWaitEvent(L.b, L.evReal[k], 0));
                groupTail(st, (int)k, f0, n, S, L.b, L.evChan[k]);
            }
        }
        mark("real");
        if (ovl)
        {
            if (S && S->posted) launchPostKernel(st, *S, f0, n, L.b);
            CUDA_OK(cudaEventRecord(L.evB, L.b));
            if (joinB) CUDA_OK(cudaStreamWaitEvent(L.a, L.evB, 0));
        }
        else
        {
            if (P.lambdasFromImage) deriveLambdas(st, n);
            for (size_t k = 0; k < P.reals.size(); k++) groupTail(st, (int)k, f0, n, nullptr, L.a);
            mark("chan");
            if (S)
            {
                for (size_t k = 0; k < P.reals.size(); k++) cascadeGroup(st, *S, (int)k, f0, n, L.a);
                mark("cascade");
                if (S->posted) { launchPostKernel(st, *S, f0, n, L.a); mark("post"); }
            }
        }
        CUDA_OK(cudaGetLastError());
    }

    // chnsCompute of the full-resolution image only (Detector::evaluate): colour conversion + real scale 0
    void colorAndReal0(SizeState& st, const uint8_t* dFrames)
    {
        const Plan& P = st.plan;
        const bool keepOverlap = overlap;
        overlap = false; // single stream, no group tails
        std::vector<RealScale> saved(P.reals.begin() + 1, P.reals.end());
        st.plan.reals.resize(1);
        std::vector<SizeState::Group> savedG = st.groups;
        for (auto& g : st.groups) { g.jobEnd = g.jobBeg; g.mJobEnd = g.mJobBeg; g.padEnd = g.padBeg; g.mRanges.clear(); }
        try { pyramidRange(st, dFrames, 0, 1, nullptr, 0); }
        catch (...) { st.plan.reals.insert(st.plan.reals.end(), saved.begin(), saved.end()); st.groups = savedG; overlap = keepOverlap; throw; }
        st.plan.reals.insert(st.plan.reals.end(), saved.begin(), saved.end());
        st.groups = savedG;
        overlap = keepOverlap;
    }

    // final channels (+ border fill, + cascade when S is given) of octave group k on stream s
    void groupTail(SizeState& st, int k, int f0, int n, Slot* S, cudaStream_t s, cudaEvent_t afterChan = nullptr)
    {
        const Plan& P = st.plan;
        const SizeState::Group& G = st.groups[k];
        const double sm = opt.smooth;
        ChanArgs c{};
        c.src = st.R.p + (size_t)f0 * st.rFloatsPerFrame; c.dst = st.pyr.p + (size_t)f0 * P.floatsPerFrame;
        c.srcFrameStride = st.rFloatsPerFrame; c.dstFrameStride = P.floatsPerFrame;
        c.axes = st.axes.p; c.n = n;
        if (sm > 0) { c.p = (float)(12.0 / sm / (sm + 2.0) - 2.0); c.nrm = 1.0f / ((c.p + 2) * (c.p + 2)); }
        else { c.p = 0; c.nrm = 0; }
        for (const SizeState::MRange& mr : G.mRanges)
        {   // planes taller than 128 rows: one block per plane, one launch per strip count
            c.jobs = st.chanJobs.p + mr.beg; c.nJobs = mr.end - mr.beg; c.blockWarps = mr.warps;
            launchChan(c, s); launches++;
        }
        c.jobs = st.chanJobs.p + G.jobBeg; c.nJobs = G.jobEnd - G.jobBeg; c.blockWarps = 0;
        if (c.nJobs > 0) { launchChan(c, s); launches++; }
        if (G.padEnd > G.padBeg)
        {
            PadArgs pa{ st.pyr.p + (size_t)f0 * P.floatsPerFrame, P.floatsPerFrame, st.padJobs.p + G.padBeg, G.padEnd - G.padBeg, n, G.padTotal };
            launchPad(pa, s); launches++;
        }
        if (afterChan) CUDA_OK(cudaEventRecord(afterChan, s)); // R_k may be overwritten from here on
        if (S) cascadeGroup(st, *S, k, f0, n, s);
    }

    // k_cascade_tail_win finishes what a k_cascade_tile launch handed over, with that launch's scales and buffers and the
    // window-footprint tensor maps of the same scales
    CascTailWinArgs tailWinArgs(const CascTileArgs& t, const CUtensorMap* winMaps) const
    {
        CascTailWinArgs q{};
        q.maps = winMaps; q.scales = t.scales; q.frame0 = t.frame0;
        const size_t nT = (size_t)t.nTrees;
        q.tabA = reinterpret_cast<const uint4*>(cascTabSoA.p); q.tabB = reinterpret_cast<const uint4*>(cascTabSoA.p + nT * 4);
        q.tabC = reinterpret_cast<const uint2*>(cascTabSoA.p + nT * 8);
        q.nTrees = t.nTrees; q.step = t.step; q.footBytes = winGeom.boxBytes; q.cascThr = t.cascThr;
        q.tail = t.tail; q.tailCount = t.tailCount; q.tailCap = t.tailCap;
        q.hitCount = t.hitCount; q.hits = t.hits; q.cap = t.cap; q.stats = t.stats;
        return q;
    }

    void cascadeGroup(SizeState& st, Slot& S, int k, int f0, int n, cudaStream_t s)
    {
        const SizeState::Group& G = st.groups[k];
        if (G.cascTasks <= 0) return;
        if (tileCascade)
        {
            CascTileArgs t{};
            t.maps = st.tmaps.p + G.sBeg; t.scales = st.ct.p + G.sBeg; t.nScales = G.sEnd - G.sBeg; t.tilesPerFrame = G.cascTiles; t.n = n; t.frame0 = f0;
            t.tab = cascTabTile.p; t.nTrees = model.nTrees();
            t.Wc = tileGeom.Wc; t.Wr = tileGeom.Wr; t.BY = tileGeom.BY; t.step = tileGeom.step; t.tileBytes = tileGeom.tileBytes; t.boxBytes = tileGeom.boxBytes;
            t.listCap = tileGeom.listCap; t.smemBytes = tileGeom.smemBytes; t.sparseMax = cascSparseMax; t.blocksPerSm = cascBlocksPerSm; t.cascThr = (float)opt.cascThr;
            t.headLevels = cascHeadLevels; t.headTrees = (int)tileHead.size(); if (!tileHead.empty()) memcpy(t.head, tileHead.data(), sizeof(t.head));
            t.hitCount = S.hitCount.p + f0; t.hits = S.hits.p + (size_t)f0 * hitCap; t.cap = hitCap; t.stats = S.stats.p;
            const int kLaunch = S.nextCounter++;
            t.taskCounter = S.stats.p + 2 + kLaunch;
            if ((size_t)S.nextCounter + 2 > S.statsWords) throw std::runtime_error("engine: cascade task counters exhausted");
            // hand-over entries pack (frame | scale << 24) and (c | r << 16)
            bool packs = t.nScales <= 256 && n < (1 << 24);
            for (int i = G.sBeg; i < G.sEnd; i++) packs = packs && st.cascHost[i].width1 < 65536 && st.cascHost[i].height1 < 65536;
            t.exportMax = (packs && S.tail.p) ? cascExportMax : 0;
            t.tail = S.tail.p ? S.tail.p + (size_t)kLaunch * cascTailCap : nullptr; t.tailCount = S.tailCount.p ? S.tailCount.p + kLaunch : nullptr; t.tailCap = cascTailCap;
            launchCascadeTile(t, s); launches++;
            if (t.exportMax > 0 && tailOnWindows)
            {
                launchCascadeTailWin(tailWinArgs(t, st.tmaps.p + st.plan.geom.size() + G.sBeg), s); launches++;
            }
            else if (t.exportMax > 0)
            {
                CascTailArgs q{};
                q.pyr = st.pyr.p + (size_t)f0 * st.plan.floatsPerFrame; q.frameStride = st.plan.floatsPerFrame; q.scales = st.casc.p + G.sBeg;
                q.tab = cascTab.p; q.nTrees = model.nTrees(); q.stride = opt.stride; q.shrink = opt.shrink; q.cascThr = (float)opt.cascThr;
                q.tail = t.tail; q.tailCount = t.tailCount; q.tailCap = cascTailCap;
                q.hitCount = t.hitCount; q.hits = t.hits; q.cap = hitCap; q.stats = S.stats.p;
                launchCascadeTail(q, s); launches++;
            }
            return;
        }
        CascArgs a{};
        a.pyr = st.pyr.p + (size_t)f0 * st.plan.floatsPerFrame; a.frameStride = st.plan.floatsPerFrame;
        a.scales = st.casc.p + G.sBeg; a.nScales = G.sEnd - G.sBeg;
        a.nBlocksPerFrame = G.cascTasks; a.n = n; a.tab = cascTab.p; a.nTrees = model.nTrees(); a.depth = model.clf.treeDepth;
        a.recWords = recWords; a.stride = opt.stride; a.shrink = opt.shrink; a.cascThr = (float)opt.cascThr;
        a.hitCount = S.hitCount.p + f0; a.hits = S.hits.p + (size_t)f0 * hitCap; a.cap = hitCap; a.stats = S.stats.p; a.prefetch = cascPrefetch; a.blocksPerSm = cascBlocksPerSm;
        a.taskCounter = S.stats.p + 2 + (S.nextCounter++);
        if ((size_t)S.nextCounter + 2 > S.statsWords) throw std::runtime_error("engine: cascade task counters exhausted");
        // k_cascade packs (frame << 8 | scale-in-group) and (c | r << 16) into its queue entries
        if (a.nScales > 256) throw std::runtime_error("engine: more than 256 scales in one octave group (nApprox too large for the global-gather cascade)");
        for (int i = G.sBeg; i < G.sEnd; i++)
            if (st.cascHost[i].width1 >= 65536 || st.cascHost[i].height1 >= 65536) throw std::runtime_error("engine: window grid of a scale exceeds 65535");
        if (n >= (1 << 24)) throw std::runtime_error("engine: batch too large for the cascade's queue entries");
        launchCascade(a, s); launches++;
    }

    // chnsPyramid.cpp:341-374; per-frame lambdas are only meaningful frame by frame, so the batch must be 1
    void deriveLambdas(SizeState& st, int n)
    {
        Plan& P = st.plan;
        if (n != 1) throw std::runtime_error("engine: a model without lambdas derives them per image; call with one frame at a time");
        std::vector<int> is;
        for (int i = 1 + opt.nOctUp * opt.nPerOct; i <= (int)P.scales.size(); i += opt.nApprox + 1) is.push_back(i - 1);
        if (is.size() < 2) throw std::runtime_error("engine: need at least two real scales to derive lambdas");
        if (is.size() > 2) is = { is[1], is[2] };
        DevBuf<double> sums;
        sums.ensure(6);
        CUDA_OK(cudaMemsetAsync(sums.p, 0, 6 * sizeof(double), stream));
        double numel[2][3] = {};
        for (int q = 0; q < 2; q++)
        {
            int k = -1;
            for (size_t kk = 0; kk < P.reals.size(); kk++) if (P.reals[kk].scaleIdx == is[q]) k = (int)kk;
            if (k < 0) throw std::runtime_error("engine: lambda scale is not a real scale");
            const RealScale& r = P.reals[k];
            for (int type = 0; type < 3; type++)
            {
                if (!P.typeCount[type]) continue;
                SumArgs sa{ st.R.p, st.rFloatsPerFrame, st.realOff[k] + (int64_t)P.typeFirst[type] * r.cw * r.cP, r.cP, r.ch, r.cw, P.typeCount[type], sums.p + q * 3 + type, 1 };
                launchPlaneSum(sa, stream); launches++;
                numel[q][type] = (double)P.typeCount[type] * r.cw * r.ch;
            }
        }
        double h[6];
        CUDA_OK(cudaMemcpyAsync(h, sums.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        lambdasFromImage.clear();
        for (int type = 0; type < 3; type++)
        {
            if (!P.typeCount[type]) continue;
            const double f0 = h[type] / numel[0][type], f1 = h[3 + type] / numel[1][type];
            lambdasFromImage.push_back(-(std::log(f0 / f1) / std::log(2.0)) / (std::log(P.scales[is[0]] / P.scales[is[1]]) / std::log(2.0)));
        }
        setRatios(P, opt, lambdasFromImage.data(), (int)lambdasFromImage.size());
        buildJobs(st);
    }

    void resetHits(Slot& S, int n, size_t cascLaunches = 0)
    {
        S.statsWords = std::max<size_t>(64, 2 + cascLaunches); // [0] trees, [1] windows, then one task counter per cascade launch
        S.stats.ensure(S.statsWords);
        if (tileCascade && cascExportMax > 0)
        {
            S.tail.ensure(S.statsWords * (size_t)cascTailCap);
            S.tailCount.ensure(S.statsWords);
            CUDA_OK(cudaMemsetAsync(S.tailCount.p, 0, S.statsWords * sizeof(int), stream));
        }
        S.hitCount.ensure(n);
        S.hits.ensure((size_t)n * hitCap);
        if (S.hCountCap < n)
        {
            if (S.hCount) cudaFreeHost(S.hCount);
            CUDA_OK(cudaMallocHost(&S.hCount, (size_t)n * sizeof(int)));
            S.hCountCap = n;
        }
        CUDA_OK(cudaMemsetAsync(S.hitCount.p, 0, n * sizeof(int), stream));
        CUDA_OK(cudaMemsetAsync(S.stats.p, 0, S.statsWords * sizeof(unsigned long long), stream));
        S.nextCounter = 0;
    }

    void cascadeRange(SizeState& st, Slot& S, int f0, int n)
    {
        for (size_t k = 0; k < st.groups.size(); k++) cascadeGroup(st, S, (int)k, f0, n, stream);
        mark("cascade");
        CUDA_OK(cudaGetLastError());
    }

    void fetchCounters(Slot& S, int n, cudaStream_t s)
    {
        CUDA_OK(cudaMemcpyAsync(S.hCount, S.hitCount.p, n * sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaMemcpyAsync(S.hStats, S.stats.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        if (S.posted) CUDA_OK(cudaMemcpyAsync(S.hPost, S.post.p, S.postBytes, cudaMemcpyDeviceToHost, s));
        if (comm && S.posted && distEarly)
        {   // every rank's boxes to every rank, straight from the device buffer k_post wrote, on the communication stream
            CUDA_OK(cudaEventRecord(S.evPost, s));
            CUDA_OK(cudaStreamWaitEvent(commStream, S.evPost, 0));
            allGatherPost(S);
        }
    }

    // ---- multi-GPU -------------------------------------------------------------------------------------------------
    void distCreateLocals()
    {
        CUDA_OK(cudaSetDevice(device));
        if (!commStream) CUDA_OK(cudaStreamCreateWithFlags(&commStream, cudaStreamNonBlocking));
        for (int i = 0; i < kSlots; i++)
        {
            if (!slots[i].evPost) CUDA_OK(cudaEventCreateWithFlags(&slots[i].evPost, cudaEventDisableTiming));
            if (!slots[i].gathDone) CUDA_OK(cudaEventCreateWithFlags(&slots[i].gathDone, cudaEventDisableTiming));
        }
    }
    // The gather moves tens of kilobytes: one CTA is plenty, and every further CTA of the collective's kernel would sit on an SM
    // spinning for the slowest rank while this rank's own kernels want that SM (ACFB_NCCL_MAX_CTAS=0: NCCL's default)
    // The exchange: ncclAllGather of the device buffers k_post wrote (default whenever libnccl.so.2 can be loaded; the form that also
    // crosses nodes), or a shared-memory ring between the ranks of one box (ACFB_DIST_EXCHANGE=shm, and the fallback without NCCL).
    // Measured equal: N = 8 17.10 vs 17.06 ms per step, N = 1 16.85 (profiles/r2_scaling.md).
    static bool distUseNccl()
    {
        const char* x = getenv("ACFB_DIST_EXCHANGE");
        if (x && std::string(x) == "shm") return false;
        if (x && std::string(x) == "nccl") return true;
        try { NcclApi::get(); return true; } catch (...) { return false; }
    }
    size_t distSlotBytes() const { return 16 + (size_t)maxBatch * sizeof(int) + (size_t)maxBatch * 64 * sizeof(acfb_det); }
    static NcclConfig distConfig()
    {
        NcclConfig cfg = ncclDefaultConfig();
        int maxCtas = 1;
        if (const char* mc = getenv("ACFB_NCCL_MAX_CTAS")) maxCtas = atoi(mc);
        if (maxCtas > 0) { cfg.minCTAs = 1; cfg.maxCTAs = maxCtas; }
        return cfg;
    }
    void distInitRank(const uint8_t* id, int rank, int world)
    {
        if (comm || exch) throw std::runtime_error("engine: the communicator exists already");
        if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("engine: bad rank / world size");
        if (anyPending()) throw std::runtime_error("engine: collect the submitted batches first");
        if (distUseNccl())
        {
            distCreateLocals();
            NcclUniqueId uid;
            memcpy(uid.internal, id, sizeof(uid.internal));
            NcclConfig cfg = distConfig();
            NCCL_OK(NcclApi::get().CommInitRankConfig(&comm, world, uid, rank, &cfg));
        }
        else exch = ShmExchange::open(id, rank, world, distSlotBytes());
        distRank = rank; distWorld = world;
    }
    // S.post of every rank -> S.gath (rank-major) -> S.hGath, on commStream
    void allGatherPost(Slot& S)
    {
        const size_t tot = (size_t)distWorld * S.postBytes;
        S.gath.ensure(tot);
        if (S.hGathCap < tot)
        {
            if (S.hGath) cudaFreeHost(S.hGath);
            CUDA_OK(cudaMallocHost(&S.hGath, tot));
            S.hGathCap = tot;
        }
        NCCL_OK(NcclApi::get().AllGather(S.post.p, S.gath.p, S.postBytes, kNcclUint8, comm, commStream));
        CUDA_OK(cudaMemcpyAsync(S.hGath, S.gath.p, tot, cudaMemcpyDeviceToHost, commStream));
        CUDA_OK(cudaEventRecord(S.gathDone, commStream));
        S.gathered = true;
    }
    // Shared-memory exchange: every rank finishes its own batch (device tail or host tail, it does not matter which), publishes
    // {n, counts[n], boxes} in its slot of batch `batchNo`; rank 0 takes every rank's slot in rank order.
    void distCollectShm(acfb_det* dets, int cap, int* counts, int* total, unsigned long long& batchNo)
    {
        Slot& S = slots[colSlot];
        if (!S.pending) throw std::runtime_error("engine: nothing submitted");
        if (!doNms || maxDet < 1 || maxDet > 64)
            throw std::runtime_error("engine: the gather carries the boxes bbNms + prune leave, at most 64 per frame: setDoNonMaximaSuppression(true), maxDetectionCount <= 64");
        const int n = S.n, po = std::max(1, std::min(maxDet, 64));
        distBuf.resize(16 + (size_t)n * sizeof(int) + (size_t)n * po * sizeof(acfb_det));
        int* hc = reinterpret_cast<int*>(distBuf.data() + 16);
        acfb_det* hd = reinterpret_cast<acfb_det*>(distBuf.data() + 16 + (size_t)n * sizeof(int));
        int tot = 0;
        collect(hd, n * po, hc, &tot); // this rank's own result; boxes arrive compacted in frame order
        reinterpret_cast<int*>(distBuf.data())[0] = n; reinterpret_cast<int*>(distBuf.data())[1] = tot;
        exch->publish(batchNo, distBuf.data(), 16 + (size_t)n * sizeof(int) + (size_t)tot * sizeof(acfb_det));
        int written = 0, all = 0;
        if (distRank == 0)
        {
            for (int r = 0; r < distWorld; r++)
            {
                size_t bytes = 0;
                const unsigned char* p = exch->wait(batchNo, r, &bytes);
                const int rn = reinterpret_cast<const int*>(p)[0], rt = reinterpret_cast<const int*>(p)[1];
                if (rn != n || bytes != 16 + (size_t)n * sizeof(int) + (size_t)rt * sizeof(acfb_det))
                    throw std::runtime_error("engine: the ranks submitted different numbers of frames");
                const int* rc = reinterpret_cast<const int*>(p + 16);
                const acfb_det* rd = reinterpret_cast<const acfb_det*>(p + 16 + (size_t)n * sizeof(int));
                int k = 0;
                for (int f = 0; f < n; f++)
                {
                    if (counts) counts[(size_t)r * n + f] = rc[f];
                    for (int j = 0; j < rc[f]; j++, k++, all++)
                        if (written < cap && dets) { dets[written] = rd[k]; dets[written].frame = r * n + f; written++; }
                }
            }
            exch->consumed(batchNo);
        }
        batchNo++;
        if (total) *total = all;
    }

    // acfb_collect on every rank + the gather: rank 0 receives the boxes of all ranks' batches in global frame order (frame =
    // rank * n + local frame), the other ranks receive nothing (*total = 0).  Every rank must have submitted the same number of
    // frames with the same options.  The fast path only waits for the all-gather enqueued at submit time; when any rank had to
    // leave a frame to its host tail (more raw hits than k_post sorts) -- every rank sees every rank's flag -- all ranks repeat
    // the gather with the host results.
    void distCollect(acfb_det* dets, int cap, int* counts, int* total, unsigned long long& batchNo)
    {
        if (exch) { distCollectShm(dets, cap, counts, total, batchNo); return; }
        if (!comm) throw std::runtime_error("engine: acfb_dist_init_rank / acfb_dist_init_all first");
        Slot& S = slots[colSlot];
        if (!S.pending) throw std::runtime_error("engine: nothing submitted");
        if (!doNms || maxDet < 1 || maxDet > 64)
            throw std::runtime_error("engine: the gather carries the boxes bbNms + prune leave, at most 64 per frame: setDoNonMaximaSuppression(true), maxDetectionCount <= 64");
        const int n = S.n, po = std::max(1, std::min(maxDet, 64));
        const bool gathered = S.gathered;
        S.gathered = false;
        std::vector<acfb_det> loc((size_t)n * po);
        std::vector<int> cnt(n);
        int tot = 0;
        collect(loc.data(), (int)loc.size(), cnt.data(), &tot); // this rank's own result; releases the slot, its buffers stay ours until the next submit
        const size_t hdr = postHeaderBytes(n), bytes = hdr + (size_t)n * po * sizeof(PostDet);
        bool gatheredNow = gathered;
        if (!gathered && S.posted)
        {   // the usual case: the gather is enqueued here, when this rank's batch is done and -- the hosts collect in step -- the other
            // ranks' batches are too, so the collective's kernel does not sit on an SM waiting for a rank that is still computing
            if (S.postBytes != bytes) throw std::runtime_error("engine: detection options changed between submit and the gather");
            allGatherPost(S);
            S.gathered = false;
            gatheredNow = true;
        }
        bool slow = !gatheredNow;
        if (gatheredNow)
        {
            if (S.postBytes != bytes) throw std::runtime_error("engine: detection options changed between submit and the gather");
            CUDA_OK(cudaEventSynchronize(S.gathDone));
            for (int r = 0; r < distWorld; r++) slow = slow || reinterpret_cast<const int*>(S.hGath + (size_t)r * bytes)[n] != 0;
        }
        if (slow)
        {
            S.postBytes = bytes;
            S.post.ensure(bytes);
            if (S.hPostCap < bytes)
            {
                if (S.hPost) cudaFreeHost(S.hPost);
                CUDA_OK(cudaMallocHost(&S.hPost, bytes));
                S.hPostCap = bytes;
            }
            memset(S.hPost, 0, bytes);
            int* hc = reinterpret_cast<int*>(S.hPost);
            PostDet* hd = reinterpret_cast<PostDet*>(S.hPost + hdr);
            int k = 0;
            for (int f = 0; f < n; f++)
            {
                hc[f] = cnt[f];
                for (int j = 0; j < cnt[f]; j++, k++) memcpy(&hd[(size_t)f * po + j], &loc[k], sizeof(PostDet));
            }
            CUDA_OK(cudaMemcpyAsync(S.post.p, S.hPost, bytes, cudaMemcpyHostToDevice, commStream));
            allGatherPost(S);
            S.gathered = false;
            CUDA_OK(cudaEventSynchronize(S.gathDone));
        }
        int written = 0, all = 0;
        if (distRank == 0)
            for (int r = 0; r < distWorld; r++)
            {
                const int* hc = reinterpret_cast<const int*>(S.hGath + (size_t)r * bytes);
                const acfb_det* hd = reinterpret_cast<const acfb_det*>(S.hGath + (size_t)r * bytes + hdr);
                for (int f = 0; f < n; f++)
                {
                    if (counts) counts[(size_t)r * n + f] = hc[f];
                    for (int j = 0; j < hc[f]; j++, all++)
                        if (written < cap && dets) { dets[written] = hd[(size_t)f * po + j]; dets[written].frame = r * n + f; written++; }
                }
            }
        if (total) *total = all;
    }

    void runCascade()
    {
        if (!cur) throw std::runtime_error("engine: no pyramid resident");
        if (anyPending()) throw std::runtime_error("engine: collect the submitted batches first");
        Slot& S = slots[subSlot];
        resetHits(S, curN, cur->groups.size());
        S.posted = false;
        cascadeRange(*cur, S, 0, curN);
        fetchCounters(S, curN, stream);
        S.st = cur; S.n = curN; S.pending = true;
        CUDA_OK(cudaEventRecord(S.done, stream));
        subSlot = (subSlot + 1) % kSlots;
    }

    // window (c, r) of a scale -> box in image coordinates (acfDetect1.cpp:326-332, ACF.cpp:302-311)
    acfb_det boxOf(int c, int r, float score, int frame, double scale, double scalehw_w, double scalehw_h) const
    {
        const int shift_w = (opt.modelDsPad_w - opt.modelDs_w) / 2 - opt.pad_w;
        const int shift_h = (opt.modelDsPad_h - opt.modelDs_h) / 2 - opt.pad_h;
        int rx = r * opt.stride, ry = c * opt.stride; // x/y swapped once there
        const int sw = (int)std::nearbyint(double(opt.modelDs_w) / scale); // cvRound
        const int sh = (int)std::nearbyint(double(opt.modelDs_h) / scale);
        rx = (int)(double(rx + shift_w) / scalehw_w);  // int truncation (A.2 Q8)
        ry = (int)(double(ry + shift_h) / scalehw_h);
        return acfb_det{ ry, rx, sh, sw, score, frame };
    }

    // The cascade on caller-provided channel buffers of one frame, one buffer per scale (nchn planes of w columns x h
    // values, y contiguous; float or byte): what Detector::operator()(const Pyramid&) does with a pyramid some other
    // producer filled (GLDetector.cpp:124).  Returns (scale index, c, r, score bits) in the reference's order.
    struct ChannelScale { const void* data; int h, w, nchn; };
    std::vector<int4> cascadeOnChannels(const std::vector<ChannelScale>& sc, bool u8, unsigned long long* trees)
    {
        if (anyPending()) throw std::runtime_error("collect the submitted batches first");
        CUDA_OK(cudaSetDevice(device));
        const size_t esz = u8 ? 1 : 4;
        const int modelHt = opt.modelDsPad_w, modelWd = opt.modelDsPad_h;
        const int needC = (opt.color_enabled ? (opt.color_space == 0 ? 1 : 3) : 0) + 1 + opt.gh_nOrients; // chnsCompute.cpp:146-338
        const bool tiled = tileCascade && !u8;
        std::vector<CascScale> cs(sc.size());
        std::vector<CascTileScale> ts(sc.size());
        size_t elems = 0; int64_t tasks = 0, windows = 0; int tiles = 0;
        for (size_t i = 0; i < sc.size(); i++)
        {
            const ChannelScale& q = sc[i];
            if (!q.data || q.h < 1 || q.w < 1) throw std::runtime_error("engine: empty channel buffer");
            if (q.nchn != needC) throw std::runtime_error("engine: channel count differs from the model's");
            CascScale& c = cs[i];
            // float channels are staged with 16-byte aligned columns (what the tensor maps of k_cascade_tile need); bytes as they come
            c.off = (int64_t)elems; c.P = u8 ? q.h : (q.h + 3) & ~3; c.planeStride = q.w * c.P;
            c.height1 = std::max(0, (int)ceil(float(q.h * opt.shrink - modelHt + 1) / opt.stride));
            c.width1 = std::max(0, (int)ceil(float(q.w * opt.shrink - modelWd + 1) / opt.stride));
            c.blk0 = (int)tasks; c.scaleIdx = (int)i;
            const int64_t nwin = (int64_t)c.height1 * c.width1;
            tasks += (nwin + kCascTask - 1) / kCascTask; windows += nwin;
            elems += ((size_t)q.nchn * c.planeStride + 31) & ~(size_t)31;
            if (tiled)
            {
                CascTileScale& t = ts[i];
                t.tile0 = tiles; t.width1 = c.width1; t.height1 = c.height1; t.scaleIdx = (int)i;
                t.nTx = nwin > 0 ? (c.width1 + tileGeom.Wc - 1) / tileGeom.Wc : 0; t.nTy = nwin > 0 ? (c.height1 + tileGeom.Wr - 1) / tileGeom.Wr : 1;
                tiles += t.nTx * t.nTy;
            }
        }
        scratch.ensure((elems * esz + 3) / 4 + 4);
        scratchScale.ensure(std::max<size_t>(1, cs.size()));
        for (size_t i = 0; i < sc.size(); i++)
            CUDA_OK(cudaMemcpy2DAsync((uint8_t*)scratch.p + (size_t)cs[i].off * esz, (size_t)cs[i].P * esz, sc[i].data, (size_t)sc[i].h * esz, (size_t)sc[i].h * esz,
                                      (size_t)sc[i].nchn * sc[i].w, cudaMemcpyHostToDevice, stream));
        if (!cs.empty()) CUDA_OK(cudaMemcpyAsync(scratchScale.p, cs.data(), cs.size() * sizeof(CascScale), cudaMemcpyHostToDevice, stream));
        const int hcap = (int)std::max<int64_t>(1, windows);
        scratchHits.ensure(hcap);
        CUDA_OK(cudaMemsetAsync(scratchCount.p, 0, 2 * sizeof(int), stream));
        CUDA_OK(cudaMemsetAsync(scratchStats.p, 0, 4 * sizeof(unsigned long long), stream));
        if (tiled)
        {
            std::vector<CUtensorMap> maps(2 * sc.size()); // box = a cascade tile, then box = one window's footprint
            for (size_t i = 0; i < sc.size(); i++)
            {
                maps[i] = channelTileMap(scratch.p + cs[i].off, sc[i].h, sc[i].w, sc[i].nchn, 1, cs[i].P, cs[i].planeStride, 0, tileGeom);
                maps[sc.size() + i] = channelTileMap(scratch.p + cs[i].off, sc[i].h, sc[i].w, sc[i].nchn, 1, cs[i].P, cs[i].planeStride, 0, winGeom);
            }
            scratchMaps.ensure(std::max<size_t>(1, maps.size()));
            scratchTileScale.ensure(std::max<size_t>(1, ts.size()));
            if (!maps.empty())
            {
                CUDA_OK(cudaMemcpyAsync(scratchMaps.p, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, stream));
                CUDA_OK(cudaMemcpyAsync(scratchTileScale.p, ts.data(), ts.size() * sizeof(CascTileScale), cudaMemcpyHostToDevice, stream));
                CUDA_OK(cudaStreamSynchronize(stream)); // pageable locals
            }
            CascTileArgs t{};
            t.maps = scratchMaps.p; t.scales = scratchTileScale.p; t.nScales = (int)ts.size(); t.tilesPerFrame = tiles; t.n = 1; t.frame0 = 0;
            t.tab = cascTabTile.p; t.nTrees = model.nTrees();
            t.Wc = tileGeom.Wc; t.Wr = tileGeom.Wr; t.BY = tileGeom.BY; t.step = tileGeom.step; t.tileBytes = tileGeom.tileBytes; t.boxBytes = tileGeom.boxBytes;
            t.listCap = tileGeom.listCap; t.smemBytes = tileGeom.smemBytes; t.sparseMax = cascSparseMax; t.blocksPerSm = cascBlocksPerSm; t.cascThr = (float)opt.cascThr;
            t.headLevels = cascHeadLevels; t.headTrees = (int)tileHead.size(); if (!tileHead.empty()) memcpy(t.head, tileHead.data(), sizeof(t.head));
            t.hitCount = scratchCount.p; t.hits = scratchHits.p; t.cap = hcap; t.stats = scratchStats.p; t.taskCounter = scratchStats.p + 2;
            bool packs = sc.size() <= 256;
            for (const CascScale& c : cs) packs = packs && c.width1 < 65536 && c.height1 < 65536;
            t.exportMax = packs ? cascExportMax : 0;
            if (t.exportMax > 0)
            {
                scratchTail.ensure(1 << 16);
                t.tail = scratchTail.p; t.tailCount = scratchCount.p + 1; t.tailCap = 1 << 16;
            }
            if (tiles > 0)
            {
                launchCascadeTile(t, stream); launches++;
                if (t.exportMax > 0 && tailOnWindows)
                {
                    launchCascadeTailWin(tailWinArgs(t, scratchMaps.p + sc.size()), stream); launches++;
                }
                else if (t.exportMax > 0)
                {
                    CascTailArgs q{};
                    q.pyr = scratch.p; q.frameStride = 0; q.scales = scratchScale.p; q.tab = cascTab.p; q.nTrees = model.nTrees();
                    q.stride = opt.stride; q.shrink = opt.shrink; q.cascThr = (float)opt.cascThr;
                    q.tail = t.tail; q.tailCount = t.tailCount; q.tailCap = t.tailCap; q.hitCount = t.hitCount; q.hits = t.hits; q.cap = hcap; q.stats = scratchStats.p;
                    launchCascadeTail(q, stream); launches++;
                }
            }
        }
        else
        {
        CascArgs a{};
        a.pyr = scratch.p; a.u8 = u8 ? 1 : 0; a.frameStride = 0; a.scales = scratchScale.p; a.nScales = (int)cs.size();
        a.nBlocksPerFrame = (int)tasks; a.n = 1;
        a.tab = u8 ? cascTabU8.p : cascTab.p; a.nTrees = model.nTrees(); a.depth = model.clf.treeDepth; a.recWords = recWords;
        a.stride = opt.stride; a.shrink = opt.shrink; a.cascThr = (float)opt.cascThr;
        a.hitCount = scratchCount.p; a.hits = scratchHits.p; a.cap = hcap; a.stats = scratchStats.p; a.taskCounter = scratchStats.p + 2;
        if (tasks > 0) { launchCascade(a, stream); launches++; }
        }
        int cnt = 0;
        unsigned long long st[2] = { 0, 0 };
        CUDA_OK(cudaMemcpyAsync(&cnt, scratchCount.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaMemcpyAsync(st, scratchStats.p, sizeof(st), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        std::vector<int4> hh(cnt);
        if (cnt) CUDA_OK(cudaMemcpy(hh.data(), scratchHits.p, (size_t)cnt * sizeof(int4), cudaMemcpyDeviceToHost));
        std::sort(hh.begin(), hh.end(), [](const int4& x, const int4& y) { return x.x != y.x ? x.x < y.x : x.y != y.y ? x.y < y.y : x.z < y.z; });
        if (trees) *trees = st[0];
        hStats[0] = st[0]; hStats[1] = st[1];
        return hh;
    }

    // host tail: order hits like the reference's loops, rescale (ACF.cpp:302-311), optional NMS + prune
    void collect(acfb_det* dets, int cap, int* counts, int* total)
    {
        Slot& S = slots[colSlot];
        if (!S.pending) throw std::runtime_error("engine: nothing submitted");
        SizeState& st = *S.st;
        const int n = S.n;
        CUDA_OK(cudaSetDevice(device));
        const auto tEnter = std::chrono::steady_clock::now();
        CUDA_OK(cudaEventSynchronize(S.done));
        hHitCount.assign(S.hCount, S.hCount + n);
        hStats[0] = S.hStats[0]; hStats[1] = S.hStats[1];
        int maxCount = 0;
        for (int f = 0; f < n; f++)
        {
            if (hHitCount[f] > hitCap)
            {   // the batch is consumed (its records are incomplete): release the slot so the engine stays usable
                S.pending = false;
                colSlot = (colSlot + 1) % kSlots;
                throw std::runtime_error("engine: per-frame hit buffer overflow; raise it with acfb_set_hit_capacity");
            }
            maxCount = std::max(maxCount, hHitCount[f]);
        }
        lastHitTotal = 0;
        for (int f = 0; f < n; f++) lastHitTotal += hHitCount[f];
        if (S.posted && reinterpret_cast<const int*>(S.hPost)[n] == 0)
        {   // ordering, rescale, bbNms and prune already ran on the device (k_post): copy its boxes out
            const auto tData = std::chrono::steady_clock::now();
            const int* dc = reinterpret_cast<const int*>(S.hPost);
            const acfb_det* dd = reinterpret_cast<const acfb_det*>(S.hPost + postHeaderBytes(n));
            const int po = (int)((S.postBytes - postHeaderBytes(n)) / sizeof(PostDet) / n);
            S.pending = false;
            colSlot = (colSlot + 1) % kSlots;
            lastHits.clear(); lastHitsValid = false;
            int written = 0, tot = 0;
            for (int f = 0; f < n; f++)
            {
                if (counts) counts[f] = dc[f];
                for (int k = 0; k < dc[f]; k++)
                {
                    if (written < cap && dets) dets[written++] = dd[(size_t)f * po + k];
                    tot++;
                }
            }
            if (total) *total = tot;
            const auto tEnd = std::chrono::steady_clock::now();
            collectWaitMs = std::chrono::duration<double, std::milli>(tData - tEnter).count();
            collectTailMs = std::chrono::duration<double, std::milli>(tEnd - tData).count();
            return;
        }
        lastHitsValid = true;
        hHits.resize((size_t)n * std::max(1, maxCount));
        if (maxCount > 0)
        {
            // separate stream: must not queue behind the next batch's kernels
            CUDA_OK(cudaMemcpy2DAsync(hHits.data(), (size_t)maxCount * sizeof(int4), S.hits.p, (size_t)hitCap * sizeof(int4),
                                      (size_t)maxCount * sizeof(int4), n, cudaMemcpyDeviceToHost, d2hStream));
            CUDA_OK(cudaStreamSynchronize(d2hStream));
        }
        if (!anyPending()) { mark("d2h"); }
        const auto tData = std::chrono::steady_clock::now();
        finishTiming();
        S.pending = false;
        colSlot = (colSlot + 1) % kSlots;
        lastHits.clear();
        const Plan& P = st.plan;
        int written = 0, tot = 0;
        std::vector<acfb_det> frameDets;
        for (int f = 0; f < n; f++)
        {
            int4* hb = hHits.data() + (size_t)f * std::max(1, maxCount);
            const int cnt = hHitCount[f];
            std::sort(hb, hb + cnt, [](const int4& a, const int4& b) {
                if (a.x != b.x) return a.x < b.x;
                if (a.y != b.y) return a.y < b.y;
                return a.z < b.z;
            });
            frameDets.clear();
            for (int k = 0; k < cnt; k++)
            {
                const int s = hb[k].x, c = hb[k].y, r = hb[k].z;
                float score;
                memcpy(&score, &hb[k].w, 4);
                lastHits.push_back(acfb_hit{ f, s, c, r, score });
                frameDets.push_back(boxOf(c, r, score, f, P.scales[s], P.scaleshw[s].first, P.scaleshw[s].second));
            }
            if (doNms && !frameDets.empty()) nmsAndPrune(frameDets);
            if (counts) counts[f] = (int)frameDets.size();
            for (auto& d : frameDets)
            {
                if (written < cap && dets) dets[written++] = d;
                tot++;
            }
        }
        if (total) *total = tot;
        const auto tEnd = std::chrono::steady_clock::now();
        collectWaitMs = std::chrono::duration<double, std::milli>(tData - tEnter).count();
        collectTailMs = std::chrono::duration<double, std::milli>(tEnd - tData).count();
    }

    // bbNms.cpp:229-304 -> nmsMax :111-192, then ObjectDetector::prune :28-44
    void nmsAndPrune(std::vector<acfb_det>& bbs)
    {
        const std::string type = opt.nms_type;
        const bool greedy = (type == "maxg");
        // 'none' returns the boxes as they are and 'ms' / 'cover' are identity stubs in the reference (bbNms.cpp:100-108,
        // 229-304): only the suppression is skipped, ObjectDetector::prune still runs (ACF.cpp:334-351)
        const bool suppress = (type == "max" || type == "maxg");
        const bool ovrUnion = std::string(opt.nms_ovrDnm) != "min";
        if (suppress) std::stable_sort(bbs.begin(), bbs.end(), [](const acfb_det& a, const acfb_det& b) { return a.score > b.score; });
        const size_t n = bbs.size();
        std::vector<char> kp(n, 1);
        if (suppress)
        for (size_t i = 0; i < n; i++)
        {
            if (greedy && !kp[i]) continue;
            const int ixe = bbs[i].x + bbs[i].w, iye = bbs[i].y + bbs[i].h, ias = bbs[i].w * bbs[i].h;
            for (size_t j = i + 1; j < n; j++)
            {
                if (!kp[j]) continue;
                const int iw = std::min(ixe, bbs[j].x + bbs[j].w) - std::max(bbs[i].x, bbs[j].x);
                if (iw <= 0) continue;
                const int ih = std::min(iye, bbs[j].y + bbs[j].h) - std::max(bbs[i].y, bbs[j].y);
                if (ih <= 0) continue;
                double o = (iw * ih);
                const int jas = bbs[j].w * bbs[j].h;
                const double u = ovrUnion ? (ias + jas - o) : std::min(ias, jas);
                o /= u;
                if (o > opt.nms_overlap) kp[j] = 0;
            }
        }
        size_t m = 0;
        for (size_t i = 0; i < n; i++) if (kp[i]) bbs[m++] = bbs[i];
        bbs.resize(m);
        if (bbs.size() > 1)
        {
            size_t cutoff = 1;
            for (size_t i = 1; i < std::min<size_t>((size_t)maxDet, bbs.size()); i++)
            {
                cutoff = i + 1;
                if ((double)bbs[i].score < ((double)bbs[0].score * pruneRatio)) break;
            }
            bbs.resize(cutoff);
        }
    }

    void finishTiming()
    {
        stageMs.clear(); stageNames.clear();
        if (!timing || evs.size() < 2) return;
        for (size_t i = 1; i < evs.size(); i++)
        {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, evs[i - 1], evs[i]) != cudaSuccess) continue;
            size_t k = 0;
            while (k < stageNames.size() && strcmp(stageNames[k], evNames[i]) != 0) k++;
            if (k == stageNames.size()) { stageNames.push_back(evNames[i]); stageMs.push_back(0.f); }
            stageMs[k] += ms;
        }
    }
};

} // namespace acfb

// The handle: the engine, plus -- created the first time a batch is submitted while others are still in flight -- up to two
// more PIPELINES: complete further sets of buffers and streams on the same device.  Batches submitted back to back go round the
// pipelines, so the kernels of the following batches fill what the current one leaves idle (k_front's 256 plane marches occupy 108
// SMs twice and 40 once, every kernel's last wave runs partly empty, the cascade's thin late levels): measured 18.4 ms per 256
// frames with one pipeline, 16.76 with two, 16.40 with three, 16.50 with four.  Results are collected in submission order; up to two
// batches per pipeline may be in flight.  Everything that is not submit / collect runs on the first pipeline.
// ACFB_PIPELINES=1 / 2 keep fewer (a pipeline is ~94 MB of device memory per 1080p frame of batch capacity).
struct acfb_engine
{
    acfb::Engine e;
    std::vector<std::unique_ptr<acfb::Engine>> extra; // pipelines 1 .. pipes - 1, created on demand
    int pipes = 3;
    int maxInFlight = 6; // ACFB_MAX_IN_FLIGHT (at most kSlots per pipeline)
    int lastSubmitted = 0, lastCollected = 0;
    unsigned long long distBatch = 0; // batches handed to acfb_dist_collect (the exchange's sequence number, same on every rank)
    std::vector<int> order; // pipeline of every batch not yet collected, oldest first

    acfb::Engine& pipe(int k) { return k ? *extra[k - 1] : e; }
    bool pending() const
    {
        if (e.anyPending()) return true;
        for (const auto& x : extra) if (x && x->anyPending()) return true;
        return false;
    }
    void linkComm()
    {
        if (!e.comm && !e.exch) return;
        for (auto& x : extra)
        {
            if (!x || x->comm || x->exch) continue;
            x->comm = e.comm; x->exch = e.exch; x->ownsComm = false; x->commStream = e.commStream; x->distRank = e.distRank; x->distWorld = e.distWorld;
            if (e.comm) x->distCreateLocals(); // its own events; the gathers of all pipelines run in submission order on the one communication stream
        }
    }
    void syncSettings(acfb::Engine& x)
    {
        x.doNms = e.doNms; x.maxDet = e.maxDet; x.pruneRatio = e.pruneRatio; x.pixfmt = e.pixfmt; x.isTranspose = e.isTranspose;
        x.isLuv = e.isLuv; x.keepC = e.keepC;
        if (x.hitCap != e.hitCap)
        {
            x.hitCap = e.hitCap;
            for (auto& s : x.slots) s.hits.release();
        }
    }
    acfb::Engine& forSubmit()
    {
        int k = 0;
        if (pipes > 1 && !e.timing && !order.empty()) k = (lastSubmitted + 1) % pipes;
        if (k > 0)
        {
            if ((int)extra.size() < pipes - 1) extra.resize(pipes - 1);
            if (!extra[k - 1])
            {
                std::unique_ptr<acfb::Engine> x(new acfb::Engine());
                x->model = e.model; x->opt = e.opt; x->device = e.device; x->maxRows = e.maxRows; x->maxCols = e.maxCols; x->maxBatch = e.maxBatch;
                x->init();
                extra[k - 1] = std::move(x);
                linkComm();
            }
            syncSettings(*extra[k - 1]);
        }
        lastSubmitted = k;
        return pipe(k);
    }
    acfb::Engine& forCollect()
    {
        if (order.empty()) throw std::runtime_error("engine: nothing submitted");
        lastCollected = order.front();
        order.erase(order.begin());
        return pipe(lastCollected);
    }
};

using namespace acfb;

#define API_BEGIN try {
#define API_END                                                       \
    return 0; }                                                       \
    catch (const std::exception& ex) { g_err = ex.what(); return 1; } \
    catch (...) { g_err = "unknown error"; return 2; }

extern "C" {

const char* acfb_last_error(void) { return g_err.c_str(); }
const char* acfb_version(void) { return "acf_b200 0.1 (sm_100a)"; }

int acfb_model_load(const void* cpb, size_t nbytes, acfb_model** out)
{
    API_BEGIN
    if (!cpb || !out) throw std::runtime_error("null argument");
    std::unique_ptr<acfb_model> m(new acfb_model());
    m->m = cpbRead((const uint8_t*)cpb, nbytes);
    *out = m.release();
    API_END
}

int acfb_model_load_file(const char* path, acfb_model** out)
{
    API_BEGIN
    if (!path || !out) throw std::runtime_error("null argument");
    if (std::string(path).find(".cpb") == std::string::npos) // ACFIO.cpp:204-208 picks the parser by file-name substring
        throw std::runtime_error("only .cpb models are supported (the .mat path needs cvmatio)");
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::unique_ptr<acfb_model> m(new acfb_model());
    m->m = cpbRead((const uint8_t*)buf.data(), buf.size());
    *out = m.release();
    API_END
}

int acfb_model_create(const acfb_options* opts, const acfb_classifier* clf, acfb_model** out)
{
    API_BEGIN
    if (!opts || !clf || !out) throw std::runtime_error("null argument");
    std::unique_ptr<acfb_model> m(new acfb_model());
    m->m = Model::fromFlat(*opts, *clf);
    *out = m.release();
    API_END
}

int acfb_model_save(const acfb_model* m, void* buf, size_t cap, size_t* nbytes)
{
    API_BEGIN
    if (!m || !nbytes) throw std::runtime_error("null argument");
    const std::vector<uint8_t> b = cpbWrite(m->m);
    *nbytes = b.size();
    if (buf && cap >= b.size()) memcpy(buf, b.data(), b.size());
    else if (buf) throw std::runtime_error("buffer too small");
    API_END
}

int acfb_model_save_file(const acfb_model* m, const char* path)
{
    API_BEGIN
    if (!m || !path) throw std::runtime_error("null argument");
    const std::vector<uint8_t> b = cpbWrite(m->m);
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    f.write((const char*)b.data(), (std::streamsize)b.size());
    API_END
}

int acfb_model_options(const acfb_model* m, acfb_options* out)
{
    API_BEGIN
    if (!m || !out) throw std::runtime_error("null argument");
    *out = m->m.flat();
    API_END
}

int acfb_model_classifier(const acfb_model* m, acfb_classifier* out)
{
    API_BEGIN
    if (!m || !out) throw std::runtime_error("null argument");
    const Classifier& c = m->m.clf;
    out->nTrees = c.fids.rows; out->nTreeNodes = c.fids.cols; out->treeDepth = c.treeDepth;
    out->fids = c.fids.ptr<uint32_t>(); out->thrs = c.thrs.ptr<float>(); out->child = c.child.ptr<uint32_t>();
    out->hs = c.hs.ptr<float>();
    out->weights = c.weights.bytes.empty() ? nullptr : c.weights.ptr<float>();
    out->depth = c.depth.bytes.empty() ? nullptr : c.depth.ptr<uint32_t>();
    API_END
}

int acfb_model_modify(acfb_model* m, double cascCal, double cascThr, int stride)
{
    API_BEGIN
    if (!m) throw std::runtime_error("null argument");
    Options& o = m->m.opts;
    if (!std::isnan(cascThr)) o.cascThr.set("cascThr", cascThr);
    if (stride > 0) o.stride.set("stride", stride);
    o.cascCal.set("cascCal", cascCal);
    const double shrink = o.pPyramid.value.pChns.value.shrink.value;
    o.stride.value = (int)(std::max(1.0, std::round(double(o.stride.value) / shrink)) * shrink); // acfModify.cpp:139
    float* hs = m->m.clf.hs.ptr<float>();
    const size_t n = (size_t)m->m.clf.hs.rows * m->m.clf.hs.cols;
    for (size_t i = 0; i < n; i++) hs[i] = (float)((double)hs[i] + cascCal); // cv::Mat += scalar (saturate_cast<float>(double sum))
    API_END
}

int acfb_model_modify_ex(acfb_model* m, const acfb_modify* p)
{
    API_BEGIN
    if (!m || !p) throw std::runtime_error("null argument");
    Model w = m->m; // work on a copy: a rejected modification leaves the model as it was
    Options& o = w.opts;
    acfb::Pyramid& py = o.pPyramid.value;
    if (p->has_nPerOct) py.nPerOct.set("nPerOct", p->nPerOct);
    if (p->has_nOctUp) py.nOctUp.set("nOctUp", p->nOctUp);
    if (p->has_nApprox) py.nApprox.set("nApprox", p->nApprox);
    if (p->has_lambdas)
    {
        if (p->nLambdas < 0 || p->nLambdas > 8) throw std::runtime_error("modify: at most 8 lambdas");
        py.lambdas.set("lambdas", std::vector<double>(p->lambdas, p->lambdas + p->nLambdas));
    }
    if (p->has_pad) { Size z; z.width = p->pad_w; z.height = p->pad_h; py.pad.set("pad", z); }
    if (p->has_minDs) { Size z; z.width = p->minDs_w; z.height = p->minDs_h; py.minDs.set("minDs", z); }
    if (p->has_nms)
    {
        Nms& n = o.pNms.value;
        o.pNms.has = true; if (o.pNms.name.empty()) o.pNms.name = "pNms";
        n.type.set("type", std::string(p->nms_type, strnlen(p->nms_type, sizeof(p->nms_type))));
        n.overlap.set("overlap", p->nms_overlap);
        n.ovrDnm.set("ovrDnm", std::string(p->nms_ovrDnm, strnlen(p->nms_ovrDnm, sizeof(p->nms_ovrDnm))));
    }
    if (p->has_stride) o.stride.set("stride", p->stride);
    if (p->has_cascThr) o.cascThr.set("cascThr", p->cascThr);
    o.cascCal.set("cascCal", p->cascCal);
    const double shrink = py.pChns.value.shrink.value;
    o.stride.value = (int)(std::max(1.0, std::round(double(o.stride.value) / shrink)) * shrink); // acfModify.cpp:139
    float* hs = w.clf.hs.ptr<float>();
    const size_t n = (size_t)w.clf.hs.rows * w.clf.hs.cols;
    for (size_t i = 0; i < n; i++) hs[i] = (float)((double)hs[i] + p->cascCal); // cv::Mat += scalar (saturate_cast<float>(double sum))
    w.validate();
    m->m = std::move(w);
    API_END
}

void acfb_model_destroy(acfb_model* m) { delete m; }

int acfb_engine_create(const acfb_model* m, int device, int max_rows, int max_cols, int max_batch, acfb_engine** out)
{
    API_BEGIN
    if (!m || !out) throw std::runtime_error("null argument");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        throw std::runtime_error(std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) throw std::runtime_error("bad device index");
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) throw std::runtime_error("device is not sm_100-class; the kernels are built for sm_100a only");
    std::unique_ptr<acfb_engine> e(new acfb_engine());
    e->e.model = m->m;
    e->e.model.validate();
    e->e.opt = m->m.flat();
    e->e.device = device; e->e.maxRows = max_rows; e->e.maxCols = max_cols; e->e.maxBatch = std::max(1, max_batch);
    e->e.init();
    if (const char* pp = getenv("ACFB_PIPELINES")) e->pipes = std::max(1, std::min(atoi(pp), 3));
    e->maxInFlight = e->pipes == 1 ? Engine::kSlots : 2 * e->pipes; // two per pipeline keep every stream fed
    if (const char* mf = getenv("ACFB_MAX_IN_FLIGHT")) e->maxInFlight = std::max(1, std::min(atoi(mf), Engine::kSlots * e->pipes));
    *out = e.release();
    API_END
}

void acfb_engine_destroy(acfb_engine* e)
{
    if (!e) return;
    cudaSetDevice(e->e.device);
    if (e->e.stream) cudaStreamSynchronize(e->e.stream);
    for (auto& x : e->extra) if (x && x->stream) cudaStreamSynchronize(x->stream);
    delete e;
}

int acfb_set_nms(acfb_engine* e, int enable) { API_BEGIN if (!e) throw std::runtime_error("null engine"); e->e.doNms = enable != 0; API_END }
int acfb_set_max_detection_count(acfb_engine* e, int n) { API_BEGIN if (!e) throw std::runtime_error("null engine"); e->e.maxDet = n; API_END }
int acfb_set_detection_score_prune_ratio(acfb_engine* e, double r) { API_BEGIN if (!e) throw std::runtime_error("null engine"); e->e.pruneRatio = r; API_END }
int acfb_set_input_format(acfb_engine* e, int format)
{
    API_BEGIN
    if (!e || format < 0 || format > 7) throw std::runtime_error("bad pixel format (0 RGB24, 1 BGR24, 2 RGBA32, 3 BGRA32, 4 GRAY8, 5 RGB32F, 6 PLANAR32F, 7 NV12)");
    if (e->pending()) throw std::runtime_error("collect the submitted batches first");
    const int old = e->e.pixfmt;
    e->e.pixfmt = format;
    try { e->e.colorMode(); } catch (...) { e->e.pixfmt = old; throw; }
    API_END
}

int acfb_set_is_transpose(acfb_engine* e, int flag)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    if (e->pending()) throw std::runtime_error("collect the submitted batches first");
    e->e.isTranspose = flag != 0;
    API_END
}

int acfb_set_is_luv(acfb_engine* e, int flag)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    if (e->pending()) throw std::runtime_error("collect the submitted batches first");
    const bool old = e->e.isLuv;
    e->e.isLuv = flag != 0;
    try { e->e.colorMode(); } catch (...) { e->e.isLuv = old; throw; }
    API_END
}

int acfb_set_hit_capacity(acfb_engine* e, int cap)
{
    API_BEGIN
    if (!e || cap < 1) throw std::runtime_error("bad argument");
    CUDA_OK(cudaSetDevice(e->e.device));
    CUDA_OK(cudaStreamSynchronize(e->e.stream));
    e->e.hitCap = cap;
    if (e->pending()) throw std::runtime_error("collect the submitted batches first");
    for (auto& s : e->e.slots) s.hits.release();
    API_END
}

int acfb_get_scales(const acfb_options* o, int rows, int cols, double* scales, double* scaleshw, int cap, int* nscales)
{
    API_BEGIN
    if (!o || !nscales) throw std::runtime_error("null argument");
    if (o->nPerOct < 1 || o->nPerOct > 64 || o->nOctUp < 0 || o->nOctUp > 8 || o->minDs_w < 1 || o->minDs_h < 1 || o->shrink < 1)
        throw std::runtime_error("acfb_get_scales: nPerOct in 1..64, nOctUp in 0..8, minDs >= 1, shrink >= 1 expected");
    std::vector<double> s; std::vector<std::pair<double, double>> hw;
    getScales(o->nPerOct, o->nOctUp, o->minDs_w, o->minDs_h, o->shrink, rows, cols, s, hw);
    *nscales = (int)s.size();
    for (int i = 0; i < (int)s.size() && i < cap; i++)
    {
        if (scales) scales[i] = s[i];
        if (scaleshw) { scaleshw[2 * i] = hw[i].first; scaleshw[2 * i + 1] = hw[i].second; }
    }
    API_END
}

int acfb_plan(acfb_engine* e, int rows, int cols, acfb_scale_info* out, int cap, int* nscales, int64_t* floats_per_frame)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    CUDA_OK(cudaSetDevice(e->e.device));
    SizeState& st = e->e.sizeState(rows, cols);
    const Plan& P = st.plan;
    if (nscales) *nscales = (int)P.geom.size();
    if (floats_per_frame) *floats_per_frame = P.floatsPerFrame;
    for (int i = 0; i < (int)P.geom.size() && i < cap && out; i++)
    {
        const ScaleGeom& g = P.geom[i];
        out[i].scale = g.scale; out[i].scalehw_w = g.shw_w; out[i].scalehw_h = g.shw_h;
        out[i].h = g.H; out[i].w = g.W; out[i].pitch = g.P; out[i].nchn = P.nChns; out[i].is_real = g.isReal;
        out[i].real_index = P.reals[g.realK].scaleIdx; out[i].offset = g.offset;
    }
    API_END
}

int acfb_pyramid(acfb_engine* e, const uint8_t* frames, int n, int rows, int cols, int on_device)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    e->e.runPyramid(frames, n, rows, cols, on_device != 0);
    CUDA_OK(cudaStreamSynchronize(e->e.stream));
    API_END
}

int acfb_pyramid_device_ptr(acfb_engine* e, int frame, const float** dptr)
{
    API_BEGIN
    if (!e || !e->e.cur || frame < 0 || frame >= e->e.curN) throw std::runtime_error("no such frame");
    *dptr = e->e.cur->pyr.p + (size_t)frame * e->e.cur->plan.floatsPerFrame;
    API_END
}

int acfb_pyramid_read(acfb_engine* e, int frame, int scale, float* host_out, size_t cap_floats)
{
    API_BEGIN
    if (!e || !e->e.cur || frame < 0 || frame >= e->e.curN) throw std::runtime_error("no such frame");
    SizeState& st = *e->e.cur;
    const Plan& P = st.plan;
    if (scale < 0 || scale >= (int)P.geom.size()) throw std::runtime_error("no such scale");
    const ScaleGeom& g = P.geom[scale];
    const size_t need = (size_t)P.nChns * g.W * g.H;
    if (cap_floats < need) throw std::runtime_error("output buffer too small");
    CUDA_OK(cudaSetDevice(e->e.device));
    const float* src = st.pyr.p + (size_t)frame * P.floatsPerFrame + g.offset;
    CUDA_OK(cudaMemcpy2DAsync(host_out, (size_t)g.H * sizeof(float), src, (size_t)g.P * sizeof(float), (size_t)g.H * sizeof(float),
                              (size_t)P.nChns * g.W, cudaMemcpyDeviceToHost, e->e.stream));
    CUDA_OK(cudaStreamSynchronize(e->e.stream));
    API_END
}

int acfb_pyramid_lambdas(acfb_engine* e, double* out, int cap, int* n)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    std::vector<double> l = e->e.lambdasFromImage;
    if (l.empty()) l.assign(e->e.opt.lambdas, e->e.opt.lambdas + e->e.opt.nLambdas);
    if (n) *n = (int)l.size();
    for (int i = 0; i < (int)l.size() && i < cap && out; i++) out[i] = l[i];
    API_END
}

int acfb_detect_pyramid(acfb_engine* e, acfb_det* dets, int cap, int* counts, int* total)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    e->e.runCascade();
    e->e.collect(dets, cap, counts, total);
    API_END
}

int acfb_submit(acfb_engine* e, const uint8_t* frames, int n, int rows, int cols, int on_device)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    if (e->order.size() >= (size_t)e->maxInFlight) throw std::runtime_error("engine: " + std::to_string(e->maxInFlight) + " batches already in flight; call acfb_collect first");
    Engine& P = e->forSubmit();
    P.submitAll(frames, n, rows, cols, on_device != 0);
    e->order.push_back(e->lastSubmitted);
    API_END
}

int acfb_collect(acfb_engine* e, acfb_det* dets, int cap, int* counts, int* total)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    e->forCollect().collect(dets, cap, counts, total);
    API_END
}

// 128 bytes that name one communicator: NCCL's unique id when the exchange is NCCL, else random bytes (they name the shared segment)
static void distMakeId(uint8_t id[128])
{
    if (Engine::distUseNccl())
    {
        NcclUniqueId uid;
        NCCL_OK(NcclApi::get().GetUniqueId(&uid));
        memcpy(id, uid.internal, sizeof(uid.internal));
        return;
    }
    std::ifstream ur("/dev/urandom", std::ios::binary);
    ur.read(reinterpret_cast<char*>(id), 128);
    if (!ur) throw std::runtime_error("cannot read /dev/urandom");
}

int acfb_dist_unique_id(uint8_t id[128])
{
    API_BEGIN
    if (!id) throw std::runtime_error("null id");
    distMakeId(id);
    API_END
}

int acfb_dist_init_rank(acfb_engine* e, const uint8_t id[128], int rank, int world)
{
    API_BEGIN
    if (!e || !id) throw std::runtime_error("null engine / id");
    if (e->pending()) throw std::runtime_error("collect the submitted batches first");
    e->e.distInitRank(id, rank, world);
    e->linkComm();
    API_END
}

int acfb_dist_init_all(acfb_engine** engines, int n)
{
    API_BEGIN
    if (!engines || n < 1) throw std::runtime_error("no engines");
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++)
    {
        if (!engines[i]) throw std::runtime_error("null engine");
        if (engines[i]->e.comm || engines[i]->e.exch) throw std::runtime_error("engine: the communicator exists already");
        devs[i] = engines[i]->e.device;
        for (int j = 0; j < i; j++)
            if (devs[j] == devs[i]) throw std::runtime_error("acfb_dist_init_all: one engine per device");
    }
    uint8_t id[128];
    distMakeId(id);
    if (!Engine::distUseNccl())
    {
        for (int i = 0; i < n; i++) { engines[i]->e.distInitRank(id, i, n); engines[i]->linkComm(); } // rank 0 creates the segment, the others attach
        return 0;
    }
    // ncclCommInitAll with a configuration: one process creates the id and joins every rank inside a group
    std::vector<NcclComm> comms(n, nullptr);
    NcclUniqueId uid;
    memcpy(uid.internal, id, sizeof(uid.internal));
    NcclConfig cfg = Engine::distConfig();
    NCCL_OK(NcclApi::get().GroupStart());
    for (int i = 0; i < n; i++)
    {
        CUDA_OK(cudaSetDevice(devs[i]));
        NCCL_OK(NcclApi::get().CommInitRankConfig(&comms[i], n, uid, i, &cfg));
    }
    NCCL_OK(NcclApi::get().GroupEnd());
    for (int i = 0; i < n; i++)
    {
        engines[i]->e.distCreateLocals();
        engines[i]->e.comm = comms[i]; engines[i]->e.distRank = i; engines[i]->e.distWorld = n;
        engines[i]->linkComm();
    }
    API_END
}

// Host-only self test of the shared-memory exchange (no device involved): `world` threads publish `batches` records each through
// one segment, rank 0 takes them in order; the ring wraps (batches > 8), so the flow control is exercised too.
int acfb_selftest_exchange(int world, int batches, int slot_bytes)
{
    API_BEGIN
    if (world < 1 || world > 64 || batches < 1 || slot_bytes < 64) throw std::runtime_error("bad argument");
    uint8_t id[128];
    std::ifstream ur("/dev/urandom", std::ios::binary);
    ur.read(reinterpret_cast<char*>(id), 128);
    if (!ur) throw std::runtime_error("cannot read /dev/urandom");
    std::vector<std::unique_ptr<ShmExchange>> ex(world);
    for (int r = 0; r < world; r++) ex[r].reset(ShmExchange::open(id, r, world, (size_t)slot_bytes));
    std::vector<std::string> errs(world);
    std::vector<std::thread> th;
    for (int r = 0; r < world; r++)
        th.emplace_back([&, r] {
            try
            {
                std::vector<unsigned char> buf(slot_bytes);
                for (int b = 0; b < batches; b++)
                {
                    const size_t bytes = 64 + (size_t)((b * 131 + r * 17) % (slot_bytes - 63));
                    for (size_t i = 0; i < bytes; i++) buf[i] = (unsigned char)(b * 7 + r * 13 + i);
                    ex[r]->publish((unsigned long long)b, buf.data(), bytes);
                    if (r == 0)
                    {
                        for (int q = 0; q < world; q++)
                        {
                            size_t got = 0;
                            const unsigned char* p = ex[0]->wait((unsigned long long)b, q, &got);
                            const size_t want = 64 + (size_t)((b * 131 + q * 17) % (slot_bytes - 63));
                            if (got != want) throw std::runtime_error("record size differs");
                            for (size_t i = 0; i < got; i++)
                                if (p[i] != (unsigned char)(b * 7 + q * 13 + i)) throw std::runtime_error("record bytes differ");
                        }
                        ex[0]->consumed((unsigned long long)b);
                    }
                }
            }
            catch (const std::exception& e) { errs[r] = e.what(); }
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < world; r++)
        if (!errs[r].empty()) throw std::runtime_error("rank " + std::to_string(r) + ": " + errs[r]);
    API_END
}

int acfb_dist_collect(acfb_engine* e, acfb_det* dets, int cap, int* counts, int* total)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    e->forCollect().distCollect(dets, cap, counts, total, e->distBatch);
    API_END
}

int acfb_dist_info(acfb_engine* e, int* rank, int* world, int* nccl_version)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    if (rank) *rank = e->e.distRank;
    if (world) *world = (e->e.comm || e->e.exch) ? e->e.distWorld : 0;
    if (nccl_version) { *nccl_version = 0; if (e->e.comm) NCCL_OK(NcclApi::get().GetVersion(nccl_version)); }
    API_END
}

int acfb_detect(acfb_engine* e, const uint8_t* frames, int n, int rows, int cols, int on_device, acfb_det* dets, int cap, int* counts, int* total)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    if (!e->order.empty()) throw std::runtime_error("collect the submitted batches first");
    e->e.submitAll(frames, n, rows, cols, on_device != 0);
    e->lastCollected = 0;
    e->e.collect(dets, cap, counts, total);
    API_END
}

int acfb_synchronize(acfb_engine* e)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    CUDA_OK(cudaSetDevice(e->e.device));
    e->e.syncAll();
    for (auto& x : e->extra) if (x) x->syncAll();
    API_END
}

int acfb_last_hits(acfb_engine* e, acfb_hit* hits, int cap, int* total, uint64_t* trees_evaluated, uint64_t* windows)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    Engine& E = e->pipe(e->lastCollected);
    if (total) *total = E.lastHitsValid ? (int)E.lastHits.size() : (int)E.lastHitTotal;
    if (!E.lastHitsValid && hits && cap > 0)
        throw std::runtime_error("acfb_last_hits: the last batch was ordered / suppressed on the device (k_post), which keeps the raw hit count but not "
                                 "the list; turn NMS off (or ACFB_DEVICE_POST=0) to read raw hits");
    for (int i = 0; i < (int)E.lastHits.size() && i < cap && hits && E.lastHitsValid; i++) hits[i] = E.lastHits[i];
    if (trees_evaluated) *trees_evaluated = E.hStats[0];
    if (windows) *windows = E.hStats[1];
    API_END
}

static void rawHitsOut(const std::vector<int4>& hh, int32_t* hit_c, int32_t* hit_r, float* hit_score, int cap, int* total)
{
    if (total) *total = (int)hh.size();
    for (int i = 0; i < (int)hh.size() && i < cap; i++)
    {
        if (hit_c) hit_c[i] = hh[i].y;
        if (hit_r) hit_r[i] = hh[i].z;
        if (hit_score) memcpy(&hit_score[i], &hh[i].w, 4);
    }
}

int acfb_acf_detect1(acfb_engine* e, const float* chns, int h, int w, int nchn, int32_t* hit_c, int32_t* hit_r, float* hit_score,
                     int cap, int* total, uint64_t* trees_evaluated)
{
    API_BEGIN
    if (!e || !chns) throw std::runtime_error("null argument");
    unsigned long long te = 0;
    const std::vector<int4> hh = e->e.cascadeOnChannels({ Engine::ChannelScale{ chns, h, w, nchn } }, false, &te);
    if (trees_evaluated) *trees_evaluated = te;
    rawHitsOut(hh, hit_c, hit_r, hit_score, cap, total);
    API_END
}

int acfb_acf_detect1_u8(acfb_engine* e, const uint8_t* chns, int h, int w, int nchn, int32_t* hit_c, int32_t* hit_r, float* hit_score,
                        int cap, int* total, uint64_t* trees_evaluated)
{
    API_BEGIN
    if (!e || !chns) throw std::runtime_error("null argument");
    unsigned long long te = 0;
    const std::vector<int4> hh = e->e.cascadeOnChannels({ Engine::ChannelScale{ chns, h, w, nchn } }, true, &te);
    if (trees_evaluated) *trees_evaluated = te;
    rawHitsOut(hh, hit_c, hit_r, hit_score, cap, total);
    API_END
}

int acfb_detect_channels(acfb_engine* e, const acfb_channels* scales, int nscales, int is_u8, acfb_det* out, int cap, int* total)
{
    API_BEGIN
    if (!e || (!scales && nscales > 0) || nscales < 0) throw std::runtime_error("bad argument");
    Engine& E = e->e;
    std::vector<Engine::ChannelScale> sc(nscales);
    for (int i = 0; i < nscales; i++)
    {
        sc[i] = Engine::ChannelScale{ scales[i].data, scales[i].h, scales[i].w, scales[i].nchn };
        if (!(scales[i].scale > 0) || !(scales[i].scalehw_w > 0) || !(scales[i].scalehw_h > 0)) throw std::runtime_error("engine: scale factors must be positive");
    }
    const std::vector<int4> hh = E.cascadeOnChannels(sc, is_u8 != 0, nullptr);
    std::vector<acfb_det> dets;
    E.lastHits.clear();
    for (const int4& q : hh)
    {
        float score;
        memcpy(&score, &q.w, 4);
        E.lastHits.push_back(acfb_hit{ 0, q.x, q.y, q.z, score });
        dets.push_back(E.boxOf(q.y, q.z, score, 0, scales[q.x].scale, scales[q.x].scalehw_w, scales[q.x].scalehw_h));
    }
    if (E.doNms && !dets.empty()) E.nmsAndPrune(dets);
    if (total) *total = (int)dets.size();
    for (int i = 0; i < (int)dets.size() && i < cap && out; i++) out[i] = dets[i];
    API_END
}

int acfb_evaluate(acfb_engine* e, const uint8_t* frame, int rows, int cols, float* score)
{
    API_BEGIN
    if (!e || !frame || !score) throw std::runtime_error("null argument");
    Engine& E = e->e;
    const acfb_options& o = E.opt;
    // Detector::evaluate builds its channels with computeChannels' FIXED defaults (ACF.cpp:165-240): LUV colour enabled,
    // smooth 1, normRad 5, normConst .005, full 0, 6 orientations, softBin 0, shrink 4 -- meaningful only for such models
    if (o.color_space != 2 || !o.color_enabled || o.color_smooth != 1.0 || o.gm_normRad != 5 || std::fabs(o.gm_normConst - 0.005) > 1e-12 ||
        o.gm_full != 0 || o.gh_nOrients != 6 || o.shrink != 4)
        throw std::runtime_error("acfb_evaluate: Detector::evaluate computes channels with computeChannels' default options (ACF.cpp:165-240); "
                                 "this model's channel options differ, so its feature ids would not address those channels");
    if (E.anyPending()) throw std::runtime_error("collect the submitted batches first");
    CUDA_OK(cudaSetDevice(E.device));
    SizeState& st = E.beginBatch(frame, 1, rows, cols, false);
    const Plan& P = st.plan;
    if (P.reals.empty() || P.reals[0].mode != RealScale::ALIAS) throw std::runtime_error("acfb_evaluate: frame size must be a multiple of shrink");
    Engine::Slot& S = E.slots[0];
    S.frames.ensure(E.frameBytes(rows, cols));
    CUDA_OK(cudaMemcpyAsync(S.frames.p, frame, E.frameBytes(rows, cols), cudaMemcpyHostToDevice, E.stream));
    E.colorAndReal0(st, S.frames.p);
    const RealScale& r = P.reals[0];
    const int mH = o.modelDsPad_w / o.shrink, mW = o.modelDsPad_h / o.shrink;
    if (mH > r.ch || mW > r.cw) throw std::runtime_error("acfb_evaluate: image smaller than the model window");
    E.scratch.ensure(1);
    launchEval1(st.R.p + st.realOff[0], r.cP, r.cw * r.cP, E.cascTab.p, E.model.nTrees(), E.model.clf.treeDepth, E.recWords, E.scratch.p, E.stream);
    E.launches++;
    CUDA_OK(cudaMemcpyAsync(score, E.scratch.p, sizeof(float), cudaMemcpyDeviceToHost, E.stream));
    CUDA_OK(cudaStreamSynchronize(E.stream));
    API_END
}

int acfb_compute_channels(acfb_engine* e, const uint8_t* frame, int rows, int cols, float* out, size_t cap_floats, int* d, int* w, int* h)
{
    API_BEGIN
    if (!e || !frame) throw std::runtime_error("null argument");
    Engine& E = e->e;
    const acfb_options& o = E.opt;
    // Detector::computeChannels builds the channels with FIXED defaults (ACF.cpp:183-240): LUV colour enabled, smooth 1,
    // normRad 5, normConst .005, full 0, 6 orientations, softBin 0, shrink 4.  The engine's buffers and kernels are set up
    // for the MODEL's options, so the call is served only for models whose channel options are those defaults.
    if (o.color_space != 2 || !o.color_enabled || o.color_smooth != 1.0 || o.gm_normRad != 5 || std::fabs(o.gm_normConst - 0.005) > 1e-12 ||
        o.gm_full != 0 || o.gh_nOrients != 6 || o.shrink != 4 || o.gm_colorChn != 0)
        throw std::runtime_error("acfb_compute_channels: Detector::computeChannels uses fixed default channel options (ACF.cpp:183-240); "
                                 "this model's channel options differ");
    if (E.anyPending()) throw std::runtime_error("collect the submitted batches first");
    CUDA_OK(cudaSetDevice(E.device));
    SizeState& st = E.beginBatch(frame, 1, rows, cols, false);
    const Plan& P = st.plan;
    if (P.reals.empty() || P.reals[0].mode != RealScale::ALIAS) throw std::runtime_error("acfb_compute_channels: frame size must be a multiple of shrink");
    const RealScale& r = P.reals[0];
    const size_t need = (size_t)P.nChns * r.cw * r.ch;
    if (d) *d = P.nChns; if (w) *w = r.cw; if (h) *h = r.ch;
    if (!out) return 0; // size query
    if (cap_floats < need) throw std::runtime_error("output buffer too small");
    Engine::Slot& S = E.slots[0];
    S.frames.ensure(E.frameBytes(rows, cols));
    CUDA_OK(cudaMemcpyAsync(S.frames.p, frame, E.frameBytes(rows, cols), cudaMemcpyHostToDevice, E.stream));
    E.colorAndReal0(st, S.frames.p);
    CUDA_OK(cudaMemcpy2DAsync(out, (size_t)r.ch * sizeof(float), st.R.p + st.realOff[0], (size_t)r.cP * sizeof(float), (size_t)r.ch * sizeof(float),
                              (size_t)P.nChns * r.cw, cudaMemcpyDeviceToHost, E.stream));
    CUDA_OK(cudaStreamSynchronize(E.stream));
    API_END
}

// ---- stand-alone channel operators: the reference's static Detector:: functions on one image (ACF.h:416-490).
// Host planes in, host planes out, transposed planar layout [z][x][y] (h = contiguous extent).

int acfb_op_rgb_convert(acfb_engine* e, const float* I, int h, int w, int colorspace, float* J, int* nplanes_out)
{
    API_BEGIN
    if (!e || !I || !J) throw std::runtime_error("null argument");
    if (h <= 0 || w <= 0 || colorspace < 0 || colorspace > 4) throw std::runtime_error("acfb_op_rgb_convert: bad size / colour space");
    Engine& E = e->e;
    CUDA_OK(cudaSetDevice(E.device));
    const size_t plane = (size_t)h * w;
    const int mode = colorspace == 0 ? 0 : colorspace == 2 ? 2 : colorspace == 3 ? 3 : 1; // rgbConvert.cpp:109-130
    const int np = mode == 0 ? 1 : 3;
    E.opA.ensure(3 * plane); E.opB.ensure(3 * plane);
    CUDA_OK(cudaMemcpyAsync(E.opA.p, I, 3 * plane * sizeof(float), cudaMemcpyHostToDevice, E.stream));
    ColorArgs ca{ reinterpret_cast<const uint8_t*>(E.opA.p), E.opB.p, E.lut.p, h, w, 1, mode, 12, 0, 1, 2, 2, 0 };
    launchColor(ca, E.stream); E.launches++;
    CUDA_OK(cudaMemcpyAsync(J, E.opB.p, np * plane * sizeof(float), cudaMemcpyDeviceToHost, E.stream));
    CUDA_OK(cudaStreamSynchronize(E.stream));
    if (nplanes_out) *nplanes_out = np;
    API_END
}

// device form shared by acfb_op_conv_tri and acfb_op_gradient_mag: out = convTri(in, r) (convTri.cpp:204-253), in == out allowed
static void opConvTri(Engine& E, float* in, float* out, int h, int w, int d, double r)
{
    const int m = std::min(h, w);
    if (m < 4 || 2 * r + 1 >= m) throw std::runtime_error("convTri: image too small for the radius (the reference leaves its toolbox path here, convTri.cpp:224-251)");
    if (r > 0 && r <= 1.0)
    {   // convTri1 (convConst.cpp:494-525); in place it is a recurrence along x, marched by k_smooth exactly
        if (h % 4 || h > 4096) throw std::runtime_error("convTri: r <= 1 needs h % 4 == 0 and h <= 4096");
        SmoothArgs sa{};
        float* dst = out;
        if (in == out) { E.opC.ensure((size_t)h * w * d); dst = E.opC.p; } // k_smooth reads columns ahead of the ones it writes
        sa.src = in; sa.dst = dst; sa.H = h; sa.W = w; sa.nPlanes = d; sa.plain = (in == out) ? 0 : 1;
        sa.p = (float)(12.0 / r / (r + 2.0) - 2.0); sa.nrm = 1.0f / ((sa.p + 2) * (sa.p + 2));
        launchSmooth(sa, E.stream); E.launches++;
        if (in == out) CUDA_OK(cudaMemcpyAsync(out, dst, (size_t)h * w * d * sizeof(float), cudaMemcpyDeviceToDevice, E.stream));
    }
    else
    {
        if (r != std::floor(r)) throw std::runtime_error("convTri: r > 1 must be an integer (convConst.cpp:178)");
        E.opC.ensure((size_t)h * w * d);
        launchTriAny(in, E.opC.p, out, h, w, d, (int)r, E.stream); E.launches += 2; // x pass reads in, y pass writes out: aliasing is harmless
    }
}

int acfb_op_conv_tri(acfb_engine* e, const float* I, int h, int w, int d, double r, float* J)
{
    API_BEGIN
    if (!e || !I || !J) throw std::runtime_error("null argument");
    if (h <= 0 || w <= 0 || d <= 0 || r < 0) throw std::runtime_error("acfb_op_conv_tri: bad arguments");
    Engine& E = e->e;
    const size_t n = (size_t)h * w * d;
    if (r == 0) { if (J != I) std::memcpy(J, I, n * sizeof(float)); return 0; } // convTri.cpp:206-210
    CUDA_OK(cudaSetDevice(E.device));
    E.opA.ensure(n); E.opB.ensure(n);
    CUDA_OK(cudaMemcpyAsync(E.opA.p, I, n * sizeof(float), cudaMemcpyHostToDevice, E.stream));
    // J aliasing I is the reference's in-place call (chnsCompute.cpp:239): the smoothing then feeds on its own output
    float* out = (J == I) ? E.opA.p : E.opB.p;
    opConvTri(E, E.opA.p, out, h, w, d, r);
    CUDA_OK(cudaMemcpyAsync(J, out, n * sizeof(float), cudaMemcpyDeviceToHost, E.stream));
    CUDA_OK(cudaStreamSynchronize(E.stream));
    API_END
}

int acfb_op_gradient_mag(acfb_engine* e, const float* I, int h, int w, int d, int channel, int normRad, double normConst, int full,
                         float* M, float* O)
{
    API_BEGIN
    if (!e || !I || !M) throw std::runtime_error("null argument");
    if (h < 4 || w < 2 || h % 4 || d <= 0 || channel < 0 || channel >= d || normRad < 0) throw std::runtime_error("acfb_op_gradient_mag: needs h % 4 == 0, 0 <= channel < d");
    Engine& E = e->e;
    CUDA_OK(cudaSetDevice(E.device));
    const size_t plane = (size_t)h * w;
    E.opA.ensure(plane); E.opB.ensure(plane); E.opO.ensure(plane);
    CUDA_OK(cudaMemcpyAsync(E.opA.p, I + (size_t)channel * plane, plane * sizeof(float), cudaMemcpyHostToDevice, E.stream)); // gradientMag.cpp:118, d = 1
    GradArgs ga{};
    ga.src = E.opA.p; ga.outM = E.opB.p; ga.outO = E.opO.p; ga.acosTab = E.acosTab.p; ga.srcFrameStride = plane; ga.moFrameStride = plane;
    ga.H = h; ga.W = w; ga.n = 1; ga.full = full; ga.colsPerThread = E.gradCols;
    launchGradMag(ga, E.stream); E.launches++;
    if (O)
    {
        launchOrientFloat(E.opO.p, E.opA.p, (int64_t)plane, E.acosTab.p, E.stream); E.launches++;
        CUDA_OK(cudaMemcpyAsync(O, E.opA.p, plane * sizeof(float), cudaMemcpyDeviceToHost, E.stream));
    }
    if (normRad != 0)
    {   // S = convTri(M, normRad); M = M / (S + normConst)  (gradientMag.cpp:125-131)
        E.opS.ensure(plane);
        opConvTri(E, E.opB.p, E.opS.p, h, w, 1, (double)normRad);
        launchMagNorm(E.opB.p, E.opS.p, (int64_t)plane, (float)normConst, E.stream); E.launches++;
        CUDA_OK(cudaMemcpyAsync(M, E.opB.p, plane * sizeof(float), cudaMemcpyDeviceToHost, E.stream));
        CUDA_OK(cudaStreamSynchronize(E.stream));
    }
    else
    {
        CUDA_OK(cudaMemcpyAsync(M, E.opB.p, plane * sizeof(float), cudaMemcpyDeviceToHost, E.stream));
        CUDA_OK(cudaStreamSynchronize(E.stream));
    }
    API_END
}

int acfb_op_gradient_hist(acfb_engine* e, const float* M, const float* O, int h, int w, int binSize, int nOrients, int softBin, int useHog,
                          double clipHog, int full, float* H)
{
    API_BEGIN
    (void)useHog; (void)clipHog; // ignored by the reference too (gradientHist.cpp:109-114)
    if (!e || !M || !O || !H) throw std::runtime_error("null argument");
    if (binSize != 4 || h % 4 || w % 4 || h < 4 || w < 4) throw std::runtime_error("acfb_op_gradient_hist: binSize 4 and sizes that are multiples of 4 only");
    if (softBin != 0 || nOrients < 1 || nOrients > 8) throw std::runtime_error("acfb_op_gradient_hist: softBin 0 and 1..8 orientations only");
    Engine& E = e->e;
    CUDA_OK(cudaSetDevice(E.device));
    const size_t plane = (size_t)h * w;
    const int ch = h / 4, cw = w / 4;
    E.opA.ensure(plane); E.opB.ensure(plane); E.opC.ensure((size_t)(nOrients + 1) * ch * cw);
    CUDA_OK(cudaMemcpyAsync(E.opA.p, M, plane * sizeof(float), cudaMemcpyHostToDevice, E.stream));
    CUDA_OK(cudaMemcpyAsync(E.opB.p, O, plane * sizeof(float), cudaMemcpyHostToDevice, E.stream));
    HistArgs ha{};
    ha.M = E.opA.p; ha.Of = E.opB.p; ha.acosTab = E.acosTab.p; ha.outR = E.opC.p; ha.moFrameStride = plane; ha.rFrameStride = 0;
    ha.H = h; ha.W = w; ha.n = 1; ha.cP = ch; ha.firstPlane = 0; ha.nOrients = nOrients; ha.doMag = 1;
    const float PI = 3.14159265f;
    ha.oMult = (float)nOrients / (full ? 2 * PI : PI);
    { const float sh = (float)binSize; ha.sInv2 = 1 / sh / sh; }
    { float q = 1.0f; q /= 4; q /= float(1 + 1e-6); ha.shrinkMul = q / 4; }
    launchHist(ha, E.stream); E.launches++;
    // plane 0 of the kernel's output is the shrunk magnitude (a channel of its own in chnsCompute); the histogram follows
    CUDA_OK(cudaMemcpyAsync(H, E.opC.p + (size_t)ch * cw, (size_t)nOrients * ch * cw * sizeof(float), cudaMemcpyDeviceToHost, E.stream));
    CUDA_OK(cudaStreamSynchronize(E.stream));
    API_END
}

int acfb_op_im_resample(acfb_engine* e, const float* A, int ha, int wa, int d, int hb, int wb, double nrm, float* B)
{
    API_BEGIN
    if (!e || !A || !B) throw std::runtime_error("null argument");
    if (ha <= 0 || wa <= 0 || hb <= 0 || wb <= 0 || d <= 0) throw std::runtime_error("acfb_op_im_resample: bad size");
    Engine& E = e->e;
    CUDA_OK(cudaSetDevice(E.device));
    const size_t na = (size_t)ha * wa * d, nb = (size_t)hb * wb * d;
    E.opA.ensure(na); E.opB.ensure(nb);
    CUDA_OK(cudaMemcpyAsync(E.opA.p, A, na * sizeof(float), cudaMemcpyHostToDevice, E.stream));
    const AxisCoef cx = makeAxisX(wa, wb), cy = makeAxisY(ha, hb); // resampleCoef (imResampleMex.cpp:25-121), throws beyond kMaxTaps
    AxisUpload ux, uy;
    ux.upload(cx, E.stream); uy.upload(cy, E.stream);
    ResampleArgs ra{};
    ra.src = E.opA.p; ra.dst = E.opB.p; ra.srcFrameStride = (int64_t)na; ra.dstFrameStride = (int64_t)nb;
    ra.ha = ha; ra.wa = wa; ra.hb = hb; ra.wb = wb; ra.d = d; ra.n = 1; ra.cx = ux.dev; ra.cy = uy.dev;
    float rr = (float)nrm; // imResampleMex.cpp:153-157: r /= 2|3|4 for the integer x fast paths, then r /= 1 + 1e-6
    rr /= cx.rdiv;
    rr /= float(1 + 1e-6);
    ra.r = rr;
    E.opC.ensure((size_t)d * wb * ha + 4);
    ra.tmp = E.opC.p; ra.tmpFrameStride = ((int64_t)d * wb * ha + 3) / 4 * 4;
    launchResample(ra, E.stream); E.launches += 2;
    CUDA_OK(cudaMemcpyAsync(B, E.opB.p, nb * sizeof(float), cudaMemcpyDeviceToHost, E.stream));
    CUDA_OK(cudaStreamSynchronize(E.stream)); // also keeps the tap tables alive until the kernel is done
    API_END
}

int acfb_selftest_math(acfb_engine* e, uint64_t n, uint32_t seed, uint64_t* mismatches)
{
    API_BEGIN
    if (!e || !mismatches) throw std::runtime_error("null argument");
    CUDA_OK(cudaSetDevice(e->e.device));
    *mismatches = selftestMath(n, seed, e->e.stream);
    e->e.launches++;
    API_END
}

uint64_t acfb_launch_count(acfb_engine* e)
{
    if (!e) return 0;
    uint64_t n = e->e.launches;
    for (auto& x : e->extra) if (x) n += x->launches;
    return n;
}
uint64_t acfb_stream(acfb_engine* e) { return e ? (uint64_t)(uintptr_t)e->e.stream : 0; }

int acfb_set_debug_taps(acfb_engine* e, int enable) { API_BEGIN if (!e) throw std::runtime_error("null engine"); e->e.keepC = enable != 0; API_END }

int acfb_enable_stage_timing(acfb_engine* e, int enable) { API_BEGIN if (!e) throw std::runtime_error("null engine"); e->e.timing = enable != 0; API_END }

int acfb_collect_times(acfb_engine* e, double* wait_ms, double* tail_ms)
{
    API_BEGIN
    if (!e) throw std::runtime_error("null engine");
    if (wait_ms) *wait_ms = e->pipe(e->lastCollected).collectWaitMs;
    if (tail_ms) *tail_ms = e->pipe(e->lastCollected).collectTailMs;
    API_END
}

int acfb_stage_times(acfb_engine* e, const char** names, float* ms, int cap)
{
    if (!e) return 0;
    const int n = (int)e->e.stageMs.size();
    for (int i = 0; i < n && i < cap; i++) { if (names) names[i] = e->e.stageNames[i]; if (ms) ms[i] = e->e.stageMs[i]; }
    return n;
}

int acfb_tap(acfb_engine* e, const char* tag, int frame, int real_k, float* out, size_t cap_floats, int* d, int* w, int* h)
{
    API_BEGIN
    if (!e || !e->e.cur || !tag) throw std::runtime_error("no pyramid resident");
    Engine& E = e->e;
    SizeState& st = *E.cur;
    const Plan& P = st.plan;
    if (frame < 0 || frame >= E.curN || real_k < 0 || real_k >= (int)P.reals.size()) throw std::runtime_error("bad frame / real scale index");
    CUDA_OK(cudaSetDevice(E.device));
    const RealScale& r = P.reals[real_k];
    const std::string t = tag;
    if (t == "C" && E.useFront && !E.keepC && E.opt.color_smooth > 0)
        throw std::runtime_error("acfb_tap: the smoothed image is only kept when debug taps are enabled (acfb_set_debug_taps) before the pyramid is computed");
    if (t == "I" || t == "C")
    {
        const float* src = nullptr;
        int hh = r.h, ww = r.w;
        src = (t == "I" ? E.imgIn(st, real_k) : E.imgSmooth(st, real_k)) + (size_t)frame * P.nImgPlanes * r.h * r.w;
        const size_t need = (size_t)P.nImgPlanes * hh * ww;
        if (cap_floats < need) throw std::runtime_error("output buffer too small");
        CUDA_OK(cudaMemcpy(out, src, need * sizeof(float), cudaMemcpyDeviceToHost));
        if (d) *d = P.nImgPlanes; if (w) *w = ww; if (h) *h = hh;
    }
    else if (t == "R")
    {
        const size_t need = (size_t)P.nChns * r.cw * r.ch;
        if (cap_floats < need) throw std::runtime_error("output buffer too small");
        const float* src = st.R.p + (size_t)frame * st.rFloatsPerFrame + st.realOff[real_k];
        CUDA_OK(cudaMemcpy2D(out, (size_t)r.ch * sizeof(float), src, (size_t)r.cP * sizeof(float), (size_t)r.ch * sizeof(float),
                             (size_t)P.nChns * r.cw, cudaMemcpyDeviceToHost));
        if (d) *d = P.nChns; if (w) *w = r.cw; if (h) *h = r.ch;
    }
    else throw std::runtime_error("unknown tap tag");
    API_END
}

} // extern "C"
