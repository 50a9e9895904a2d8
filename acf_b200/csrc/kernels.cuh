// kernels.cuh -- launch-side declarations of the sm_100a kernels (kernels.cu).
//
// Data layout in HBM (all float planes): element (x, y) of a plane at [x*pitch + y] -- contiguous
// along the ORIGINAL image's y axis, exactly the reference's transposed cv::Mat, so the
// x axis is the slow axis of every recurrence and a warp marching along x emits whole y-contiguous
// columns.  Frames arrive as ordinary HWC u8 and are transposed once by k_color.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

struct CUtensorMap_st; // <cuda.h>; only cascade_tile.cu and the engine's encoder need the definition

namespace acfb
{

constexpr int kStripRows = 128;   // rows (orig y) one warp owns while marching along x: 32 lanes x 4 rows
constexpr int kMaxTapsDev = 12;   // must equal kMaxTaps in plan.cpp
constexpr int kCascTask = 256;    // windows per cascade task (fetched by one warp from a global counter)

struct AxisDev // device view of plan.h's AxisCoef
{
    const int* start;
    const int* cnt;
    const float* wt; // nOut * kMaxTapsDev
    int nOut, nIn, mode, ymul;
};

struct ColorArgs
{
    const uint8_t* frames; // [n] frames of rows*cols*bpp bytes each (layout below)
    float* out;            // [n][nPlanes][cols][rows]
    const float* lut;      // 1064-entry L table (luv only)
    int rows, cols, n;     // of the UPRIGHT image, whatever the memory layout
    int mode;              // 0 gray, 1 pass-through (rgb / orig / input already LUV), 2 luv, 3 hsv
    int bpp, ri, gi, bi;   // bytes per pixel and byte offsets of R, G, B inside a pixel (RGB24: 3,0,1,2; BGRA32: 4,2,1,0; GRAY8: 1,0,0,0)
    int srcKind;           // 0 u8 interleaved, 1 f32 interleaved RGB (bpp 12), 2 f32 planar [3][cols][rows] (bpp 12), 3 NV12 (rows * cols * 3 / 2 bytes per frame)
    int transposed;        // interleaved frame is stored [cols][rows][bpp] (Detector::setIsTranspose)
};
void launchColor(const ColorArgs& a, cudaStream_t s);

struct ResampleArgs
{
    const float* src; // [n][d][wa][ha]
    float* dst;       // [n][d][wb][hb]
    int64_t srcFrameStride, dstFrameStride;
    int ha, wa, hb, wb, d, n;
    AxisDev cx, cy;
    float r;
    float* tmp;       // optional [n][d][wb][ha] scratch: selects the two-pass form (x pass into tmp, y pass out of it)
    int64_t tmpFrameStride;
};
void launchResample(const ResampleArgs& a, cudaStream_t s);
void launchDown2(const ResampleArgs& a, cudaStream_t s); // wa == 2 wb, ha == 2 hb, hb % 4 == 0: the reference's /2 fast path

struct SmoothArgs
{
    const float* src;   // [n][nc][W][H]
    float* dst;         // [n][nc][W][H]  (must not alias src)
    int H, W, nPlanes;  // nPlanes = n * nc planes, one thread block each; H % 4 == 0, H <= 4096
    float p, nrm;       // [1 p 1] x [1 p 1]^T, nrm = 1 / (p + 2)^2
    float* dst2;        // optional [n][nc][W/2][H/2]: the smoothed plane resampled by exactly 1/2 (k_down2's arithmetic) on the fly
    float r2;           // k_down2's multiplier r / 2
    int pfAhead;        // columns past the register banks that are prefetched into L2 (0 = none)
    int plain;          // 0: the reference's in-place call (recurrence along x, the hot path); 1: distinct input / output, plain filter
};
void launchSmooth(const SmoothArgs& a, cudaStream_t s);

struct GradArgs
{
    const float* src;     // plane pGradMag.colorChn of the smoothed image, frame 0: [W][H]
    float* outM;          // raw gradient magnitude [n][W][H]
    uint16_t* outO;       // orientation [n][W][H] as the acos-table INDEX (0..20019) the reference looks up, bit 15 = "add pi"
                          // (full orientation, Gy < 0); the table value itself is fetched where the orientation is binned
    const float* acosTab; // 20020-entry table, pointer to element 0 (index range -10010..10009 via +10010)
    int64_t srcFrameStride, moFrameStride;
    int H, W, n, full;
    int colsPerThread;    // > 1: a thread walks eight columns (ACFB_GRAD_COLS=1 keeps one column per thread)
};
void launchGradMag(const GradArgs& a, cudaStream_t s);

struct TrixArgs
{
    const float* M;  // raw magnitude [n][W][H]
    float* U;        // x pass of the radius-5 triangle [n][W][H]
    int64_t frameStride;
    int H, W, n;
    int pfAhead;     // see SmoothArgs
};
void launchTrix(const TrixArgs& a, cudaStream_t s);

struct FrontArgs // k_front: smoothing + gradMag + x pass of the normalisation triangle in one march
{
    const float* src;   // [n][nc][W][H] image planes of the real scale
    float* dstC;        // optional smoothed planes [n][nc][W][H] (only when something still reads them)
    float* dst2;        // optional [n][nc][W/2][H/2]: the smoothed planes resampled by exactly 1/2 (next octave's input)
    float* outM;        // raw gradient magnitude [n][W][H] of plane gradPlane
    uint16_t* outO;     // its orientation as acos-table index (GradArgs::outO)
    float* outU;        // x pass of the radius-5 triangle of M [n][W][H]; nullptr for models without normalisation
    int64_t moFrameStride;
    int H, W, nc, nPlanes, gradPlane, full; // nPlanes = n * nc blocks; H % 4 == 0, 16 <= H <= 2304, W >= 16
    float p, nrm, r2;   // smoothing [1 p 1], nrm = 1 / (p + 2)^2; r2 = k_down2's multiplier
};
void launchFront(const FrontArgs& a, cudaStream_t s);

struct HistArgs
{
    const float* M;     // gradient magnitude [n][W][H] (raw; normalised on the fly when the caller is k_triyhist)
    const uint16_t* O;  // orientation as acos-table index (GradArgs::outO)
    const float* Of;    // or, when not null, as floats (stand-alone gradientHist operator)
    const float* acosTab;
    const float* C;     // smoothed image planes [n][firstPlane][W][H] (colour channels; unused when firstPlane == 0)
    int64_t cFrameStride;
    float* outR;        // real-scale channels: plane firstPlane = shrunk magnitude, then nOrients histogram planes
    int64_t moFrameStride, rFrameStride;
    int H, W, n, cP, firstPlane, nOrients;
    int doMag;          // 0: only the colour planes are shrunk (magnitude + histogram come from k_triyhist)
    float oMult, sInv2, shrinkMul;
};
void launchHist(const HistArgs& a, cudaStream_t s);

struct TriyArgs
{
    const float* U;  // x-filtered magnitude [n][W][H]
    HistArgs h;      // raw magnitude, orientation indices and the real-scale channel planes they are binned into
    int64_t frameStride;
    int H, W, n;
    float normConst;
    int blocksPerSm; // persistent blocks per SM (4 warps, 51 KB shared memory each)
    int fastScan;    // steady-state emissions use compile-time ring rows (ACFB_TRIY_FAST=0 keeps the generic scan everywhere)
    int frame0;      // k_triyhist_tma: frame coordinate of the launch's first frame inside the tensor maps (U, M then point at frame 0)
};
void launchTriyHist(const TriyArgs& a, cudaStream_t s);
// the same with U chunks and the emissions' magnitudes staged by 3-D tensor copies (maps: float, dims y | x | frame, box 32 x 32 x 1, 128-byte swizzle)
void launchTriyHistTma(const TriyArgs& a, const CUtensorMap_st& mapU, const CUtensorMap_st& mapM, cudaStream_t s);

struct ChanJob // one (scale, channel, strip) unit of the final-channel kernel
{
    int64_t srcOff;  // float offset of the source plane inside one frame's real-channel block
    int64_t dstOff;  // float offset of the destination plane (padded origin) inside one frame's pyramid block
    int srcH, srcW, srcP;
    int h, w, P;     // unpadded dst dims and pitch
    int padX, padY;
    int strip;
    int kind;        // -1 padding (multi-strip job lists), 0 generic (<= 3 taps per axis), 1 identity (real scale), 2 bilinear up-sample (2 taps, <= 128 source rows per strip)
    int axis;        // index into the AxisDev pair table (2*axis = x, 2*axis+1 = y)
    float r;
};
struct ChanArgs
{
    const float* src; // real-scale channels, all real scales in one block per frame
    float* dst;       // pyramid
    int64_t srcFrameStride, dstFrameStride;
    const ChanJob* jobs;
    const AxisDev* axes;
    int nJobs, n;
    int blockWarps;   // 0: planes of at most 128 rows, four independent warps per block; k > 0: every plane is a block of k
                      // warps (its 128-row strips; job list padded with kind = -1), smoothing exact across the strips
    float p, nrm;     // final smoothing
};
void launchChan(const ChanArgs& a, cudaStream_t s);

struct PadJob
{
    int64_t off;   // float offset of the first plane of this type inside one frame's pyramid block
    int h, w, P, W, H, padX, padY, d;
    int64_t cum;   // cumulative BORDER element count before this job
};
struct PadArgs
{
    float* pyr;
    int64_t frameStride;
    const PadJob* jobs;
    int nJobs, n;
    int64_t total;
};
void launchPad(const PadArgs& a, cudaStream_t s);

struct CascScale
{
    int64_t off;       // float offset of the scale inside one frame's pyramid block
    int P, planeStride; // column pitch, floats per plane
    int width1, height1; // window grid (c extent, r extent)
    int blk0;          // first task index of this scale inside its launch (tasks of kCascTask windows)
    int scaleIdx;      // index of the scale in the pyramid (reported with every hit)
};
struct CascArgs
{
    const void* pyr;    // float channels, or uint8_t channels when u8 != 0 (offsets / strides count elements)
    int u8;
    int64_t frameStride;
    const CascScale* scales;
    int nScales, nBlocksPerFrame, n;
    const uint32_t* tab; // per tree (recWords words, a multiple of 4): (2^D - 1) x {z, c, r, thr bits} then 2^D leaf values
    int nTrees, depth, recWords;
    int stride, shrink;
    float cascThr;
    int* hitCount;      // [n]
    int4* hits;         // [n][cap]  (scale, c, r, score bits)
    int cap;
    unsigned long long* stats; // [0] trees evaluated, [1] windows
    unsigned long long* taskCounter; // zeroed before every launch
    int blocksPerSm;    // 0 = as many as fit (2); 1 leaves half of every SM to the kernels of the other streams
    int prefetch;       // 0 none, 1 L2, 2 L1: fresh batches prefetch the next 32 rows of every line they gather
};
void launchCascade(const CascArgs& a, cudaStream_t s);

// ---- k_cascade_tile (cascade_tile.cu): depth-2 float cascade on TMA-staged shared-memory channel tiles
struct CascTileGeom
{
    int Wc, Wr;        // windows per tile along x (c) and y (r); Wr is a multiple of 32
    int BX, BY;        // TMA box = shared-memory tile: BX columns x BY rows per channel (BY % 4 == 0)
    int step, nChns;   // channel pixels between neighbouring windows (stride / shrink), channels
    int tileBytes, boxBytes, listCap, smemBytes;
};
bool cascTileGeometry(int mH, int mW, int nChns, int step, int blocksPerSm, CascTileGeom& g); // false: the window does not fit a tile
int cascTileRecWords(); // words per tree of the tile-local table: {off0, off1, off2, thr0} {thr1, thr2, leaf0, leaf1} {leaf2, leaf3, 0, 0}

constexpr int kCascHeadTrees = 64; // trees whose records travel in the kernel parameters (constant bank operands)
struct CascHeadRec { uint32_t off[3]; float thr[3]; float leaf[4]; }; // the first 10 words of a tile-table record

struct CascTileScale
{
    int tile0;           // first tile of the scale inside its launch (per frame)
    int nTx, nTy;        // tiles along c and r
    int width1, height1; // window grid
    int scaleIdx;        // index of the scale in the pyramid (reported with every hit)
};
struct CascTileArgs
{
    const CUtensorMap_st* maps; // [nScales] 4-D tensor maps (y | x | channel | frame) of the launch's scales, device memory
    const CascTileScale* scales;
    int nScales, tilesPerFrame, n;
    int frame0;          // frame coordinate of the launch's first frame inside the tensor maps
    const uint32_t* tab; // tile-local tree table, cascTileRecWords() words per tree (byte offsets inside a tile)
    int nTrees;
    int Wc, Wr, BY, step, tileBytes, boxBytes, listCap, smemBytes;
    int blocksPerSm;     // 1: one block per SM (half of every SM stays free for the kernels of the other streams), 3: three blocks of 384
                         // threads on smaller tiles (the geometry must have been chosen for it), else two of 512
    int sparseMax;       // levels past tree 64 with at most this many survivors run one window per warp (lanes = trees)
    int exportMax;       // ... and with at most this many, the survivors are handed to k_cascade_tail instead (0 = never)
    int4* tail;          // hand-over list: (frame | scale-in-launch << 24, c | r << 16, score bits, first tree still to run)
    int* tailCount;      // zeroed before the launch
    int tailCap;
    float cascThr;
    int* hitCount;       // [n]
    int4* hits;          // [n][cap]  (scale, c, r, score bits)
    int cap;
    unsigned long long* stats;       // [0] trees evaluated, [1] windows
    unsigned long long* taskCounter; // zeroed before every launch
    int headLevels;      // 5: levels [0,4) [4,8) [8,16) [16,32) [32,64) before the 64-tree chunks; 3: [0,4) [4,8) [8,64)
    int headTrees;       // kCascHeadTrees when the model has at least that many trees (levels [0,4) .. [32,64) then run fully unrolled
                         // with head[] as constant operands), else 0: every level reads its records from shared memory
    CascHeadRec head[kCascHeadTrees];
};
void launchCascadeTile(const CascTileArgs& a, cudaStream_t s);

struct CascTailArgs // k_cascade_tail: the windows a k_cascade_tile launch handed over, one window per warp, lanes = trees
{
    const float* pyr;   // frame 0 of the launch
    int64_t frameStride;
    const CascScale* scales; // the launch's scales (same indexing as CascTileArgs::scales)
    const uint32_t* tab;     // the global-gather table of k_cascade: depth 2, 16 words per tree
    int nTrees, stride, shrink;
    float cascThr;
    const int4* tail;
    const int* tailCount;
    int tailCap;
    int* hitCount;
    int4* hits;
    int cap;
    unsigned long long* stats;
};
void launchCascadeTail(const CascTailArgs& a, cudaStream_t s);

struct CascTailWinArgs // k_cascade_tail_win: the same windows, each on its own TMA-staged channel footprint
{
    const CUtensorMap_st* maps; // [nScales] 4-D maps of the launch's scales whose box is ONE window's footprint (mHp rows x mW columns x channels)
    const CascTileScale* scales;
    int frame0;
    const uint4* tabA;   // structure-of-arrays tree table with window-local byte offsets: A[t] = {off0, off1, off2, thr0},
    const uint4* tabB;   // B[t] = {thr1, thr2, leaf0, leaf1},
    const uint2* tabC;   // C[t] = {leaf2, leaf3}: a warp's 32 trees are three contiguous segments
    int nTrees;
    int step;            // channel pixels between neighbouring windows
    int footBytes;       // bytes of one footprint (= the box of `maps`)
    float cascThr;
    const int4* tail;
    const int* tailCount;
    int tailCap;
    int* hitCount;
    int4* hits;
    int cap;
    unsigned long long* stats;
};
void launchCascadeTailWin(const CascTailWinArgs& a, cudaStream_t s);
int cascTailWinSmem(int footBytes); // dynamic shared memory of k_cascade_tail_win (eight footprint slots)

// ---- k_post (post.cu): hit ordering + box rescale + bbNms (max / maxg) + prune on the device, one block per frame
struct PostScale { double scale, shw_w, shw_h; };
struct PostDet { int32_t x, y, w, h; float score; int32_t frame; }; // == acfb_det
struct PostArgs
{
    const int4* hits;      // [n][hitCap] raw hits of the cascade (scale, c, r, score bits)
    const int* hitCount;   // [n]
    int hitCap, n, frame0;
    int cap2;              // power of two: most hits per frame the shared-memory sort holds (frames beyond it raise `fallback`)
    const PostScale* scales;
    int stride, modelDs_w, modelDs_h, shift_w, shift_h;
    int greedy, ovrUnion, maxDet, maxOut; // maxOut = records per frame in `dets` (>= min(maxDet, 64))
    double overlap, pruneRatio;
    PostDet* dets;         // [n][maxOut]
    int* detCount;         // [n]; -1 = this frame was left to the host
    int* fallback;         // set to 1 when any frame was left to the host
};
size_t postSmemBytes(int cap2);
void launchPost(const PostArgs& a, cudaStream_t s);

struct SumArgs
{
    const float* src; int64_t frameStride; int64_t off; int P, h, w, d; double* out; // out[frame]
    int n;
};
void launchPlaneSum(const SumArgs& a, cudaStream_t s);

// stand-alone operators (any radius): O = convTri(I, r) through the scratch plane set tmp; M *= 1/(S+norm); index -> float orientation
void launchTriAny(const float* I, float* tmp, float* O, int h, int w, int d, int r, cudaStream_t s);
void launchMagNorm(float* M, const float* S, int64_t n, float norm, cudaStream_t s);
void launchOrientFloat(const uint16_t* Oi, float* O, int64_t n, const float* acosTab, cudaStream_t s);

void launchEval1(const float* chns, int P, int planeStride, const uint32_t* tab, int nTrees, int depth, int recWords, float* out, cudaStream_t s);

// number of inputs (out of 2n checks) where the in-kernel reciprocal / square root differ from the IEEE operators
unsigned long long selftestMath(unsigned long long n, unsigned seed, cudaStream_t s);

} // namespace acfb
