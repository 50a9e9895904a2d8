// plan.h -- host-side planning for one frame size: the scale schedule, the real/approximated
// split, buffer geometry and the resampling coefficient tables the kernels consume.
//
// Restates Detector::getScales (chnsPyramid.cpp:461-529), the isR/isA/isN logic of chnsPyramid
// (chnsPyramid.cpp:272-292), the real-scale source aliasing (chnsPyramid.cpp:297-316)
// and resampleCoef (toolbox/imResampleMex.cpp:25-121) -- all scalar fp32/fp64 host work that
// drives every buffer size and box coordinate, so it stays on the host exactly as in the reference.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>
#include "model.h"

namespace acfb
{

// One axis of a separable resample, flattened for the device.
//   out[b] = post( sum_{k<cnt[b]} in[start[b]+k] * wt[b*maxTaps+k] )      (ordered left to right)
// mode 0: weights are final                      (x axis; y axis general down-sampling)
// mode 1: integer-ratio y fast path: (sum of taps) * scaleMul, weights are 1
// mode 2: bilinear y: in[s]*(w*r) + in[s+1]*(r - w*r), boundary rows use the first tap only
struct AxisCoef
{
    int nOut = 0, nIn = 0, maxTaps = 0, mode = 0;
    float rdiv = 1.0f; // factor folded into r by the x axis (wa == 2/3/4 * wb  =>  r /= 2/3/4)
    int ymul = 1;      // y fast path: number of summed rows (2, 3 or 4)
    std::vector<int> start, cnt;
    std::vector<float> wt; // nOut * maxTaps
};
AxisCoef makeAxisX(int wa, int wb);
AxisCoef makeAxisY(int ha, int hb);

struct RealScale
{
    int scaleIdx = 0;   // index into scales
    int h = 0, w = 0;   // image-resolution size of this real scale (orig-y extent, orig-x extent)
    // where its input image comes from
    enum Src { FROM_I0 = 0, FROM_C = 1 } srcKind = FROM_I0;
    int srcReal = -1;   // when FROM_C: which real scale's smoothed image C_k
    enum Mode { ALIAS = 0, DOWN2 = 1, GENERIC = 2 } mode = ALIAS;
    bool writeC = false; // a later real scale resamples from this scale's smoothed image
    int srcH = 0, srcW = 0;
    AxisCoef cx, cy;     // GENERIC only
    float r = 1.0f;      // GENERIC / DOWN2: final multiplier (after the /2 and /(1+1e-6) adjustments)
    int ch = 0, cw = 0;  // channel-resolution size (h/shrink, w/shrink)
    int cP = 0;          // pitch (floats) of a channel-resolution column: ch rounded up to 4
};

struct ScaleGeom
{
    double scale = 0, shw_w = 0, shw_h = 0;
    int realK = 0;       // index into reals
    bool isReal = false;
    int h = 0, w = 0;    // unpadded channel dims
    int H = 0, W = 0;    // padded dims
    int P = 0;           // column pitch in floats (H rounded up to 4 so columns stay 16-byte aligned)
    int64_t offset = 0;  // float offset inside one frame's pyramid block
    AxisCoef cx, cy;     // approximated scales: resample from the real scale's channel planes
    bool identity = false;
    float ratio[3] = { 1, 1, 1 }; // per channel type: final r handed to the y weights
};

struct Plan
{
    int rows = 0, cols = 0;       // frame size
    int shrink = 4, nColor = 1, nImgPlanes = 1, nChns = 0, nTypes = 0;
    int typeFirst[3] = { 0, 0, 0 }, typeCount[3] = { 0, 0, 0 }; // channel ranges of colour / M / H
    int padX = 0, padY = 0;       // pad in channel pixels along orig x / orig y
    std::vector<double> scales;
    std::vector<std::pair<double, double>> scaleshw;
    std::vector<RealScale> reals;
    std::vector<ScaleGeom> geom;
    int64_t floatsPerFrame = 0;
    bool lambdasFromImage = false;
};

void getScales(int nPerOct, int nOctUp, int minDs_w, int minDs_h, int shrink, int sz_w, int sz_h,
               std::vector<double>& scales, std::vector<std::pair<double, double>>& scaleshw);

// lambdas: per-type exponents (from the model, or image-derived ones filled in later)
Plan makePlan(const acfb_options& o, int rows, int cols);
void setRatios(Plan& p, const acfb_options& o, const double* lambdas, int nLambdas);

} // namespace acfb
