// acf_oracle.cpp -- restated orchestration (L2) of the reference's chnsPyramid + acfDetect path.
//
// TEST INFRASTRUCTURE ONLY (see acf_oracle.h).  Follows, line by line, including the aliasing
// and border quirks catalogued in SURVEY.md A.2 (reference file:line under src/lib/acf/acf/):
//   get_scales        chnsPyramid.cpp:461-529
//   chns_compute      chnsCompute.cpp:146-370 (+ gradientMag.cpp:109-135, gradientHist.cpp:92-114, convTri.cpp:204-253)
//   Pyr::build        ACF.cpp:116-141 (u8 -> f32, transpose, planar), chnsPyramid.cpp:160-456, MatP.cpp:122-129, ACF.h:653-672
//   detect1           toolbox/acfDetect1.cpp:72-144, 231-335, 390-406
//   oracle_detect     ACF.cpp:268-367
//   nms / prune       bbNms.cpp:111-192, 229-304 ; ObjectDetector.cpp:28-44
// The arithmetic itself is reached through the OracleL1 table (port or reference objects).
#include "acf_oracle.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <cstdlib>
#include <string>
#include <vector>

namespace
{

// A planar image in the reference's transposed layout: d planes, each w (orig-x) by h (orig-y),
// element (z,x,y) at z*w*h + x*h + y.  shared_ptr gives MatP's shallow-copy aliasing.
struct Planes
{
    int h = 0, w = 0, d = 0;
    std::shared_ptr<float> buf;
    float* p() const { return buf.get(); }
    bool empty() const { return !buf || d == 0; }
    // cv::Mat::create does not clear its buffer; only callers that need zeros (gradientHist.cpp:95) ask for them
    static Planes make(int h, int w, int d, bool zero = false)
    {
        Planes r; r.h = h; r.w = w; r.d = d;
        const size_t n = (size_t)h * w * d;
        float* q = static_cast<float*>(zero ? calloc(n ? n : 1, sizeof(float)) : malloc((n ? n : 1) * sizeof(float)));
        if (!q) throw std::bad_alloc();
        r.buf = std::shared_ptr<float>(q, free);
        return r;
    }
};

double log2_ref(double x) { return std::log(x) / std::log(2.0); } // util/acf_math.h:20-29

void get_scales(int nPerOct, int nOctUp, int minDs_w, int minDs_h, int shrink, int sz_w, int sz_h,
                std::vector<double>& scales, std::vector<std::pair<double, double>>& scaleshw)
{
    scales.clear(); scaleshw.clear();
    if (!(sz_w * sz_h)) return;
    const double rw = double(sz_w) / double(minDs_w), rh = double(sz_h) / double(minDs_h);
    const int nScales = (int)std::floor(double(nPerOct) * (double(nOctUp) + log2_ref(std::min(rw, rh))) + 1.0);
    double d0 = sz_h, d1 = sz_w;
    if (sz_h >= sz_w) std::swap(d0, d1);
    std::vector<double> tmp;
    for (int i = 0; i < nScales; i++)
    {
        const double s = std::pow(2.0, -double(i) / double(nPerOct) + double(nOctUp));
        const double s0 = (std::round(d0 * s / shrink) * shrink - 0.25 * shrink) / d0;
        const double s1 = (std::round(d0 * s / shrink) * shrink + 0.25 * shrink) / d0;
        double bestS = 0, bestE = std::numeric_limits<double>::max();
        for (double j = 0.0; j < 1.0 - std::numeric_limits<double>::epsilon(); j += 0.01)
        {
            const double ss = (j * (s1 - s0) + s0);
            double es0 = d0 * ss; es0 = std::abs(es0 - std::round(es0 / shrink) * shrink);
            double es1 = d1 * ss; es1 = std::abs(es1 - std::round(es1 / shrink) * shrink);
            const double es = std::max(es0, es1);
            if (es < bestE) { bestS = ss; bestE = es; }
        }
        tmp.push_back(bestS);
    }
    tmp.push_back(0);
    for (size_t i = 1; i < tmp.size(); i++)
    {
        if (tmp[i] != tmp[i - 1])
        {
            const double s = tmp[i - 1];
            scales.push_back(s);
            const double x = std::round(double(sz_w) * s / shrink) * shrink / sz_w;
            const double y = std::round(double(sz_h) * s / shrink) * shrink / sz_h;
            scaleshw.emplace_back(x, y);
        }
    }
}

// Detector::convTri (convTri.cpp:204-253) restricted to the branches the pyramid reaches.
void conv_tri_inplace(const Planes& I, double r)
{
    const OracleL1& L = oracle_l1();
    if (I.empty() || r == 0) return;
    const int m = std::min(I.h, I.w);
    if (m < 4 || (2 * r + 1) >= m) throw std::runtime_error("oracle: plane too small for convConst (sepFilter2D fallback not restated)");
    if (!(r > 0 && r <= 1.0)) throw std::runtime_error("oracle: in-place convTri with r>1 not restated");
    const float p = (float)(12.0 / r / (r + 2.0) - 2.0);
    L.convTri1(I.p(), I.p(), I.h, I.w, I.d, p, 1); // in == out: SURVEY A.2 Q1
}

// imResample(MatP) imResampleMex.cpp:385-420 ; size given as (h, w) in original-image orientation
Planes im_resample(const Planes& A, int hb, int wb, double nrm)
{
    Planes B = Planes::make(hb, wb, A.d);
    oracle_l1().resample(A.p(), B.p(), A.h, hb, A.w, wb, A.d, float(nrm));
    return B;
}

struct Chns { std::vector<Planes> data; }; // types: [color] [M] [H]

void chns_compute(const Planes& I, const oracle_opts& o, Chns& out, oracle_tap_fn tap, void* user, int scaleIdx)
{
    const OracleL1& L = oracle_l1();
    const int shrink = o.shrink;
    if (I.h % shrink || I.w % shrink) throw std::runtime_error("oracle: chnsCompute crop branch not reachable from the pyramid (A.2 Q11)");
    const int h = I.h / shrink, w = I.w / shrink;
    // colour: conversion already done (colorSpace forced to "orig", chnsPyramid.cpp:263); smooth in place
    conv_tri_inplace(I, o.color_smooth);
    if (tap) tap("C", scaleIdx, I.p(), I.h, I.w, I.d, user);
    if (o.color_enabled) out.data.push_back((I.h != h || I.w != w) ? im_resample(I, h, w, 1.0) : I);
    Planes M, O;
    if (o.gh_enabled || o.gm_enabled)
    {
        M = Planes::make(I.h, I.w, 1);
        O = Planes::make(I.h, I.w, 1);
        float* src = I.p() + (size_t)o.gm_colorChn * I.h * I.w;
        L.gradMag(src, M.p(), O.p(), I.h, I.w, 1, o.gm_full != 0);
        if (tap) tap("M", scaleIdx, M.p(), I.h, I.w, 1, user);
        if (o.gm_normRad != 0)
        {
            Planes S = Planes::make(I.h, I.w, 1);
            const double r = o.gm_normRad;
            const int m = std::min(I.h, I.w);
            if (m < 4 || (2 * r + 1) >= m) throw std::runtime_error("oracle: plane too small for convTri(normRad)");
            if (r > 0 && r <= 1.0) L.convTri1(M.p(), S.p(), I.h, I.w, 1, (float)(12.0 / r / (r + 2.0) - 2.0), 1);
            else L.convTri(M.p(), S.p(), I.h, I.w, 1, (int)std::round((float)r), 1);
            if (tap) tap("S", scaleIdx, S.p(), I.h, I.w, 1, user);
            L.gradMagNorm(M.p(), S.p(), I.h, I.w, (float)o.gm_normConst);
        }
        if (tap) { tap("Mnorm", scaleIdx, M.p(), I.h, I.w, 1, user); tap("O", scaleIdx, O.p(), I.h, I.w, 1, user); }
    }
    if (o.gm_enabled) out.data.push_back((I.h != h || I.w != w) ? im_resample(M, h, w, 1.0) : M);
    if (o.gh_enabled)
    {
        const int bin = o.gh_binSize > 0 ? o.gh_binSize : shrink;
        Planes H = Planes::make(I.h / bin, I.w / bin, o.gh_nOrients, true); // zero-filled (gradientHist.cpp:95)
        L.gradHist(M.p(), O.p(), H.p(), I.h, I.w, bin, o.gh_nOrients, o.gh_softBin, o.gm_full != 0);
        if (tap) tap("H", scaleIdx, H.p(), H.h, H.w, H.d, user);
        if (H.h != h || H.w != w) H = im_resample(H, h, w, 1.0);
        out.data.push_back(H);
    }
}

// cv::copyMakeBorder(BORDER_REFLECT) on every plane of one multi-plane type, including the
// submatrix rule: rows missing above/below plane k are taken from the parent
// buffer (= neighbouring planes) when it has them; only the rest is reflected.
// px = pad along orig x (Mat rows; pad.height/shrink), py = pad along orig y (Mat cols; pad.width/shrink).
Planes pad_reflect(const Planes& A, int px, int py)
{
    const int h = A.h, w = A.w, d = A.d, H = h + 2 * py, W = w + 2 * px;
    Planes B = Planes::make(H, W, d);
    const float* a = A.p();
    const int totalRows = w * d; // rows of the tall parent Mat
    const bool sub = d > 1;      // full-size ROI of a one-plane Mat is not a submatrix
    for (int k = 0; k < d; k++)
    {
        // grown source rows [r0, r1) in parent coordinates
        int r0 = k * w, r1 = (k + 1) * w;
        int top = px, bottom = px;
        if (sub)
        {
            const int dtop = std::min(r0, top), dbottom = std::min(totalRows - r1, bottom);
            r0 -= dtop; r1 += dbottom; top -= dtop; bottom -= dbottom;
        }
        const int srows = r1 - r0;
        float* b = B.p() + (size_t)k * W * H;
        for (int X = 0; X < W; X++)
        {
            int sr = X - top; // row in grown source
            if (sr < 0) sr = -sr - 1;                 // fedcba|abc
            else if (sr >= srows) sr = 2 * srows - sr - 1;
            const float* srow = a + (size_t)(r0 + sr) * h;
            float* drow = b + (size_t)X * H;
            for (int Y = 0; Y < H; Y++)
            {
                int sc = Y - py;
                if (sc < 0) sc = -sc - 1;
                else if (sc >= h) sc = 2 * h - sc - 1;
                drow[Y] = srow[sc];
            }
        }
    }
    return B;
}

struct Pyr
{
    int nScales = 0, nTypes = 0;
    std::vector<double> scales, lambdas;
    std::vector<std::pair<double, double>> scaleshw;
    std::vector<Planes> fused; // per scale, all channels stacked

    void build(const oracle_opts& o, const void* img, int rows, int cols, bool isF32, oracle_tap_fn tap, void* user)
    {
        const int h = rows, w = cols;
        Planes I = convert_input(o, img, rows, cols, isF32);
        if (tap) tap("I", -1, I.p(), I.h, I.w, I.d, user);
        build_from(o, I, h, w, tap, user);
    }

    // ACF.cpp:135-141 : transpose, u8 -> f32 via convertTo(1/255.) (one float multiply), planar split; then the colour
    // conversion of chnsPyramid.cpp:231-261 / rgbConvert.cpp:102-170
    static Planes convert_input(const oracle_opts& o, const void* img, int rows, int cols, bool isF32)
    {
        const OracleL1& L = oracle_l1();
        const int h = rows, w = cols;
        Planes rgb = Planes::make(h, w, 3);
        const float k255 = (float)(1.0 / 255.0);
        // cv::Mat::t() + convertTo + extractChannel, walked in 32 x 32 blocks the way an optimised transpose does (values are
        // identical whatever the order; the timed CPU baseline should not pay for a cache-hostile loop the reference does not have)
        for (int y0 = 0; y0 < h; y0 += 32)
            for (int x0 = 0; x0 < w; x0 += 32)
                for (int x = x0; x < std::min(x0 + 32, w); x++)
                    for (int c = 0; c < 3; c++)
                    {
                        float* dst = rgb.p() + (size_t)c * w * h + (size_t)x * h;
                        for (int y = y0; y < std::min(y0 + 32, h); y++)
                        {
                            const size_t si = ((size_t)y * w + x) * 3 + c;
                            dst[y] = isF32 ? ((const float*)img)[si] : (float)((const uint8_t*)img)[si] * k255;
                        }
                    }
        // colour conversion (chnsPyramid.cpp:231-261, rgbConvert.cpp:102-170)
        Planes I;
        if (o.color_space == 0) { I = Planes::make(h, w, 1); L.rgbConvert(rgb.p(), I.p(), h * w, 3, 0, 1.0f); }
        else if (o.color_space == 2) { I = Planes::make(h, w, 3); L.rgbConvert(rgb.p(), I.p(), h * w, 3, 2, 1.0f); }
        else if (o.color_space == 3) { I = Planes::make(h, w, 3); L.rgbConvert(rgb.p(), I.p(), h * w, 3, 3, 1.0f); }
        else if (o.color_space == 1 || o.color_space == 4) I = rgb; // pass-through (aliases the caller's planes, A.2 Q13)
        else throw std::runtime_error("oracle: colour space not restated");
        return I;
    }

    void build_from(const oracle_opts& o, Planes I, int h, int w, oracle_tap_fn tap, void* user)
    {
        const int shrink = o.shrink;
        get_scales(o.nPerOct, o.nOctUp, o.minDs_w, o.minDs_h, shrink, /*sz.width=*/h, /*sz.height=*/w, scales, scaleshw);
        nScales = (int)scales.size();
        std::vector<int> isR, isA, isN(nScales, 0);
        for (int i = 0; i < nScales; i++) ((i % (o.nApprox + 1)) > 0 ? isA : isR).push_back(i + 1);
        std::vector<int> isH(isR.size() + 1, 0);
        isH.back() = nScales;
        for (int i = 0; i < std::max(int(isR.size()) - 1, 0); i++) isH[i + 1] = (isR[i] + isR[i + 1]) / 2;
        for (size_t i = 0; i < isR.size(); i++)
            for (int j = isH[i]; j < isH[i + 1]; j++) isN[j] = isR[i];

        std::vector<std::vector<Planes>> data(nScales);
        for (int i : isR)
        {
            const double s = scales[i - 1];
            const int h1 = (int)(std::round(double(h) * s / double(shrink))) * shrink;
            const int w1 = (int)(std::round(double(w) * s / double(shrink))) * shrink;
            Planes I1;
            if (h1 == I.h && w1 == I.w) I1 = I; // shallow (A.2 Q2)
            else I1 = im_resample(I, h1, w1, 1.0);
            if (s == 0.5 && (o.nApprox > 0 || o.nPerOct == 1)) I = I1; // shallow
            Chns ch;
            chns_compute(I1, o, ch, tap, user, i - 1);
            nTypes = (int)ch.data.size();
            data[i - 1] = ch.data;
        }
        // image-specific lambdas (chnsPyramid.cpp:341-374)
        lambdas.assign(o.lambdas, o.lambdas + o.nLambdas);
        if (nScales > 0 && o.nApprox > 0 && lambdas.empty())
        {
            std::vector<int> is;
            for (int i = 1 + o.nOctUp * o.nPerOct; i <= nScales; i += o.nApprox + 1) is.push_back(i - 1);
            if (is.size() < 2) throw std::runtime_error("oracle: need >= 2 real scales to derive lambdas");
            if (is.size() > 2) is = { is[1], is[2] };
            auto mean = [](const Planes& P) {
                double t = 0; // cv::sum accumulates in double
                const size_t n = (size_t)P.h * P.w * P.d;
                for (size_t k = 0; k < n; k++) t += P.p()[k];
                return t / double(n);
            };
            lambdas.resize(nTypes);
            for (int j = 0; j < nTypes; j++)
            {
                const double f0 = mean(data[is[0]][j]), f1 = mean(data[is[1]][j]);
                lambdas[j] = -log2_ref(f0 / f1) / log2_ref(scales[is[0]] / scales[is[1]]);
            }
        }
        if ((int)lambdas.size() < nTypes && !isA.empty()) throw std::runtime_error("oracle: fewer lambdas than channel types");
        // approximated scales (chnsPyramid.cpp:385-397)
        for (int i : isA)
        {
            const int iR = isN[i - 1];
            const int h1 = (int)std::round(double(h) * scales[i - 1] / double(shrink));
            const int w1 = (int)std::round(double(w) * scales[i - 1] / double(shrink));
            data[i - 1].resize(nTypes);
            for (int j = 0; j < nTypes; j++)
            {
                const double ratio = std::pow(scales[i - 1] / scales[iR - 1], -lambdas[j]);
                data[i - 1][j] = im_resample(data[iR - 1][j], h1, w1, ratio);
            }
        }
        // smooth every scale / type in place (chnsPyramid.cpp:399-407)
        for (int i = 0; i < nScales; i++)
            for (int j = 0; j < nTypes; j++) conv_tri_inplace(data[i][j], o.smooth);
        if (tap)
            for (int i = 0; i < nScales; i++)
                for (int j = 0; j < nTypes; j++) tap(j == 0 ? "T0" : j == 1 ? "T1" : "T2", i, data[i][j].p(), data[i][j].h, data[i][j].w, data[i][j].d, user);
        // pad (chnsPyramid.cpp:410-424): y = pad.height/shrink rows (orig x), x = pad.width/shrink cols (orig y)
        if (o.pad_w || o.pad_h)
            for (int i = 0; i < nScales; i++)
                for (int j = 0; j < nTypes; j++) data[i][j] = pad_reflect(data[i][j], o.pad_h / shrink, o.pad_w / shrink);
        // concat (ACF.h:653-672)
        fused.resize(nScales);
        for (int i = 0; i < nScales; i++)
        {
            int d = 0;
            for (auto& t : data[i]) d += t.d;
            Planes F = Planes::make(data[i][0].h, data[i][0].w, d);
            size_t off = 0;
            for (auto& t : data[i])
            {
                const size_t n = (size_t)t.h * t.w * t.d;
                memcpy(F.p() + off, t.p(), n * sizeof(float));
                off += n;
            }
            fused[i] = F;
        }
    }
};

// acfDetect1.cpp:231-335 (column-major path, m_isRowMajor == false).  T = float with thrs, or uint8_t with thrsU8
// (ParallelDetectionBody<uint8_t,k>, acfDetect1.cpp:157-166,187-191): the feature is widened to float and compared
// with the (widened) threshold, exactly the `float ftr = chns1[...]; ftr < thrs[k]` of :72-82
template <typename T, typename TT>
int detect1(const T* chns, const TT* thrs, int height, int width, int nChns, const oracle_opts& o, const oracle_clf& clf,
            std::vector<int>& hc, std::vector<int>& hr, std::vector<float>& hs_out, uint64_t* treesEval)
{
    const int shrink = o.shrink, stride = o.stride;
    const int modelHt = o.modelDsPad_w, modelWd = o.modelDsPad_h; // swapped (acfDetect1.cpp:252-256)
    const int rowStride = height;
    const int height1 = (int)ceil(float(height * shrink - modelHt + 1) / stride);
    const int width1 = (int)ceil(float(width * shrink - modelWd + 1) / stride);
    const int mW = modelWd / shrink, mH = modelHt / shrink;
    std::vector<uint32_t> cids((size_t)nChns * mW * mH);
    {
        int m = 0;
        const int area = width * height;
        for (int z = 0; z < nChns; z++)
            for (int c = 0; c < mW; c++)
                for (int r = 0; r < mH; r++) cids[m++] = z * area + c * height + r;
    }
    const float cascThr = (float)o.cascThr;
    const int nTrees = clf.nTrees, nTreeNodes = clf.nTreeNodes, depth = clf.treeDepth;
    uint64_t nEval = 0;
    for (int c = 0; c < width1; c++)
        for (int r = 0; r < height1; r++)
        {
            const int offset = (r * stride / shrink) + (c * stride / shrink) * rowStride;
            const T* chns1 = chns + offset;
            float h = 0.f;
            for (int t = 0; t < nTrees; t++)
            {
                uint32_t off = t * nTreeNodes, k = off, k0 = 0;
                if (depth == 0)
                {
                    k0 = k; // k0 = k * isZero
                    while (clf.child[k])
                    {
                        const float ftr = chns1[cids[clf.fids[k]]];
                        k = (ftr < thrs[k]) ? 1 : 0;
                        k0 = k = clf.child[k0] - k + off;
                    }
                }
                else
                {
                    for (int i = 0; i < depth; i++)
                    {
                        const float ftr = chns1[cids[clf.fids[k]]];
                        k = (ftr < thrs[k]) ? 1 : 2;
                        k0 = k += k0 * 2;
                        k += off;
                    }
                }
                h += clf.hs[k];
                nEval++;
                if (h <= cascThr) break;
            }
            if (h > cascThr) { hc.push_back(c); hr.push_back(r); hs_out.push_back(h); }
        }
    if (treesEval) *treesEval += nEval;
    return (int)hc.size();
}

int cv_round(double v) { return (int)std::nearbyint(v); } // cvRound: round-half-to-even (default FE mode)

} // namespace

extern "C" {

const char* oracle_kind(void) { return oracle_l1().kind; }

int oracle_get_scales(int nPerOct, int nOctUp, int minDs_w, int minDs_h, int shrink, int sz_w, int sz_h,
                      double* scales, double* scaleshw, int cap)
{
    std::vector<double> s; std::vector<std::pair<double, double>> hw;
    get_scales(nPerOct, nOctUp, minDs_w, minDs_h, shrink, sz_w, sz_h, s, hw);
    for (int i = 0; i < (int)s.size() && i < cap; i++) { scales[i] = s[i]; scaleshw[2 * i] = hw[i].first; scaleshw[2 * i + 1] = hw[i].second; }
    return (int)s.size();
}

static thread_local std::string g_err;
const char* oracle_last_error(void) { return g_err.c_str(); }

void* oracle_pyramid_create(const oracle_opts* o, const void* img, int rows, int cols, int is_f32, oracle_tap_fn tap, void* user)
{
    try
    {
        auto* P = new Pyr();
        try { P->build(*o, img, rows, cols, is_f32 != 0, tap, user); }
        catch (...) { delete P; throw; }
        return P;
    }
    catch (const std::exception& e) { g_err = e.what(); }
    catch (const char* e) { g_err = e; }
    return nullptr;
}
void oracle_pyramid_destroy(void* pyr) { delete (Pyr*)pyr; }
int oracle_pyramid_nscales(void* pyr) { return ((Pyr*)pyr)->nScales; }
int oracle_pyramid_ntypes(void* pyr) { return ((Pyr*)pyr)->nTypes; }
const float* oracle_pyramid_scale(void* pyr, int i, int* h, int* w, int* nchn, double* scale, double* shw_w, double* shw_h)
{
    Pyr* P = (Pyr*)pyr;
    const Planes& F = P->fused[i];
    if (h) *h = F.h; if (w) *w = F.w; if (nchn) *nchn = F.d;
    if (scale) *scale = P->scales[i];
    if (shw_w) *shw_w = P->scaleshw[i].first;
    if (shw_h) *shw_h = P->scaleshw[i].second;
    return F.p();
}
int oracle_pyramid_lambdas(void* pyr, double* out, int cap)
{
    Pyr* P = (Pyr*)pyr;
    for (int i = 0; i < (int)P->lambdas.size() && i < cap; i++) out[i] = P->lambdas[i];
    return (int)P->lambdas.size();
}

int oracle_acf_detect1(const float* chns, int h, int w, int nchn, const oracle_opts* o, const oracle_clf* clf,
                       int* hit_c, int* hit_r, float* hit_score, int cap, uint64_t* trees_evaluated)
{
    std::vector<int> hc, hr; std::vector<float> hs;
    detect1(chns, clf->thrs, h, w, nchn, *o, *clf, hc, hr, hs, trees_evaluated);
    for (int i = 0; i < (int)hc.size() && i < cap; i++) { hit_c[i] = hc[i]; hit_r[i] = hr[i]; hit_score[i] = hs[i]; }
    return (int)hc.size();
}

// Classifier::thrsU8 = thrs.convertTo(CV_8UC1, 255.0f) (ACFIOArchive.h:96-99): OpenCV scales CV_32F sources in float and
// converts with saturate_cast<uchar>(cvRound(v)) (round half to even, clamp to [0,255])
int oracle_acf_detect1_u8(const uint8_t* chns, int h, int w, int nchn, const oracle_opts* o, const oracle_clf* clf,
                          int* hit_c, int* hit_r, float* hit_score, int cap, uint64_t* trees_evaluated)
{
    const size_t n = (size_t)clf->nTrees * clf->nTreeNodes;
    std::vector<uint8_t> thrsU8(n);
    for (size_t i = 0; i < n; i++)
    {
        const float v = clf->thrs[i] * 255.0f;
        const int q = cv_round(v);
        thrsU8[i] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
    }
    std::vector<int> hc, hr; std::vector<float> hs;
    detect1(chns, thrsU8.data(), h, w, nchn, *o, *clf, hc, hr, hs, trees_evaluated);
    for (int i = 0; i < (int)hc.size() && i < cap; i++) { hit_c[i] = hc[i]; hit_r[i] = hr[i]; hit_score[i] = hs[i]; }
    return (int)hc.size();
}

int oracle_detect(void* pyr, const oracle_opts* o, const oracle_clf* clf, oracle_det* out, int cap, int* total,
                  int* hit_scale, int* hit_c, int* hit_r, uint64_t* trees_evaluated)
{
    Pyr* P = (Pyr*)pyr;
    // shift = (modelDsPad - modelDs)/2 - pad  (integer cv::Size arithmetic, ACF.cpp:275)
    const int shift_w = (o->modelDsPad_w - o->modelDs_w) / 2 - o->pad_w;
    const int shift_h = (o->modelDsPad_h - o->modelDs_h) / 2 - o->pad_h;
    int n = 0;
    if (trees_evaluated) *trees_evaluated = 0;
    for (int i = 0; i < P->nScales; i++)
    {
        const Planes& F = P->fused[i];
        std::vector<int> hc, hr; std::vector<float> hs;
        detect1(F.p(), clf->thrs, F.h, F.w, F.d, *o, *clf, hc, hr, hs, trees_evaluated);
        for (size_t k = 0; k < hc.size(); k++)
        {
            // acfDetect1.cpp:326-334: Rect(x=c*stride, y=r*stride, winSize=(modelWd, modelHt)) then swap
            int rx = hr[k] * o->stride, ry = hc[k] * o->stride;
            // ACF.cpp:302-311
            const int sw = cv_round(double(o->modelDs_w) / P->scales[i]);
            const int sh = cv_round(double(o->modelDs_h) / P->scales[i]);
            rx = (int)(double(rx + shift_w) / P->scaleshw[i].first);
            ry = (int)(double(ry + shift_h) / P->scaleshw[i].second);
            if (n < cap)
            {
                out[n].x = ry; out[n].y = rx; out[n].w = sh; out[n].h = sw; // final swap x<->y, w<->h
                out[n].score = hs[k];
                if (hit_scale) hit_scale[n] = i;
                if (hit_c) hit_c[n] = hc[k];
                if (hit_r) hit_r[n] = hr[k];
            }
            n++;
        }
    }
    if (total) *total = n;
    return std::min(n, cap);
}

// Detector::evaluate(const cv::Mat&) ACF.cpp:123-133 + acfDetect1.cpp:337-342: channels of the whole image through
// chnsCompute (no pyramid, no final smoothing, no padding), score of the single window at (0,0) with cascThr = 0.
// The reference computes these channels with computeChannels' FIXED default options (ACF.cpp:165-240); the caller passes them.
int oracle_evaluate(const oracle_opts* o, const void* img, int rows, int cols, int is_f32, const oracle_clf* clf, float* score)
{
    try
    {
        Planes I = Pyr::convert_input(*o, img, rows, cols, is_f32 != 0);
        Chns ch;
        chns_compute(I, *o, ch, nullptr, nullptr, 0);
        int d = 0;
        for (auto& t : ch.data) d += t.d;
        const int h = ch.data[0].h, w = ch.data[0].w;
        std::vector<float> F((size_t)h * w * d);
        size_t off = 0;
        for (auto& t : ch.data) { const size_t n = (size_t)t.h * t.w * t.d; memcpy(F.data() + off, t.p(), n * sizeof(float)); off += n; }
        const int mH = o->modelDsPad_w / o->shrink, mW = o->modelDsPad_h / o->shrink;
        float hs = 0.f;
        for (int t = 0; t < clf->nTrees; t++)
        {
            uint32_t offs = t * clf->nTreeNodes, k = offs, k0 = 0;
            for (int i = 0; i < clf->treeDepth; i++)
            {
                const uint32_t fid = clf->fids[k];
                const uint32_t r = fid % mH, c = (fid / mH) % mW, z = fid / (mH * mW);
                const float ftr = F[(size_t)z * w * h + (size_t)c * h + r];
                k = (ftr < clf->thrs[k]) ? 1 : 2;
                k0 = k += k0 * 2;
                k += offs;
            }
            hs += clf->hs[k];
            if (hs <= 0.f) break;
        }
        *score = hs;
        return 0;
    }
    catch (const std::exception& e) { g_err = e.what(); }
    catch (const char* e) { g_err = e; }
    return 1;
}

int oracle_nms(oracle_det* dets, int n, double overlap, int greedy, int ovr_union)
{
    // nmsMax bbNms.cpp:111-192 (std::sort on score descending; ties unordered in the reference, stable here)
    std::vector<int> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return dets[a].score > dets[b].score; });
    std::vector<oracle_det> bbs(n);
    struct Roi { int as, xs, xe, ys, ye, kp; };
    std::vector<Roi> c(n);
    for (int i = 0; i < n; i++)
    {
        bbs[i] = dets[ord[i]];
        c[i] = { bbs[i].w * bbs[i].h, bbs[i].x, bbs[i].x + bbs[i].w, bbs[i].y, bbs[i].y + bbs[i].h, 1 };
    }
    for (int i = 0; i < n; i++)
    {
        if (greedy && !c[i].kp) continue;
        for (int j = i + 1; j < n; j++)
        {
            if (c[j].kp == 0) continue;
            const int iw = std::min(c[i].xe, c[j].xe) - std::max(c[i].xs, c[j].xs);
            if (iw <= 0) continue;
            const int ih = std::min(c[i].ye, c[j].ye) - std::max(c[i].ys, c[j].ys);
            if (ih <= 0) continue;
            double ov = (iw * ih);
            const double u = ovr_union ? (c[i].as + c[j].as - ov) : std::min(c[i].as, c[j].as);
            ov /= u;
            if (ov > overlap) c[j].kp = 0;
        }
    }
    int m = 0;
    for (int i = 0; i < n; i++)
        if (c[i].kp) dets[m++] = bbs[i];
    return m;
}

int oracle_prune(oracle_det* dets, int n, int max_count, double prune_ratio)
{
    if (n > 1)
    {
        int cutoff = 1;
        for (int i = 1; i < std::min(max_count, n); i++)
        {
            cutoff = i + 1;
            if (dets[i].score < dets[0].score * prune_ratio) break;
        }
        return cutoff;
    }
    return n;
}

void oracle_rgb_convert(const float* I, float* J, int n, int flag) { oracle_l1().rgbConvert(const_cast<float*>(I), J, n, 3, flag, 1.0f); }
void oracle_conv_tri1(float* I, float* O, int h, int w, int d, float p, int s) { oracle_l1().convTri1(I, O, h, w, d, p, s); }
void oracle_conv_tri(float* I, float* O, int h, int w, int d, int r, int s) { oracle_l1().convTri(I, O, h, w, d, r, s); }
void oracle_grad_mag(float* I, float* M, float* O, int h, int w, int d, int full) { oracle_l1().gradMag(I, M, O, h, w, d, full != 0); }
void oracle_grad_mag_norm(float* M, float* S, int h, int w, float norm) { oracle_l1().gradMagNorm(M, S, h, w, norm); }
void oracle_grad_hist(float* M, float* O, float* H, int h, int w, int bin, int nOrients, int softBin, int full)
{
    oracle_l1().gradHist(M, O, H, h, w, bin, nOrients, softBin, full != 0);
}
void oracle_resample(float* A, float* B, int ha, int hb, int wa, int wb, int d, float r) { oracle_l1().resample(A, B, ha, hb, wa, wb, d, r); }

} // extern "C"
