// plan.cpp -- see plan.h.  Host-only; no CUDA here.
#include "plan.h"
#include <algorithm>
#include <cmath>
#include <limits>

namespace acfb
{

static double log2ref(double x) { return std::log(x) / std::log(2.0); } // util/acf_math.h:20-29 (ln x / ln 2)

// Detector::getScales, chnsPyramid.cpp:461-529.  sz_w = frame rows, sz_h = frame cols (the reference
// sees the transposed image, so its cv::Size width is the original row count).
void getScales(int nPerOct, int nOctUp, int minDs_w, int minDs_h, int shrink, int sz_w, int sz_h,
               std::vector<double>& scales, std::vector<std::pair<double, double>>& scaleshw)
{
    scales.clear();
    scaleshw.clear();
    if (sz_w <= 0 || sz_h <= 0) return;
    const double ratioW = double(sz_w) / double(minDs_w), ratioH = double(sz_h) / double(minDs_h);
    const int nCand = (int)std::floor(double(nPerOct) * (double(nOctUp) + log2ref(std::min(ratioW, ratioH))) + 1.0);
    double dShort = sz_h, dLong = sz_w;
    if (sz_h >= sz_w) std::swap(dShort, dLong);
    std::vector<double> cand;
    for (int i = 0; i < nCand; i++)
    {
        const double s = std::pow(2.0, -double(i) / double(nPerOct) + double(nOctUp));
        const double lo = (std::round(dShort * s / shrink) * shrink - 0.25 * shrink) / dShort;
        const double hi = (std::round(dShort * s / shrink) * shrink + 0.25 * shrink) / dShort;
        double best = 0, bestErr = std::numeric_limits<double>::max();
        // the reference accumulates j += 0.01 in fp64; the accumulated value (not k*0.01) decides ties
        for (double j = 0.0; j < 1.0 - std::numeric_limits<double>::epsilon(); j += 0.01)
        {
            const double ss = j * (hi - lo) + lo;
            double e0 = dShort * ss;
            e0 = std::abs(e0 - std::round(e0 / shrink) * shrink);
            double e1 = dLong * ss;
            e1 = std::abs(e1 - std::round(e1 / shrink) * shrink);
            const double e = std::max(e0, e1);
            if (e < bestErr) { best = ss; bestErr = e; }
        }
        cand.push_back(best);
    }
    cand.push_back(0);
    for (size_t i = 1; i < cand.size(); i++)
    {
        if (cand[i] == cand[i - 1]) continue; // drop duplicates
        const double s = cand[i - 1];
        scales.push_back(s);
        scaleshw.emplace_back(std::round(double(sz_w) * s / shrink) * shrink / sz_w,
                              std::round(double(sz_h) * s / shrink) * shrink / sz_h);
    }
}

// resampleCoef, imResampleMex.cpp:25-121, in its own terms: a list of (dst, src, weight) entries.
struct CoefList
{
    std::vector<int> src, dst;
    std::vector<float> wt;
    int bd0 = 0, bd1 = 0;
};

static CoefList coefList(int na, int nb, int pad)
{
    CoefList c;
    const float s = float(nb) / float(na), sInv = 1 / s;
    const float wtMin = float(1e-3) * s;
    if (na > nb)
    {
        for (int b = 0; b < nb; b++)
        {
            const float a0f = b * sInv, a1f = a0f + sInv;
            const int a0 = int(std::ceil(a0f)), a1 = int(a1f);
            float W = 0;
            int cnt = 0;
            for (int a = a0 - 1; a < a1 + 1; a++)
            {
                float wt = s;
                if (a == a0 - 1) wt = (a0 - a0f) * s;
                else if (a == a1) wt = (a1f - a1) * s;
                if (wt > wtMin && a >= 0)
                {
                    c.dst.push_back(b); c.src.push_back(a); c.wt.push_back(wt);
                    cnt++;
                    W += wt;
                }
            }
            if (W > 1)
                for (int i = 0; i < cnt; i++) c.wt[c.wt.size() - cnt + i] /= W;
            c.bd0 = std::max(c.bd0, cnt);
            for (; cnt < pad; cnt++) { c.dst.push_back(b); c.src.push_back(c.src.back()); c.wt.push_back(0); }
        }
    }
    else
    {
        for (int b = 0; b < nb; b++)
        {
            const float af = (float(.5) + b) * sInv - float(.5);
            int a = (int)std::floor(af);
            float wt = 1;
            if (a >= 0 && a < na - 1) wt = 1 - (af - a);
            if (a < 0) { a = 0; c.bd0++; }
            if (a >= na - 1) { a = na - 1; c.bd1++; }
            c.dst.push_back(b); c.src.push_back(a); c.wt.push_back(wt);
        }
    }
    return c;
}

static const int kMaxTaps = 12; // real-scale images are down-sampled by up to ~8x (4K: 1080 -> 136 rows)
static inline int roundUp(int v, int a) { return (v + a - 1) / a * a; }

// first pass of resample<T> (imResampleMex.cpp:184-280): along the slow (x / column) axis
AxisCoef makeAxisX(int wa, int wb)
{
    AxisCoef ax;
    ax.nIn = wa; ax.nOut = wb; ax.maxTaps = kMaxTaps; ax.mode = 0;
    ax.start.assign(wb, 0); ax.cnt.assign(wb, 0); ax.wt.assign((size_t)wb * kMaxTaps, 0.0f);
    const CoefList c = coefList(wa, wb, 0);
    const int n = (int)c.wt.size();
    int k = (wa == 2 * wb) ? 2 : (wa == 3 * wb) ? 3 : (wa == 4 * wb) ? 4 : 0;
    ax.rdiv = k ? float(k) : 1.0f;
    int i1 = 0;
    for (int x = 0; x < wb; x++)
    {
        if (i1 >= n) throw std::runtime_error("plan: resample coefficient walk ran off the table");
        const int xb = c.dst[i1];
        ax.start[xb] = c.src[i1];
        float* w = &ax.wt[(size_t)xb * kMaxTaps];
        if (k)
        { // integer ratio: plain sum of k consecutive columns, r divided by k afterwards
            ax.cnt[xb] = k;
            for (int j = 0; j < k; j++) w[j] = 1.0f;
            i1 += k;
        }
        else if (wa > wb)
        {
            int m = 1;
            while (i1 + m < n && xb == c.dst[i1 + m]) m++;
            if (m > kMaxTaps) throw std::runtime_error("plan: down-sampling ratio too large for the tap table");
            ax.cnt[xb] = m;
            for (int j = 0; j < m; j++) w[j] = c.wt[i1 + j];
            i1 += m;
        }
        else
        {
            const bool border = x < c.bd0 || x >= wb - c.bd1;
            const float wt = c.wt[i1];
            if (border) { ax.cnt[xb] = 1; w[0] = 1.0f; } // memcpy of the source column
            else { ax.cnt[xb] = 2; w[0] = wt; w[1] = 1 - wt; }
            i1++;
        }
        if (ax.start[xb] + ax.cnt[xb] > wa) throw std::runtime_error("plan: resample taps leave the source");
    }
    return ax;
}

// second pass (imResampleMex.cpp:283-372): along the contiguous (y) axis
AxisCoef makeAxisY(int ha, int hb)
{
    AxisCoef ay;
    ay.nIn = ha; ay.nOut = hb; ay.maxTaps = kMaxTaps;
    ay.start.assign(hb, 0); ay.cnt.assign(hb, 0); ay.wt.assign((size_t)hb * kMaxTaps, 0.0f);
    const int k = (ha == 2 * hb) ? 2 : (ha == 3 * hb) ? 3 : (ha == 4 * hb) ? 4 : 0;
    if (k)
    {
        ay.mode = 1; ay.ymul = k;
        for (int y = 0; y < hb; y++)
        {
            ay.start[y] = k * y; ay.cnt[y] = k;
            for (int j = 0; j < k; j++) ay.wt[(size_t)y * kMaxTaps + j] = 1.0f;
        }
        return ay;
    }
    const CoefList c = coefList(ha, hb, 4);
    if (ha > hb)
    {
        ay.mode = 0;
        // both reference forms (fixed nb taps from yas[4y], or the accumulate loop for nb > 4) are the
        // same ordered sum over the non-padded entries of row y; zero-weight pad entries add +-0.
        for (size_t i = 0; i < c.wt.size(); i++)
        {
            const int y = c.dst[i];
            if (c.wt[i] == 0.0f) continue;
            int& n = ay.cnt[y];
            if (n == 0) ay.start[y] = c.src[i];
            if (c.src[i] != ay.start[y] + n) throw std::runtime_error("plan: non-consecutive y taps");
            if (n >= kMaxTaps) throw std::runtime_error("plan: down-sampling ratio too large for the tap table");
            ay.wt[(size_t)y * kMaxTaps + n] = c.wt[i];
            n++;
        }
        for (int y = 0; y < hb; y++)
            if (ay.cnt[y] == 0 || ay.start[y] + ay.cnt[y] > ha) throw std::runtime_error("plan: bad y taps");
    }
    else
    {
        ay.mode = 2;
        for (int y = 0; y < hb; y++)
        {
            const bool border = y < c.bd0 || y >= hb - c.bd1;
            ay.start[y] = c.src[y];
            ay.cnt[y] = border ? 1 : 2;
            ay.wt[(size_t)y * kMaxTaps] = c.wt[y];
            if (!border && c.src[y] + 1 >= ha) throw std::runtime_error("plan: bilinear tap leaves the source");
        }
    }
    return ay;
}

Plan makePlan(const acfb_options& o, int rows, int cols)
{
    Plan p;
    p.rows = rows; p.cols = cols; p.shrink = o.shrink;
    if (o.shrink != 4) throw std::runtime_error("engine: only shrink == 4 is implemented (every shipped model uses 4)");
    if (o.gh_binSize != 0 && o.gh_binSize != o.shrink) throw std::runtime_error("engine: pGradHist.binSize must equal shrink");
    if (o.color_space < 0 || o.color_space > 4) throw std::runtime_error("engine: colorSpace must be gray, rgb, luv, hsv or orig");
    if (!o.gm_enabled || !o.gh_enabled) throw std::runtime_error("engine: pGradMag and pGradHist must be enabled");
    if (o.gm_colorChn < 0 || o.gm_colorChn >= ((o.color_space == 0) ? 1 : 3)) throw std::runtime_error("engine: pGradMag.colorChn outside the image planes");
    if (o.gm_normRad != 5 && o.gm_normRad != 0) throw std::runtime_error("engine: pGradMag.normRad must be 5 or 0 (k_trix / k_triyhist march the radius-5 triangle)");
    if (o.nPerOct < 1 || o.nOctUp < 0 || o.nApprox < 0 || o.minDs_w < 1 || o.minDs_h < 1 || o.pad_w < 0 || o.pad_h < 0 || o.stride < 1)
        throw std::runtime_error("engine: nPerOct >= 1, nOctUp >= 0, nApprox >= 0, minDs >= 1, pad >= 0, stride >= 1 expected");
    if (o.gh_softBin != 0) throw std::runtime_error("engine: pGradHist.softBin must be 0 (orientation-soft, spatially hard binning)");
    if (o.gh_nOrients < 1 || o.gh_nOrients > 8) throw std::runtime_error("engine: nOrients must be in 1..8");
    if (!(o.smooth >= 0 && o.smooth <= 1.0) || !(o.color_smooth >= 0 && o.color_smooth <= 1.0))
        throw std::runtime_error("engine: smooth / pColor.smooth must be in [0,1] (the [1 p 1] branch of convTri)");
    // frame sizes that are not multiples of shrink are resampled to the nearest multiple at scale 1 (chnsPyramid.cpp:300-311)
    p.nImgPlanes = (o.color_space == 0) ? 1 : 3;
    p.nColor = o.color_enabled ? p.nImgPlanes : 0;
    p.typeFirst[0] = 0; p.typeCount[0] = p.nColor;
    p.typeFirst[1] = p.nColor; p.typeCount[1] = 1;
    p.typeFirst[2] = p.nColor + 1; p.typeCount[2] = o.gh_nOrients;
    p.nChns = p.nColor + 1 + o.gh_nOrients;
    p.nTypes = (p.nColor ? 1 : 0) + 2;
    p.padX = o.pad_h / o.shrink; // Mat rows = orig x  (chnsPyramid.cpp:417)
    p.padY = o.pad_w / o.shrink; // Mat cols = orig y  (chnsPyramid.cpp:418)
    getScales(o.nPerOct, o.nOctUp, o.minDs_w, o.minDs_h, o.shrink, rows, cols, p.scales, p.scaleshw);
    const int nScales = (int)p.scales.size();
    if (nScales == 0) throw std::runtime_error("engine: frame smaller than the model (no scales)");
    // real / approximated split, chnsPyramid.cpp:272-292 (1-based there, 0-based here)
    std::vector<int> isR, nearestReal(nScales, 0);
    for (int i = 0; i < nScales; i++)
        if (i % (o.nApprox + 1) == 0) isR.push_back(i);
    std::vector<int> bound(isR.size() + 1, 0);
    bound.back() = nScales;
    for (size_t k = 0; k + 1 < isR.size(); k++) bound[k + 1] = ((isR[k] + 1) + (isR[k + 1] + 1)) / 2;
    for (size_t k = 0; k < isR.size(); k++)
        for (int j = bound[k]; j < bound[k + 1]; j++) nearestReal[j] = (int)k;
    // real scales and where their input image comes from (chnsPyramid.cpp:297-316, A.2 Q2)
    RealScale::Src curKind = RealScale::FROM_I0;
    int curReal = -1, curH = rows, curW = cols;
    for (size_t k = 0; k < isR.size(); k++)
    {
        const double s = p.scales[isR[k]];
        RealScale r;
        r.scaleIdx = isR[k];
        r.h = (int)std::round(double(rows) * s / double(o.shrink)) * o.shrink;
        r.w = (int)std::round(double(cols) * s / double(o.shrink)) * o.shrink;
        r.ch = r.h / o.shrink; r.cw = r.w / o.shrink; r.cP = roundUp(r.ch, 4);
        r.srcKind = curKind; r.srcReal = curReal; r.srcH = curH; r.srcW = curW;
        if (r.h < 16 || r.w < 16) throw std::runtime_error("engine: real scale smaller than 16 px (normalisation radius does not fit)");
        if (r.h == curH && r.w == curW)
        {
            r.mode = RealScale::ALIAS;
            curKind = RealScale::FROM_C; curReal = (int)k; // smoothed in place: later scales see C_k
        }
        else
        {
            const bool half = (curH == 2 * r.h && curW == 2 * r.w);
            r.mode = half ? RealScale::DOWN2 : RealScale::GENERIC;
            r.cx = makeAxisX(curW, r.w);
            r.cy = makeAxisY(curH, r.h);
            float rr = 1.0f;
            rr /= r.cx.rdiv;
            rr /= float(1 + 1e-6);
            r.r = rr;
        }
        if (r.srcKind == RealScale::FROM_C && r.srcReal >= 0) p.reals[r.srcReal].writeC = true;
        p.reals.push_back(r);
        if (s == 0.5 && (o.nApprox > 0 || o.nPerOct == 1)) { curKind = RealScale::FROM_C; curReal = (int)k; curH = r.h; curW = r.w; }
    }
    // per-scale channel geometry
    p.geom.resize(nScales);
    int64_t off = 0;
    for (int i = 0; i < nScales; i++)
    {
        ScaleGeom& g = p.geom[i];
        g.scale = p.scales[i]; g.shw_w = p.scaleshw[i].first; g.shw_h = p.scaleshw[i].second;
        g.realK = nearestReal[i];
        const RealScale& r = p.reals[g.realK];
        g.isReal = (r.scaleIdx == i);
        if (g.isReal) { g.h = r.ch; g.w = r.cw; g.identity = true; }
        else
        {
            g.h = (int)std::round(double(rows) * p.scales[i] / double(o.shrink));
            g.w = (int)std::round(double(cols) * p.scales[i] / double(o.shrink));
            g.cx = makeAxisX(r.cw, g.w);
            g.cy = makeAxisY(r.ch, g.h);
        }
        if (o.smooth > 0 && std::min(g.h, g.w) < 4) throw std::runtime_error("engine: channel plane smaller than 4 px (reference falls back to sepFilter2D)");
        g.H = g.h + 2 * p.padY; g.W = g.w + 2 * p.padX; g.P = roundUp(g.H, 4);
        g.offset = off;
        off += (int64_t)roundUp(g.P * g.W * p.nChns, 32); // keep every scale 128-byte aligned
    }
    p.floatsPerFrame = off;
    p.lambdasFromImage = (o.nLambdas == 0 && o.nApprox > 0);
    if (!p.lambdasFromImage && o.nLambdas < p.nTypes && nScales > (int)isR.size())
        throw std::runtime_error("engine: model has fewer lambdas than channel types");
    if (!p.lambdasFromImage) setRatios(p, o, o.lambdas, o.nLambdas);
    return p;
}

// chnsPyramid.cpp:390-394 + the r adjustments of resample<T> (imResampleMex.cpp:145-158)
void setRatios(Plan& p, const acfb_options& o, const double* lambdas, int nLambdas)
{
    (void)o;
    for (size_t i = 0; i < p.geom.size(); i++)
    {
        ScaleGeom& g = p.geom[i];
        if (g.isReal) continue;
        const double sR = p.scales[p.reals[g.realK].scaleIdx];
        int t = 0;
        for (int type = 0; type < 3; type++)
        {
            if (p.typeCount[type] == 0) continue;
            if (t >= nLambdas) throw std::runtime_error("engine: missing lambda");
            const double ratio = std::pow(p.scales[i] / sR, -lambdas[t]);
            float r = float(ratio);
            r /= g.cx.rdiv;
            r /= float(1 + 1e-6);
            g.ratio[type] = r;
            t++;
        }
    }
}

} // namespace acfb
