// l1_port.cpp -- scalar, IEEE-exact restatement of the reference's numeric kernels (L1).
//
// TEST INFRASTRUCTURE ONLY (see acf_oracle.h).  Written from the algorithm, one scalar lane at
// a time, keeping the reference's operation ORDER so that results are bit-identical with the
// reference's own object code built in "exact" mode (oracle/_ref/liboracle_ref_exact.so:
// rcpps -> 1/x, rsqrtps -> 1/sqrt(x)); tests/test_oracle_pinning.py checks that equality.
// All arrays use the toolbox's column-major convention: element (x, y) of an h-by-w plane is
// at [x*h + y].
//
// Follows (reference file:line under src/lib/acf/acf/):
//   port::rgbConvert   toolbox/rgbConvertMex.cpp:20-59 (tables), :88-190 (SSE luv order), :193-238 (hsv), :242-252 (gray), :339-380
//   port::convTri1     toolbox/convConst.cpp:445-525
//   port::convTri      toolbox/convConst.cpp:269-442
//   port::gradMag      toolbox/gradientMex.cpp:17-87 (grad1), :103-165 (acos table), :168-251
//   port::gradMagNorm  toolbox/gradientMex.cpp:254-275
//   port::gradHist     toolbox/gradientMex.cpp:278-372 (quantize), :375-509 (orientation-soft / hard branches)
//   port::resample     toolbox/imResampleMex.cpp:25-121 (coefficients), :125-383
#include "acf_oracle.h"
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace port
{

// ---------------------------------------------------------------- colour
struct LuvTables
{
    float lut[1064];
    float mr[3], mg[3], mb[3], minu, minv, un13, vn13;
    LuvTables()
    {
        const float z = 1.0f;
        const float y0 = (float)((6.0 / 29) * (6.0 / 29) * (6.0 / 29));
        const float a = (float)((29.0 / 3) * (29.0 / 3) * (29.0 / 3));
        const float un = (float)0.197833, vn = (float)0.468331;
        mr[0] = (float)0.430574 * z; mr[1] = (float)0.222015 * z; mr[2] = (float)0.020183 * z;
        mg[0] = (float)0.341550 * z; mg[1] = (float)0.706655 * z; mg[2] = (float)0.129553 * z;
        mb[0] = (float)0.178325 * z; mb[1] = (float)0.071330 * z; mb[2] = (float)0.939180 * z;
        const float maxi = (float)1.0 / 270;
        minu = -88 * maxi;
        minv = -134 * maxi;
        un13 = 13 * un;
        vn13 = 13 * vn;
        for (int i = 0; i < 1025; i++)
        {
            float y = (float)(i / 1024.0);
            float l = y > y0 ? 116 * (float)pow((double)y, 1.0 / 3.0) - 16 : y * a;
            lut[i] = l * maxi;
        }
        for (int i = 1025; i < 1064; i++) lut[i] = lut[i - 1];
    }
};

static void rgb2luv(const float* I, float* J, int n)
{
    static const LuvTables T;
    const float *R = I, *G = I + n, *B = I + 2 * n;
    float *L = J, *U = J + n, *V = J + 2 * n;
    for (int i = 0; i < n; i++)
    {
        const float r = R[i], g = G[i], b = B[i];
        const float x = (r * T.mr[0] + g * T.mg[0]) + b * T.mb[0];
        const float y = (r * T.mr[1] + g * T.mg[1]) + b * T.mb[1];
        const float z = (r * T.mr[2] + g * T.mg[2]) + b * T.mb[2];
        // SSE path order (rgbConvertMex.cpp:161): x + (eps + (15 y + 3 z)), then reciprocal
        const float den = x + (1e-35f + (15.0f * y + 3.0f * z));
        const float zi = 1.0f / den;
        const float li = 1024.0f * y;
        const float up = (52.0f * x) * zi - T.un13;
        const float vp = (117.0f * y) * zi - T.vn13;
        const float l = T.lut[(int)li];
        L[i] = l;
        U[i] = l * up - T.minu;
        V[i] = l * vp - T.minv;
    }
}

static void rgb2gray(const float* I, float* J, int n)
{
    const float nrm = 1.0f;
    const float mr = (float).2989360213 * nrm, mg = (float).5870430745 * nrm, mb = (float).1140209043 * nrm;
    const float *R = I, *G = I + n, *B = I + 2 * n;
    for (int i = 0; i < n; i++) J[i] = R[i] * mr + G[i] * mg + B[i] * mb;
}

// hue / saturation / value (rgbConvertMex.cpp:193-238, nrm == 1): the sector is chosen red first, then green, then blue
// (ties go to the earlier test); hue = (sector offset + chroma difference / range) / 6, wrapped into [0,1) for red
static void rgb2hsv(const float* I, float* J, int n)
{
    const float *R = I, *G = I + n, *B = I + 2 * n;
    float *Hh = J, *Ss = J + n, *Vv = J + 2 * n;
    const float sixth = (float)(1 / 6.0);
    for (int i = 0; i < n; i++)
    {
        const float r = R[i], g = G[i], b = B[i];
        if (r == g && g == b) { Hh[i] = 0; Ss[i] = 0; Vv[i] = r; continue; }
        float hi, lo, hue;
        if (r >= g && r >= b)
        {
            hi = r; lo = (g < b) ? g : b;
            hue = (g - b) / (hi - lo) + 6;
            if (hue >= 6) hue -= 6;
        }
        else if (g >= r && g >= b) { hi = g; lo = (r < b) ? r : b; hue = (b - r) / (hi - lo) + 2; }
        else { hi = b; lo = (r < g) ? r : g; hue = (r - g) / (hi - lo) + 4; }
        Hh[i] = hue * sixth;
        Ss[i] = 1 - lo / hi;
        Vv[i] = hi;
    }
}

static void rgbConvert(float* I, float* J, int n, int d, int flag, float nrm)
{
    if (nrm != 1.0f) throw std::runtime_error("port::rgbConvert: nrm must be 1");
    if (flag == 2 && d == 3) rgb2luv(I, J, n);
    else if (flag == 0 && d == 3) rgb2gray(I, J, n);
    else if ((flag == 0 && d == 1) || flag == 1) { for (int i = 0; i < n * d; i++) J[i] = I[i] * nrm; }
    else if (flag == 3 && d == 3) rgb2hsv(I, J, n);
    else throw std::runtime_error("port::rgbConvert: unsupported flag/d");
}

// ---------------------------------------------------------------- [1 p 1] smoothing
static void convTri1(float* I, float* O, int h, int w, int d, float p, int s)
{
    if (s != 1) throw std::runtime_error("port::convTri1: only s==1 is on the hot path");
    const float nrm = 1.0f / ((p + 2) * (p + 2));
    std::vector<float> T(h);
    for (int d0 = 0; d0 < d; d0++)
    {
        for (int i = 0; i < w; i++)
        {
            // NOTE: when O aliases I (the reference's in-place call) column i-1
            // already holds output; the reads below deliberately see that.
            const float* Im = I + (size_t)i * h + (size_t)d0 * h * w;
            const float* Il = i > 0 ? Im - h : Im;
            const float* Ir = i < w - 1 ? Im + h : Im;
            for (int j = 0; j < h; j++) T[j] = nrm * ((Il[j] + p * Im[j]) + Ir[j]);
            float* Oc = O + (size_t)i * h + (size_t)d0 * h * w;
            Oc[0] = (1 + p) * T[0] + T[1];
            for (int j = 1; j < h - 1; j++) Oc[j] = (T[j - 1] + p * T[j]) + T[j + 1];
            Oc[h - 1] = T[h - 2] + (1 + p) * T[h - 1];
        }
    }
}

// ---------------------------------------------------------------- triangle filter via running sums
static void convTriY(const float* I, float* O, int h, int r)
{
    r++;
    float t, u;
    const int r0 = r - 1, r1 = r + 1, r2 = 2 * h - r, h0 = r + 1, h1 = h - r + 1, h2 = h;
    u = t = I[0];
    for (int j = 1; j < r; j++) { t += I[j]; u += t; }
    u = 2 * u - t;
    t = 0;
    O[0] = u;
    int j = 1;
    for (; j < h0; j++) { t += (I[r - j] + I[r0 + j]) - 2 * I[j - 1]; u += t; O[j] = u; }
    for (; j < h1; j++) { t += (I[j - r1] + I[r0 + j]) - 2 * I[j - 1]; u += t; O[j] = u; }
    for (; j < h2; j++) { t += (I[j - r1] + I[r2 - j]) - 2 * I[j - 1]; u += t; O[j] = u; }
}

static void convTri(float* I, float* O, int h, int w, int d, int r, int s)
{
    if (s != 1) throw std::runtime_error("port::convTri: only s==1 is on the hot path");
    r++;
    const float nrm = 1.0f / (r * r * r * r);
    std::vector<float> T(h), U(h);
    while (d-- > 0)
    {
        for (int j = 0; j < h; j++) U[j] = T[j] = I[j];
        for (int i = 1; i < r; i++)
            for (int j = 0; j < h; j++) { T[j] += I[j + (size_t)i * h]; U[j] += T[j]; }
        for (int j = 0; j < h; j++) { U[j] = nrm * (2 * U[j] - T[j]); T[j] = 0; }
        convTriY(U.data(), O, h, r - 1);
        O += h;
        for (int i = 1; i < w; i++)
        {
            const float* Il = I + (size_t)(i - 1 - r) * h;
            if (i <= r) Il = I + (size_t)(r - i) * h;
            const float* Im = I + (size_t)(i - 1) * h;
            const float* Ir = I + (size_t)(i - 1 + r) * h;
            if (i > w - r) Ir = I + (size_t)(2 * w - r - i) * h;
            for (int j = 0; j < h; j++)
            {
                T[j] += (Il[j] + Ir[j]) + (-2.0f * Im[j]);
                U[j] += nrm * T[j];
            }
            convTriY(U.data(), O, h, r - 1);
            O += h;
        }
        I += (size_t)w * h;
    }
}

// ---------------------------------------------------------------- gradient magnitude / orientation
struct AcosTable
{
    enum { n = 10000, b = 10 };
    std::vector<float> a;
    AcosTable() : a(2 * (n + b))
    {
        const float PI = 3.14159265f;
        float* a1 = a.data() + n + b;
        for (int i = -n - b; i < -n; i++) a1[i] = PI;
        for (int i = -n; i < n; i++) a1[i] = float(std::acos(i / float(n)));
        for (int i = n; i < n + b; i++) a1[i] = 0;
        for (int i = -n - b; i < n / 10; i++)
            if (a1[i] > PI - 1e-6f) a1[i] = PI - 1e-6f;
    }
    float operator[](int i) const { return a[i + n + b]; }
};

static void gradMag(float* I, float* M, float* O, int h, int w, int d, bool full)
{
    if (d != 1) throw std::runtime_error("port::gradMag: the reference always passes d==1 (chnsCompute.cpp:278)");
    static const AcosTable acosT;
    const float PI = 3.14159265f;
    const float upper = (float)(AcosTable::n + AcosTable::b - 1), lower = -upper;
    for (int x = 0; x < w; x++)
    {
        const float* Ic = I + (size_t)x * h;
        const float* Ip = Ic - h;
        const float* In = Ic + h;
        float rx = .5f;
        if (x == 0) { rx = 1; Ip += h; }
        else if (x == w - 1) { rx = 1; In -= h; }
        for (int y = 0; y < h; y++)
        {
            const float gx = (In[y] - Ip[y]) * rx;
            float gy;
            if (y == 0) gy = (Ic[1] - Ic[0]) * 1.0f;
            else if (y == h - 1) gy = (Ic[h - 1] - Ic[h - 2]) * 1.0f;
            else gy = (Ic[y + 1] - Ic[y - 1]) * .5f;
            const float m2 = gx * gx + gy * gy;
            float m = 1.0f / std::sqrt(m2);
            m = (m < 1e10f) ? m : 1e10f; // _mm_min_ps(a, b): a < b ? a : b
            M[(size_t)x * h + y] = 1.0f / m;
            if (O)
            {
                float g = (gx * m) * 10000.0f;
                if (std::signbit(gy)) g = -g;
                g = (g < upper) ? g : upper;
                g = (g > lower) ? g : lower;
                float o = acosT[(int)g];
                if (full) o += (gy < 0) * PI;
                O[(size_t)x * h + y] = o;
            }
        }
    }
}

static void gradMagNorm(float* M, float* S, int h, int w, float norm)
{
    const int n = h * w, n4 = (n / 4) * 4;
    int i = 0;
    for (; i < n4; i++) M[i] = M[i] * (1.0f / (S[i] + norm)); // MUL(M, RCP(S + norm)) with exact RCP
    for (; i < n; i++) M[i] /= (S[i] + norm);
}

// ---------------------------------------------------------------- gradient histograms
static void gradHist(float* M, float* O, float* H, int h, int w, int bin, int nOrients, int softBin, bool full)
{
    if (!(softBin % 2 == 0 || bin == 1))
        throw std::runtime_error("port::gradHist: trilinear (odd softBin) branch is outside the hot path");
    const float PI = 3.14159265f;
    const int hb = h / bin, wb = w / bin, h0 = hb * bin, w0 = wb * bin, nb = wb * hb;
    const float s = (float)bin, sInv2 = 1 / s / s;
    const float oMult = (float)nOrients / (full ? 2 * PI : PI);
    const int oMax = nOrients * nb;
    const bool interpolate = softBin >= 0;
    for (int x = 0; x < w0; x++)
    {
        float* H1 = H + (size_t)(x / bin) * hb;
        for (int y = 0; y < h0; y++)
        {
            const float ov = O[(size_t)x * h + y], mv = M[(size_t)x * h + y];
            if (interpolate)
            {
                const float o = ov * oMult;
                int o0 = (int)o;
                const float od = o - (float)o0;
                o0 *= nb;
                if (o0 >= oMax) o0 = 0;
                int o1 = o0 + nb;
                if (o1 >= oMax) o1 = 0;
                const float m = mv * sInv2;
                const float m1 = od * m;
                const float m0 = m - m1;
                H1[o0 + y / bin] += m0;
                H1[o1 + y / bin] += m1;
            }
            else
            {
                const float o = ov * oMult;
                int o0 = (int)(o + .5f);
                o0 *= nb;
                if (o0 >= oMax) o0 = 0;
                H1[o0 + y / bin] += mv * sInv2;
            }
        }
    }
}

// ---------------------------------------------------------------- resampling
struct Coef
{
    int n = 0;
    std::vector<int> yas, ybs;
    std::vector<float> wts;
    int bd[2] = { 0, 0 };
};

static Coef resampleCoef(int ha, int hb, int pad)
{
    Coef c;
    const float s = float(hb) / float(ha), sInv = 1 / s;
    const float wt0 = float(1e-3) * s;
    const bool ds = ha > hb;
    if (ds)
    {
        for (int yb = 0; yb < hb; yb++)
        {
            const float ya0f = yb * sInv, ya1f = ya0f + sInv;
            float W = 0;
            const int ya0 = int(std::ceil(ya0f)), ya1 = int(ya1f);
            int n1 = 0;
            for (int ya = ya0 - 1; ya < ya1 + 1; ya++)
            {
                float wt = s;
                if (ya == ya0 - 1) wt = (ya0 - ya0f) * s;
                else if (ya == ya1) wt = (ya1f - ya1) * s;
                if (wt > wt0 && ya >= 0)
                {
                    c.ybs.push_back(yb); c.yas.push_back(ya); c.wts.push_back(wt);
                    n1++;
                    W += wt;
                }
            }
            if (W > 1)
                for (int i = 0; i < n1; i++) c.wts[c.wts.size() - n1 + i] /= W;
            if (n1 > c.bd[0]) c.bd[0] = n1;
            while (n1 < pad)
            {
                c.ybs.push_back(yb); c.yas.push_back(c.yas.back()); c.wts.push_back(0);
                n1++;
            }
        }
        c.n = (int)c.wts.size();
    }
    else
    {
        c.n = hb;
        c.yas.resize(hb); c.ybs.resize(hb); c.wts.resize(hb);
        for (int yb = 0; yb < hb; yb++)
        {
            const float yaf = (float(.5) + yb) * sInv - float(.5);
            int ya = (int)std::floor(yaf);
            float wt = 1;
            if (ya >= 0 && ya < ha - 1) wt = 1 - (yaf - ya);
            if (ya < 0) { ya = 0; c.bd[0]++; }
            if (ya >= ha - 1) { ya = ha - 1; c.bd[1]++; }
            c.ybs[yb] = yb; c.yas[yb] = ya; c.wts[yb] = wt;
        }
    }
    return c;
}

static void resample(float* A, float* B, int ha, int hb, int wa, int wb, int d, float r)
{
    if (A == B) throw std::runtime_error("port::resample: A == B");
    std::vector<float> C(ha + 4, 0.0f);
    Coef cx = resampleCoef(wa, wb, 0);
    Coef cy = resampleCoef(ha, hb, 4);
    const int wn = cx.n, hn = cy.n;
    if (wa == 2 * wb) r /= 2;
    if (wa == 3 * wb) r /= 3;
    if (wa == 4 * wb) r /= 4;
    r /= float(1 + 1e-6);
    for (int y = 0; y < hn; y++) cy.wts[y] *= r;
    // the padded 4-tap form reads up to yas[4y]+3 and the bilinear form yas[y]+1
    const int *xas = cx.yas.data(), *xbs = cx.ybs.data(), *yas = cy.yas.data(), *ybs = cy.ybs.data();
    const float *xwts = cx.wts.data(), *ywts = cy.wts.data();
    int x1 = 0;
    for (int z = 0; z < d; z++)
    {
        for (int x = 0; x < wb; x++)
        {
            if (x == 0) x1 = 0;
            const int xa = xas[x1], xb = xbs[x1];
            const float wt = xwts[x1], wt1 = 1 - wt;
            const float* A0 = A + (size_t)z * ha * wa + (size_t)xa * ha;
            const float *A1 = A0 + ha, *A2 = A1 + ha, *A3 = A2 + ha;
            float* B0 = B + (size_t)z * hb * wb + (size_t)xb * hb;
            // ---- along x: A -> C
            if (wa == 2 * wb) { for (int y = 0; y < ha; y++) C[y] = A0[y] + A1[y]; x1 += 2; }
            else if (wa == 3 * wb) { for (int y = 0; y < ha; y++) C[y] = (A0[y] + A1[y]) + A2[y]; x1 += 3; }
            else if (wa == 4 * wb) { for (int y = 0; y < ha; y++) C[y] = ((A0[y] + A1[y]) + A2[y]) + A3[y]; x1 += 4; }
            else if (wa > wb)
            {
                int m = 1;
                while (x1 + m < wn && xb == xbs[x1 + m]) m++;
                const float w0 = xwts[x1], w1 = m > 1 ? xwts[x1 + 1] : 0, w2 = m > 2 ? xwts[x1 + 2] : 0, w3 = m > 3 ? xwts[x1 + 3] : 0;
                if (m == 1) for (int y = 0; y < ha; y++) C[y] = A0[y] * w0;
                if (m == 2) for (int y = 0; y < ha; y++) C[y] = A0[y] * w0 + A1[y] * w1;
                if (m == 3) for (int y = 0; y < ha; y++) C[y] = (A0[y] * w0 + A1[y] * w1) + A2[y] * w2;
                if (m >= 4) for (int y = 0; y < ha; y++) C[y] = ((A0[y] * w0 + A1[y] * w1) + A2[y] * w2) + A3[y] * w3;
                for (int x0 = 4; x0 < m; x0++)
                {
                    const float* Ak = A0 + (size_t)x0 * ha;
                    const float wk = xwts[x1 + x0];
                    for (int y = 0; y < ha; y++) C[y] = C[y] + Ak[y] * wk;
                }
                x1 += m;
            }
            else
            {
                const bool xBd = x < cx.bd[0] || x >= wb - cx.bd[1];
                x1++;
                if (xBd) memcpy(C.data(), A0, ha * sizeof(float));
                else for (int y = 0; y < ha; y++) C[y] = A0[y] * wt + A1[y] * wt1;
            }
            // ---- along y: C -> B
            if (ha == hb * 2)
            {
                const float r2 = r / 2;
                for (int y = 0; y < hb; y++) B0[y] = (C[2 * y] + C[2 * y + 1]) * r2;
            }
            else if (ha == hb * 3)
            {
                for (int y = 0; y < hb; y++) B0[y] = (C[3 * y] + C[3 * y + 1] + C[3 * y + 2]) * (r / 3);
            }
            else if (ha == hb * 4)
            {
                for (int y = 0; y < hb; y++) B0[y] = (C[4 * y] + C[4 * y + 1] + C[4 * y + 2] + C[4 * y + 3]) * (r / 4);
            }
            else if (ha > hb)
            {
                const int nb = cy.bd[0];
                if (nb <= 4)
                {
                    for (int y = 0; y < hb; y++)
                    {
                        const int ya = yas[y * 4];
                        float v = C[ya] * ywts[y * 4];
                        if (nb >= 2) v = v + C[ya + 1] * ywts[y * 4 + 1];
                        if (nb >= 3) v = v + C[ya + 2] * ywts[y * 4 + 2];
                        if (nb >= 4) v = v + C[ya + 3] * ywts[y * 4 + 3];
                        if (nb >= 2) B0[y] = v; // nb == 1 cannot occur for ha > hb; the reference writes nothing then
                    }
                }
                else
                {
                    memset(B0, 0, hb * sizeof(float));
                    for (int y = 0; y < hn; y++) B0[ybs[y]] += C[yas[y]] * ywts[y];
                }
            }
            else
            {
                int y = 0;
                for (; y < cy.bd[0]; y++) B0[y] = C[yas[y]] * ywts[y];
                for (; y < hb - cy.bd[1]; y++) B0[y] = C[yas[y]] * ywts[y] + C[yas[y] + 1] * (r - ywts[y]);
                for (; y < hb; y++) B0[y] = C[yas[y]] * ywts[y];
            }
        }
    }
}

} // namespace port

#ifdef ORACLE_L1_PORT
const OracleL1& oracle_l1()
{
    static const OracleL1 t = { "port", port::rgbConvert, port::convTri1, port::convTri, port::gradMag,
                                port::gradMagNorm, port::gradHist, port::resample };
    return t;
}
#endif
