// ref_glue.cpp -- binds the oracle's L1 table to the REFERENCE's own toolbox code.
// TEST INFRASTRUCTURE ONLY; built only where /root/reference exists (oracle/Makefile, target ref),
// output to oracle/_ref/.  No reference source is copied: the two template-only translation
// units are #included by path at compile time, the other two are compiled where they lie.
//   toolbox/rgbConvertMex.cpp  -> rgbConvert<float,float>   (rgbConvertMex.cpp:339-380)
//   toolbox/imResampleMex.cpp  -> resample<float>           (imResampleMex.cpp:125-383)
//   toolbox/convConst.cpp      -> convTri1, convTri         (convConst.cpp:494-525, 347-442)
//   toolbox/gradientMex.cpp    -> gradMag, gradMagNorm, gradHist (gradientMex.cpp:168-275, 375-664)
#include REF_RGBCONVERT_CPP
#include REF_IMRESAMPLE_CPP
#include "acf_oracle.h"

void convTri1(float* I, float* O, int h, int w, int d, float p, int s);
void convTri(float* I, float* O, int h, int w, int d, int r, int s);
void gradMag(float* I, float* M, float* O, int h, int w, int d, bool full);
void gradMagNorm(float* M, float* S, int h, int w, float norm);
void gradHist(float* M, float* O, float* H, int h, int w, int bin, int nOrients, int softBin, bool full);

static void ref_rgbConvert(float* I, float* J, int n, int d, int flag, float nrm) { rgbConvert<float, float>(I, J, n, d, flag, nrm); }
static void ref_resample(float* A, float* B, int ha, int hb, int wa, int wb, int d, float r) { resample<float>(A, B, ha, hb, wa, wb, d, r); }

const OracleL1& oracle_l1()
{
    static const OracleL1 t = { ORACLE_KIND, ref_rgbConvert, convTri1, convTri, gradMag, gradMagNorm, gradHist, ref_resample };
    return t;
}
