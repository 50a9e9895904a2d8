/*
 * acf_oracle.h -- CPU oracle for the chnsPyramid + acfDetect hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or the timed CPU baseline.  The product (acf_b200/) never links
 * or calls it and fails loudly when its CUDA library is missing.
 *
 * Three builds share this C API (see oracle/Makefile):
 *   liboracle_port.so            my scalar, IEEE-exact restatement of the reference's L1
 *                                arithmetic (l1_port.cpp) + restated orchestration (acf_oracle.cpp)
 *   _ref/liboracle_ref_exact.so  the REFERENCE's own toolbox objects compiled from
 *                                /root/reference (rcpps/rsqrtps replaced by IEEE 1/x, 1/sqrt)
 *                                under the same restated orchestration  -> pins the port
 *   _ref/liboracle_ref_native.so the reference toolbox objects exactly as shipped (SSE
 *                                approximations) -> the honest CPU baseline that gets timed
 *
 * Memory convention (same as the reference): every plane is stored
 * "transposed": index = x*h + y with h = original image rows, i.e. contiguous along the
 * original y axis; multi-plane buffers are planes stacked one after another.
 */
#ifndef ACF_ORACLE_H
#define ACF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* mirrors acf::Detector::Options (ACF.h:68-275) restricted to fields that change results */
typedef struct oracle_opts {
    int shrink;            /* pChns.shrink */
    int color_enabled;     /* pChns.pColor.enabled */
    double color_smooth;   /* pChns.pColor.smooth */
    int color_space;       /* 0 gray, 1 rgb, 2 luv, 4 orig (rgbConvert.cpp:109-130) */
    int gm_enabled, gm_colorChn, gm_normRad;
    double gm_normConst;
    int gm_full;
    int gh_enabled, gh_binSize /*0 = unset -> shrink*/, gh_nOrients, gh_softBin;
    int nPerOct, nOctUp, nApprox;
    int nLambdas;          /* 0 = derive from image (chnsPyramid.cpp:341-374) */
    double lambdas[8];
    int pad_w, pad_h;      /* cv::Size as loaded: width <- MATLAB h (orig y), height <- MATLAB w (orig x) */
    int minDs_w, minDs_h;
    double smooth;
    int modelDs_w, modelDs_h, modelDsPad_w, modelDsPad_h;
    int stride;
    double cascThr;
} oracle_opts;

typedef struct oracle_det { int x, y, w, h; double score; } oracle_det;

typedef void (*oracle_tap_fn)(const char* tag, int scale, const float* data, int h, int w, int d, void* user);

const char* oracle_kind(void); /* "port" | "ref_exact" | "ref_native" */

/* Detector::getScales chnsPyramid.cpp:461-529. sz_w = image rows (orig H), sz_h = image cols (orig W)
 * exactly as the reference sees the transposed input. Returns nScales; scaleshw = (w,h) pairs. */
int oracle_get_scales(int nPerOct, int nOctUp, int minDs_w, int minDs_h, int shrink, int sz_w, int sz_h,
                      double* scales, double* scaleshw, int cap);

/* Detector::operator()(cv::Mat) front end + chnsPyramid (ACF.cpp:135-141, chnsPyramid.cpp:160-456).
 * img: HWC u8 RGB (is_f32 = 0) or HWC float RGB in [0,1] (is_f32 = 1), rows x cols. */
void* oracle_pyramid_create(const oracle_opts* o, const void* img, int rows, int cols, int is_f32,
                            oracle_tap_fn tap, void* user);
void oracle_pyramid_destroy(void* pyr);
int oracle_pyramid_nscales(void* pyr);
int oracle_pyramid_ntypes(void* pyr);
/* plane geometry of scale i after concat: nchn planes of w (orig-x extent) x h (orig-y extent), y contiguous */
const float* oracle_pyramid_scale(void* pyr, int i, int* h, int* w, int* nchn, double* scale, double* scalehw_w, double* scalehw_h);
int oracle_pyramid_lambdas(void* pyr, double* out, int cap);

/* Detector::operator()(Pyramid) ACF.cpp:268-367 + acfDetect1.cpp:72-144,231-335.
 * Tree tables are [nTrees x nTreeNodes] row major.  Returns the number of detections written
 * (<= cap); *total receives the full count.  hit_* (optional, may be NULL) receive per-hit
 * (scale, c, r, exit-tree-count) in the same order.  do_nms: bbNms.cpp:229-304 'maxg'|'max'. */
typedef struct oracle_clf {
    int nTrees, nTreeNodes, treeDepth;
    const uint32_t* fids; const float* thrs; const uint32_t* child; const float* hs;
} oracle_clf;
int oracle_detect(void* pyr, const oracle_opts* o, const oracle_clf* clf, oracle_det* out, int cap, int* total,
                  int* hit_scale, int* hit_c, int* hit_r, uint64_t* trees_evaluated);
/* acfDetect1 on a caller-provided channel buffer (nchn planes, w x h, y contiguous). Raw hits (c,r,score) in
 * reference order (c outer, r inner). */
int oracle_acf_detect1(const float* chns, int h, int w, int nchn, const oracle_opts* o, const oracle_clf* clf,
                       int* hit_c, int* hit_r, float* hit_score, int cap, uint64_t* trees_evaluated);
/* the byte-channel detector (ParallelDetectionBody<uint8_t,k>, acfDetect1.cpp:157-191) with Classifier::thrsU8 */
int oracle_acf_detect1_u8(const uint8_t* chns, int h, int w, int nchn, const oracle_opts* o, const oracle_clf* clf,
                          int* hit_c, int* hit_r, float* hit_score, int cap, uint64_t* trees_evaluated);
/* Detector::evaluate(const cv::Mat&) ACF.cpp:123-133: score of the window at (0,0) of chnsCompute(image), cascThr = 0 */
int oracle_evaluate(const oracle_opts* o, const void* img, int rows, int cols, int is_f32, const oracle_clf* clf, float* score);
const char* oracle_last_error(void);
/* bbNms (type "max"/"maxg", ovrDnm union/min) + ObjectDetector::prune. In place; returns new count. */
int oracle_nms(oracle_det* dets, int n, double overlap, int greedy, int ovr_union);
int oracle_prune(oracle_det* dets, int n, int max_count, double prune_ratio);

/* L1 entry points (column-major convention of the toolbox), for stage-level checks */
void oracle_rgb_convert(const float* I, float* J, int n, int flag);               /* planar RGB -> gray(0)/luv(2) */
void oracle_conv_tri1(float* I, float* O, int h, int w, int d, float p, int s);   /* may alias I==O */
void oracle_conv_tri(float* I, float* O, int h, int w, int d, int r, int s);
void oracle_grad_mag(float* I, float* M, float* O, int h, int w, int d, int full);
void oracle_grad_mag_norm(float* M, float* S, int h, int w, float norm);
void oracle_grad_hist(float* M, float* O, float* H, int h, int w, int bin, int nOrients, int softBin, int full);
void oracle_resample(float* A, float* B, int ha, int hb, int wa, int wb, int d, float r);

#ifdef __cplusplus
}
/* L1 function table: filled by l1_port.cpp (port build) or ref_glue.cpp (_ref builds) */
struct OracleL1 {
    const char* kind;
    void (*rgbConvert)(float* I, float* J, int n, int d, int flag, float nrm);
    void (*convTri1)(float* I, float* O, int h, int w, int d, float p, int s);
    void (*convTri)(float* I, float* O, int h, int w, int d, int r, int s);
    void (*gradMag)(float* I, float* M, float* O, int h, int w, int d, bool full);
    void (*gradMagNorm)(float* M, float* S, int h, int w, float norm);
    void (*gradHist)(float* M, float* O, float* H, int h, int w, int bin, int nOrients, int softBin, bool full);
    void (*resample)(float* A, float* B, int ha, int hb, int wa, int wb, int d, float r);
};
const OracleL1& oracle_l1();
#endif
#endif
