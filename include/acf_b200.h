/*
 * acf_b200.h -- C ABI of libacf_b200.so: the B200 (sm_100a) replacement for the reference's
 * chnsPyramid + acfDetect hot path (elucideye/acf).  Plain pointers and sizes only; no C++,
 * OpenCV or torch types.  Every call returns 0 on success, non-zero on error with the message
 * available from acfb_last_error() (thread local).  There is NO CPU fallback: creating an
 * engine without a usable CUDA device fails.
 *
 * Each entry point names the reference interface it stands in for (file:line under
 * /root/reference/src/lib/acf/acf/).  The header-only C++ facade include/acf/ACF.h rebuilds
 * acf::Detector on top of these calls; INTEGRATION.md shows the reference-side binding.
 *
 * Conventions kept from the reference:
 *  - cv::Size fields of a loaded model hold (width <- MATLAB h = extent along original image
 *    rows/y, height <- MATLAB w = extent along original columns/x).  They are passed through
 *    unchanged here as *_w / *_h.
 *  - Channel planes are stored "transposed": element (x, y) of a plane of h rows (orig y) and
 *    w columns (orig x) is at [x*h + y]; the planes of one scale are stacked back to back in the
 *    order colour, gradient magnitude, gradient histogram bins (chnsCompute.cpp:255,306,332).
 *  - Frames are passed UNtransposed: HWC uint8 RGB, rows x cols x 3, the layout callers hold
 *    before Detector::operator()(cv::Mat) transposes it (ACF.cpp:135-141).
 */
#ifndef ACF_B200_H
#define ACF_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ACFB_API __attribute__((visibility("default")))
#else
#define ACFB_API
#endif

typedef struct acfb_model acfb_model;   /* host-side model: Detector::{clf,opts} (ACF.h:68-310) */
typedef struct acfb_engine acfb_engine; /* one per (host thread, device): buffers + plan + streams */

/* The subset of acf::Detector::Options that changes results (ACF.h:68-275), as plain data. */
typedef struct acfb_options {
    int32_t shrink;                       /* pPyramid.pChns.shrink */
    int32_t color_enabled;                /* pChns.pColor.enabled */
    double color_smooth;                  /* pChns.pColor.smooth */
    int32_t color_space;                  /* 0 gray 1 rgb 2 luv 3 hsv 4 orig (rgbConvert.cpp:109-130) */
    int32_t gm_enabled, gm_colorChn, gm_normRad;
    double gm_normConst;
    int32_t gm_full;
    int32_t gh_enabled, gh_binSize /* 0 = unset (use shrink) */, gh_nOrients, gh_softBin, gh_useHog;
    double gh_clipHog;
    int32_t nPerOct, nOctUp, nApprox;
    int32_t nLambdas;                     /* 0: derive from the image (chnsPyramid.cpp:341-374) */
    double lambdas[8];
    int32_t pad_w, pad_h, minDs_w, minDs_h;
    double smooth;
    int32_t concat;
    int32_t modelDs_w, modelDs_h, modelDsPad_w, modelDsPad_h;
    int32_t stride;
    double cascThr, cascCal;
    char nms_type[16];                    /* pNms.type: "max" | "maxg" | "none" (bbNms.cpp:194-225) */
    double nms_overlap;
    char nms_ovrDnm[16];                  /* "union" | "min" */
} acfb_options;

/* Detector::Classifier (ACF.h:292-310): tables are [nTrees x nTreeNodes], row major. */
typedef struct acfb_classifier {
    int32_t nTrees, nTreeNodes, treeDepth;
    const uint32_t* fids;
    const float* thrs;
    const uint32_t* child;
    const float* hs;
    const float* weights;   /* may be NULL */
    const uint32_t* depth;  /* may be NULL */
} acfb_classifier;

/* One detection, 24 bytes: box in original image coordinates (ACF.cpp:302-311), cascade score,
 * frame index inside the batch. */
typedef struct acfb_det { int32_t x, y, w, h; float score; int32_t frame; } acfb_det;

/* One raw cascade hit before rescaling (acfDetect1.cpp:84-98): window column c (orig x / stride),
 * row r, scale index, score. */
typedef struct acfb_hit { int32_t frame, scale, c, r; float score; } acfb_hit;

/* Geometry of one pyramid scale (Detector::Pyramid, ACF.h:364-389). */
typedef struct acfb_scale_info {
    double scale, scalehw_w, scalehw_h;
    int32_t h, w;            /* padded plane dims: h along orig y (contiguous), w along orig x */
    int32_t pitch;           /* device column pitch in floats (>= h); host copies are dense (pitch == h) */
    int32_t nchn;
    int32_t is_real;         /* computed (1) or approximated (0) scale */
    int32_t real_index;      /* scale index of the real scale it derives from */
    int64_t offset;          /* float offset of this scale inside one frame's pyramid block */
} acfb_scale_info;

ACFB_API const char* acfb_last_error(void);
ACFB_API const char* acfb_version(void);

/* ---- model: replaces Detector(const std::string&) / deserializeAny / load_cpb
 *      (ACF.cpp:38-46, ACFIO.cpp:202-217, io/cereal_pba.h:41-54, ACFIOArchive.h:75-216) */
ACFB_API int acfb_model_load(const void* cpb, size_t nbytes, acfb_model** out);
ACFB_API int acfb_model_load_file(const char* path, acfb_model** out);
/* build a model from plain tables (what acf-mat2cpb would hold after reading a .mat; mat2cpb.cpp:75-85) */
ACFB_API int acfb_model_create(const acfb_options* opts, const acfb_classifier* clf, acfb_model** out);
/* save_cpb (io/cereal_pba.h:56-85): writes at most cap bytes, *nbytes = size needed */
ACFB_API int acfb_model_save(const acfb_model* m, void* buf, size_t cap, size_t* nbytes);
ACFB_API int acfb_model_save_file(const acfb_model* m, const char* path);
ACFB_API int acfb_model_options(const acfb_model* m, acfb_options* out);
/* borrowed pointers into the model, valid until acfb_model_destroy */
ACFB_API int acfb_model_classifier(const acfb_model* m, acfb_classifier* out);
/* Detector::acfModify (acfModify.cpp:83-152): cascCal is ADDED to every hs (cumulative, A.2 Q14);
 * pass NaN for cascThr / a negative stride to leave them unchanged. */
ACFB_API int acfb_model_modify(acfb_model* m, double cascCal, double cascThr, int stride);
/* The whole Detector::Modify field set (ACF.h:392-408; merged into the options by acfModify.cpp:99-123): every group is applied only
 * when its has_* flag is set; stride is re-rounded to a multiple of shrink and cascCal ADDED to every hs as above.  `rescale` is not
 * offered: the reference asserts on it (acfModify.cpp:145-149).  The model is re-validated; on failure nothing is changed. */
typedef struct acfb_modify {
    int32_t has_nPerOct, nPerOct, has_nOctUp, nOctUp, has_nApprox, nApprox;
    int32_t has_lambdas, nLambdas;        /* nLambdas = 0: derive them from the image */
    double lambdas[8];
    int32_t has_pad, pad_w, pad_h, has_minDs, minDs_w, minDs_h;
    int32_t has_nms;
    char nms_type[16];
    double nms_overlap;
    char nms_ovrDnm[16];
    int32_t has_stride, stride, has_cascThr;
    double cascThr, cascCal;
} acfb_modify;
ACFB_API int acfb_model_modify_ex(acfb_model* m, const acfb_modify* p);
ACFB_API void acfb_model_destroy(acfb_model* m);

/* ---- engine */
ACFB_API int acfb_engine_create(const acfb_model* m, int device, int max_rows, int max_cols, int max_batch,
                                acfb_engine** out);
ACFB_API void acfb_engine_destroy(acfb_engine* e);
/* ObjectDetector::setDoNonMaximaSuppression / setMaxDetectionCount / setDetectionScorePruneRatio
 * (ObjectDetector.h:37-47) */
ACFB_API int acfb_set_nms(acfb_engine* e, int enable);
ACFB_API int acfb_set_max_detection_count(acfb_engine* e, int n);
ACFB_API int acfb_set_detection_score_prune_ratio(acfb_engine* e, double ratio);
/* layout of the frames handed to every call below (the `frames` pointers are typed uint8_t* for all of them):
 *   0 RGB24 (default, what Detector::operator() expects, ACF.cpp:137-139), 1 BGR24 and 3 BGRA32 (what OpenCV / video
 *   sources hold before the apps' cvtColor, acf.cpp:334-346, pipeline.cpp:319-335), 2 RGBA32,
 *   4 GRAY8 (replicated to three planes like chnsPyramid.cpp:234-244; colorSpace gray or orig only),
 *   5 RGB32F  interleaved float RGB in [0,1], used as it is (the CV_32F branch of ACF.cpp:137),
 *   6 PLANAR32F  three float planes of the TRANSPOSED image, [3][cols][rows] -- the MatP overloads
 *                (ACF.h:423-427, ACF.cpp:161-165); never written (the reference smooths such input in place, A.2 Q13),
 *   7 NV12  rows x cols luma bytes followed by rows/2 x cols interleaved (U, V) bytes (what cameras / decoders deliver and the
 *           reference's GL front end ingests, GPUACF.cpp:438-476): 1.5 bytes per pixel over PCIe.  Converted on the device to
 *           RGB8 by the integer ITU-R BT.601 formula of cv::cvtColor(COLOR_YUV2RGB_NV12) -- bit-identical to running that call
 *           on the host and handing the result over as RGB24; rows and cols must be even. */
ACFB_API int acfb_set_input_format(acfb_engine* e, int format);
/* Detector::setIsTranspose (ACF.h:569-576): interleaved frames are already transposed, i.e. stored [cols][rows][pixel];
 * rows / cols arguments and returned boxes still refer to the upright image (ACF.cpp:310-311). */
ACFB_API int acfb_set_is_transpose(acfb_engine* e, int flag);
/* Detector::setIsLuv (ACF.h:560-567): the three input channels already hold L, u, v (colorSpace must be luv;
 * rgbConvert.cpp:150-155 passes them through). */
ACFB_API int acfb_set_is_luv(acfb_engine* e, int flag);
/* capacity of the per-frame raw-hit buffer on the device (default 4096) */
ACFB_API int acfb_set_hit_capacity(acfb_engine* e, int cap);

/* Detector::getScales (chnsPyramid.cpp:461-529) alone: host only, no engine or device needed.  scaleshw = (w, h) pairs. */
ACFB_API int acfb_get_scales(const acfb_options* opts, int rows, int cols, double* scales, double* scaleshw, int cap,
                             int* nscales);
/* Detector::getScales + the real/approximate split of chnsPyramid (chnsPyramid.cpp:270-292,461-529)
 * for a frame size.  Returns the number of scales; fills at most cap entries. */
ACFB_API int acfb_plan(acfb_engine* e, int rows, int cols, acfb_scale_info* out, int cap, int* nscales,
                       int64_t* floats_per_frame);

/* Detector::computePyramid / chnsPyramid (ACF.cpp:147-159, chnsPyramid.cpp:160-456) for a batch of
 * n frames of identical size.  frames: HWC u8 RGB, n*rows*cols*3 bytes; on_device != 0 means a
 * device pointer on the engine's device (no copy).  The pyramid stays resident on the device. */
ACFB_API int acfb_pyramid(acfb_engine* e, const uint8_t* frames, int n, int rows, int cols, int on_device);
/* device pointer to the resident pyramid block of frame f (floats_per_frame floats) */
ACFB_API int acfb_pyramid_device_ptr(acfb_engine* e, int frame, const float** dptr);
/* copy one scale of one frame to host memory in the reference's Pyramid layout (data[scale][0]) */
ACFB_API int acfb_pyramid_read(acfb_engine* e, int frame, int scale, float* host_out, size_t cap_floats);
/* image-derived lambdas of frame 0 (only meaningful when the model has none) */
ACFB_API int acfb_pyramid_lambdas(acfb_engine* e, double* out, int cap, int* n);

/* Detector::operator()(const Pyramid&) (ACF.cpp:268-367) on the resident pyramid: cascade on the
 * device, box rescale / optional bbNms + prune on the host.  dets: capacity cap, filled frame by
 * frame in the reference's order (scale-major, then window column, then row; with NMS: score
 * descending).  counts[n] receives per-frame counts. */
ACFB_API int acfb_detect_pyramid(acfb_engine* e, acfb_det* dets, int cap, int* counts, int* total);
/* Detector::operator()(const cv::Mat&) for a batch: acfb_pyramid + acfb_detect_pyramid */
ACFB_API int acfb_detect(acfb_engine* e, const uint8_t* frames, int n, int rows, int cols, int on_device,
                         acfb_det* dets, int cap, int* counts, int* total);
/* raw hits of the last detect call, sorted like the reference's per-scale loops; also reports the
 * number of trees evaluated over all windows (for the trees/window figure) */
ACFB_API int acfb_last_hits(acfb_engine* e, acfb_hit* hits, int cap, int* total, uint64_t* trees_evaluated,
                            uint64_t* windows);

/* Detector::acfDetect1 (acfDetect1.cpp:309-335) on caller-provided channels: nchn planes of w x h
 * floats (host memory, reference layout).  Hits come back in the reference's order (c outer, r inner). */
ACFB_API int acfb_acf_detect1(acfb_engine* e, const float* chns, int h, int w, int nchn,
                              int32_t* hit_c, int32_t* hit_r, float* hit_score, int cap, int* total,
                              uint64_t* trees_evaluated);
/* The byte-channel detector, ParallelDetectionBody<uint8_t,k> (acfDetect1.cpp:157-166,187-191): channels are
 * uint8 (what the reference's GPU producer delivers) and are compared with thresholds pre-scaled by 255 exactly as
 * Classifier::thrsU8 is built (ACFIOArchive.h:96-99, ACF.h:305).  Scores are the same float sums. */
ACFB_API int acfb_acf_detect1_u8(acfb_engine* e, const uint8_t* chns, int h, int w, int nchn,
                                 int32_t* hit_c, int32_t* hit_r, float* hit_score, int cap, int* total,
                                 uint64_t* trees_evaluated);
/* Detector::operator()(const Pyramid&) (ACF.cpp:268-367) on a pyramid some OTHER producer filled -- the seam the
 * reference's GL pipeline uses (GLDetector.cpp:124, GPUACF fills Detector::Pyramid, ACF.h:364-389).  One entry per
 * scale: Pyramid::data[i][0] after concat (nchn planes of w columns x h values, planes stacked; float, or uint8 when
 * is_u8), Pyramid::scales[i], Pyramid::scaleshw[i].  Boxes are rescaled, then NMS / prune as configured. */
typedef struct acfb_channels
{
    const void* data; /* host memory */
    int32_t h, w, nchn;
    double scale, scalehw_w, scalehw_h;
} acfb_channels;
ACFB_API int acfb_detect_channels(acfb_engine* e, const acfb_channels* scales, int nscales, int is_u8,
                                  acfb_det* out, int cap, int* total);
/* Detector::evaluate(const cv::Mat&) (ACF.cpp:123-133): score of the single window at (0,0) of
 * chnsCompute(frame), no pyramid */
ACFB_API int acfb_evaluate(acfb_engine* e, const uint8_t* frame, int rows, int cols, float* score);

/* Detector::computeChannels(const cv::Mat&, MatP&) (ACF.h:419-420, ACF.cpp:164-240): chnsCompute of the frame with the
 * reference's fixed default channel options, fused into one plane stack: out = d planes of w x h floats ([z][x][y], the
 * reference's transposed layout), d = 10 (L, U, V, M, 6 orientation bins), w = cols / 4, h = rows / 4.  Served for models
 * whose channel options equal those defaults (the engine is built for the model's options).  out == NULL queries the size. */
ACFB_API int acfb_compute_channels(acfb_engine* e, const uint8_t* frame, int rows, int cols, float* out, size_t cap_floats,
                                   int* d, int* w, int* h);

/* ---- stand-alone channel operators: the reference's static Detector:: functions on ONE image (ACF.h:416-490), on the
 * GPU with the reference's arithmetic (bit-identical to its exact-math build).  Host pointers; all planes are float in
 * the reference's transposed planar layout (MatP of the transposed image): element (x, y) of plane z at
 * [z*w*h + x*h + y], h = contiguous extent (MatP::cols, the original image's rows), w = MatP::rows. */
/* Detector::rgbConvert(I, J, cs, useSingle=true) (ACF.h:443-450, rgbConvert.cpp:102-170; toolbox rgbConvertMex.cpp:382-423):
 * I = 3 planes RGB in [0,1]; colorspace 0 gray 1 rgb 2 luv 3 hsv 4 orig; J gets *nplanes_out (1 or 3) planes */
ACFB_API int acfb_op_rgb_convert(acfb_engine* e, const float* I, int h, int w, int colorspace, float* J, int* nplanes_out);
/* Detector::convTri(I, J, r, s=1) (ACF.h:464, convTri.cpp:204-253 -> convConst.cpp:494-525 for r <= 1, :347-442 otherwise).
 * J == I selects the reference's IN-PLACE behaviour (chnsCompute.cpp:239: for r <= 1 column x is then filtered from the
 * already filtered column x-1); distinct buffers give the plain filter.  r <= 1 needs h % 4 == 0. */
ACFB_API int acfb_op_conv_tri(acfb_engine* e, const float* I, int h, int w, int d, double r, float* J);
/* Detector::gradientMag(I, M, O, channel, normRad, normConst, full) (ACF.h:467-478, gradientMag.cpp:109-135 ->
 * gradMag gradientMex.cpp:168-251, convTri, gradMagNorm :254-275).  I has d planes, plane `channel` is used; O may be
 * NULL.  Needs h % 4 == 0. */
ACFB_API int acfb_op_gradient_mag(acfb_engine* e, const float* I, int h, int w, int d, int channel, int normRad, double normConst,
                                  int full, float* M, float* O);
/* Detector::gradientHist(M, O, H, binSize, nOrients, softBin, useHog, clipHog, full) (ACF.h:480-491,
 * gradientHist.cpp:92-114 -> gradHist gradientMex.cpp:375-664).  H = nOrients planes of (h/bin) x (w/bin).  binSize 4 and
 * softBin 0 (the shipped models' values); useHog / clipHog are ignored exactly as the reference ignores them. */
ACFB_API int acfb_op_gradient_hist(acfb_engine* e, const float* M, const float* O, int h, int w, int binSize, int nOrients, int softBin,
                                   int useHog, double clipHog, int full, float* H);
/* imResample(A, B, size, nrm) (ACF.h:676, imResampleMex.cpp:385-420 -> resample<float> :125-383): d planes ha x wa -> hb x wb,
 * multiplied by nrm.  At most 12 taps per axis (down-sampling by up to ~10x). */
ACFB_API int acfb_op_im_resample(acfb_engine* e, const float* A, int ha, int wa, int d, int hb, int wb, double nrm, float* B);

/* ---- asynchronous / benchmark surface.  acfb_submit enqueues pyramid + cascade (+ ordering / rescale / bbNms / prune when they
 * run on the device) for n frames and returns; acfb_collect waits for the OLDEST submitted batch and returns its boxes (results
 * come back in submission order).  Batches submitted while others are in flight go round up to three pipelines -- complete
 * buffer / stream sets on the device, created on demand (ACFB_PIPELINES=1|2 in the environment keep fewer) -- so the kernels of
 * consecutive batches overlap; at most two batches per pipeline may be in flight (the next acfb_submit fails until one is
 * collected).  Device-resident frames make the timed region kernel-only. */
ACFB_API int acfb_submit(acfb_engine* e, const uint8_t* frames, int n, int rows, int cols, int on_device);
ACFB_API int acfb_collect(acfb_engine* e, acfb_det* dets, int cap, int* counts, int* total);
ACFB_API int acfb_synchronize(acfb_engine* e);
/* diagnostic: runs n random normal-range inputs through the kernels' reciprocal / square-root sequences and counts the
 * results that differ from the IEEE operators (must be 0) */
ACFB_API int acfb_selftest_math(acfb_engine* e, uint64_t n, uint32_t seed, uint64_t* mismatches);
/* number of kernel launches issued by this engine since creation (bench.py's gpu_launches claim) */
ACFB_API uint64_t acfb_launch_count(acfb_engine* e);
/* cudaStream_t of the engine as an integer, so callers can record CUDA events on it */
ACFB_API uint64_t acfb_stream(acfb_engine* e);
/* per-stage device time of the last submit in milliseconds (CUDA events on the engine's stream):
 * names[i] is a static string; returns the number of stages */
ACFB_API int acfb_stage_times(acfb_engine* e, const char** names, float* ms, int cap);
ACFB_API int acfb_enable_stage_timing(acfb_engine* e, int enable);
/* host-side wall time of the last acfb_collect in milliseconds: wait_ms = blocked until the batch's kernels and the hit
 * read-back were done, tail_ms = ordering the hits like the reference's loops, box rescale (ACF.cpp:302-311), bbNms / prune
 * (the per-stage wall-clock logging of src/app/common/ScopeTimeLogger.h, for the host half of the path) */
ACFB_API int acfb_collect_times(acfb_engine* e, double* wait_ms, double* tail_ms);

/* ---- multi-GPU (the reference's frame parallelism, src/app/acf/acf.cpp:443-455, across devices).  Frames are
 * independent (ACF.cpp:249), so a batch shards contiguously over the ranks -- rank r runs frames [r*n, (r+1)*n) of the global
 * batch on its own engine / device -- with no collective on the data path.  The one exchange is the gather of the boxes that
 * bbNms + prune leave (per-frame counts + 24-byte records, at most 64 per frame: tens of KB per rank and batch):
 * acfb_dist_collect is acfb_collect on every rank plus that gather; rank 0 receives the boxes of ALL ranks in global frame order
 * (frame = r*n + local frame), the other ranks nothing (*total = 0).  Two exchanges:
 *   NCCL (default whenever libnccl.so.2 can be loaded at run time; also across nodes): ncclAllGather of the device buffer k_post
 *       wrote (fixed-capacity records) on a communication stream, one CTA (ncclConfig_t.maxCTAs);
 *   shared memory (ACFB_DIST_EXCHANGE=shm in the environment of every rank, and the fallback without NCCL; one box): a ring of
 *       per-rank slots in a POSIX shared-memory segment named after the 128-byte id, two atomics per slot; nothing on the device.
 *   Both cost nothing measurable per step (N = 8: 17.10 / 17.06 ms against 16.85 ms at N = 1).
 * One engine per device:
 *   process per GPU : rank 0 calls acfb_dist_unique_id, the host's own plumbing broadcasts the 128 bytes, every rank calls
 *                     acfb_dist_init_rank;
 *   one process     : acfb_dist_init_all(engines, n), then ONE HOST THREAD PER ENGINE for submit / collect.
 * Needs setDoNonMaximaSuppression(true) with maxDetectionCount <= 64, and the same options, the same n and the same sequence of
 * batches on every rank. */
ACFB_API int acfb_dist_unique_id(uint8_t id[128]);
ACFB_API int acfb_dist_init_rank(acfb_engine* e, const uint8_t id[128], int rank, int world);
ACFB_API int acfb_dist_init_all(acfb_engine** engines, int n);
/* counts: [world * n] on rank 0 (may be NULL elsewhere) */
ACFB_API int acfb_dist_collect(acfb_engine* e, acfb_det* dets, int cap, int* counts, int* total);
/* host-only self test of the shared-memory exchange (threads as ranks, ring wrap-around, flow control); no device needed */
ACFB_API int acfb_selftest_exchange(int world, int batches, int slot_bytes);
/* world = 0: no communicator; nccl_version as ncclGetVersion reports it */
ACFB_API int acfb_dist_info(acfb_engine* e, int* rank, int* world, int* nccl_version);

/* ---- debug taps (the reference's MatLoggerType hook, ACF.h:57,578-581): copy an intermediate
 * plane set of frame f at real scale index k to host.  tag: "I" converted image, "C" smoothed
 * image, "R" real-scale channels before the final smoothing.  dims returned as (d, w, h). */
/* the smoothed image ("C") is an on-chip intermediate of the fused march k_front; it is written to memory for every real
 * scale only after acfb_set_debug_taps(e, 1) (taps "I" and "R" are always available) */
ACFB_API int acfb_set_debug_taps(acfb_engine* e, int enable);
ACFB_API int acfb_tap(acfb_engine* e, const char* tag, int frame, int real_k, float* out, size_t cap_floats,
                      int* d, int* w, int* h);

#ifdef __cplusplus
}
#endif
#endif
