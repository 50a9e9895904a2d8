// include/acf/ACF.h -- header-only C++ facade: acf::Detector / acf::Detector::Pyramid over libacf_b200.so.
//
// Same class and method names, argument meaning and error behaviour as the reference's public header
// (src/lib/acf/acf/ACF.h:50-624, ObjectDetector.h:31-48, MatP.h:25-190) for the chnsPyramid + acfDetect
// hot path, so a caller of the reference (acf-detect, GPUDetectionPipeline, drishti) recompiles against
// this header and links libacf_b200.so instead of libacf.  Everything below the interface runs on the
// B200 through the C ABI in include/acf_b200.h; nothing here computes on the CPU except the
// rescale / NMS tail the library already performs.
//
// OpenCV is optional: with ACF_B200_WITH_OPENCV defined (and <opencv2/core.hpp> available) the facade
// accepts cv::Mat / returns cv::Rect exactly like the reference; without it the minimal stand-ins below
// (acf::cv::Mat view, Rect, Size) keep the same member names.
//
// Differences a maintainer must know (also listed in INTEGRATION.md):
//  * computePyramid returns channel planes copied back from the device in the reference's layout
//    (Pyramid::data[scale][0], planes stacked, transposed, float);
//  * images are never modified: the reference smooths a float MatP input in place, this engine
//    works on its own device copy; CV_8UC3 / CV_8UC1 / CV_32FC3 images and planar float MatP input are all ingested
//    by the device directly (acfb_set_input_format), transposed or not (setIsTranspose);
//  * .mat models are not supported (ACFIO.cpp:202-232 needs cvmatio): use acf-mat2cpb output (.cpb).
#ifndef ACF_B200_ACF_H
#define ACF_B200_ACF_H

#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <istream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

#include "../acf_b200.h"

#ifdef ACF_B200_WITH_OPENCV
#include <opencv2/core.hpp>
#define ACF_CV ::cv
#else
namespace acf { namespace cv {
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Size2d { double width = 0, height = 0; Size2d() {} Size2d(double w, double h) : width(w), height(h) {} };
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
// non-owning view of an image: rows x cols x channels, 8-bit unsigned (depth 0) or float (depth 5), row step in bytes
struct Mat
{
    const void* data = nullptr; int rows = 0, cols = 0, chans = 0, dep = 0; size_t step = 0;
    Mat() {}
    Mat(int r, int c, int ch, int depth_, const void* p, size_t step_ = 0)
        : data(p), rows(r), cols(c), chans(ch), dep(depth_), step(step_ ? step_ : (size_t)c * ch * (depth_ == 0 ? 1 : 4)) {}
    bool empty() const { return !data || rows == 0 || cols == 0; }
    int channels() const { return chans; }
    int depth() const { return dep; }
};
}} // namespace acf::cv
#define ACF_CV ::acf::cv
#endif

namespace acf
{

// MatP.h:25-190 -- planar image: channel planes of identical size stored back to back
class MatP
{
public:
    MatP() {}
    MatP(int rows, int cols, int channels) { create(rows, cols, channels); }
    void create(int rows, int cols, int channels) { m_rows = rows; m_cols = cols; m_channels = channels; m_data.assign((size_t)rows * cols * channels, 0.f); }
    int rows() const { return m_rows; }
    int cols() const { return m_cols; }
    int channels() const { return m_channels; }
    bool empty() const { return m_data.empty(); }
    float* ptr(int plane = 0) { return m_data.data() + (size_t)plane * m_rows * m_cols; }
    const float* ptr(int plane = 0) const { return m_data.data() + (size_t)plane * m_rows * m_cols; }
    std::vector<float>& base() { return m_data; }
private:
    int m_rows = 0, m_cols = 0, m_channels = 0;
    std::vector<float> m_data;
};

// ObjectDetector.h:31-48
class ObjectDetector
{
public:
    virtual ~ObjectDetector() {}
    virtual int operator()(const ACF_CV::Mat& image, std::vector<ACF_CV::Rect>& objects, std::vector<double>* scores = nullptr) = 0;
    virtual void setDoNonMaximaSuppression(bool flag) { m_doNms = flag; }
    virtual bool getDoNonMaximaSuppression() const { return m_doNms; }
    virtual void setMaxDetectionCount(size_t maxCount) { m_maxDetectionCount = maxCount; }
    virtual void setDetectionScorePruneRatio(double ratio) { m_detectionScorePruneRatio = ratio; }
    virtual ACF_CV::Size getWindowSize() const = 0;
protected:
    bool m_doNms = false;
    double m_detectionScorePruneRatio = 0.0;
    size_t m_maxDetectionCount = 10;
};

class Detector : public ObjectDetector
{
public:
    using RectVec = std::vector<ACF_CV::Rect>;
    using RealVec = std::vector<double>;
    using Size2dVec = std::vector<ACF_CV::Size2d>;

    // ACF.h:364-389
    struct Pyramid
    {
        int nTypes = 0, nScales = 0;
        std::vector<std::vector<MatP>> data; // [scale][0] after concat: planes (w x h, transposed) stacked
        std::vector<double> lambdas, scales;
        Size2dVec scaleshw;
        void clear() { data.clear(); lambdas.clear(); scales.clear(); scaleshw.clear(); }
    };
    struct Detection { ACF_CV::Rect roi; double score = 0; };
    // ACF.h:392-408: every field optional (negative / NaN / empty = leave as it is), merged like acfModify.cpp:99-123
    struct Modify
    {
        double cascThr = std::nan(""); double cascCal = 0.0; int stride = -1;
        int nPerOct = -1, nOctUp = -1, nApprox = -2;   // nApprox = -1 is a legal value of the reference (= nPerOct - 1)
        bool hasLambdas = false; std::vector<double> lambdas;
        int padWidth = -1, padHeight = -1, minDsWidth = -1, minDsHeight = -1;
        std::string nmsType; double nmsOverlap = 0.65; std::string nmsOvrDnm = "min";
    };

    Detector() {}
    // Detector(const std::string&) ACF.cpp:43-46 ; success reported through good() (ACF.h:65-66)
    explicit Detector(const std::string& filename, int device = 0, int maxRows = 2160, int maxCols = 3840, int maxBatch = 1)
    {
        m_good = acfb_model_load_file(filename.c_str(), &m_model) == 0 && init(device, maxRows, maxCols, maxBatch);
        if (!m_good && m_error.empty()) m_error = acfb_last_error();
    }
    // Detector(std::istream&, hint) ACF.cpp:38-41
    explicit Detector(std::istream& is, const std::string& hint = {}, int device = 0, int maxRows = 2160, int maxCols = 3840, int maxBatch = 1)
    {
        (void)hint;
        std::vector<char> buf((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
        m_good = acfb_model_load(buf.data(), buf.size(), &m_model) == 0 && init(device, maxRows, maxCols, maxBatch);
        if (!m_good && m_error.empty()) m_error = acfb_last_error();
    }
    Detector(const Detector&) = delete;
    Detector& operator=(const Detector&) = delete;
    ~Detector() override
    {
        if (m_engine) acfb_engine_destroy(m_engine);
        if (m_model) acfb_model_destroy(m_model);
    }

    bool good() const { return m_good; }
    explicit operator bool() const { return m_good; }
    const std::string& error() const { return m_error; }

    ACF_CV::Size getWindowSize() const override { return ACF_CV::Size(m_opts.modelDs_w, m_opts.modelDs_h); } // ACF.h:410-413
    void setDoNonMaximaSuppression(bool flag) override { m_doNms = flag; check(acfb_set_nms(m_engine, flag)); }
    void setMaxDetectionCount(size_t n) override { m_maxDetectionCount = n; check(acfb_set_max_detection_count(m_engine, (int)n)); }
    void setDetectionScorePruneRatio(double r) override { m_detectionScorePruneRatio = r; check(acfb_set_detection_score_prune_ratio(m_engine, r)); }
    void setIsTranspose(bool flag) { m_isTranspose = flag; } // ACF.h:569-576
    bool getIsTranspose() const { return m_isTranspose; }
    void setIsLuv(bool flag) { check(acfb_set_is_luv(m_engine, flag)); m_isLuv = flag; } // ACF.h:560-567
    bool getIsLuv() const { return m_isLuv; }

    // Detector::operator()(const cv::Mat&, RectVec&, RealVec*) ACF.cpp:135-141: RGB u8 image, returns 0, appends boxes
    int operator()(const ACF_CV::Mat& I, RectVec& objects, RealVec* scores = nullptr) override
    {
        std::vector<uint8_t> packed;
        int rows = 0, cols = 0;
        const uint8_t* p = pack(I, packed, rows, cols);
        return detectOne(p, rows, cols, objects, scores);
    }
    // Detector::operator()(const MatP&, RectVec&, RealVec*) ACF.h:423-427: planar float planes of the TRANSPOSED image
    int operator()(const MatP& Ip, RectVec& objects, RealVec* scores = nullptr)
    {
        int rows = 0, cols = 0;
        const uint8_t* p = packPlanar(Ip, rows, cols);
        return detectOne(p, rows, cols, objects, scores);
    }
    // Detector::evaluate(const cv::Mat&) ACF.cpp:123-133: score of the window at (0,0) of the image's channels
    float evaluate(const ACF_CV::Mat& I)
    {
        std::vector<uint8_t> packed;
        int rows = 0, cols = 0;
        const uint8_t* p = pack(I, packed, rows, cols);
        float score = 0;
        check(acfb_evaluate(m_engine, p, rows, cols, &score));
        return score;
    }
    // batch form: frames[i] are images of identical size; results per frame
    int operator()(const std::vector<ACF_CV::Mat>& frames, std::vector<RectVec>& objects, std::vector<RealVec>* scores = nullptr)
    {
        if (frames.empty()) return 0;
        std::vector<uint8_t> all, one;
        int rows = 0, cols = 0;
        for (const auto& f : frames)
        {
            int r, c;
            const int fmtBefore = m_fmt;
            const uint8_t* p = pack(f, one, r, c);
            if (rows && (r != rows || c != cols || m_fmt != fmtBefore)) throw std::runtime_error("acf::Detector: frames of a batch must share one size and type");
            rows = r; cols = c;
            all.insert(all.end(), p, p + (size_t)r * c * bytesPerPixel());
        }
        std::vector<acfb_det> dets(m_cap * frames.size());
        std::vector<int> counts(frames.size());
        int total = 0;
        check(acfb_detect(m_engine, all.data(), (int)frames.size(), rows, cols, 0, dets.data(), (int)dets.size(), counts.data(), &total));
        if (total > (int)dets.size()) throw std::runtime_error("acf::Detector: detection buffer too small");
        objects.assign(frames.size(), {});
        if (scores) scores->assign(frames.size(), {});
        size_t k = 0;
        for (size_t f = 0; f < frames.size(); f++)
            for (int j = 0; j < counts[f]; j++, k++)
            {
                objects[f].push_back(ACF_CV::Rect(dets[k].x, dets[k].y, dets[k].w, dets[k].h));
                if (scores) (*scores)[f].push_back(dets[k].score);
            }
        return 0;
    }
    // Detector::operator()(const Pyramid&) ACF.cpp:268-367: detection on the channels P holds, whoever filled them
    // (computePyramid, or another producer as in GLDetector.cpp:124).  The planes are uploaded; to run on the pyramid that
    // is still resident on the device after computePyramid, without a copy, use detectResident().
    int operator()(const Pyramid& P, RectVec& objects, RealVec* scores = nullptr)
    {
        const int nchn = (m_opts.color_enabled ? (m_opts.color_space == 0 ? 1 : 3) : 0) + 1 + m_opts.gh_nOrients;
        std::vector<acfb_channels> sc;
        for (int i = 0; i < P.nScales; i++)
        {
            if (P.data[i].empty() || P.data[i][0].empty()) continue; // scales a producer skipped (ACF.cpp:283-287)
            const MatP& m = P.data[i][0];
            if (m.rows() % nchn) throw std::runtime_error("acf::Detector: pyramid planes do not match the model's channel count");
            acfb_channels c{};
            c.data = m.ptr(); c.h = m.cols(); c.w = m.rows() / nchn; c.nchn = nchn;
            c.scale = P.scales[i]; c.scalehw_w = P.scaleshw[i].width; c.scalehw_h = P.scaleshw[i].height;
            sc.push_back(c);
        }
        std::vector<acfb_det> dets(m_cap);
        int total = 0;
        check(acfb_detect_channels(m_engine, sc.data(), (int)sc.size(), 0, dets.data(), (int)dets.size(), &total));
        if (total > (int)dets.size())
        {
            dets.resize(total);
            check(acfb_detect_channels(m_engine, sc.data(), (int)sc.size(), 0, dets.data(), (int)dets.size(), &total));
        }
        append(dets, total, objects, scores);
        return 0;
    }
    // the same on the pyramid left on the device by the last computePyramid (no upload)
    int detectResident(RectVec& objects, RealVec* scores = nullptr)
    {
        std::vector<acfb_det> dets(m_cap);
        int count = 0, total = 0;
        check(acfb_detect_pyramid(m_engine, dets.data(), (int)dets.size(), &count, &total));
        append(dets, count, objects, scores);
        return 0;
    }

    // Detector::computePyramid ACF.cpp:147-159
    void computePyramid(const ACF_CV::Mat& I, Pyramid& P)
    {
        std::vector<uint8_t> packed;
        int rows = 0, cols = 0;
        const uint8_t* p = pack(I, packed, rows, cols);
        pyramidOf(p, rows, cols, P);
    }
    // Detector::computePyramid(const MatP&, Pyramid&) ACF.cpp:161-165
    void computePyramid(const MatP& Ip, Pyramid& P)
    {
        int rows = 0, cols = 0;
        const uint8_t* p = packPlanar(Ip, rows, cols);
        pyramidOf(p, rows, cols, P);
    }

    // Detector::acfModify acfModify.cpp:83-152 (cascCal cumulative, stride re-rounded); rebuilds the engine tables
    int acfModify(const Modify& p)
    {
        acfb_modify q{};
        if (p.nPerOct >= 0) { q.has_nPerOct = 1; q.nPerOct = p.nPerOct; }
        if (p.nOctUp >= 0) { q.has_nOctUp = 1; q.nOctUp = p.nOctUp; }
        if (p.nApprox >= -1) { q.has_nApprox = 1; q.nApprox = p.nApprox; }
        if (p.hasLambdas)
        {
            if (p.lambdas.size() > 8) throw std::runtime_error("acfModify: at most 8 lambdas");
            q.has_lambdas = 1; q.nLambdas = (int)p.lambdas.size();
            for (size_t i = 0; i < p.lambdas.size(); i++) q.lambdas[i] = p.lambdas[i];
        }
        if (p.padWidth >= 0 && p.padHeight >= 0) { q.has_pad = 1; q.pad_w = p.padWidth; q.pad_h = p.padHeight; }
        if (p.minDsWidth >= 0 && p.minDsHeight >= 0) { q.has_minDs = 1; q.minDs_w = p.minDsWidth; q.minDs_h = p.minDsHeight; }
        if (!p.nmsType.empty())
        {
            q.has_nms = 1; q.nms_overlap = p.nmsOverlap;
            std::snprintf(q.nms_type, sizeof(q.nms_type), "%s", p.nmsType.c_str());
            std::snprintf(q.nms_ovrDnm, sizeof(q.nms_ovrDnm), "%s", p.nmsOvrDnm.c_str());
        }
        if (p.stride > 0) { q.has_stride = 1; q.stride = p.stride; }
        if (!std::isnan(p.cascThr)) { q.has_cascThr = 1; q.cascThr = p.cascThr; }
        q.cascCal = p.cascCal;
        check(acfb_model_modify_ex(m_model, &q));
        acfb_engine_destroy(m_engine); m_engine = nullptr;
        if (!init(m_device, m_maxRows, m_maxCols, m_maxBatch)) throw std::runtime_error(m_error);
        return 0;
    }

    // Detector::getScales chnsPyramid.cpp:461-529 for a frame size (rows x cols)
    int getScales(int rows, int cols, RealVec& scales, Size2dVec& scaleshw)
    {
        int n = 0; int64_t fl = 0;
        check(acfb_plan(m_engine, rows, cols, nullptr, 0, &n, &fl));
        std::vector<acfb_scale_info> info(n);
        check(acfb_plan(m_engine, rows, cols, info.data(), n, &n, &fl));
        scales.clear(); scaleshw.clear();
        for (auto& s : info) { scales.push_back(s.scale); scaleshw.push_back(ACF_CV::Size2d(s.scalehw_w, s.scalehw_h)); }
        return 0;
    }

    // Detector::computeChannels(const cv::Mat&, MatP&) ACF.cpp:164-240: the ten default channels of I at 1/shrink resolution
    void computeChannels(const ACF_CV::Mat& I, MatP& Ip2)
    {
        std::vector<uint8_t> packed;
        int rows = 0, cols = 0, d = 0, w = 0, h = 0;
        const uint8_t* p = pack(I, packed, rows, cols);
        check(acfb_compute_channels(m_engine, p, rows, cols, nullptr, 0, &d, &w, &h));
        Ip2.create(w, h, d);
        check(acfb_compute_channels(m_engine, p, rows, cols, Ip2.ptr(), (size_t)d * w * h, &d, &w, &h));
    }

    // ---- the reference's static channel operators (ACF.h:416-490, 676), here members because they run on this
    //      detector's GPU engine.  MatP planes follow the reference: rows() x cols() per plane, cols() contiguous
    //      (for the transposed image the library works on: cols() = original rows).
    // Detector::rgbConvert rgbConvert.cpp:102-170 (useSingle must be true: planes are float; isLuv = input already LUV)
    int rgbConvert(const MatP& I, MatP& J, const std::string& cs, bool useSingle = true, bool isLuv = false)
    {
        if (!useSingle) throw std::runtime_error("rgbConvert: float planes only");
        int code = cs == "gray" ? 0 : cs == "rgb" ? 1 : cs == "luv" ? 2 : cs == "hsv" ? 3 : cs == "orig" ? 4 : -1;
        if (code < 0) throw std::runtime_error("rgbConvert: unknown colour space " + cs);
        if (isLuv && code == 2) code = 1; // rgbConvert.cpp:150-155: already converted, passed through
        if (I.channels() != 3) throw std::runtime_error("rgbConvert: three input planes expected");
        MatP out(I.rows(), I.cols(), 3);
        int np = 0;
        check(acfb_op_rgb_convert(m_engine, I.ptr(), I.cols(), I.rows(), code, out.ptr(), &np));
        J.create(I.rows(), I.cols(), np);
        std::copy(out.ptr(), out.ptr() + (size_t)np * I.rows() * I.cols(), J.ptr());
        return 0;
    }
    // Detector::convTri convTri.cpp:204-253; &J == &I is the reference's in-place call (chnsCompute.cpp:239)
    int convTri(const MatP& I, MatP& J, double r = 1.0, int s = 1)
    {
        if (s != 1) throw std::runtime_error("convTri: s == 1 only");
        if (&J != &I) J.create(I.rows(), I.cols(), I.channels());
        check(acfb_op_conv_tri(m_engine, I.ptr(), I.cols(), I.rows(), I.channels(), r, J.ptr()));
        return 0;
    }
    // Detector::gradientMag gradientMag.cpp:109-135
    int gradientMag(const MatP& I, MatP& M, MatP& O, int channel = 0, int normRad = 0, double normConst = 0.005, int full = 0)
    {
        if (I.empty()) return 0;
        M.create(I.rows(), I.cols(), 1); O.create(I.rows(), I.cols(), 1);
        check(acfb_op_gradient_mag(m_engine, I.ptr(), I.cols(), I.rows(), I.channels(), channel, normRad, normConst, full, M.ptr(), O.ptr()));
        return 0;
    }
    // Detector::gradientHist gradientHist.cpp:109-114 (returns 1 like the reference)
    int gradientHist(const MatP& M, const MatP& O, MatP& H, int binSize, int nOrients, int softBin, int useHog, double clipHog, int full)
    {
        H.create(M.rows() / std::max(1, binSize), M.cols() / std::max(1, binSize), nOrients);
        check(acfb_op_gradient_hist(m_engine, M.ptr(), O.ptr(), M.cols(), M.rows(), binSize, nOrients, softBin, useHog, clipHog, full, H.ptr()));
        return 1;
    }
    // imResample(A, B, size, nrm) imResampleMex.cpp:385-420; size = (rows, cols) of a plane of B
    void imResample(const MatP& A, MatP& B, int rows, int cols, double nrm = 1.0)
    {
        B.create(rows, cols, A.channels());
        check(acfb_op_im_resample(m_engine, A.ptr(), A.cols(), A.rows(), A.channels(), cols, rows, nrm, B.ptr()));
    }

    acfb_engine* engine() { return m_engine; }
    const acfb_options& options() const { return m_opts; }

private:
    int detectOne(const uint8_t* p, int rows, int cols, RectVec& objects, RealVec* scores)
    {
        std::vector<acfb_det> dets(m_cap);
        int count = 0, total = 0;
        check(acfb_detect(m_engine, p, 1, rows, cols, 0, dets.data(), (int)dets.size(), &count, &total));
        if (total > (int)dets.size())
        {
            dets.resize(total);
            check(acfb_detect(m_engine, p, 1, rows, cols, 0, dets.data(), (int)dets.size(), &count, &total));
        }
        append(dets, count, objects, scores);
        return 0;
    }
    void pyramidOf(const uint8_t* p, int rows, int cols, Pyramid& P)
    {
        check(acfb_pyramid(m_engine, p, 1, rows, cols, 0));
        int n = 0; int64_t fl = 0;
        check(acfb_plan(m_engine, rows, cols, nullptr, 0, &n, &fl));
        std::vector<acfb_scale_info> info(n);
        check(acfb_plan(m_engine, rows, cols, info.data(), n, &n, &fl));
        P.clear();
        P.nScales = n; P.nTypes = (m_opts.color_enabled ? 1 : 0) + 2;
        P.data.resize(n);
        for (int i = 0; i < n; i++)
        {
            MatP m(info[i].nchn * info[i].w, info[i].h, 1); // planes stacked vertically (fuseChannels, ACF.h:653-672)
            check(acfb_pyramid_read(m_engine, 0, i, m.ptr(), m.base().size()));
            P.data[i].push_back(std::move(m));
            P.scales.push_back(info[i].scale);
            P.scaleshw.push_back(ACF_CV::Size2d(info[i].scalehw_w, info[i].scalehw_h));
        }
        double lam[8]; int nl = 0;
        check(acfb_pyramid_lambdas(m_engine, lam, 8, &nl));
        P.lambdas.assign(lam, lam + nl);
    }

    bool init(int device, int maxRows, int maxCols, int maxBatch)
    {
        m_device = device; m_maxRows = maxRows; m_maxCols = maxCols; m_maxBatch = maxBatch;
        if (!m_model) { m_error = acfb_last_error(); return false; }
        if (acfb_model_options(m_model, &m_opts) != 0 || acfb_engine_create(m_model, device, maxRows, maxCols, maxBatch, &m_engine) != 0)
        {
            m_error = acfb_last_error();
            return false;
        }
        acfb_set_nms(m_engine, m_doNms);
        acfb_set_max_detection_count(m_engine, (int)m_maxDetectionCount);
        acfb_set_detection_score_prune_ratio(m_engine, m_detectionScorePruneRatio);
        m_fmt = 0; m_engineTransposed = false; // a fresh engine starts with upright RGB24 frames
        if (m_isLuv) acfb_set_is_luv(m_engine, 1);
        return true;
    }
    static void check(int rc) { if (rc != 0) throw std::runtime_error(acfb_last_error()); } // CV_Assert -> exception in the reference
    int bytesPerPixel() const { return m_fmt == 0 ? 3 : m_fmt == 4 ? 1 : 12; }
    void useFormat(int fmt, bool transposed)
    {
        if (fmt != m_fmt) { check(acfb_set_input_format(m_engine, fmt)); m_fmt = fmt; }
        if (transposed != m_engineTransposed) { check(acfb_set_is_transpose(m_engine, transposed)); m_engineTransposed = transposed; }
    }
    // Dense view of an interleaved image for the engine: CV_8UC3 (RGB24), CV_8UC1 (GRAY8, chnsPyramid.cpp:234-244) or
    // CV_32FC3 (ACF.cpp:137); a transposed image (setIsTranspose) is passed through as it is and transposed on the device.
    const uint8_t* pack(const ACF_CV::Mat& I, std::vector<uint8_t>& tmp, int& rows, int& cols)
    {
        if (I.empty()) throw std::runtime_error("acf::Detector: empty image");
        const int ch = I.channels(), dep = I.depth();
        int fmt;
        if (dep == 0 && ch == 3) fmt = 0;
        else if (dep == 0 && ch == 1) fmt = 4;
        else if (dep == 5 && ch == 3) fmt = 5;
        else throw std::runtime_error("acf::Detector: expected CV_8UC3, CV_8UC1 or CV_32FC3");
        useFormat(fmt, m_isTranspose);
        rows = m_isTranspose ? I.cols : I.rows; cols = m_isTranspose ? I.rows : I.cols; // of the upright image
        const size_t line = (size_t)I.cols * bytesPerPixel(), step = (size_t)I.step;
        if (step == line) return (const uint8_t*)I.data;
        tmp.resize((size_t)I.rows * line);
        for (int y = 0; y < I.rows; y++) memcpy(&tmp[(size_t)y * line], (const uint8_t*)I.data + y * step, line);
        return tmp.data();
    }
    const uint8_t* packPlanar(const MatP& Ip, int& rows, int& cols)
    {
        if (Ip.empty() || Ip.channels() != 3) throw std::runtime_error("acf::Detector: expected three float planes");
        useFormat(6, false);
        rows = Ip.cols(); cols = Ip.rows(); // MatP holds the transposed image (ACF.cpp:135-141)
        return (const uint8_t*)Ip.ptr();
    }
    static void append(const std::vector<acfb_det>& dets, int count, RectVec& objects, RealVec* scores)
    {
        for (int i = 0; i < count; i++)
        {
            objects.push_back(ACF_CV::Rect(dets[i].x, dets[i].y, dets[i].w, dets[i].h));
            if (scores) scores->push_back(dets[i].score);
        }
    }

    acfb_model* m_model = nullptr;
    acfb_engine* m_engine = nullptr;
    acfb_options m_opts{};
    bool m_good = false, m_isTranspose = false, m_isLuv = false, m_engineTransposed = false;
    int m_fmt = 0; // acfb_set_input_format code the engine currently holds
    std::string m_error;
    int m_device = 0, m_maxRows = 0, m_maxCols = 0, m_maxBatch = 1;
    size_t m_cap = 1 << 16;
};

} // namespace acf
#endif
