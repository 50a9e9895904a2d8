// model.h -- host-side mirror of acf::Detector::{Classifier, Options} and the .cpb archive.
//
// Plain structs with the reference's field names (ACF.h:68-310, ACFField.h:24-136) so that a .cpb
// written by acf-mat2cpb (cereal 1.2.2 PortableBinary, ACFIOArchive.h:75-216, io/cvmat_cereal.h:18-73)
// round-trips byte for byte.  No OpenCV, no cereal: the archive walker in cpb.cpp restates cereal's
// published wire rules.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/acf_b200.h"

namespace acfb
{

template <class T>
struct Field // ACFField.h:24-136
{
    T value{};
    std::string name;
    bool has = false;
    bool isLeaf = true;
    void set(const std::string& n, const T& v) { name = n; value = v; has = true; }
};

struct Size { int width = 0, height = 0; };

struct MatBlob // cv::Mat as serialised by io/cvmat_cereal.h:18-46
{
    int rows = 0, cols = 0, type = 0; // OpenCV type code: CV_8U=0, CV_32S=4, CV_32F=5
    std::vector<uint8_t> bytes;
    static int elemSize(int type)
    {
        const int depth = type & 7, cn = (type >> 3) + 1;
        static const int sz[8] = { 1, 1, 2, 2, 4, 4, 8, 2 };
        return sz[depth] * cn;
    }
    template <class T> const T* ptr() const { return reinterpret_cast<const T*>(bytes.data()); }
    template <class T> T* ptr() { return reinterpret_cast<T*>(bytes.data()); }
};

struct Color { Field<int> enabled; Field<double> smooth; Field<std::string> colorSpace; };
struct GradMag { Field<int> enabled, colorChn, normRad; Field<double> normConst; Field<int> full; };
struct GradHist { Field<int> enabled, binSize, nOrients, softBin, useHog; Field<double> clipHog; };
struct Chns
{
    Field<int> shrink, complete;
    Field<Color> pColor;
    Field<GradMag> pGradMag;
    Field<GradHist> pGradHist;
};
struct Pyramid
{
    Field<Chns> pChns;
    Field<int> nPerOct, nOctUp, nApprox;
    Field<std::vector<double>> lambdas;
    Field<Size> pad, minDs;
    Field<double> smooth;
    Field<int> concat, complete;
};
struct Nms { Field<std::string> type; Field<double> overlap; Field<std::string> ovrDnm; };
struct Tree { Field<int> nBins, maxDepth; Field<double> minWeight, fracFtrs; Field<int> nThreads; };
struct Boost { Field<Tree> pTree; Field<int> nWeak, discrete, verbose; };
struct Jitter { Field<int> flip; };

struct Options
{
    Field<Pyramid> pPyramid;
    Field<Size> modelDs, modelDsPad;
    Field<Nms> pNms;
    Field<int> stride;
    Field<double> cascThr, cascCal;
    Field<std::vector<int>> nWeak;
    Field<Boost> pBoost;
    Field<std::string> posGtDir, posImgDir, negImgDir, posWinDir, negWinDir;
    Field<int> nPos, nNeg, nPerNeg, nAccNeg;
    Field<Jitter> pJitter;
    Field<int> winsSave;
};

struct Classifier
{
    MatBlob fids, thrs, child, hs, weights, depth;
    std::vector<double> errs, losses;
    int treeDepth = 0;
};

struct Model
{
    Classifier clf;
    Options opts;
    uint32_t detectorVersion = 1; // CEREAL_CLASS_VERSION(acf::Detector, 1), ACFIOArchiveCereal.cpp:7

    int nTrees() const { return clf.fids.rows; }
    int nTreeNodes() const { return clf.fids.cols; }
    void validate() const;                 // throws std::runtime_error on an unusable model
    acfb_options flat() const;             // plain view used by the planner / C ABI
    static Model fromFlat(const acfb_options& o, const acfb_classifier& c);
};

std::vector<uint8_t> cpbWrite(const Model& m);
Model cpbRead(const uint8_t* data, size_t n);

} // namespace acfb

struct acfb_model { acfb::Model m; };
