// cpb.cpp -- reader and writer for the reference's .cpb model files.
//
// .cpb = cereal 1.2.2 PortableBinaryOutputArchive of acf::Detector (io/cereal_pba.h:41-85).
// cereal is not available here, so this restates its wire rules:
//   * 1 byte endianness flag first (1 = little endian writer)
//   * arithmetic types raw; bool 1 byte; std::string / std::vector<arithmetic> = uint64 count + raw
//   * a type whose serialize() takes a version emits a uint32 class version ONCE, immediately
//     before its first instance in the stream
// and walks the fields in the order of the reference's serialisers:
//   Detector ACFIOArchive.h:75-80, Classifier :82-100, Options :102-127, Boost :129-136,
//   Tree :138-146, Pyramid :148-161, Nms :163-169, Chns :171-181, Color :183-189,
//   GradMag :191-199, GradHist :201-210, Jitter :212-216, Field<T> ACFField.h:123-130,
//   cv::Mat io/cvmat_cereal.h:18-73, cv::Size ACFIOArchive.h:47-52.
// One templated walker serves both directions so reader and writer cannot drift apart.
#include "model.h"
#include <cmath>
#include <cstring>
#include <set>

namespace acfb
{
namespace
{

struct Writer
{
    static constexpr bool loading = false;
    std::vector<uint8_t> out;
    std::set<std::string> seen;
    void raw(const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; out.insert(out.end(), b, b + n); }
    template <class T> void pod(T& v) { raw(&v, sizeof(T)); }
    void version(const char* type, uint32_t& v)
    {
        if (seen.insert(type).second) pod(v);
    }
    size_t remaining() const { return (size_t)-1; }
};

struct Reader
{
    static constexpr bool loading = true;
    const uint8_t* p; size_t n, pos = 0;
    std::set<std::string> seen;
    bool swap = false;
    Reader(const uint8_t* d, size_t len) : p(d), n(len) {}
    size_t remaining() const { return n - pos; }
    void raw(void* dst, size_t k)
    {
        if (k > n - pos) throw std::runtime_error("cpb: truncated archive");
        memcpy(dst, p + pos, k);
        pos += k;
    }
    template <class T> void pod(T& v)
    {
        raw(&v, sizeof(T));
        if (swap && sizeof(T) > 1)
        {
            uint8_t* b = (uint8_t*)&v;
            for (size_t i = 0; i < sizeof(T) / 2; i++) std::swap(b[i], b[sizeof(T) - 1 - i]);
        }
    }
    void version(const char* type, uint32_t& v)
    {
        if (seen.insert(type).second) pod(v);
    }
};

template <class Ar> void io(Ar& ar, int& v) { ar.pod(v); }
template <class Ar> void io(Ar& ar, double& v) { ar.pod(v); }
template <class Ar> void io(Ar& ar, bool& v)
{
    uint8_t b = v ? 1 : 0;
    ar.pod(b);
    v = b != 0;
}
template <class Ar> void io(Ar& ar, std::string& s)
{
    uint64_t n = s.size();
    ar.pod(n);
    if (Ar::loading)
    {
        if (n > (1u << 20) || n > ar.remaining()) throw std::runtime_error("cpb: implausible string length");
        s.resize((size_t)n);
    }
    if (n) ar.raw(&s[0], (size_t)n);
}
template <class Ar, class T> void ioVec(Ar& ar, std::vector<T>& v)
{
    uint64_t n = v.size();
    ar.pod(n);
    if (Ar::loading)
    {
        // a length field is only believed when the archive still holds that many elements (no allocation on a corrupt count)
        if (n > (1u << 28) || n * sizeof(T) > ar.remaining()) throw std::runtime_error("cpb: implausible vector length");
        v.resize((size_t)n);
    }
    for (auto& e : v) ar.pod(e);
}
template <class Ar> void io(Ar& ar, std::vector<double>& v) { ioVec(ar, v); }
template <class Ar> void io(Ar& ar, std::vector<int>& v) { ioVec(ar, v); }

template <class Ar> void io(Ar& ar, Size& s)
{
    uint32_t v = 0;
    ar.version("cv::Size", v);
    io(ar, s.width);
    io(ar, s.height);
}

template <class Ar> void io(Ar& ar, MatBlob& m)
{
    uint32_t v = 0;
    ar.version("cv::Mat", v);
    io(ar, m.rows); io(ar, m.cols); io(ar, m.type);
    bool continuous = true;
    io(ar, continuous);
    if (m.rows < 0 || m.cols < 0 || (int64_t)m.rows * m.cols > (1 << 28)) throw std::runtime_error("cpb: implausible Mat size");
    const size_t nbytes = (size_t)m.rows * m.cols * MatBlob::elemSize(m.type);
    if (Ar::loading)
    {
        if (nbytes > ar.remaining()) throw std::runtime_error("cpb: truncated archive (Mat payload larger than the rest of the file)");
        m.bytes.resize(nbytes);
    }
    if (m.bytes.size() != nbytes) throw std::runtime_error("cpb: Mat payload size mismatch");
    if (nbytes) ar.raw(m.bytes.data(), nbytes); // rows written back to back either way (continuous or not)
}

// forward declarations of the struct walkers
template <class Ar> void io(Ar& ar, Color& c);
template <class Ar> void io(Ar& ar, GradMag& c);
template <class Ar> void io(Ar& ar, GradHist& c);
template <class Ar> void io(Ar& ar, Chns& c);
template <class Ar> void io(Ar& ar, Pyramid& c);
template <class Ar> void io(Ar& ar, Nms& c);
template <class Ar> void io(Ar& ar, Tree& c);
template <class Ar> void io(Ar& ar, Boost& c);
template <class Ar> void io(Ar& ar, Jitter& c);

template <class T> struct FieldTag;
#define ACFB_TAG(T, S) template <> struct FieldTag<T> { static const char* name() { return "Field<" S ">"; } };
ACFB_TAG(int, "int") ACFB_TAG(double, "double") ACFB_TAG(std::string, "string") ACFB_TAG(std::vector<double>, "vector<double>")
ACFB_TAG(std::vector<int>, "vector<int>") ACFB_TAG(Size, "cv::Size") ACFB_TAG(Color, "Color") ACFB_TAG(GradMag, "GradMag")
ACFB_TAG(GradHist, "GradHist") ACFB_TAG(Chns, "Chns") ACFB_TAG(Pyramid, "Pyramid") ACFB_TAG(Nms, "Nms") ACFB_TAG(Tree, "Tree")
ACFB_TAG(Boost, "Boost") ACFB_TAG(Jitter, "Jitter")
#undef ACFB_TAG

template <class Ar, class T> void io(Ar& ar, Field<T>& f)
{
    uint32_t v = 0;
    ar.version(FieldTag<T>::name(), v);
    io(ar, f.value);
    io(ar, f.name);
    io(ar, f.has);
    io(ar, f.isLeaf);
}

#define ACFB_VERSIONED(NAME) uint32_t v = 0; ar.version(NAME, v);
template <class Ar> void io(Ar& ar, Color& c) { ACFB_VERSIONED("Color") io(ar, c.enabled); io(ar, c.smooth); io(ar, c.colorSpace); }
template <class Ar> void io(Ar& ar, GradMag& c)
{
    ACFB_VERSIONED("GradMag") io(ar, c.enabled); io(ar, c.colorChn); io(ar, c.normRad); io(ar, c.normConst); io(ar, c.full);
}
template <class Ar> void io(Ar& ar, GradHist& c)
{
    ACFB_VERSIONED("GradHist") io(ar, c.enabled); io(ar, c.binSize); io(ar, c.nOrients); io(ar, c.softBin); io(ar, c.useHog); io(ar, c.clipHog);
}
template <class Ar> void io(Ar& ar, Chns& c)
{
    ACFB_VERSIONED("Chns") io(ar, c.shrink); io(ar, c.complete); io(ar, c.pColor); io(ar, c.pGradMag); io(ar, c.pGradHist);
}
template <class Ar> void io(Ar& ar, Pyramid& c)
{
    ACFB_VERSIONED("Pyramid") io(ar, c.pChns); io(ar, c.nPerOct); io(ar, c.nOctUp); io(ar, c.nApprox); io(ar, c.lambdas);
    io(ar, c.pad); io(ar, c.minDs); io(ar, c.smooth); io(ar, c.concat); io(ar, c.complete);
}
template <class Ar> void io(Ar& ar, Nms& c) { ACFB_VERSIONED("Nms") io(ar, c.type); io(ar, c.overlap); io(ar, c.ovrDnm); }
template <class Ar> void io(Ar& ar, Tree& c)
{
    ACFB_VERSIONED("Tree") io(ar, c.nBins); io(ar, c.maxDepth); io(ar, c.minWeight); io(ar, c.fracFtrs); io(ar, c.nThreads);
}
template <class Ar> void io(Ar& ar, Boost& c) { ACFB_VERSIONED("Boost") io(ar, c.pTree); io(ar, c.nWeak); io(ar, c.discrete); io(ar, c.verbose); }
template <class Ar> void io(Ar& ar, Jitter& c) { ACFB_VERSIONED("Jitter") io(ar, c.flip); }

template <class Ar> void io(Ar& ar, Classifier& c)
{
    ACFB_VERSIONED("Classifier")
    io(ar, c.fids); io(ar, c.thrs); io(ar, c.child); io(ar, c.hs); io(ar, c.weights); io(ar, c.depth);
    io(ar, c.errs); io(ar, c.losses); io(ar, c.treeDepth);
}
template <class Ar> void io(Ar& ar, Options& o)
{
    ACFB_VERSIONED("Options")
    io(ar, o.pPyramid); io(ar, o.modelDs); io(ar, o.modelDsPad); io(ar, o.pNms); io(ar, o.stride); io(ar, o.cascThr);
    io(ar, o.cascCal); io(ar, o.nWeak); io(ar, o.pBoost);
    io(ar, o.posGtDir); io(ar, o.posImgDir); io(ar, o.negImgDir); io(ar, o.posWinDir); io(ar, o.negWinDir);
    io(ar, o.nPos); io(ar, o.nNeg); io(ar, o.nPerNeg); io(ar, o.nAccNeg); io(ar, o.pJitter); io(ar, o.winsSave);
}
template <class Ar> void io(Ar& ar, Model& m)
{
    ar.version("Detector", m.detectorVersion);
    io(ar, m.clf);
    io(ar, m.opts);
}

} // namespace

std::vector<uint8_t> cpbWrite(const Model& m)
{
    Writer w;
    uint8_t little = 1;
    w.pod(little);
    Model copy = m;
    io(w, copy);
    return std::move(w.out);
}

Model cpbRead(const uint8_t* data, size_t n)
{
    Reader r(data, n);
    uint8_t flag = 0;
    r.pod(flag);
    if (flag > 1) throw std::runtime_error("cpb: bad endianness flag (not a cereal PortableBinary archive)");
    r.swap = (flag != 1); // this host is little endian
    Model m;
    io(r, m);
    if (r.pos != n) throw std::runtime_error("cpb: trailing bytes after the Detector record");
    if (m.detectorVersion != 1) throw std::runtime_error("cpb: unsupported acf::Detector class version");
    m.validate();
    return m;
}

// ---------------------------------------------------------------------------------------------
void Model::validate() const
{
    auto fail = [](const std::string& s) { throw std::runtime_error("model: " + s); };
    const int CV_32S = 4, CV_32F = 5;
    if (clf.fids.type != CV_32S || clf.child.type != CV_32S) fail("fids/child must be CV_32S");
    if (clf.thrs.type != CV_32F || clf.hs.type != CV_32F) fail("thrs/hs must be CV_32F");
    const int nt = clf.fids.rows, nn = clf.fids.cols;
    if (nt <= 0 || nn <= 0) fail("empty classifier");
    for (const MatBlob* b : { &clf.thrs, &clf.child, &clf.hs })
        if (b->rows != nt || b->cols != nn) fail("tree tables disagree in shape");
    if (clf.treeDepth < 0 || clf.treeDepth > 8) fail("treeDepth must be in 0..8 (acfDetect1.cpp:201-228)");
    if (clf.treeDepth > 0 && nn < (1 << (clf.treeDepth + 1)) - 1) fail("too few nodes per tree for treeDepth");
    const Pyramid& p = opts.pPyramid.value;
    const Chns& c = p.pChns.value;
    if (!p.complete.has || p.complete.value != 1 || !c.complete.has || c.complete.value != 1)
        fail("pPyramid.complete / pChns.complete must be 1 (defaults are not re-merged, chnsPyramid.cpp:177-205)");
    if (c.shrink.value < 1 || c.shrink.value > 64) fail("shrink outside 1..64");
    if (opts.stride.value < 1) fail("stride < 1");
    // values that drive loop counts and divisions on the host (getScales, the real / approximated split): a corrupt archive
    // must be refused here, not discovered as a hang or a division by zero later
    if (p.nPerOct.value < 1 || p.nPerOct.value > 64) fail("nPerOct outside 1..64");
    if (p.nOctUp.value < 0 || p.nOctUp.value > 8) fail("nOctUp outside 0..8");
    if (p.nApprox.value < 0 || p.nApprox.value > 1024) fail("nApprox must be >= 0 once options are complete (chnsPyramid.cpp:201-204 resolves -1 only while merging defaults)");
    if (p.minDs.value.width < 1 || p.minDs.value.height < 1) fail("minDs must be positive");
    if (p.pad.value.width < 0 || p.pad.value.height < 0 || p.pad.value.width > 4096 || p.pad.value.height > 4096) fail("pad outside 0..4096");
    if (opts.modelDs.value.width < 1 || opts.modelDs.value.height < 1 || opts.modelDsPad.value.width < 1 || opts.modelDsPad.value.height < 1 ||
        opts.modelDsPad.value.width > 4096 || opts.modelDsPad.value.height > 4096) fail("modelDs / modelDsPad outside 1..4096");
    if (c.pGradHist.value.nOrients.value < 0 || c.pGradHist.value.nOrients.value > 64) fail("nOrients outside 0..64");
    if (p.lambdas.value.size() > 8) fail("more than 8 lambdas");
    if (opts.modelDsPad.value.width % c.shrink.value || opts.modelDsPad.value.height % c.shrink.value) fail("modelDsPad not a multiple of shrink");
    // every fid must address a feature inside the window (acfDetect1.cpp:269, 390-406)
    int nChns = 0;
    const std::string cs = c.pColor.value.colorSpace.value;
    if (c.pColor.value.enabled.value) nChns += (cs == "gray") ? 1 : 3;
    if (c.pGradMag.value.enabled.value) nChns += 1;
    if (c.pGradHist.value.enabled.value) nChns += c.pGradHist.value.nOrients.value;
    const int64_t nFtrs = (int64_t)nChns * (opts.modelDsPad.value.width / c.shrink.value) * (opts.modelDsPad.value.height / c.shrink.value);
    const uint32_t* f = clf.fids.ptr<uint32_t>();
    const uint32_t* ch = clf.child.ptr<uint32_t>();
    for (int i = 0; i < nt * nn; i++)
    {
        const bool internal = clf.treeDepth > 0 ? ((i % nn) < (1 << clf.treeDepth) - 1) : (ch[i] != 0);
        if (internal && f[i] >= nFtrs) fail("feature id outside the model window");
    }
    // leaf outputs and thresholds must be finite: the cascade sums leaf values and marks a rejected window by a score of -inf
    // (k_cascade_tile), which only works when no finite score can be followed by +inf or NaN
    {
        const float* hsv = clf.hs.ptr<float>();
        const float* thv = clf.thrs.ptr<float>();
        for (int i = 0; i < nt * nn; i++)
            if (!std::isfinite(hsv[i]) || !std::isfinite(thv[i])) fail("non-finite leaf output or threshold");
    }
    // variable-depth trees are walked by following child links (acfDetect1.cpp:146-155: k = child[k] - (ftr < thr)): the
    // 1-based link `child` selects node child-1 (left) or child (right), so both must lie inside the tree and strictly
    // after the node itself -- otherwise a corrupt archive reads outside the record or never reaches a leaf
    if (clf.treeDepth == 0)
        for (int t = 0; t < nt; t++)
            for (int k = 0; k < nn; k++)
            {
                const uint32_t c0 = ch[(size_t)t * nn + k];
                if (c0 != 0 && !(c0 > (uint32_t)k + 1 && c0 <= (uint32_t)nn - 1)) fail("child link does not point forward inside the tree");
            }
    // optional tables (Classifier::weights / depth, ACF.h:292-310): empty, or one 4-byte entry per node
    for (const MatBlob* b : { &clf.weights, &clf.depth })
        if (!b->bytes.empty() && (b->rows != nt || b->cols != nn || MatBlob::elemSize(b->type) != 4)) fail("weights/depth must be empty or nTrees x nTreeNodes of a 4-byte type");
}

static void copyStr(char* dst, size_t cap, const std::string& s)
{
    memset(dst, 0, cap);
    strncpy(dst, s.c_str(), cap - 1);
}

acfb_options Model::flat() const
{
    acfb_options o;
    memset(&o, 0, sizeof(o));
    const Pyramid& p = opts.pPyramid.value;
    const Chns& c = p.pChns.value;
    o.shrink = c.shrink.value;
    o.color_enabled = c.pColor.value.enabled.value;
    o.color_smooth = c.pColor.value.smooth.value;
    std::string cs = c.pColor.value.colorSpace.value;
    for (auto& ch : cs) ch = (char)tolower(ch);
    o.color_space = cs == "gray" ? 0 : cs == "rgb" ? 1 : cs == "luv" ? 2 : cs == "hsv" ? 3 : cs == "orig" ? 4 : -1;
    const GradMag& gm = c.pGradMag.value;
    o.gm_enabled = gm.enabled.value; o.gm_colorChn = gm.colorChn.value; o.gm_normRad = gm.normRad.value;
    o.gm_normConst = gm.normConst.value; o.gm_full = gm.full.has ? gm.full.value : 0; // chnsCompute.cpp:266
    const GradHist& gh = c.pGradHist.value;
    o.gh_enabled = gh.enabled.value; o.gh_binSize = gh.binSize.has ? gh.binSize.value : 0; // chnsCompute.cpp:316
    o.gh_nOrients = gh.nOrients.value; o.gh_softBin = gh.softBin.value; o.gh_useHog = gh.useHog.value; o.gh_clipHog = gh.clipHog.value;
    o.nPerOct = p.nPerOct.value; o.nOctUp = p.nOctUp.value; o.nApprox = p.nApprox.value;
    o.nLambdas = (int)std::min<size_t>(p.lambdas.value.size(), 8);
    for (int i = 0; i < o.nLambdas; i++) o.lambdas[i] = p.lambdas.value[i];
    o.pad_w = p.pad.value.width; o.pad_h = p.pad.value.height;
    o.minDs_w = p.minDs.value.width; o.minDs_h = p.minDs.value.height;
    o.smooth = p.smooth.value; o.concat = p.concat.value;
    o.modelDs_w = opts.modelDs.value.width; o.modelDs_h = opts.modelDs.value.height;
    o.modelDsPad_w = opts.modelDsPad.value.width; o.modelDsPad_h = opts.modelDsPad.value.height;
    o.stride = opts.stride.value; o.cascThr = opts.cascThr.value; o.cascCal = opts.cascCal.value;
    copyStr(o.nms_type, sizeof(o.nms_type), opts.pNms.value.type.has ? opts.pNms.value.type.value : "max");
    o.nms_overlap = opts.pNms.value.overlap.has ? opts.pNms.value.overlap.value : 0.5; // bbNms.cpp:231-238 defaults
    copyStr(o.nms_ovrDnm, sizeof(o.nms_ovrDnm), opts.pNms.value.ovrDnm.has ? opts.pNms.value.ovrDnm.value : "union");
    return o;
}

Model Model::fromFlat(const acfb_options& o, const acfb_classifier& c)
{
    Model m;
    auto mat = [&](MatBlob& b, const void* src, int type) {
        b.rows = c.nTrees; b.cols = c.nTreeNodes; b.type = type;
        b.bytes.assign((size_t)c.nTrees * c.nTreeNodes * 4, 0);
        if (src) memcpy(b.bytes.data(), src, b.bytes.size());
    };
    if (!c.fids || !c.thrs || !c.child || !c.hs) throw std::runtime_error("model: fids/thrs/child/hs are required");
    mat(m.clf.fids, c.fids, 4); mat(m.clf.thrs, c.thrs, 5); mat(m.clf.child, c.child, 4); mat(m.clf.hs, c.hs, 5);
    mat(m.clf.weights, c.weights, 5); mat(m.clf.depth, c.depth, 4);
    m.clf.treeDepth = c.treeDepth;
    Options& op = m.opts;
    op.pPyramid.name = "pPyramid"; op.pPyramid.has = true; op.pPyramid.isLeaf = true; // ParserNode::create names and marks a struct field, isLeaf keeps its default (ACFIO.h:311-323)
    Pyramid& p = op.pPyramid.value;
    p.pChns.name = "pChns"; p.pChns.has = true; p.pChns.isLeaf = true; // ParserNode::create names and marks a struct field, isLeaf keeps its default (ACFIO.h:311-323)
    Chns& ch = p.pChns.value;
    ch.shrink.set("shrink", o.shrink); ch.complete.set("complete", 1);
    ch.pColor.name = "pColor"; ch.pColor.has = true; ch.pColor.isLeaf = true; // ParserNode::create names and marks a struct field, isLeaf keeps its default (ACFIO.h:311-323)
    ch.pColor.value.enabled.set("enabled", o.color_enabled); ch.pColor.value.smooth.set("smooth", o.color_smooth);
    static const char* csn[] = { "gray", "rgb", "luv", "hsv", "orig" };
    if (o.color_space < 0 || o.color_space > 4) throw std::runtime_error("model: bad color_space");
    ch.pColor.value.colorSpace.set("colorSpace", csn[o.color_space]);
    ch.pGradMag.name = "pGradMag"; ch.pGradMag.has = true; ch.pGradMag.isLeaf = true; // ParserNode::create names and marks a struct field, isLeaf keeps its default (ACFIO.h:311-323)
    GradMag& gm = ch.pGradMag.value;
    gm.enabled.set("enabled", o.gm_enabled); gm.colorChn.set("colorChn", o.gm_colorChn); gm.normRad.set("normRad", o.gm_normRad);
    gm.normConst.set("normConst", o.gm_normConst); gm.full.set("full", o.gm_full);
    ch.pGradHist.name = "pGradHist"; ch.pGradHist.has = true; ch.pGradHist.isLeaf = true; // ParserNode::create names and marks a struct field, isLeaf keeps its default (ACFIO.h:311-323)
    GradHist& gh = ch.pGradHist.value;
    gh.enabled.set("enabled", o.gh_enabled);
    if (o.gh_binSize > 0) gh.binSize.set("binSize", o.gh_binSize); else gh.binSize.name = "binSize";
    gh.nOrients.set("nOrients", o.gh_nOrients); gh.softBin.set("softBin", o.gh_softBin); gh.useHog.set("useHog", o.gh_useHog);
    gh.clipHog.set("clipHog", o.gh_clipHog);
    p.nPerOct.set("nPerOct", o.nPerOct); p.nOctUp.set("nOctUp", o.nOctUp); p.nApprox.set("nApprox", o.nApprox);
    p.lambdas.set("lambdas", std::vector<double>(o.lambdas, o.lambdas + std::max(0, std::min(o.nLambdas, 8))));
    p.pad.set("pad", Size{ o.pad_w, o.pad_h }); p.minDs.set("minDs", Size{ o.minDs_w, o.minDs_h });
    p.smooth.set("smooth", o.smooth); p.concat.set("concat", o.concat ? o.concat : 1); p.complete.set("complete", 1);
    op.modelDs.set("modelDs", Size{ o.modelDs_w, o.modelDs_h }); op.modelDsPad.set("modelDsPad", Size{ o.modelDsPad_w, o.modelDsPad_h });
    op.pNms.name = "pNms"; op.pNms.has = true; op.pNms.isLeaf = true; // ParserNode::create names and marks a struct field, isLeaf keeps its default (ACFIO.h:311-323)
    op.pNms.value.type.set("type", o.nms_type[0] ? std::string(o.nms_type) : std::string("maxg"));
    op.pNms.value.overlap.set("overlap", o.nms_overlap > 0 ? o.nms_overlap : 0.65);
    op.pNms.value.ovrDnm.set("ovrDnm", o.nms_ovrDnm[0] ? std::string(o.nms_ovrDnm) : std::string("min"));
    op.stride.set("stride", o.stride); op.cascThr.set("cascThr", o.cascThr); op.cascCal.set("cascCal", o.cascCal);
    op.nWeak.set("nWeak", std::vector<int>{ c.nTrees });
    // what a synthetic model does not carry is written the way acf-mat2cpb leaves fields that were missing from the .mat:
    // named, has = false, zero value (ACFIO.cpp:139-184, ACFIO.h:200-210)
    op.pBoost.name = "pBoost"; op.pJitter.name = "pJitter";
    Boost& bo = op.pBoost.value;
    bo.pTree.name = "pTree";
    bo.pTree.value.nBins.name = "nBins"; bo.pTree.value.maxDepth.name = "maxDepth"; bo.pTree.value.minWeight.name = "minWeight";
    bo.pTree.value.fracFtrs.name = "fracFtrs"; bo.pTree.value.nThreads.name = "nThreads";
    bo.nWeak.name = "nWeak"; bo.discrete.name = "discrete"; bo.verbose.name = "verbose";
    op.posGtDir.name = "posGtDir"; op.posImgDir.name = "posImgDir"; op.negImgDir.name = "negImgDir"; op.posWinDir.name = "posWinDir"; op.negWinDir.name = "negWinDir";
    op.nPos.name = "nPos"; op.nNeg.name = "nNeg"; op.nPerNeg.name = "nPerNeg"; op.nAccNeg.name = "nAccNeg";
    op.pJitter.value.flip.name = "flip"; op.winsSave.name = "winsSave";
    m.validate();
    return m;
}

} // namespace acfb
