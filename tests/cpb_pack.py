"""An INDEPENDENT writer of the reference's .cpb model format, for tests only.

Written from the reference's serialisers and cereal's archive rules -- NOT from acf_b200/csrc/cpb.cpp -- so that a
misreading of the format shared by that file's reader and writer (one templated walker serves both) shows up as a byte
difference in tests/test_host.py::test_cpb_bytes_match_independent_packer.

Sources (under /root/reference/src/lib/acf/):
  acf/ACFIOArchive.h:75-80    Detector::serialize      : clf, opts
  acf/ACFIOArchive.h:82-100   Classifier::serialize    : fids thrs child hs weights depth (cv::Mat) errs losses (vector<double>) treeDepth (int)
  acf/ACFIOArchive.h:102-127  Options::serialize       : pPyramid modelDs modelDsPad pNms stride cascThr cascCal nWeak pBoost
                                                          posGtDir posImgDir negImgDir posWinDir negWinDir nPos nNeg nPerNeg nAccNeg pJitter winsSave
  acf/ACFIOArchive.h:129-216  Boost, Tree, Pyramid, Nms, Chns, Color, GradMag, GradHist, Jitter
  acf/ACFIOArchive.h:47-52    cv::Size                 : width, height
  acf/ACFField.h:123-130      Field<T>::serialize      : value, name, has, isLeaf   (value is written even when has == false)
  io/cvmat_cereal.h:18-46     cv::Mat save             : rows, cols, type, continuous, then rows*cols*elemSize raw bytes
  acf/ACFIOArchiveCereal.cpp:7  CEREAL_CLASS_VERSION(acf::Detector, 1)
  acf/ACFIO.cpp:46-187, acf/ACFIO.h:200-210,311-323: what acf-mat2cpb leaves in the Fields -- every parsed field gets its MATLAB
      name, has = "was in the .mat", isLeaf = true; struct fields get name + has = true through ParserNode::create and keep the
      default isLeaf = true (nothing ever calls setIsLeaf(false)); pCustom is not serialised.
cereal 1.2.2 PortableBinaryOutputArchive: one leading byte = 1 on a little-endian writer; arithmetic types raw little endian
(bool 1 byte, int 4, double 8); std::string and std::vector<arithmetic> = uint64 size + raw elements; a type whose
serialize / save takes a version argument emits its uint32 class version ONCE, immediately before the first instance of that
type in the stream (versions are 0 unless CEREAL_CLASS_VERSION says otherwise).
"""
import struct

import numpy as np

CV_32S, CV_32F = 4, 5


class _Out:
    def __init__(self):
        self.b = bytearray()
        self.seen = set()

    def version(self, type_name, v=0):
        if type_name not in self.seen:
            self.seen.add(type_name)
            self.b += struct.pack("<I", v)

    def i32(self, v): self.b += struct.pack("<i", int(v))
    def f64(self, v): self.b += struct.pack("<d", float(v))
    def boolean(self, v): self.b += struct.pack("<B", 1 if v else 0)

    def string(self, s):
        raw = s.encode()
        self.b += struct.pack("<Q", len(raw)) + raw

    def vec_f64(self, v):
        self.b += struct.pack("<Q", len(v)) + b"".join(struct.pack("<d", float(x)) for x in v)

    def vec_i32(self, v):
        self.b += struct.pack("<Q", len(v)) + b"".join(struct.pack("<i", int(x)) for x in v)

    def size(self, wh):
        self.version("cv::Size")
        self.i32(wh[0]); self.i32(wh[1])

    def mat(self, a, cv_type):
        self.version("cv::Mat")
        if a is None:
            rows = cols = 0; raw = b""
        else:
            a = np.ascontiguousarray(a)
            rows, cols = a.shape; raw = a.tobytes()
        self.i32(rows); self.i32(cols); self.i32(cv_type); self.boolean(True)
        self.b += raw

    # Field<T>: [version of Field<T>] value name has isLeaf
    def field(self, tname, write_value, name, has, is_leaf=True):
        self.version("Field<" + tname + ">")
        write_value()
        self.string(name); self.boolean(has); self.boolean(is_leaf)

    def f_int(self, name, v, has=True): self.field("int", lambda: self.i32(v if has else 0), name, has)
    def f_dbl(self, name, v, has=True): self.field("double", lambda: self.f64(v if has else 0.0), name, has)
    def f_str(self, name, v, has=True): self.field("string", lambda: self.string(v if has else ""), name, has)
    def f_size(self, name, v, has=True): self.field("cv::Size", lambda: self.size(v if has else (0, 0)), name, has)
    def f_vecd(self, name, v, has=True): self.field("vector<double>", lambda: self.vec_f64(v if has else []), name, has)
    def f_veci(self, name, v, has=True): self.field("vector<int>", lambda: self.vec_i32(v if has else []), name, has)


def pack(opts, clf, weights=None, depth=None):
    """opts / clf: the dictionaries of acf_b200.synth (same keys as acf_b200.Model.create).  Fields a synthetic model does not
    carry (training options, pBoost, pJitter) are written the way acf-mat2cpb writes fields missing from the .mat: named,
    has = false, zero value."""
    o = _Out()
    o.b += b"\x01"
    o.version("Detector", 1)
    # ---- Classifier
    o.version("Classifier")
    nt, nn = clf["fids"].shape
    o.mat(np.asarray(clf["fids"], np.uint32), CV_32S)
    o.mat(np.asarray(clf["thrs"], np.float32), CV_32F)
    o.mat(np.asarray(clf["child"], np.uint32), CV_32S)
    o.mat(np.asarray(clf["hs"], np.float32), CV_32F)
    o.mat(np.asarray(weights if weights is not None else clf.get("weights", np.zeros((nt, nn), np.float32)), np.float32), CV_32F)
    o.mat(np.asarray(depth if depth is not None else clf.get("depth", np.zeros((nt, nn), np.uint32)), np.uint32), CV_32S)
    o.vec_f64([]); o.vec_f64([])  # errs, losses
    o.i32(clf["treeDepth"])
    # ---- Options
    o.version("Options")

    def pyramid():
        o.version("Pyramid")

        def chns():
            o.version("Chns")
            o.f_int("shrink", opts["shrink"]); o.f_int("complete", 1)

            def color():
                o.version("Color")
                o.f_int("enabled", opts["color_enabled"]); o.f_dbl("smooth", opts["color_smooth"]); o.f_str("colorSpace", opts["colorSpace"])
            o.field("Color", color, "pColor", True)

            def gradmag():
                o.version("GradMag")
                o.f_int("enabled", opts["gm_enabled"]); o.f_int("colorChn", opts["gm_colorChn"]); o.f_int("normRad", opts["gm_normRad"])
                o.f_dbl("normConst", opts["gm_normConst"]); o.f_int("full", opts["gm_full"])
            o.field("GradMag", gradmag, "pGradMag", True)

            def gradhist():
                o.version("GradHist")
                o.f_int("enabled", opts["gh_enabled"])
                o.f_int("binSize", opts.get("gh_binSize", 0), has=opts.get("gh_binSize", 0) > 0)  # toolbox models leave binSize empty (= shrink)
                o.f_int("nOrients", opts["gh_nOrients"]); o.f_int("softBin", opts["gh_softBin"]); o.f_int("useHog", opts["gh_useHog"])
                o.f_dbl("clipHog", opts["gh_clipHog"])
            o.field("GradHist", gradhist, "pGradHist", True)
        o.field("Chns", chns, "pChns", True)
        o.f_int("nPerOct", opts["nPerOct"]); o.f_int("nOctUp", opts["nOctUp"]); o.f_int("nApprox", opts["nApprox"])
        o.f_vecd("lambdas", opts["lambdas"])
        o.f_size("pad", opts["pad"]); o.f_size("minDs", opts["minDs"])
        o.f_dbl("smooth", opts["smooth"]); o.f_int("concat", opts.get("concat", 1) or 1); o.f_int("complete", 1)
    o.field("Pyramid", pyramid, "pPyramid", True)
    o.f_size("modelDs", opts["modelDs"]); o.f_size("modelDsPad", opts["modelDsPad"])

    def nms():
        o.version("Nms")
        o.f_str("type", opts["nms_type"]); o.f_dbl("overlap", opts["nms_overlap"]); o.f_str("ovrDnm", opts["nms_ovrDnm"])
    o.field("Nms", nms, "pNms", True)
    o.f_int("stride", opts["stride"]); o.f_dbl("cascThr", opts["cascThr"]); o.f_dbl("cascCal", opts["cascCal"])
    o.f_veci("nWeak", [nt])

    def boost():
        o.version("Boost")

        def tree():
            o.version("Tree")
            o.f_int("nBins", 0, False); o.f_int("maxDepth", 0, False); o.f_dbl("minWeight", 0, False); o.f_dbl("fracFtrs", 0, False)
            o.f_int("nThreads", 0, False)
        o.field("Tree", tree, "pTree", False)
        o.f_int("nWeak", 0, False); o.f_int("discrete", 0, False); o.f_int("verbose", 0, False)
    o.field("Boost", boost, "pBoost", False)
    for name in ("posGtDir", "posImgDir", "negImgDir", "posWinDir", "negWinDir"):
        o.f_str(name, "", False)
    for name in ("nPos", "nNeg", "nPerNeg", "nAccNeg"):
        o.f_int(name, 0, False)

    def jitter():
        o.version("Jitter")
        o.f_int("flip", 0, False)
    o.field("Jitter", jitter, "pJitter", False)
    o.f_int("winsSave", 0, False)
    return bytes(o.b)
