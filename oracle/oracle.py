"""ctypes binding for the CPU oracle (oracle/acf_oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (acf_b200/) never imports this.

kinds:  "port"       oracle/liboracle_port.so            (this repo's restatement)
        "ref_exact"  oracle/_ref/liboracle_ref_exact.so  (reference toolbox objects, IEEE 1/x, 1/sqrt)
        "ref_native" oracle/_ref/liboracle_ref_native.so (reference toolbox objects as shipped)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "port": os.path.join(HERE, "liboracle_port.so"),
    "ref_exact": os.path.join(HERE, "_ref", "liboracle_ref_exact.so"),
    "ref_native": os.path.join(HERE, "_ref", "liboracle_ref_native.so"),
}


class Opts(C.Structure):
    _fields_ = [
        ("shrink", C.c_int), ("color_enabled", C.c_int), ("color_smooth", C.c_double), ("color_space", C.c_int),
        ("gm_enabled", C.c_int), ("gm_colorChn", C.c_int), ("gm_normRad", C.c_int), ("gm_normConst", C.c_double),
        ("gm_full", C.c_int),
        ("gh_enabled", C.c_int), ("gh_binSize", C.c_int), ("gh_nOrients", C.c_int), ("gh_softBin", C.c_int),
        ("nPerOct", C.c_int), ("nOctUp", C.c_int), ("nApprox", C.c_int),
        ("nLambdas", C.c_int), ("lambdas", C.c_double * 8),
        ("pad_w", C.c_int), ("pad_h", C.c_int), ("minDs_w", C.c_int), ("minDs_h", C.c_int), ("smooth", C.c_double),
        ("modelDs_w", C.c_int), ("modelDs_h", C.c_int), ("modelDsPad_w", C.c_int), ("modelDsPad_h", C.c_int),
        ("stride", C.c_int), ("cascThr", C.c_double),
    ]


class Det(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int), ("score", C.c_double)]


class Clf(C.Structure):
    _fields_ = [("nTrees", C.c_int), ("nTreeNodes", C.c_int), ("treeDepth", C.c_int),
                ("fids", C.c_void_p), ("thrs", C.c_void_p), ("child", C.c_void_p), ("hs", C.c_void_p)]


TAP_FN = C.CFUNCTYPE(None, C.c_char_p, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_void_p)


def available(kind):
    return os.path.exists(_PATHS[kind])


def build(ref=None):
    """make the port library (and the _ref libraries when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref is None:
        ref = os.path.isdir("/root/reference/src/lib/acf")
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def opts_from_dict(d):
    """d: the plain dict form used by acf_b200.model (same key names as the reference's options)."""
    o = Opts()
    cs = {"gray": 0, "rgb": 1, "luv": 2, "hsv": 3, "orig": 4}[d["colorSpace"].lower()]
    o.shrink = d["shrink"]; o.color_enabled = d["color_enabled"]; o.color_smooth = d["color_smooth"]; o.color_space = cs
    o.gm_enabled = d["gm_enabled"]; o.gm_colorChn = d["gm_colorChn"]; o.gm_normRad = d["gm_normRad"]
    o.gm_normConst = d["gm_normConst"]; o.gm_full = d["gm_full"]
    o.gh_enabled = d["gh_enabled"]; o.gh_binSize = d.get("gh_binSize", 0); o.gh_nOrients = d["gh_nOrients"]
    o.gh_softBin = d["gh_softBin"]
    o.nPerOct = d["nPerOct"]; o.nOctUp = d["nOctUp"]; o.nApprox = d["nApprox"]
    lam = list(d.get("lambdas", []))
    o.nLambdas = len(lam)
    for i, v in enumerate(lam):
        o.lambdas[i] = v
    o.pad_w, o.pad_h = d["pad"]; o.minDs_w, o.minDs_h = d["minDs"]; o.smooth = d["smooth"]
    o.modelDs_w, o.modelDs_h = d["modelDs"]; o.modelDsPad_w, o.modelDsPad_h = d["modelDsPad"]
    o.stride = d["stride"]; o.cascThr = d["cascThr"]
    return o


class Oracle:
    def __init__(self, kind="port"):
        path = _PATHS[kind]
        if not os.path.exists(path):
            raise FileNotFoundError(f"oracle library {path} missing (run `make -C oracle`)")
        self.kind = kind
        L = self.lib = C.CDLL(path)
        L.oracle_kind.restype = C.c_char_p
        assert L.oracle_kind().decode() == kind
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_pyramid_create.restype = C.c_void_p
        L.oracle_pyramid_create.argtypes = [C.POINTER(Opts), C.c_void_p, C.c_int, C.c_int, C.c_int, TAP_FN, C.c_void_p]
        L.oracle_pyramid_destroy.argtypes = [C.c_void_p]
        L.oracle_pyramid_nscales.argtypes = [C.c_void_p]
        L.oracle_pyramid_ntypes.argtypes = [C.c_void_p]
        L.oracle_pyramid_scale.restype = C.POINTER(C.c_float)
        L.oracle_pyramid_scale.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_double)] * 3
        L.oracle_pyramid_lambdas.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.oracle_detect.argtypes = [C.c_void_p, C.POINTER(Opts), C.POINTER(Clf), C.POINTER(Det), C.c_int,
                                    C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_acf_detect1.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Opts), C.POINTER(Clf),
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.oracle_acf_detect1_u8.argtypes = L.oracle_acf_detect1.argtypes
        L.oracle_evaluate.argtypes = [C.POINTER(Opts), C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Clf), C.POINTER(C.c_float)]
        L.oracle_nms.argtypes = [C.POINTER(Det), C.c_int, C.c_double, C.c_int, C.c_int]
        L.oracle_prune.argtypes = [C.POINTER(Det), C.c_int, C.c_int, C.c_double]
        L.oracle_get_scales.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        fp = C.c_void_p
        L.oracle_rgb_convert.argtypes = [fp, fp, C.c_int, C.c_int]
        L.oracle_conv_tri1.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
        L.oracle_conv_tri.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_grad_mag.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_grad_mag_norm.argtypes = [fp, fp, C.c_int, C.c_int, C.c_float]
        L.oracle_grad_hist.argtypes = [fp, fp, fp] + [C.c_int] * 6
        L.oracle_resample.argtypes = [fp, fp] + [C.c_int] * 5 + [C.c_float]

    # ------------------------------------------------------------------ L2
    def get_scales(self, nPerOct, nOctUp, minDs, shrink, sz):
        s = (C.c_double * 256)(); hw = (C.c_double * 512)()
        n = self.lib.oracle_get_scales(nPerOct, nOctUp, minDs[0], minDs[1], shrink, sz[0], sz[1], s, hw, 256)
        return np.array(s[:n]), np.array(hw[:2 * n]).reshape(n, 2)

    def pyramid(self, opts, img, taps=None):
        """img: HWC uint8 or float32 RGB.  Returns Pyramid. taps: optional dict filled with
        {(tag, scale): array[d, w, h]} of intermediate planes."""
        img = np.ascontiguousarray(img)
        assert img.ndim == 3 and img.shape[2] == 3 and img.dtype in (np.uint8, np.float32)
        o = opts if isinstance(opts, Opts) else opts_from_dict(opts)

        def _tap(tag, scale, data, h, w, d, user):
            arr = np.ctypeslib.as_array(data, shape=(d, w, h)).copy()
            taps[(tag.decode(), scale)] = arr
        cb = TAP_FN(_tap) if taps is not None else C.cast(None, TAP_FN)
        h = self.lib.oracle_pyramid_create(C.byref(o), img.ctypes.data, img.shape[0], img.shape[1],
                                           int(img.dtype == np.float32), cb, None)
        if not h:
            raise RuntimeError("oracle: " + self.lib.oracle_last_error().decode())
        return Pyramid(self, h, o)

    def acf_detect1(self, chns, opts, clf):
        """chns: float32 [nchn, w, h]; returns (c, r, score) arrays in reference order, trees evaluated."""
        o = opts if isinstance(opts, Opts) else opts_from_dict(opts)
        chns = np.ascontiguousarray(chns, dtype=np.float32)
        nchn, w, h = chns.shape
        cap = max(1, w * h)
        hc = np.zeros(cap, np.int32); hr = np.zeros(cap, np.int32); hs = np.zeros(cap, np.float32)
        ne = C.c_uint64(0)
        c = make_clf(clf)
        n = self.lib.oracle_acf_detect1(chns.ctypes.data, h, w, nchn, C.byref(o), C.byref(c[0]), hc.ctypes.data,
                                        hr.ctypes.data, hs.ctypes.data, cap, C.byref(ne))
        return hc[:n], hr[:n], hs[:n], ne.value

    def acf_detect1_u8(self, chns, opts, clf):
        """chns: uint8 [nchn, w, h]; the byte-channel detector with thrsU8 (acfDetect1.cpp:157-191)."""
        o = opts if isinstance(opts, Opts) else opts_from_dict(opts)
        chns = np.ascontiguousarray(chns, dtype=np.uint8)
        nchn, w, h = chns.shape
        cap = max(1, w * h)
        hc = np.zeros(cap, np.int32); hr = np.zeros(cap, np.int32); hs = np.zeros(cap, np.float32)
        ne = C.c_uint64(0)
        c = make_clf(clf)
        n = self.lib.oracle_acf_detect1_u8(chns.ctypes.data, h, w, nchn, C.byref(o), C.byref(c[0]), hc.ctypes.data,
                                           hr.ctypes.data, hs.ctypes.data, cap, C.byref(ne))
        return hc[:n], hr[:n], hs[:n], ne.value

    def evaluate(self, opts, img, clf):
        """Detector::evaluate(cv::Mat): score of the window at (0,0) of chnsCompute(img)"""
        o = opts if isinstance(opts, Opts) else opts_from_dict(opts)
        img = np.ascontiguousarray(img)
        c, keep = make_clf(clf)
        s = C.c_float(0)
        if self.lib.oracle_evaluate(C.byref(o), img.ctypes.data, img.shape[0], img.shape[1], int(img.dtype == np.float32), C.byref(c), C.byref(s)):
            raise RuntimeError("oracle: " + self.lib.oracle_last_error().decode())
        return float(s.value)

    def nms(self, dets, overlap=0.65, greedy=True, ovr_union=True):
        arr = (Det * max(1, len(dets)))(*[Det(*d) for d in dets])
        n = self.lib.oracle_nms(arr, len(dets), overlap, int(greedy), int(ovr_union))
        return [(a.x, a.y, a.w, a.h, a.score) for a in arr[:n]]

    def prune(self, dets, max_count=10, ratio=0.0):
        arr = (Det * max(1, len(dets)))(*[Det(*d) for d in dets])
        n = self.lib.oracle_prune(arr, len(dets), max_count, ratio)
        return [(a.x, a.y, a.w, a.h, a.score) for a in arr[:n]]

    # ------------------------------------------------------------------ L1 (arrays are [d, w, h], y contiguous)
    def rgb_convert(self, I, flag):
        I = np.ascontiguousarray(I, np.float32)
        n = I.shape[1] * I.shape[2]
        J = np.zeros((1 if flag == 0 else 3,) + I.shape[1:], np.float32)
        self.lib.oracle_rgb_convert(I.ctypes.data, J.ctypes.data, n, flag)
        return J

    def conv_tri1(self, I, p, inplace=False):
        I = np.array(I, np.float32, order="C")
        d, w, h = I.shape
        O = I if inplace else np.zeros_like(I)
        self.lib.oracle_conv_tri1(I.ctypes.data, O.ctypes.data, h, w, d, p, 1)
        return O

    def conv_tri(self, I, r):
        I = np.array(I, np.float32, order="C")
        d, w, h = I.shape
        O = np.zeros_like(I)
        self.lib.oracle_conv_tri(I.ctypes.data, O.ctypes.data, h, w, d, r, 1)
        return O

    def grad_mag(self, I, full=0):
        I = np.array(I, np.float32, order="C")
        w, h = I.shape
        M = np.zeros_like(I); O = np.zeros_like(I)
        self.lib.oracle_grad_mag(I.ctypes.data, M.ctypes.data, O.ctypes.data, h, w, 1, full)
        return M, O

    def grad_mag_norm(self, M, S, norm):
        M = np.array(M, np.float32, order="C"); S = np.ascontiguousarray(S, np.float32)
        w, h = M.shape
        self.lib.oracle_grad_mag_norm(M.ctypes.data, S.ctypes.data, h, w, norm)
        return M

    def grad_hist(self, M, O, bin, nOrients, softBin=0, full=0):
        M = np.ascontiguousarray(M, np.float32); O = np.ascontiguousarray(O, np.float32)
        w, h = M.shape
        H = np.zeros((nOrients, w // bin, h // bin), np.float32)
        self.lib.oracle_grad_hist(M.ctypes.data, O.ctypes.data, H.ctypes.data, h, w, bin, nOrients, softBin, full)
        return H

    def resample(self, A, hb, wb, r=1.0):
        A = np.ascontiguousarray(A, np.float32)
        d, wa, ha = A.shape
        B = np.zeros((d, wb, hb), np.float32)
        self.lib.oracle_resample(A.ctypes.data, B.ctypes.data, ha, hb, wa, wb, d, r)
        return B


def make_clf(clf):
    """clf: dict with fids(uint32 [nTrees,nNodes]), thrs(f32), child(uint32), hs(f32), treeDepth.
    Returns (Clf, keepalive)."""
    fids = np.ascontiguousarray(clf["fids"], np.uint32); thrs = np.ascontiguousarray(clf["thrs"], np.float32)
    child = np.ascontiguousarray(clf["child"], np.uint32); hs = np.ascontiguousarray(clf["hs"], np.float32)
    c = Clf(fids.shape[0], fids.shape[1], int(clf["treeDepth"]), fids.ctypes.data, thrs.ctypes.data,
            child.ctypes.data, hs.ctypes.data)
    return c, (fids, thrs, child, hs)


class Pyramid:
    def __init__(self, orc, handle, opts):
        self.orc, self.h, self.opts = orc, handle, opts
        L = orc.lib
        self.nScales = L.oracle_pyramid_nscales(handle)
        self.nTypes = L.oracle_pyramid_ntypes(handle)
        self.scales, self.scaleshw, self.data = [], [], []
        for i in range(self.nScales):
            h = C.c_int(); w = C.c_int(); d = C.c_int(); s = C.c_double(); sw = C.c_double(); sh = C.c_double()
            p = L.oracle_pyramid_scale(handle, i, C.byref(h), C.byref(w), C.byref(d), C.byref(s), C.byref(sw), C.byref(sh))
            self.data.append(np.ctypeslib.as_array(p, shape=(d.value, w.value, h.value)).copy())
            self.scales.append(s.value); self.scaleshw.append((sw.value, sh.value))
        lam = (C.c_double * 8)()
        n = L.oracle_pyramid_lambdas(handle, lam, 8)
        self.lambdas = list(lam[:n])

    def detect(self, clf, cap=1 << 20):
        c, keep = make_clf(clf)
        out = (Det * cap)()
        total = C.c_int(0); ne = C.c_uint64(0)
        hs = np.zeros(cap, np.int32); hc = np.zeros(cap, np.int32); hr = np.zeros(cap, np.int32)
        n = self.orc.lib.oracle_detect(self.h, C.byref(self.opts), C.byref(c), out, cap, C.byref(total),
                                       hs.ctypes.data, hc.ctypes.data, hr.ctypes.data, C.byref(ne))
        dets = [(o.x, o.y, o.w, o.h, o.score) for o in out[:n]]
        return dets, (hs[:n], hc[:n], hr[:n]), ne.value, total.value

    def close(self):
        if self.h:
            self.orc.lib.oracle_pyramid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
