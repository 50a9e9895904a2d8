"""Host-side mirror of the reference's detector interface over the C ABI (libacf_b200.so).

Same names and argument meaning as acf::Detector / acf::ObjectDetector (ACF.h:50-624,
ObjectDetector.h:31-48) so the parity tests read like the reference's own API tests
(src/test/test-acf-api.cpp).  All computation happens in the CUDA library; nothing here
falls back to the CPU.
"""
import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import check, lib

_CS = {"gray": 0, "rgb": 1, "luv": 2, "hsv": 3, "orig": 4}
_CS_INV = {v: k for k, v in _CS.items()}


def options_to_struct(d):
    o = _capi.Options()
    o.shrink = d["shrink"]; o.color_enabled = d["color_enabled"]; o.color_smooth = d["color_smooth"]
    o.color_space = _CS[d["colorSpace"].lower()]
    o.gm_enabled = d["gm_enabled"]; o.gm_colorChn = d["gm_colorChn"]; o.gm_normRad = d["gm_normRad"]
    o.gm_normConst = d["gm_normConst"]; o.gm_full = d["gm_full"]
    o.gh_enabled = d["gh_enabled"]; o.gh_binSize = d.get("gh_binSize", 0); o.gh_nOrients = d["gh_nOrients"]
    o.gh_softBin = d["gh_softBin"]; o.gh_useHog = d.get("gh_useHog", 0); o.gh_clipHog = d.get("gh_clipHog", 0.2)
    o.nPerOct = d["nPerOct"]; o.nOctUp = d["nOctUp"]; o.nApprox = d["nApprox"]
    lam = list(d.get("lambdas", []))
    o.nLambdas = len(lam)
    for i, v in enumerate(lam):
        o.lambdas[i] = v
    o.pad_w, o.pad_h = d["pad"]; o.minDs_w, o.minDs_h = d["minDs"]; o.smooth = d["smooth"]; o.concat = d.get("concat", 1)
    o.modelDs_w, o.modelDs_h = d["modelDs"]; o.modelDsPad_w, o.modelDsPad_h = d["modelDsPad"]
    o.stride = d["stride"]; o.cascThr = d["cascThr"]; o.cascCal = d.get("cascCal", 0.0)
    o.nms_type = d.get("nms_type", "maxg").encode(); o.nms_overlap = d.get("nms_overlap", 0.65)
    o.nms_ovrDnm = d.get("nms_ovrDnm", "min").encode()
    return o


def options_from_struct(o):
    return dict(
        shrink=o.shrink, color_enabled=o.color_enabled, color_smooth=o.color_smooth, colorSpace=_CS_INV[o.color_space],
        gm_enabled=o.gm_enabled, gm_colorChn=o.gm_colorChn, gm_normRad=o.gm_normRad, gm_normConst=o.gm_normConst,
        gm_full=o.gm_full, gh_enabled=o.gh_enabled, gh_binSize=o.gh_binSize, gh_nOrients=o.gh_nOrients,
        gh_softBin=o.gh_softBin, gh_useHog=o.gh_useHog, gh_clipHog=o.gh_clipHog,
        nPerOct=o.nPerOct, nOctUp=o.nOctUp, nApprox=o.nApprox, lambdas=list(o.lambdas[:o.nLambdas]),
        pad=(o.pad_w, o.pad_h), minDs=(o.minDs_w, o.minDs_h), smooth=o.smooth, concat=o.concat,
        modelDs=(o.modelDs_w, o.modelDs_h), modelDsPad=(o.modelDsPad_w, o.modelDsPad_h), stride=o.stride,
        cascThr=o.cascThr, cascCal=o.cascCal, nms_type=o.nms_type.decode(), nms_overlap=o.nms_overlap,
        nms_ovrDnm=o.nms_ovrDnm.decode())


class Model:
    """Detector::{clf, opts}: loaded from a .cpb archive or built from plain tables."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    @classmethod
    def load(cls, path_or_bytes):
        h = C.c_void_p()
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            b = bytes(path_or_bytes)
            check(lib().acfb_model_load(b, len(b), C.byref(h)))
        else:
            check(lib().acfb_model_load_file(str(path_or_bytes).encode(), C.byref(h)))
        return cls(h.value)

    @classmethod
    def create(cls, opts, clf):
        """opts: dict (reference option names); clf: dict(fids, thrs, child, hs, treeDepth[, weights, depth])."""
        o = options_to_struct(opts)
        fids = np.ascontiguousarray(clf["fids"], np.uint32); thrs = np.ascontiguousarray(clf["thrs"], np.float32)
        child = np.ascontiguousarray(clf["child"], np.uint32); hs = np.ascontiguousarray(clf["hs"], np.float32)
        weights = np.ascontiguousarray(clf["weights"], np.float32) if clf.get("weights") is not None else None
        depth = np.ascontiguousarray(clf["depth"], np.uint32) if clf.get("depth") is not None else None
        c = _capi.Classifier(fids.shape[0], fids.shape[1], int(clf["treeDepth"]), fids.ctypes.data, thrs.ctypes.data,
                             child.ctypes.data, hs.ctypes.data,
                             weights.ctypes.data if weights is not None else None,
                             depth.ctypes.data if depth is not None else None)
        h = C.c_void_p()
        check(lib().acfb_model_create(C.byref(o), C.byref(c), C.byref(h)))
        return cls(h.value)

    def to_bytes(self):
        n = C.c_size_t(0)
        check(lib().acfb_model_save(self._h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        check(lib().acfb_model_save(self._h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def save(self, path):
        check(lib().acfb_model_save_file(self._h, str(path).encode()))

    @property
    def options(self):
        o = _capi.Options()
        check(lib().acfb_model_options(self._h, C.byref(o)))
        return options_from_struct(o)

    @property
    def classifier(self):
        c = _capi.Classifier()
        check(lib().acfb_model_classifier(self._h, C.byref(c)))
        shape = (c.nTrees, c.nTreeNodes)

        def arr(p, dt):
            if not p:
                return None
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(dt)), shape=shape).copy()
        return dict(fids=arr(c.fids, C.c_uint32), thrs=arr(c.thrs, C.c_float), child=arr(c.child, C.c_uint32),
                    hs=arr(c.hs, C.c_float), weights=arr(c.weights, C.c_float), depth=arr(c.depth, C.c_uint32),
                    treeDepth=c.treeDepth)

    def acfModify(self, cascCal=0.0, cascThr=None, stride=None, nPerOct=None, nOctUp=None, nApprox=None, lambdas=None, pad=None,
                  minDs=None, pNms=None):
        """Detector::acfModify (acfModify.cpp:83-152) with the whole Detector::Modify field set (ACF.h:392-408): pad / minDs are
        (width, height) like the reference's cv::Size, pNms = dict(type=, overlap=, ovrDnm=), lambdas = [] derives them from the
        image.  cascCal is added to every hs (cumulative), stride is re-rounded to a multiple of shrink."""
        if all(v is None for v in (nPerOct, nOctUp, nApprox, lambdas, pad, minDs, pNms)):
            check(lib().acfb_model_modify(self._h, float(cascCal), float("nan") if cascThr is None else float(cascThr),
                                          -1 if stride is None else int(stride)))
            return
        p = _capi.Modify()
        for name, v in (("nPerOct", nPerOct), ("nOctUp", nOctUp), ("nApprox", nApprox), ("stride", stride)):
            if v is not None:
                setattr(p, "has_" + name, 1); setattr(p, name, int(v))
        if lambdas is not None:
            p.has_lambdas, p.nLambdas = 1, len(lambdas)
            for i, v in enumerate(lambdas):
                p.lambdas[i] = float(v)
        if pad is not None:
            p.has_pad, p.pad_w, p.pad_h = 1, int(pad[0]), int(pad[1])
        if minDs is not None:
            p.has_minDs, p.minDs_w, p.minDs_h = 1, int(minDs[0]), int(minDs[1])
        if pNms is not None:
            p.has_nms = 1
            p.nms_type = pNms.get("type", "maxg").encode(); p.nms_overlap = float(pNms.get("overlap", 0.65)); p.nms_ovrDnm = pNms.get("ovrDnm", "min").encode()
        if cascThr is not None:
            p.has_cascThr, p.cascThr = 1, float(cascThr)
        p.cascCal = float(cascCal)
        check(lib().acfb_model_modify_ex(self._h, C.byref(p)))

    def close(self):
        if self._h:
            lib().acfb_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_scales(opts, rows, cols):
    """Detector::getScales (chnsPyramid.cpp:461-529) through the C ABI; host only."""
    o = options_to_struct(opts)
    s = (C.c_double * 256)(); hw = (C.c_double * 512)(); n = C.c_int(0)
    check(lib().acfb_get_scales(C.byref(o), rows, cols, s, hw, 256, C.byref(n)))
    return np.array(s[:n.value]), np.array(hw[:2 * n.value]).reshape(n.value, 2)


class Pyramid:
    """Detector::Pyramid (ACF.h:364-389): data[i] is the concatenated plane stack of scale i,
    shape [nChns, w, h] (reference memory order: y contiguous)."""

    def __init__(self):
        self.nScales = 0; self.nTypes = 0
        self.scales = []; self.scaleshw = []; self.data = []; self.lambdas = []; self.info = []


DET_DTYPE = np.dtype([("x", np.int32), ("y", np.int32), ("w", np.int32), ("h", np.int32), ("score", np.float32), ("frame", np.int32)])


class Detector:
    """acf::Detector over the B200 engine."""

    def __init__(self, model, device=0, max_rows=2160, max_cols=3840, max_batch=1):
        if not isinstance(model, Model):
            model = Model.load(model)
        self.model = model
        self.opts = model.options
        h = C.c_void_p()
        check(lib().acfb_engine_create(model._h, device, max_rows, max_cols, max_batch, C.byref(h)))
        self._e = h
        self.max_batch = max_batch
        self._good = True

    # ---- ObjectDetector interface (ObjectDetector.h:31-48)
    def good(self):
        return self._good

    def setDoNonMaximaSuppression(self, flag):
        check(lib().acfb_set_nms(self._e, int(bool(flag))))

    def setMaxDetectionCount(self, n):
        check(lib().acfb_set_max_detection_count(self._e, int(n)))

    def setDetectionScorePruneRatio(self, r):
        check(lib().acfb_set_detection_score_prune_ratio(self._e, float(r)))

    def getWindowSize(self):
        return self.opts["modelDs"]

    # name -> (format code, trailing shape of one frame given (rows, cols), dtype)
    FORMATS = {"rgb": (0, 3, np.uint8), "bgr": (1, 3, np.uint8), "rgba": (2, 4, np.uint8), "bgra": (3, 4, np.uint8),
               "gray": (4, 1, np.uint8), "rgb32f": (5, 3, np.float32), "planar32f": (6, 3, np.float32), "nv12": (7, 1, np.uint8)}

    def setInputFormat(self, name):
        """layout of the frames: 'rgb' (default), 'bgr', 'rgba', 'bgra', 'gray' (uint8 [rows, cols, c]); 'rgb32f'
        (float32 [rows, cols, 3] in [0,1]); 'planar32f' (float32 [3, cols, rows]: the reference's MatP overloads); 'nv12'
        (uint8 [rows * 3 / 2, cols]: luma rows, then interleaved (U, V) rows; converted on the device like
        cv2.cvtColor(COLOR_YUV2RGB_NV12))"""
        check(lib().acfb_set_input_format(self._e, self.FORMATS[name][0]))
        self._fmt = name

    def setIsTranspose(self, flag):
        """Detector::setIsTranspose (ACF.h:569-576): interleaved frames arrive as [cols, rows, c]"""
        check(lib().acfb_set_is_transpose(self._e, 1 if flag else 0))
        self._transposed = bool(flag)

    def setIsLuv(self, flag):
        """Detector::setIsLuv (ACF.h:560-567): the three input channels already hold L, u, v"""
        check(lib().acfb_set_is_luv(self._e, 1 if flag else 0))

    def setHitCapacity(self, cap):
        check(lib().acfb_set_hit_capacity(self._e, int(cap)))

    # ---- planning
    def plan(self, rows, cols):
        n = C.c_int(0); fl = C.c_int64(0)
        check(lib().acfb_plan(self._e, rows, cols, None, 0, C.byref(n), C.byref(fl)))
        arr = (_capi.ScaleInfo * n.value)()
        check(lib().acfb_plan(self._e, rows, cols, arr, n.value, C.byref(n), C.byref(fl)))
        return list(arr), fl.value

    # ---- detection
    _fmt = "rgb"
    _transposed = False

    def _frames(self, I):
        """returns (contiguous array, n, rows, cols) with rows / cols of the UPRIGHT image"""
        _, c, dt = self.FORMATS[self._fmt]
        I = np.asarray(I)
        if self._fmt == "nv12":
            if I.ndim == 2:
                I = I[None]
            if I.ndim != 3 or I.dtype != np.uint8 or I.shape[1] % 3:
                raise ValueError("nv12 frames must be uint8 [rows * 3 / 2, cols] (or a batch of them)")
            return np.ascontiguousarray(I), I.shape[0], I.shape[1] * 2 // 3, I.shape[2]
        if I.ndim == 3:
            I = I[None]
        if I.ndim != 4 or I.dtype != dt:
            raise ValueError(f"frames must be {np.dtype(dt).name} with 3 or 4 dimensions for format '{self._fmt}'")
        if self._fmt == "planar32f":
            if I.shape[1] != 3:
                raise ValueError("planar32f frames are [3, cols, rows] or [n, 3, cols, rows]")
            n, _, cols, rows = I.shape
        else:
            if I.shape[3] != c:
                raise ValueError(f"frames must have {c} interleaved channels for format '{self._fmt}'")
            n, a, b, _ = I.shape
            rows, cols = (b, a) if self._transposed else (a, b)
        return np.ascontiguousarray(I), n, rows, cols

    def __call__(self, I, cap=1 << 16):
        """Detector::operator()(const cv::Mat&, RectVec&, RealVec*): returns (rects, scores) for one frame,
        or a list of such pairs for a batch."""
        single = np.asarray(I).ndim == (2 if self._fmt == "nv12" else 3)
        res = self.detect_batch(I, cap=cap)
        return res[0] if single else res

    def detect_batch(self, frames, cap=1 << 16):
        frames, n, rows, cols = self._frames(frames)
        dets = (_capi.Det * cap)()
        counts = (C.c_int * n)(); total = C.c_int(0)
        check(lib().acfb_detect(self._e, frames.ctypes.data, n, rows, cols, 0, dets, cap, counts, C.byref(total)))
        if total.value > cap:
            raise _capi.AcfError("detection buffer too small")
        return self._split(dets, counts, n)

    @staticmethod
    def _split(dets, counts, n):
        out, k = [], 0
        for f in range(n):
            rects = [(dets[k + j].x, dets[k + j].y, dets[k + j].w, dets[k + j].h) for j in range(counts[f])]
            scores = [float(dets[k + j].score) for j in range(counts[f])]
            out.append((rects, scores))
            k += counts[f]
        return out

    def submit(self, ptr, n, rows, cols, on_device):
        """asynchronous: enqueue pyramid + cascade for n frames (host or device pointer)"""
        check(lib().acfb_submit(self._e, C.c_void_p(ptr), n, rows, cols, int(on_device)))

    def collect(self, n, cap=1 << 16):
        dets = (_capi.Det * cap)()
        counts = (C.c_int * n)(); total = C.c_int(0)
        check(lib().acfb_collect(self._e, dets, cap, counts, C.byref(total)))
        if total.value > cap:
            raise _capi.AcfError("detection buffer too small")
        return self._split(dets, counts, n), total.value

    def collect_arrays(self, n, cap=1 << 16):
        """acfb_collect into numpy buffers that are kept between calls (no per-detection Python objects): returns
        (dets[:total] as a structured array with fields x, y, w, h, score, frame; counts[n]; total)."""
        buf = getattr(self, "_collect_buf", None)
        if buf is None or len(buf[0]) < cap or len(buf[1]) < n:
            buf = (np.zeros(cap, DET_DTYPE), np.zeros(n, np.int32))
            self._collect_buf = buf
        dets, counts = buf
        total = C.c_int(0)
        check(lib().acfb_collect(self._e, dets.ctypes.data_as(C.POINTER(_capi.Det)), len(dets), counts.ctypes.data_as(C.POINTER(C.c_int)), C.byref(total)))
        if total.value > len(dets):
            raise _capi.AcfError("detection buffer too small")
        return dets[:total.value], counts[:n], total.value

    # ---- multi-GPU (include/acf_b200.h: acfb_dist_*)
    @staticmethod
    def dist_unique_id():
        """128-byte NCCL unique id (rank 0 creates it; the caller's plumbing hands it to every rank)"""
        buf = (C.c_uint8 * 128)()
        check(lib().acfb_dist_unique_id(buf))
        return bytes(buf)

    def dist_init_rank(self, unique_id, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(lib().acfb_dist_init_rank(self._e, buf, rank, world))
        self._dist = (rank, world)

    @staticmethod
    def dist_init_all(detectors):
        """one process, one engine per device (ncclCommInitAll); drive every engine from its own host thread afterwards"""
        arr = (C.c_void_p * len(detectors))(*[d._e for d in detectors])
        check(lib().acfb_dist_init_all(arr, len(detectors)))
        for r, d in enumerate(detectors):
            d._dist = (r, len(detectors))

    def dist_info(self):
        r, w, v = C.c_int(0), C.c_int(0), C.c_int(0)
        check(lib().acfb_dist_info(self._e, C.byref(r), C.byref(w), C.byref(v)))
        return r.value, w.value, v.value

    def dist_collect_arrays(self, n, cap=1 << 16):
        """acfb_dist_collect: every rank calls it after submit; rank 0 gets (dets of ALL ranks in global frame order as a
        structured array, counts[world * n], total), the other ranks (empty, empty, 0)."""
        rank, world = self._dist
        buf = getattr(self, "_dist_buf", None)
        if buf is None or len(buf[0]) < cap or len(buf[1]) < world * n:
            buf = (np.zeros(cap, DET_DTYPE), np.zeros(world * n, np.int32))
            self._dist_buf = buf
        dets, counts = buf
        total = C.c_int(0)
        check(lib().acfb_dist_collect(self._e, dets.ctypes.data_as(C.POINTER(_capi.Det)), len(dets), counts.ctypes.data_as(C.POINTER(C.c_int)), C.byref(total)))
        if total.value > len(dets):
            raise _capi.AcfError("detection buffer too small")
        if rank != 0:
            return dets[:0], counts[:0], 0
        return dets[:total.value], counts[:world * n], total.value

    def synchronize(self):
        check(lib().acfb_synchronize(self._e))

    def last_hits(self):
        total = C.c_int(0); te = C.c_uint64(0); nw = C.c_uint64(0)
        check(lib().acfb_last_hits(self._e, None, 0, C.byref(total), C.byref(te), C.byref(nw)))
        arr = (_capi.Hit * max(1, total.value))()
        check(lib().acfb_last_hits(self._e, arr, total.value, C.byref(total), C.byref(te), C.byref(nw)))
        hits = [(h.frame, h.scale, h.c, h.r, float(h.score)) for h in arr[:total.value]]
        return hits, te.value, nw.value

    def last_hit_count(self):
        """(raw hits, trees evaluated, windows) of the last collected batch, without copying the hit records"""
        total = C.c_int(0); te = C.c_uint64(0); nw = C.c_uint64(0)
        check(lib().acfb_last_hits(self._e, None, 0, C.byref(total), C.byref(te), C.byref(nw)))
        return total.value, te.value, nw.value

    # ---- pyramid
    def computePyramid(self, I, frame=0):
        """Detector::computePyramid(const cv::Mat&, Pyramid&) (ACF.cpp:147-159) for frame `frame` of the batch."""
        frames, n, rows, cols = self._frames(I)
        check(lib().acfb_pyramid(self._e, frames.ctypes.data, n, rows, cols, 0))
        return self.readPyramid(rows, cols, frame)

    def readPyramid(self, rows, cols, frame=0):
        info, _ = self.plan(rows, cols)
        P = Pyramid()
        P.nScales = len(info)
        P.nTypes = (1 if self.opts["color_enabled"] else 0) + 2
        for i, s in enumerate(info):
            buf = np.empty((s.nchn, s.w, s.h), np.float32)
            check(lib().acfb_pyramid_read(self._e, frame, i, buf.ctypes.data, buf.size))
            P.data.append(buf); P.scales.append(s.scale); P.scaleshw.append((s.scalehw_w, s.scalehw_h)); P.info.append(s)
        lam = (C.c_double * 8)(); nl = C.c_int(0)
        check(lib().acfb_pyramid_lambdas(self._e, lam, 8, C.byref(nl)))
        P.lambdas = list(lam[:nl.value])
        return P

    def detectPyramid(self, n=1, cap=1 << 16):
        """Detector::operator()(const Pyramid&) on the resident pyramid."""
        dets = (_capi.Det * cap)()
        counts = (C.c_int * n)(); total = C.c_int(0)
        check(lib().acfb_detect_pyramid(self._e, dets, cap, counts, C.byref(total)))
        if total.value > cap:
            raise _capi.AcfError("detection buffer too small")
        return self._split(dets, counts, n)

    def acfDetect1(self, chns):
        """Detector::acfDetect1 on caller-provided channels [nchn, w, h]: (c, r, score) in reference order."""
        chns = np.ascontiguousarray(chns, np.float32)
        nchn, w, h = chns.shape
        cap = max(1, w * h)
        hc = np.zeros(cap, np.int32); hr = np.zeros(cap, np.int32); hs = np.zeros(cap, np.float32)
        total = C.c_int(0); te = C.c_uint64(0)
        check(lib().acfb_acf_detect1(self._e, chns.ctypes.data, h, w, nchn, hc.ctypes.data, hr.ctypes.data,
                                     hs.ctypes.data, cap, C.byref(total), C.byref(te)))
        n = total.value
        return hc[:n], hr[:n], hs[:n], te.value

    def acfDetect1U8(self, chns):
        """ParallelDetectionBody<uint8_t,k> (acfDetect1.cpp:157-191) on caller-provided uint8 channels [nchn, w, h]."""
        chns = np.ascontiguousarray(chns, np.uint8)
        nchn, w, h = chns.shape
        cap = max(1, w * h)
        hc = np.zeros(cap, np.int32); hr = np.zeros(cap, np.int32); hs = np.zeros(cap, np.float32)
        total = C.c_int(0); te = C.c_uint64(0)
        check(lib().acfb_acf_detect1_u8(self._e, chns.ctypes.data, h, w, nchn, hc.ctypes.data, hr.ctypes.data,
                                        hs.ctypes.data, cap, C.byref(total), C.byref(te)))
        n = total.value
        return hc[:n], hr[:n], hs[:n], te.value

    def detectChannels(self, data, scales, scaleshw, cap=1 << 16):
        """Detector::operator()(const Pyramid&) on a caller-provided pyramid: data[i] = [nchn, w, h] float32 or uint8
        arrays, scales[i], scaleshw[i] = (w, h) as in Detector::Pyramid (ACF.h:364-389)."""
        u8 = data[0].dtype == np.uint8
        keep = [np.ascontiguousarray(d, np.uint8 if u8 else np.float32) for d in data]
        arr = (_capi.Channels * len(keep))()
        for i, d in enumerate(keep):
            arr[i] = _capi.Channels(d.ctypes.data, d.shape[2], d.shape[1], d.shape[0], scales[i], scaleshw[i][0], scaleshw[i][1])
        dets = (_capi.Det * cap)()
        total = C.c_int(0)
        check(lib().acfb_detect_channels(self._e, arr, len(keep), int(u8), dets, cap, C.byref(total)))
        if total.value > cap:
            raise _capi.AcfError("detection buffer too small")
        counts = (C.c_int * 1)(total.value)
        return self._split(dets, counts, 1)[0]

    def evaluate(self, I):
        """Detector::evaluate(const cv::Mat&) (ACF.cpp:123-133): score of the single window at (0,0) of the
        channels of I (no pyramid), cascThr = 0."""
        fr, _, rows, cols = self._frames(I)
        s = C.c_float(0)
        check(lib().acfb_evaluate(self._e, fr.ctypes.data, rows, cols, C.byref(s)))
        return float(s.value)

    def computeChannels(self, I):
        """Detector::computeChannels(I, Ip2) (ACF.cpp:164-240): the fused real-scale channels of one frame, [10, w/4, h/4]
        (L, U, V, M, six orientation bins) -- for models whose channel options are computeChannels' fixed defaults."""
        fr, _, rows, cols = self._frames(I)
        d = C.c_int(); w = C.c_int(); h = C.c_int()
        check(lib().acfb_compute_channels(self._e, fr.ctypes.data, rows, cols, None, 0, C.byref(d), C.byref(w), C.byref(h)))
        out = np.empty((d.value, w.value, h.value), np.float32)
        check(lib().acfb_compute_channels(self._e, fr.ctypes.data, rows, cols, out.ctypes.data, out.size, C.byref(d), C.byref(w), C.byref(h)))
        return out

    # ---- the reference's static channel operators (ACF.h:416-490) on one image; arrays are [d, w, h] float32, y contiguous
    #      (the reference's transposed planar MatP)
    _CS = {"gray": 0, "rgb": 1, "luv": 2, "hsv": 3, "orig": 4}

    def rgbConvert(self, I, cs):
        """Detector::rgbConvert(I, J, cs, useSingle=true) (rgbConvert.cpp:102-170): I = [3, w, h] RGB in [0, 1]."""
        I = np.ascontiguousarray(I, np.float32)
        if I.ndim != 3 or I.shape[0] != 3:
            raise ValueError("rgbConvert: I must be [3, w, h]")
        J = np.empty_like(I)
        npl = C.c_int(0)
        check(lib().acfb_op_rgb_convert(self._e, I.ctypes.data, I.shape[2], I.shape[1], self._CS[cs] if isinstance(cs, str) else int(cs),
                                        J.ctypes.data, C.byref(npl)))
        return J[: npl.value].copy()

    def convTri(self, I, r=1.0, inplace=False):
        """Detector::convTri(I, J, r) (convTri.cpp:204-253).  inplace=True is the reference's aliased call (J is I,
        chnsCompute.cpp:239), whose r <= 1 form feeds on its own output; the result is returned either way."""
        I = np.array(I, np.float32, order="C")
        if I.ndim == 2:
            I = I[None]
        d, w, h = I.shape
        J = I if inplace else np.empty_like(I)
        check(lib().acfb_op_conv_tri(self._e, I.ctypes.data, h, w, d, float(r), J.ctypes.data))
        return J

    def gradientMag(self, I, channel=0, normRad=0, normConst=0.005, full=0):
        """Detector::gradientMag(I, M, O, channel, normRad, normConst, full) (gradientMag.cpp:109-135) -> (M, O), [w, h]."""
        I = np.ascontiguousarray(I, np.float32)
        if I.ndim == 2:
            I = I[None]
        d, w, h = I.shape
        M = np.empty((w, h), np.float32); O = np.empty((w, h), np.float32)
        check(lib().acfb_op_gradient_mag(self._e, I.ctypes.data, h, w, d, channel, normRad, float(normConst), int(full), M.ctypes.data, O.ctypes.data))
        return M, O

    def gradientHist(self, M, O, binSize=4, nOrients=6, softBin=0, useHog=0, clipHog=0.2, full=0):
        """Detector::gradientHist (gradientHist.cpp:109-114) -> H [nOrients, w / bin, h / bin]."""
        M = np.ascontiguousarray(M, np.float32); O = np.ascontiguousarray(O, np.float32)
        w, h = M.shape
        H = np.empty((nOrients, w // max(1, binSize), h // max(1, binSize)), np.float32)
        check(lib().acfb_op_gradient_hist(self._e, M.ctypes.data, O.ctypes.data, h, w, binSize, nOrients, softBin, useHog, float(clipHog), int(full),
                                          H.ctypes.data))
        return H

    def imResample(self, A, hb, wb, nrm=1.0):
        """imResample(A, B, size, nrm) (imResampleMex.cpp:385-420): [d, wa, ha] -> [d, wb, hb]."""
        A = np.ascontiguousarray(A, np.float32)
        if A.ndim == 2:
            A = A[None]
        d, wa, ha = A.shape
        B = np.empty((d, wb, hb), np.float32)
        check(lib().acfb_op_im_resample(self._e, A.ctypes.data, ha, wa, d, hb, wb, float(nrm), B.ctypes.data))
        return B

    def enable_taps(self, flag=True):
        """keep the smoothed image of every real scale in memory so that tap("C", ...) can read it (off by default: it is an
        on-chip intermediate of the fused march)"""
        check(lib().acfb_set_debug_taps(self._e, int(bool(flag))))

    def tap(self, tag, frame, real_k, shape_hint):
        """debug tap (reference's MatLoggerType hook): 'I', 'C' or 'R' planes of a real scale, [d, w, h]."""
        buf = np.empty(int(np.prod(shape_hint)), np.float32)
        d = C.c_int(); w = C.c_int(); h = C.c_int()
        check(lib().acfb_tap(self._e, tag.encode(), frame, real_k, buf.ctypes.data, buf.size, C.byref(d), C.byref(w), C.byref(h)))
        return buf[: d.value * w.value * h.value].reshape(d.value, w.value, h.value).copy()

    # ---- instrumentation
    def selftest_math(self, n=1 << 26, seed=1):
        bad = C.c_uint64(0)
        check(lib().acfb_selftest_math(self._e, n, seed, C.byref(bad)))
        return int(bad.value)

    def launch_count(self):
        return int(lib().acfb_launch_count(self._e))

    def stream(self):
        return int(lib().acfb_stream(self._e))

    def enable_stage_timing(self, flag=True):
        check(lib().acfb_enable_stage_timing(self._e, int(flag)))

    def stage_times(self):
        names = (C.c_char_p * 32)(); ms = (C.c_float * 32)()
        n = lib().acfb_stage_times(self._e, names, ms, 32)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def collect_times(self):
        """(wait_ms, tail_ms) of the last collect: blocked on the device, then the host tail (ordering, rescale, NMS)"""
        w = C.c_double(0); t = C.c_double(0)
        check(lib().acfb_collect_times(self._e, C.byref(w), C.byref(t)))
        return w.value, t.value

    def close(self):
        if getattr(self, "_e", None):
            lib().acfb_engine_destroy(self._e)
            self._e = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
