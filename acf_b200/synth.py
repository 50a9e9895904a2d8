"""Seeded synthetic frames and models.

The reference ships no models or images in-tree (Hunter downloads them), so tests and the
benchmark use stand-ins of the right SHAPE:
  * frames: "shapes" -- the recipe of acf-detect's --random mode (src/app/acf/acf.cpp:65-86,553-612:
    black canvas, rand()%32 rounds of {line, rectangle, ellipse} with switch fall-through, rectangles
    noise-filled 3 times in 8, line thickness 1-16) driven by a seeded numpy generator instead of
    the unseeded rand(); "noise" -- a smooth random field (4 octaves of bilinear noise).
  * models: FACE80 / FACE64 (gray, 7 or 8 channels) and INRIA-shaped (LUV, 10 channels, 128x64
    window, pad 16x12) option sets with random depth-2 trees whose thresholds follow the empirical
    per-channel value quantiles below, and leaf outputs tuned for a stated mean trees/window.
Pure numpy; no CUDA, no oracle.
"""
import numpy as np

# --------------------------------------------------------------------------------------------- frames


def _draw_line(img, p0, p1, color, thick):
    x0, y0 = p0; x1, y1 = p1
    r = max(1.0, thick / 2.0)
    H, W, _ = img.shape
    xa, xb = int(max(0, min(x0, x1) - r - 1)), int(min(W, max(x0, x1) + r + 2))
    ya, yb = int(max(0, min(y0, y1) - r - 1)), int(min(H, max(y0, y1) + r + 2))
    if xa >= xb or ya >= yb:
        return
    yy, xx = np.mgrid[ya:yb, xa:xb].astype(np.float32)
    dx, dy = float(x1 - x0), float(y1 - y0)
    L2 = dx * dx + dy * dy
    if L2 == 0:
        d2 = (xx - x0) ** 2 + (yy - y0) ** 2
    else:
        t = np.clip(((xx - x0) * dx + (yy - y0) * dy) / L2, 0, 1)
        d2 = (xx - (x0 + t * dx)) ** 2 + (yy - (y0 + t * dy)) ** 2
    img[ya:yb, xa:xb][d2 <= r * r] = color


def _draw_rect(img, p0, p1, color, rng, noise):
    H, W, _ = img.shape
    xa, xb = sorted((int(p0[0]), int(p1[0]))); ya, yb = sorted((int(p0[1]), int(p1[1])))
    xa, ya = max(0, xa), max(0, ya); xb, yb = min(W, xb + 1), min(H, yb + 1)
    if xa >= xb or ya >= yb:
        return
    if noise:
        img[ya:yb, xa:xb] = rng.integers(0, 256, (yb - ya, xb - xa, 3), dtype=np.uint8)
    else:
        img[ya:yb, xa:xb] = color


def _draw_ellipse(img, center, axes, angle, color):
    H, W, _ = img.shape
    cx, cy = center; a, b = max(1, axes[0]), max(1, axes[1])
    r = max(a, b)
    xa, xb = int(max(0, cx - r - 1)), int(min(W, cx + r + 2)); ya, yb = int(max(0, cy - r - 1)), int(min(H, cy + r + 2))
    if xa >= xb or ya >= yb:
        return
    yy, xx = np.mgrid[ya:yb, xa:xb].astype(np.float32)
    ca, sa = np.cos(angle), np.sin(angle)
    u = (xx - cx) * ca + (yy - cy) * sa; v = -(xx - cx) * sa + (yy - cy) * ca
    img[ya:yb, xa:xb][(u / a) ** 2 + (v / b) ** 2 <= 1.0] = color


def shapes_frame(seed, rows, cols):
    """acf-detect --random recipe with a seeded generator; returns HWC uint8 RGB."""
    rng = np.random.default_rng(seed)
    img = np.zeros((rows, cols, 3), np.uint8)
    for _ in range(int(rng.integers(0, 32))):
        kind = int(rng.integers(0, 3))
        color = rng.integers(0, 256, 3).astype(np.uint8)
        pts = [(int(rng.integers(0, cols)), int(rng.integers(0, rows))) for _ in range(2)]
        if kind <= 0:
            _draw_line(img, pts[0], pts[1], color, int(rng.integers(1, 17)))
        if kind <= 1:  # switch fall-through: case 0 draws all three, case 1 rect + ellipse
            _draw_rect(img, pts[0], pts[1], color, rng, noise=int(rng.integers(0, 8)) < 3)
        _draw_ellipse(img, pts[0], (int(rng.integers(1, max(2, cols // 4))), int(rng.integers(1, max(2, rows // 4)))),
                      float(rng.uniform(0, np.pi)), color)
    return img


def _bilinear_up(a, rows, cols):
    r0, c0 = a.shape
    yi = np.linspace(0, r0 - 1, rows); xi = np.linspace(0, c0 - 1, cols)
    y0 = np.floor(yi).astype(int); x0 = np.floor(xi).astype(int)
    y1 = np.minimum(y0 + 1, r0 - 1); x1 = np.minimum(x0 + 1, c0 - 1)
    wy = (yi - y0)[:, None]; wx = (xi - x0)[None, :]
    return (a[y0][:, x0] * (1 - wy) * (1 - wx) + a[y0][:, x1] * (1 - wy) * wx
            + a[y1][:, x0] * wy * (1 - wx) + a[y1][:, x1] * wy * wx)


def noise_frame(seed, rows, cols):
    """smooth random colour field: 4 octaves of bilinearly up-sampled uniform noise; HWC uint8 RGB."""
    rng = np.random.default_rng(seed)
    out = np.zeros((rows, cols, 3), np.float64)
    amp, tot = 1.0, 0.0
    for k in (64, 32, 16, 8):
        g = rng.random((max(2, rows // k), max(2, cols // k), 3))
        for c in range(3):
            out[:, :, c] += amp * _bilinear_up(g[:, :, c], rows, cols)
        tot += amp; amp *= 0.5
    return np.clip(out / tot * 255.0 + 0.5, 0, 255).astype(np.uint8)


def frames(kind, n, rows, cols, seed0=0):
    gen = {"shapes": shapes_frame, "noise": noise_frame}[kind]
    return np.stack([gen(seed0 + i, rows, cols) for i in range(n)])


# --------------------------------------------------------------------------------------------- NV12


def rgb_to_nv12(rgb):
    """HWC uint8 RGB -> NV12 uint8 [rows * 3 / 2, cols] (BT.601 limited range, 2 x 2 chroma means); rows, cols even.
    Only a way to make NV12 test frames: any byte content is valid NV12."""
    f = rgb.astype(np.float64)
    r, g, b = f[:, :, 0], f[:, :, 1], f[:, :, 2]
    y = 16 + (65.481 * r + 128.553 * g + 24.966 * b) / 255.0
    u = 128 + (-37.797 * r - 74.203 * g + 112.0 * b) / 255.0
    v = 128 + (112.0 * r - 93.786 * g - 18.214 * b) / 255.0
    rows, cols = y.shape
    sub = lambda p: p.reshape(rows // 2, 2, cols // 2, 2).mean(axis=(1, 3))
    out = np.empty((rows * 3 // 2, cols), np.uint8)
    out[:rows] = np.clip(np.rint(y), 0, 255)
    uv = np.stack([sub(u), sub(v)], axis=-1)
    out[rows:] = np.clip(np.rint(uv), 0, 255).reshape(rows // 2, cols)
    return out


def nv12_to_rgb(nv12):
    """NV12 -> HWC uint8 RGB by the integer ITU-R BT.601 formula of cv2.cvtColor(COLOR_YUV2RGB_NV12) (20-bit fixed point);
    tests/test_host.py checks it against cv2 where cv2 is importable.  This is the DEFINITION the device conversion must equal."""
    h = nv12.shape[0] * 2 // 3
    w = nv12.shape[1]
    Y = nv12[:h].astype(np.int64)
    UV = nv12[h:].reshape(h // 2, w // 2, 2).astype(np.int64)
    U = np.repeat(np.repeat(UV[:, :, 0], 2, 0), 2, 1) - 128
    V = np.repeat(np.repeat(UV[:, :, 1], 2, 0), 2, 1) - 128
    y = np.maximum(0, Y - 16) * 1220542
    half = 1 << 19
    r = (y + half + 1673527 * V) >> 20
    g = (y + half - 852492 * V - 409993 * U) >> 20
    b = (y + half + 2116026 * U) >> 20
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------- options

def face_opts(size=80, color=False):
    """FACE80 / FACE64 stand-in (gray, shrink 4, nPerOct 8, nApprox 7, pad 0, 7 or 8 channels)."""
    return dict(shrink=4, color_enabled=1 if color else 0, color_smooth=1.0, colorSpace="gray",
                gm_enabled=1, gm_colorChn=0, gm_normRad=5, gm_normConst=0.005, gm_full=0,
                gh_enabled=1, gh_binSize=0, gh_nOrients=6, gh_softBin=0, gh_useHog=0, gh_clipHog=0.2,
                nPerOct=8, nOctUp=0, nApprox=7, lambdas=[0.0, 0.11, 0.11] if color else [0.11, 0.11],
                pad=(0, 0), minDs=(size, size), smooth=1.0, concat=1,
                modelDs=(size, size), modelDsPad=(size, size), stride=4, cascThr=-1.0, cascCal=0.0,
                nms_type="maxg", nms_overlap=0.65, nms_ovrDnm="min")


def inria_opts():
    """INRIA-pedestrian-shaped stand-in: LUV + M + 6 H = 10 channels, modelDs (100,41), modelDsPad (128,64), pad (16,12)."""
    o = face_opts(80, True)
    o.update(colorSpace="luv", lambdas=[0.0, 0.11, 0.11], pad=(16, 12), minDs=(100, 41), modelDs=(100, 41), modelDsPad=(128, 64))
    return o


def n_channels(opts):
    nc = (1 if opts["colorSpace"] == "gray" else 3) if opts["color_enabled"] else 0
    return nc + (1 if opts["gm_enabled"] else 0) + (opts["gh_nOrients"] if opts["gh_enabled"] else 0)


# Empirical quantiles (5,25,50,75,95 %) of final pyramid channel values on the synthetic frames above,
# measured once with tools/calibrate_synth.py through the oracle.  Keys: channel kind.
_QUANT = {
    "gray": (0.0, 0.40, 0.50, 0.56, 0.69),
    "L": (0.0, 0.25, 0.28, 0.30, 0.33),
    "U": (0.21, 0.29, 0.33, 0.37, 0.54),
    "V": (0.32, 0.45, 0.50, 0.56, 0.74),
    "M": (0.0, 0.02, 0.17, 0.27, 0.69),
    "H": (0.0, 0.0001, 0.011, 0.042, 0.134),
}


def channel_kinds(opts):
    kinds = []
    if opts["color_enabled"]:
        kinds += ["gray"] if opts["colorSpace"] == "gray" else ["L", "U", "V"]
    kinds += ["M"] + ["H"] * opts["gh_nOrients"]
    return kinds


def make_classifier(opts, n_trees=2048, depth=2, seed=0, drift=None, gain=None, sigma=0.1, n_reject=None, confirm=0.02):
    """Random complete depth-`depth` trees in the reference's table layout:
    fids/thrs/child/hs/depth are [nTrees, 2^(depth+1)-1]; child is 1-based heap order with 0 at leaves.
    Leaf output = drift + gain * (fraction of 'feature >= threshold' turns on the path - 0.5) + N(0, sigma):
    textured windows (large M / H features) drift up and survive, flat ones are rejected after a few trees.
    """
    if drift is None or gain is None:  # 'fast-reject' operating points found with tools/calibrate_synth.py
        # face: ~13 trees/window on 1080p 'shapes' frames, (almost) no hits; inria-shape: ~7 trees/window, ~80 hits/frame
        d0, g0, nr0 = (-0.15, 0.28, 48) if opts["colorSpace"] == "luv" else (-0.15, 0.25, None)
        drift = d0 if drift is None else drift
        gain = g0 if gain is None else gain
        n_reject = nr0 if n_reject is None else n_reject
    rng = np.random.default_rng(seed)
    shrink = opts["shrink"]
    mH, mW = opts["modelDsPad"][0] // shrink, opts["modelDsPad"][1] // shrink
    kinds = channel_kinds(opts)
    n_nodes = (1 << (depth + 1)) - 1
    n_int = (1 << depth) - 1
    n_ftrs = len(kinds) * mH * mW
    fids = np.zeros((n_trees, n_nodes), np.uint32)
    thrs = np.zeros((n_trees, n_nodes), np.float32)
    child = np.zeros((n_trees, n_nodes), np.uint32)
    hs = np.zeros((n_trees, n_nodes), np.float32)
    dep = np.zeros((n_trees, n_nodes), np.uint32)
    fids[:, :n_int] = rng.integers(0, n_ftrs, (n_trees, n_int))
    z = fids[:, :n_int] // (mH * mW)
    qs = np.array([_QUANT[k] for k in kinds], np.float64)  # [nchn, 5]
    u = rng.uniform(0.25, 0.85, (n_trees, n_int))
    pos = u * 4
    lo = np.floor(pos).astype(int); fr = pos - lo
    thrs[:, :n_int] = (qs[z, lo] * (1 - fr) + qs[z, np.minimum(lo + 1, 4)] * fr).astype(np.float32)
    for k in range(n_nodes):
        dep[:, k] = int(np.floor(np.log2(k + 1)))
        if k < n_int:
            child[:, k] = 2 * k + 2  # 1-based index of the left child (toolbox convention)
    rng_confirm = np.random.default_rng(seed + 7919)  # its own stream: n_reject must not change the rejector trees
    for leaf in range(n_int, n_nodes):
        k, rights = leaf, 0
        while k > 0:
            rights += 1 if (k % 2 == 0) else 0  # even heap index = right child = feature >= threshold
            k = (k - 1) // 2
        hs[:, leaf] = (drift + gain * (rights / depth - 0.5) + sigma * rng.standard_normal(n_trees)).astype(np.float32)
        if n_reject is not None and n_reject < n_trees:  # later trees only confirm: survivors of the rejectors become hits
            hs[n_reject:, leaf] = (confirm + 0.25 * sigma * rng_confirm.standard_normal(n_trees - n_reject)).astype(np.float32)
    return dict(fids=fids, thrs=thrs, child=child, hs=hs, depth=dep, weights=np.zeros_like(hs), treeDepth=depth)


def make_variable_classifier(opts, n_trees=64, max_depth=3, seed=0, prune=0.35, **kw):
    """Variable-depth trees (treeDepth == 0, the reference's ParallelDetectionBody<T,0> path): complete trees of
    depth `max_depth` in which a random subset of internal nodes below the root is turned into leaves (child = 0)."""
    clf = make_classifier(opts, n_trees, max_depth, seed, **kw)
    rng = np.random.default_rng(seed + 1)
    n_int = (1 << max_depth) - 1
    child, hs = clf["child"], clf["hs"]
    for t in range(n_trees):
        for k in range(1, n_int):
            if rng.random() < prune:
                child[t, k] = 0
                hs[t, k] = np.float32(rng.normal(-0.05, 0.15))
    clf["treeDepth"] = 0
    return clf
