"""NumPy / torch-CPU / cv2 restatement of the post-model crown pipeline.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Each function cites
the reference file:line it follows (paths relative to ``/root/reference``).
Where the arithmetic lives in an absent third-party package the published
algorithm is restated and the function says "parity unpinned".

Pinned here against the reference's own code through ``oracle.refshim``:
``nms_bbox``, ``crown_stats_combined``, ``crown_stats_height``,
``crown_stats_ndvi``, ``containment``, ``ndvi_from_rgbi``, ``centroids``,
``process_features`` / ``process_geojson`` (tests/test_oracle_golden.py against
tests/golden/*.npz, which tests/golden/make_golden*.py produced by running the
reference's functions).
"""
from __future__ import annotations

import math

import numpy as np

from . import geom

# ============================================================================
# P2  paste + threshold  (detectron2, external; entered at prediction.py:183)
# ============================================================================


def scale_clip_boxes(boxes, in_hw, out_hw):
    """detectron2 ``detector_postprocess``: scale boxes from the network input
    size to the tile size, clip, and flag non-empty ones.

    boxes (N,4) float32 xyxy in network-input pixels.  Returns (boxes32, keep).
    parity unpinned (detectron2 absent); follows modeling/postprocessing.py and
    structures/boxes.py (scale / clip / nonempty) of detectron2 v0.6.
    """
    b = np.array(boxes, dtype=np.float32).reshape(-1, 4).copy()
    sx = np.float32(out_hw[1] / in_hw[1])   # python float scale applied to a float32 tensor
    sy = np.float32(out_hw[0] / in_hw[0])
    b[:, 0::2] *= sx
    b[:, 1::2] *= sy
    b[:, 0] = np.clip(b[:, 0], 0, np.float32(out_hw[1]))
    b[:, 2] = np.clip(b[:, 2], 0, np.float32(out_hw[1]))
    b[:, 1] = np.clip(b[:, 1], 0, np.float32(out_hw[0]))
    b[:, 3] = np.clip(b[:, 3], 0, np.float32(out_hw[0]))
    keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    return b, keep


def paste_window(box, img_h, img_w):
    """CPU-path window of ``_do_paste_mask`` (skip_empty=True, one instance per
    chunk): [floor(x0)-1 clamp 0, ceil(x1)+1 clamp W) x [floor(y0)-1, ceil(y1)+1)."""
    x0 = max(int(math.floor(float(box[0]))) - 1, 0)
    y0 = max(int(math.floor(float(box[1]))) - 1, 0)
    x1 = min(int(math.ceil(float(box[2]))) + 1, img_w)
    y1 = min(int(math.ceil(float(box[3]))) + 1, img_h)
    return x0, y0, x1, y1


def paste_probs(box, prob, img_h, img_w):
    """``_do_paste_mask`` for one instance on its CPU-path window, restated with
    torch CPU ops in the same order (detectron2 layers/mask_ops.py).

    box (4,) float32 xyxy in tile pixels, prob (M,M) float32 probabilities.
    Returns (values float32 (h,w), (x0,y0,x1,y1)).
    """
    import torch
    import torch.nn.functional as F

    x0i, y0i, x1i, y1i = paste_window(box, img_h, img_w)
    b = torch.as_tensor(np.asarray(box, dtype=np.float32)).reshape(1, 4)
    m = torch.as_tensor(np.asarray(prob, dtype=np.float32))[None, None]
    bx0, by0, bx1, by1 = torch.split(b, 1, dim=1)
    img_y = torch.arange(y0i, y1i, dtype=torch.float32) + 0.5
    img_x = torch.arange(x0i, x1i, dtype=torch.float32) + 0.5
    img_y = (img_y - by0) / (by1 - by0) * 2 - 1
    img_x = (img_x - bx0) / (bx1 - bx0) * 2 - 1
    gx = img_x[:, None, :].expand(1, img_y.size(1), img_x.size(1))
    gy = img_y[:, :, None].expand(1, img_y.size(1), img_x.size(1))
    grid = torch.stack([gx, gy], dim=3)
    out = F.grid_sample(m, grid, align_corners=False)
    return out[0, 0].numpy(), (x0i, y0i, x1i, y1i)


def paste_threshold(boxes, probs, img_h, img_w, threshold=0.5):
    """``paste_masks_in_image`` (CPU path): bool masks (N, H, W).

    Only for small cases (materialises N*H*W bools like the reference)."""
    n = len(boxes)
    out = np.zeros((n, img_h, img_w), dtype=bool)
    for i in range(n):
        v, (x0, y0, x1, y1) = paste_probs(boxes[i], probs[i], img_h, img_w)
        out[i, y0:y1, x0:x1] = v >= np.float32(threshold)
    return out


def _fmaf(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in
    float64, so one float64 add followed by a rounding reproduces fmaf."""
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64)
            + np.asarray(c, dtype=np.float64)).astype(np.float32)


def paste_probs_closed_form(box, prob, img_h, img_w):
    """Closed-form float32 restatement of ``_do_paste_mask`` + ATen's vectorised
    CPU ``grid_sample`` (bilinear, zeros padding, align_corners=False).  This is
    the operation order the CUDA kernel reproduces; it is bit-identical to
    ``paste_probs`` (torch CPU) -- tests/test_oracle_paste.py:

        g   = ((p + 0.5) - b0) / (b1 - b0) * 2 - 1      every op rounded to f32
        u   = fma(g + 1, M/2, -0.5)                      ATen unnormalize, contracted
        w   = u - floor(u);  e = 1 - w                   (x axis; n, s on the y axis)
        out = fma(se, n*w, fma(sw, n*e, fma(ne, s*w, nw * (s*e))))
              with zero padding outside [0, M)
    """
    f32 = np.float32
    M = prob.shape[0]
    x0i, y0i, x1i, y1i = paste_window(box, img_h, img_w)
    bx0, by0, bx1, by1 = [f32(v) for v in box]
    xs = (np.arange(x0i, x1i, dtype=np.float32) + f32(0.5))
    ys = (np.arange(y0i, y1i, dtype=np.float32) + f32(0.5))
    gx = (xs - bx0) / (bx1 - bx0) * f32(2) - f32(1)
    gy = (ys - by0) / (by1 - by0) * f32(2) - f32(1)
    ux = _fmaf(gx + f32(1), f32(M / 2), f32(-0.5))
    uy = _fmaf(gy + f32(1), f32(M / 2), f32(-0.5))
    xw = np.floor(ux); yn = np.floor(uy)
    w = ux - xw; e = f32(1) - w
    n = uy - yn; s = f32(1) - n
    ixw = xw.astype(np.int64); iyn = yn.astype(np.int64)
    pad = np.zeros((M + 2, M + 2), dtype=np.float32)
    pad[1:-1, 1:-1] = prob

    def at(iy, ix):
        iy = np.clip(iy + 1, 0, M + 1); ix = np.clip(ix + 1, 0, M + 1)
        return pad[iy[:, None], ix[None, :]]

    nwv = at(iyn, ixw); nev = at(iyn, ixw + 1); swv = at(iyn + 1, ixw); sev = at(iyn + 1, ixw + 1)
    cnw = s[:, None] * e[None, :]
    cne = s[:, None] * w[None, :]
    csw = n[:, None] * e[None, :]
    cse = n[:, None] * w[None, :]
    acc = nwv * cnw
    acc = _fmaf(nev, cne, acc)
    acc = _fmaf(swv, csw, acc)
    acc = _fmaf(sev, cse, acc)
    return acc.astype(np.float32), (x0i, y0i, x1i, y1i)


# ============================================================================
# P3  mask -> polygons   (prediction.py:197-265, utilities.py:182-207)
# ============================================================================


def find_contours(mask_u8):
    """prediction.py:232-234 -- the same OpenCV call."""
    import cv2
    contours, _ = cv2.findContours(np.ascontiguousarray(mask_u8, dtype=np.uint8), cv2.RETR_TREE,
                                   cv2.CHAIN_APPROX_SIMPLE)
    return [c.reshape(-1, 2) for c in contours]


def contour_to_ring_px(contour):
    """prediction.py:235-239: keep contours with >= 4 points, flatten, close."""
    if contour.size < 8:
        return None
    flat = contour.flatten().tolist()
    if flat[:2] != flat[-2:]:
        flat.extend(flat[:2])
    return flat


def xy_affine(transform, rows, cols):
    """utilities.py:182-207 ``xy_gpu``: corner convention, float64."""
    a, b, c, d, e, f = transform[:6]
    xs = np.asarray(cols); ys = np.asarray(rows)
    return a * xs + b * ys + c, d * xs + e * ys + f


def mask_to_polygons(mask_bool, transform):
    """prediction.py:216-251 for one instance: list of rings [(X,Y),...] in CRS
    coordinates (one entry per kept contour, holes included as own polygons)."""
    out = []
    for cnt in find_contours(mask_bool.astype(np.uint8)):
        flat = contour_to_ring_px(cnt)
        if flat is None:
            continue
        X, Y = xy_affine(transform, flat[1::2], flat[::2])
        out.append(list(zip(X.tolist(), Y.tolist())))
    return out


# ============================================================================
# P4  stitch: simplify + tile box filter   (helpers.py:265-319, 419-476)
# ============================================================================


def tile_filter_box(minx, miny, width, buffer, shift=1):
    """helpers.py:280-303 ``box_make``: note ``width`` is used for both axes."""
    return (minx - buffer + shift, miny - buffer + shift, minx + width + buffer - shift, miny + width + buffer - shift)


def stitch_tile(rings, scores, tile_box, simplify_tolerance=0.2):
    """helpers.py:443-472: Polygon -> simplify(tol, preserve_topology) -> keep
    iff within the shrunk tile box.  Returns (rings_out, scores_out)."""
    ro, so = [], []
    for ring, sc in zip(rings, scores):
        p = geom.Polygon(ring)
        if simplify_tolerance > 0:
            p = p.simplify(simplify_tolerance, preserve_topology=True)
        b = p.bounds
        if b[0] >= tile_box[0] and b[1] >= tile_box[1] and b[2] <= tile_box[2] and b[3] <= tile_box[3]:
            ro.append(p.exterior.coords)
            so.append(sc)
    return ro, so


# ============================================================================
# P5  NDVI  (helpers.py:862-896) and decimated reads (postprocessing.py:780-800)
# ============================================================================


def ndvi_from_rgbi(rgbi):
    """helpers.py:880-896: bands 0 (red) and 3 (nir), /255, float64."""
    red = rgbi[0].astype(np.float64) / 255.0
    nir = rgbi[3].astype(np.float64) / 255.0
    return (nir - red) / (nir + red + 1e-10)


def _triangle_coeffs(in_size, out_size):
    """Separable triangle-filter decimation weights (float64, normalised):
    support = max(scale, 1), centre = (i + 0.5) * scale.  This is the
    convolution form GDAL uses for RasterIO(bilinear) when down-sampling;
    GDAL is absent here, so this restatement defines the result
    (parity unpinned)."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 1.0 * fscale
    bounds = []
    coeffs = []
    for i in range(out_size):
        center = (i + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size)
        ws = []
        for x in range(xmin, xmax):
            t = abs((x - center + 0.5) / fscale)
            ws.append(1.0 - t if t < 1.0 else 0.0)
        tot = sum(ws)
        ws = [w / tot for w in ws] if tot != 0.0 else ws
        bounds.append((xmin, xmax - xmin))
        coeffs.append(ws)
    return bounds, coeffs


def decimate_bilinear(band, out_h, out_w):
    """postprocessing.py:782-786 / 793-796 (``Resampling.bilinear`` with an
    ``out_shape``).  Two passes (horizontal then vertical), float32 weights and
    accumulation in tap order; uint8 results are rounded half up and clamped,
    float32 results are left as accumulated.  Identity when the shape is unchanged."""
    h, w = band.shape
    if (out_h, out_w) == (h, w):
        return band.copy()
    bx, cx = _triangle_coeffs(w, out_w)
    by, cy = _triangle_coeffs(h, out_h)
    src = band.astype(np.float32)
    tmp = np.zeros((h, out_w), dtype=np.float32)
    for j in range(out_w):
        x0, n = bx[j]
        acc = np.zeros(h, dtype=np.float32)
        for k in range(n):
            acc = acc + src[:, x0 + k] * np.float32(cx[j][k])
        tmp[:, j] = acc
    out = np.zeros((out_h, out_w), dtype=np.float32)
    for i in range(out_h):
        y0, n = by[i]
        acc = np.zeros(out_w, dtype=np.float32)
        for k in range(n):
            acc = acc + tmp[y0 + k, :] * np.float32(cy[i][k])
        out[i, :] = acc
    if band.dtype == np.uint8:
        return np.clip(np.floor(out + np.float32(0.5)), 0, 255).astype(np.uint8)
    return out


def scaled_transform(transform, width, height, out_w, out_h):
    """postprocessing.py:787-788: ``src.transform * Affine.scale(W/w', H/h')``."""
    a, b, c, d, e, f = transform[:6]
    sx = width / out_w; sy = height / out_h
    return (a * sx + b * 0.0, a * 0.0 + b * sy, a * 0.0 + b * 0.0 + c,
            d * sx + e * 0.0, d * 0.0 + e * sy, d * 0.0 + e * 0.0 + f)


# ============================================================================
# P6  bbox NMS   (postprocessing.py:349-406, utilities.py:112-144)
# ============================================================================


def nms_bbox(bounds, conf, area, iou_threshold, area_threshold):
    """Returns a bool array ``removed`` (True = suppressed), index order = input
    order.  bounds (N,4) float64, conf (N,), area (N,).

    dtypes: bbox float32, confidence float16, area float16
    (postprocessing.py:367-369); python-scalar thresholds are weak, i.e. the
    comparisons happen in float32 / float16."""
    b = np.array([[np.float32(v) for v in row] for row in bounds], dtype=np.float32).reshape(-1, 4)
    c16 = np.asarray(conf, dtype=np.float16)
    a16 = np.asarray(area, dtype=np.float16)
    n = len(b)
    xA = np.maximum(b[:, 0][:, None], b[:, 0]); yA = np.maximum(b[:, 1][:, None], b[:, 1])
    xB = np.minimum(b[:, 2][:, None], b[:, 2]); yB = np.minimum(b[:, 3][:, None], b[:, 3])
    inter = np.maximum(0, xB - xA) * np.maximum(0, yB - yA)
    ar = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    union = ar[:, None] + ar - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / union
        am = np.abs(a16[:, None] - a16) / np.maximum(a16[:, None], a16)
    mask = (iou > iou_threshold) & (am < area_threshold)
    removed = np.zeros(n, dtype=bool)
    for i in range(n):
        if removed[i]:
            continue
        connected = np.append(np.where(mask[i])[0], i)
        best = connected[np.argmax(c16[connected])]
        for j in connected:
            if j != best:
                removed[j] = True
    return removed


def nms_bbox_sparse(bounds, conf, area, iou_threshold, area_threshold):
    """Same result as :func:`nms_bbox` without N^2 memory (pairs are found with
    a sort-and-sweep on x); for parity checks at sizes the dense form cannot hold."""
    b = np.asarray(bounds, dtype=np.float64).astype(np.float32).reshape(-1, 4)
    c16 = np.asarray(conf, dtype=np.float16)
    a16 = np.asarray(area, dtype=np.float16)
    n = len(b)
    order = np.argsort(b[:, 0], kind="stable")
    nbrs = [[] for _ in range(n)]
    thr_i = np.float32(iou_threshold); thr_a = np.float16(area_threshold)
    ar = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    xs = b[order, 0]
    for oi in range(n):
        i = order[oi]
        hi = np.searchsorted(xs, b[i, 2], side="right")
        cand = order[oi:hi]
        if len(cand) == 0:
            continue
        xA = np.maximum(b[i, 0], b[cand, 0]); yA = np.maximum(b[i, 1], b[cand, 1])
        xB = np.minimum(b[i, 2], b[cand, 2]); yB = np.minimum(b[i, 3], b[cand, 3])
        inter = np.maximum(np.float32(0), xB - xA) * np.maximum(np.float32(0), yB - yA)
        union = (ar[i] + ar[cand]) - inter
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter / union
            am = np.abs(a16[i] - a16[cand]) / np.maximum(a16[i], a16[cand])
        # NOTE union is computed as ar[i] + ar[j] - inter in both orders in the
        # dense form; float32 addition is commutative so iou[i,j] == iou[j,i].
        ok = (iou > thr_i) & (am < thr_a)
        for j in cand[ok]:
            if j == i:
                nbrs[i].append(i)       # mask[i, i] True when iou(i,i)=1 > thr
            else:
                nbrs[i].append(j); nbrs[j].append(i)
    removed = np.zeros(n, dtype=bool)
    for i in range(n):
        if removed[i]:
            continue
        connected = np.array(sorted(set(nbrs[i])) + [i], dtype=np.int64)
        best = connected[np.argmax(c16[connected])]
        for j in connected:
            if j != best:
                removed[j] = True
    return removed


# ============================================================================
# P7  per-crown raster statistics   (postprocessing.py:25-347, utilities.py:38-98)
# ============================================================================


def pad_polygons(rings):
    """postprocessing.py:509-540: NaN-padded (N,V) float32 vertex arrays."""
    v = max(len(r) for r in rings)
    px = np.full((len(rings), v), np.nan)
    py = np.full((len(rings), v), np.nan)
    for i, r in enumerate(rings):
        px[i, :len(r)] = [p[0] for p in r]
        py[i, :len(r)] = [p[1] for p in r]
    return px.astype(np.float32), py.astype(np.float32)


def centroids(px32, py32):
    """utilities.py:163-180: nanmean over the padded float32 vertices."""
    return np.stack((np.nanmean(px32, axis=1), np.nanmean(py32, axis=1)), axis=1)


def crown_circle(px32_i, py32_i):
    """postprocessing.py:285-301: bbox centre and max vertex distance, float32."""
    valid = ~np.isnan(px32_i) & ~np.isnan(py32_i)
    vx = px32_i[valid]; vy = py32_i[valid]
    cx = (vx.min() + vx.max()) / 2
    cy = (vy.min() + vy.max()) / 2
    dx = vx - cx; dy = vy - cy
    r = np.sqrt(dx ** 2 + dy ** 2).max()
    return cx, cy, r


def pixel_coords(transform, h, w, dtype):
    """postprocessing.py:259-266 / 60-67: corner-convention pixel coordinates of
    the whole raster, computed in float64 then cast (float32 in the combined
    and NDVI paths, float64 in the height-only path)."""
    a, b, c, d, e, f = transform[:6]
    rows, cols = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    x = a * cols + b * rows + c
    y = d * cols + e * rows + f
    return x.astype(dtype), y.astype(dtype)


def _stats_one(vals_h, xs, ys, inside_h, vals_n, inside_n):
    out = {}
    if inside_n is not None:
        sel = vals_n[inside_n]
        if sel.shape[0] == 0:
            out.update(ndvi_min=-1.0, ndvi_max=-1.0, ndvi_mean=-1.0, ndvi_var=-1.0)
        else:
            out.update(ndvi_min=sel[np.argmin(sel)], ndvi_max=sel[np.argmax(sel)],
                       ndvi_mean=np.mean(sel), ndvi_var=np.var(sel))
    if inside_h is not None:
        sel = vals_h[inside_h]
        if sel.size == 0:
            out.update(max_h=-1.0, hx=-1.0, hy=-1.0)
        else:
            k = np.argmax(sel)
            out.update(max_h=sel[k], hx=xs[inside_h][k], hy=ys[inside_h][k])
    return out


def crown_stats_combined(px32, py32, ndvi32, height32, transform):
    """postprocessing.py:221-347 ``get_metadata_within_polygon`` with the
    raster's own bounds (subset == whole raster).
    Height set: d^2 <= r^2; NDVI set: d^2 <= (0.5 r)^2; float32 coordinates.
    Returns dict of float32 arrays."""
    h, w = ndvi32.shape
    xs, ys = pixel_coords(transform, h, w, np.float32)
    xs = xs.ravel(); ys = ys.ravel()
    hv = height32.ravel(); nv = ndvi32.ravel()
    n = px32.shape[0]
    res = {k: np.zeros(n, dtype=np.float32) for k in
           ("max_h", "hx", "hy", "ndvi_min", "ndvi_max", "ndvi_mean", "ndvi_var")}
    for i in range(n):
        cx, cy, r = crown_circle(px32[i], py32[i])
        d2 = (xs - cx) ** 2 + (ys - cy) ** 2
        inside_n = d2 <= (r * 0.5) ** 2
        inside_h = d2 <= r ** 2
        for k, v in _stats_one(hv, xs, ys, inside_h, nv, inside_n).items():
            res[k][i] = v
    return res


def crown_stats_height(px32, py32, height32, transform):
    """postprocessing.py:25-115 ``get_height_within_polygon``: float64 pixel
    coordinates (:67), full radius."""
    h, w = height32.shape
    xs, ys = pixel_coords(transform, h, w, np.float64)
    xs = xs.ravel(); ys = ys.ravel(); hv = height32.ravel()
    n = px32.shape[0]
    res = {k: np.zeros(n, dtype=np.float32) for k in ("max_h", "hx", "hy")}
    for i in range(n):
        cx, cy, r = crown_circle(px32[i], py32[i])
        inside = (xs - cx) ** 2 + (ys - cy) ** 2 <= r ** 2
        for k, v in _stats_one(hv, xs, ys, inside, None, None).items():
            res[k][i] = v
    return res


def crown_stats_ndvi(px32, py32, ndvi32, transform):
    """postprocessing.py:117-219 ``get_ndvi_within_polygon``: float32 pixel
    coordinates (:160), FULL radius (:195)."""
    h, w = ndvi32.shape
    xs, ys = pixel_coords(transform, h, w, np.float32)
    xs = xs.ravel(); ys = ys.ravel(); nv = ndvi32.ravel()
    n = px32.shape[0]
    res = {k: np.zeros(n, dtype=np.float32) for k in ("ndvi_min", "ndvi_max", "ndvi_mean", "ndvi_var")}
    for i in range(n):
        cx, cy, r = crown_circle(px32[i], py32[i])
        inside = (xs - cx) ** 2 + (ys - cy) ** 2 <= r ** 2
        for k, v in _stats_one(None, xs, ys, None, nv, inside).items():
            res[k][i] = v
    return res


def crown_height_summary(px32, py32, height32, transform, q=95.0):
    """Optional nDSM summary per crown (north_star "min / max / mean / percentile"; the reference keeps the maximum
    only): over the pixel set of ``get_height_within_polygon`` (postprocessing.py:25-115: float64 pixel
    coordinates, full radius) -> (N,4) float32 [min, mean, numpy "linear" percentile q, pixel count]; -1 for an
    empty set.  Mean and percentile are evaluated in float64 on the float32 values."""
    h, w = height32.shape
    xs, ys = pixel_coords(transform, h, w, np.float64)
    xs = xs.ravel(); ys = ys.ravel(); hv = height32.ravel()
    out = np.zeros((px32.shape[0], 4), dtype=np.float32)
    for i in range(px32.shape[0]):
        cx, cy, r = crown_circle(px32[i], py32[i])
        sel = hv[(xs - cx) ** 2 + (ys - cy) ** 2 <= r ** 2]
        if sel.size == 0:
            out[i] = (-1, -1, -1, 0)
            continue
        s64 = sel.astype(np.float64)
        out[i] = (sel.min(), s64.mean(), np.percentile(s64, q), sel.size)
    return out


# ---- windowed forms (test infrastructure for FULL-SIZE scenes) ---------------------------------
# The three functions above follow the reference literally: every crown is tested against EVERY raster
# pixel (O(N * P): 10^4 crowns x 10^8 pixels at BASELINE config 2).  The forms below evaluate the same
# arithmetic on a conservative pixel window around the crown's circle only.  The selected pixels, their
# row-major order (which fixes numpy's pairwise float32 mean / variance and the first arg-max) and every
# value are identical as long as the window covers the circle, which a margin of `_WINDOW_MARGIN` metres
# (far above the float32 rounding of UTM coordinates, <= 0.5 m) guarantees for axis-aligned transforms;
# tests/test_oracle_windowed.py checks them against the literal forms.
_WINDOW_MARGIN = 3.0


def _crown_window(cx, cy, r, transform, h, w):
    a, b, c, d, e, f = transform[:6]
    if b != 0.0 or d != 0.0 or a == 0.0 or e == 0.0 or not np.isfinite([cx, cy, r]).all():
        return 0, h, 0, w
    m = float(r) + _WINDOW_MARGIN
    xa, xb = (float(cx) - m - c) / a, (float(cx) + m - c) / a
    ya, yb = (float(cy) - m - f) / e, (float(cy) + m - f) / e
    c0 = int(max(np.floor(min(xa, xb)) - 2, 0)); c1 = int(min(np.ceil(max(xa, xb)) + 3, w))
    r0 = int(max(np.floor(min(ya, yb)) - 2, 0)); r1 = int(min(np.ceil(max(ya, yb)) + 3, h))
    return r0, max(r1, r0), c0, max(c1, c0)


def _window_coords(transform, r0, r1, c0, c1, dtype):
    a, b, c, d, e, f = transform[:6]
    rows, cols = np.meshgrid(np.arange(r0, r1), np.arange(c0, c1), indexing="ij")
    x = a * cols + b * rows + c
    y = d * cols + e * rows + f
    return x.astype(dtype).ravel(), y.astype(dtype).ravel()


def crown_stats_combined_windowed(px32, py32, ndvi32, height32, transform):
    h, w = ndvi32.shape
    n = px32.shape[0]
    res = {k: np.zeros(n, dtype=np.float32) for k in
           ("max_h", "hx", "hy", "ndvi_min", "ndvi_max", "ndvi_mean", "ndvi_var")}
    for i in range(n):
        cx, cy, r = crown_circle(px32[i], py32[i])
        r0, r1, c0, c1 = _crown_window(cx, cy, r, transform, h, w)
        xs, ys = _window_coords(transform, r0, r1, c0, c1, np.float32)
        d2 = (xs - cx) ** 2 + (ys - cy) ** 2
        inside_n = d2 <= (r * 0.5) ** 2
        inside_h = d2 <= r ** 2
        for k, v in _stats_one(height32[r0:r1, c0:c1].ravel(), xs, ys, inside_h, ndvi32[r0:r1, c0:c1].ravel(),
                               inside_n).items():
            res[k][i] = v
    return res


def crown_stats_height_windowed(px32, py32, height32, transform):
    h, w = height32.shape
    n = px32.shape[0]
    res = {k: np.zeros(n, dtype=np.float32) for k in ("max_h", "hx", "hy")}
    for i in range(n):
        cx, cy, r = crown_circle(px32[i], py32[i])
        r0, r1, c0, c1 = _crown_window(cx, cy, r, transform, h, w)
        xs, ys = _window_coords(transform, r0, r1, c0, c1, np.float64)
        inside = (xs - cx) ** 2 + (ys - cy) ** 2 <= r ** 2
        for k, v in _stats_one(height32[r0:r1, c0:c1].ravel(), xs, ys, inside, None, None).items():
            res[k][i] = v
    return res


def crown_stats_ndvi_windowed(px32, py32, ndvi32, transform):
    h, w = ndvi32.shape
    n = px32.shape[0]
    res = {k: np.zeros(n, dtype=np.float32) for k in ("ndvi_min", "ndvi_max", "ndvi_mean", "ndvi_var")}
    for i in range(n):
        cx, cy, r = crown_circle(px32[i], py32[i])
        r0, r1, c0, c1 = _crown_window(cx, cy, r, transform, h, w)
        xs, ys = _window_coords(transform, r0, r1, c0, c1, np.float32)
        inside = (xs - cx) ** 2 + (ys - cy) ** 2 <= r ** 2
        for k, v in _stats_one(None, xs, ys, None, ndvi32[r0:r1, c0:c1].ravel(), inside).items():
            res[k][i] = v
    return res


# ============================================================================
# opt-in mask-IoU cleaner (the rule of clean_crowns, helpers.py:602-701, on pixel masks)
# ============================================================================


def mask_iou_clean(masks, origins, scores, iou_threshold=0.7, confidence=0.2):
    """``clean_crowns`` (helpers.py:602-701; detectree2's, never called by the reference) restated with the
    PIXEL IoU of boolean masks: for every crown, among the crowns whose IoU with it exceeds the threshold
    (itself included) take the one with the highest confidence (first of equals, as ``nlargest(1, field)``);
    the crown survives only if that one coincides with it (IoU == 1) and its confidence exceeds ``confidence``.
    masks: list of 2-D bool arrays, origins: their (x0, y0) on the image's pixel grid.  Returns (keep, match)."""
    n = len(masks)
    area = [int(m.sum()) for m in masks]
    keep = np.zeros(n, dtype=bool)
    match = np.full(n, -1, dtype=np.int64)
    boxes = [(int(o[0]), int(o[1]), int(o[0]) + m.shape[1], int(o[1]) + m.shape[0]) for m, o in zip(masks, origins)]
    for i in range(n):
        if area[i] == 0:
            continue
        best, best_conf, best_eq = -1, 0.0, False
        for j in range(n):
            x0, y0 = max(boxes[i][0], boxes[j][0]), max(boxes[i][1], boxes[j][1])
            x1, y1 = min(boxes[i][2], boxes[j][2]), min(boxes[i][3], boxes[j][3])
            if x0 >= x1 or y0 >= y1 or area[j] == 0:
                continue
            a = masks[i][y0 - boxes[i][1]:y1 - boxes[i][1], x0 - boxes[i][0]:x1 - boxes[i][0]]
            b = masks[j][y0 - boxes[j][1]:y1 - boxes[j][1], x0 - boxes[j][0]:x1 - boxes[j][0]]
            inter = int((a & b).sum())
            uni = area[i] + area[j] - inter
            if not np.float32(inter) / np.float32(uni) > np.float32(iou_threshold):
                continue
            cj = float(scores[j])
            if best < 0 or cj > best_conf:
                best, best_conf, best_eq = j, cj, inter == uni
        match[i] = best
        keep[i] = best >= 0 and best_eq and best_conf > float(np.float32(confidence))
    return keep, match


# ============================================================================
# P8  containment   (postprocessing.py:408-476)
# ============================================================================


def containment_chunked(bounds, threshold, chunk=512):
    """:func:`containment` evaluated over row blocks of the N x N matrices (same float32 arithmetic per
    element; for crown tables whose N x N temporaries do not fit comfortably)."""
    b = np.asarray(bounds, dtype=np.float32).reshape(-1, 4)
    n = b.shape[0]
    inner = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    rmax = np.full(n, -np.inf, dtype=np.float32)
    any_c = np.zeros(n, dtype=bool)
    num = np.zeros(n, dtype=np.int64)
    has_nan = np.zeros(n, dtype=bool)
    for o0 in range(0, n, chunk):
        o = b[o0:o0 + chunk]
        ix0 = np.maximum(o[:, 0][:, None], b[:, 0][None, :]); iy0 = np.maximum(o[:, 1][:, None], b[:, 1][None, :])
        ix1 = np.minimum(o[:, 2][:, None], b[:, 2][None, :]); iy1 = np.minimum(o[:, 3][:, None], b[:, 3][None, :])
        inter = np.maximum(0, ix1 - ix0) * np.maximum(0, iy1 - iy0)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = inter / inner[None, :]
        isc = ratio >= threshold
        k = np.arange(o.shape[0])
        isc[k, o0 + k] = False
        has_nan |= np.isnan(ratio).any(axis=0)
        rmax = np.fmax(rmax, np.nanmax(np.where(np.isnan(ratio), -np.inf, ratio), axis=0))
        any_c |= isc.any(axis=0)
        num[o0:o0 + chunk] = isc.sum(axis=1)
    rmax = np.where(has_nan, np.float32(np.nan), rmax)
    return rmax, any_c, num


def containment(bounds, threshold):
    """Returns (containment_ratio (N,) f32, is_contained (N,) bool,
    num_contained (N,) int).  ratio[o,i] = area(o & i)/area(i) in float32."""
    b = np.asarray(bounds, dtype=np.float32).reshape(-1, 4)
    n = b.shape[0]
    ix0 = np.maximum(b[:, 0][:, None], b[:, 0][None, :]); iy0 = np.maximum(b[:, 1][:, None], b[:, 1][None, :])
    ix1 = np.minimum(b[:, 2][:, None], b[:, 2][None, :]); iy1 = np.minimum(b[:, 3][:, None], b[:, 3][None, :])
    iw = np.maximum(0, ix1 - ix0); ih = np.maximum(0, iy1 - iy0)
    inter = iw * ih
    inner = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = inter / inner[None, :]
    isc = ratio >= threshold
    isc[np.arange(n), np.arange(n)] = False
    num = isc.sum(axis=1)
    return ratio.max(axis=0), isc.any(axis=0), num


# ============================================================================
# stage drivers (the order of operations of the reference's own drivers)
# ============================================================================


def predict_stage(det, tiles, simplify_tolerance=0.2, shift=1, mask_threshold=0.5, paste="closed_form"):
    """prediction.py:178-265 + helpers.py:419-554 for one image.

    det: object with boxes_net, scores, probs, inst_tile, tile_dims (treedetection_b200.synth.Detections
    layout); tiles: the tiles JSON dict in tile order.  Returns (rings, conf): the crown
    table of ``geojson_predictions/<image>.gpkg`` in canonical order (tile order of the
    tiles JSON, instance order = model output order, contour order = cv2)."""
    tile_ids = list(tiles.keys())
    rings_out, conf_out = [], []
    paste_fn = paste_probs_closed_form if paste == "closed_form" else paste_probs
    n = len(det.scores)
    i = 0
    for t, tid in enumerate(tile_ids):
        th, tw, nh, nw = [int(v) for v in det.tile_dims[t]]
        j = i
        while j < n and det.inst_tile[j] == t:
            j += 1
        rings_t, scores_t = [], []
        if j > i:
            boxes, keep = scale_clip_boxes(det.boxes_net[i:j], (nh, nw), (th, tw))
            tf = tiles[tid]["transform"]
            for k in range(i, j):
                if not keep[k - i]:
                    continue
                v, (x0, y0, x1, y1) = paste_fn(boxes[k - i], det.probs[k], th, tw)
                mask = np.zeros((th, tw), dtype=bool)
                mask[y0:y1, x0:x1] = v >= np.float32(mask_threshold)
                for ring in mask_to_polygons(mask, tf):
                    rings_t.append(ring)
                    scores_t.append(float(det.scores[k]))
        parts = [int(p) for p in tid.split("_")[-5:]]
        box = tile_filter_box(parts[0], parts[1], parts[2], parts[3], shift)
        ro, so = stitch_tile(rings_t, scores_t, box, simplify_tolerance)
        rings_out.extend(ro)
        conf_out.extend(so)
        i = j
    return rings_out, conf_out


def near_border(b, bounds, eps):
    """helpers.py:501-522."""
    left, bottom, right, top = bounds
    return (b[0] < left + eps) or (b[2] > right - eps) or (b[1] < bottom + eps) or (b[3] > top - eps)


def post_process(rings, conf, ndvi32, ndvi_transform, ndvi_bounds, height32, height_transform, height_bounds,
                 pixel_x, pixel_y, cfg, large=False):
    """postprocessing.py:722-809 (process_geojson) + :478-720 (process_features), restated
    on arrays.  ``cfg``: dict with the config.yml keys.  Returns a list of feature dicts
    {poly_id, Confidence_score, Area, TreeHeight, Centroid, is_contained, num_contained,
    coords} in output order, plus a debug dict.  ``large``: full-size scenes -- the sparse NMS, the
    windowed statistics and the chunked containment (each proven equal to the literal form in the tests)
    instead of the N x N / N x P literal forms."""
    f32 = np.float32
    # 1. confidence filter, ids, simplify(2) area, area range
    feats = [(r, c) for r, c in zip(rings, conf) if c is not None and float(c) >= cfg["confidence_threshold"]]
    areas = [geom.Polygon(r).simplify(2).area for r, _ in feats]
    ids = list(range(len(feats)))
    debug = {"area_after_conf": list(areas)}
    if not feats:
        return [], debug
    keep = [i for i in ids if areas[i] >= cfg["area_threshold"]]
    keep = [i for i in keep if areas[i] <= 1000]
    debug["ids_after_area"] = list(keep)
    if not keep:
        return [], debug
    bounds = {i: geom.Polygon(feats[i][0]).bounds for i in keep}
    # 3. NMS
    removed = (nms_bbox_sparse if large else nms_bbox)([bounds[i] for i in keep], [feats[i][1] for i in keep],
                                                       [areas[i] for i in keep], cfg["iou_threshold"],
                                                       cfg["area_threshold"])
    F = [i for i, r in zip(keep, removed) if not r]
    debug["ids_after_nms"] = list(F)
    if not F:
        return [], debug
    # 4. process_features
    Frings = [feats[i][0] for i in F]
    px32, py32 = pad_polygons(Frings)
    cent = centroids(px32, py32)
    t_eq = all(abs(a - b) < 1e-5 for a, b in zip(height_transform[:6], ndvi_transform[:6]))
    b_eq = all(abs(a - b) < 1e-3 for a, b in zip(height_bounds, ndvi_bounds))
    s_comb, s_h, s_n = (crown_stats_combined_windowed, crown_stats_height_windowed, crown_stats_ndvi_windowed) \
        if large else (crown_stats_combined, crown_stats_height, crown_stats_ndvi)
    if t_eq and b_eq:
        st = s_comb(px32, py32, ndvi32, height32, ndvi_transform)
    else:
        st = dict(s_h(px32, py32, height32, height_transform))
        st.update(s_n(px32, py32, ndvi32, ndvi_transform))
    heights, mean_ndvi, var_ndvi = st["max_h"], st["ndvi_mean"], st["ndvi_var"]
    debug.update(stats=st, centroid=cent, combined=bool(t_eq and b_eq))
    pre = []
    rows, cols = ndvi32.shape
    for k, i in enumerate(F):
        b = bounds[i]
        if cfg["use_overlap"]:
            if near_border(b, ndvi_bounds, 1.0):
                continue
            vmh = ((cfg["tile_height"] + 2 * cfg["buffer"]) * cfg["overlapping_tiles_height"]) * pixel_y
            hmw = ((cfg["tile_width"] + 2 * cfg["buffer"]) * cfg["overlapping_tiles_width"]) * pixel_x
            if not (rows == vmh or cols == hmw):
                right_border = ndvi_bounds[2] - hmw / 2.0
                left_border = ndvi_bounds[0] + hmw / 2.0
                top_border = ndvi_bounds[3] - vmh / 2.0
                bottom_border = ndvi_bounds[1] + vmh / 2.0
                if top_border < b[1] or bottom_border > b[3] or left_border > b[2] or right_border < b[0]:
                    continue
        if heights[k] < f32(cfg["height_threshold"]) and heights[k] > f32(-1.0):
            continue
        if (mean_ndvi[k] < f32(cfg["ndvi_mean_threshold"]) or var_ndvi[k] > f32(cfg["ndvi_var_threshold"])) \
                and mean_ndvi[k] > f32(-1.0):
            continue
        pre.append(k)
    b32 = np.array([bounds[i] for i in F], dtype=np.float32)
    ratio, is_c, num_c = (containment_chunked if large else containment)(b32, cfg["containment_threshold"])
    debug.update(pre=list(pre), is_contained=is_c, num_contained=num_c)
    contained_idx = [k for k in range(len(F)) if is_c[k]]
    selected = []
    for rank, k in enumerate(pre):
        nc = int(num_c[k])
        if nc >= 3:
            continue
        elif nc == 2:
            continue          # postprocessing.py:642-649: no branch appends
        elif nc == 1:
            o = contained_idx[0]
            if abs(mean_ndvi[k] - mean_ndvi[o]) > f32(0.05):
                if var_ndvi[rank] < var_ndvi[o]:
                    selected.append(k)
                else:
                    selected.append(o)
            elif areas[F[k]] > 0:
                selected.append(k)
        else:
            selected.append(k)
    out = []
    for k in selected:
        i = F[k]
        ring = feats[i][0]
        out.append({
            "poly_id": str(i), "Confidence_score": feats[i][1], "Area": areas[i],
            "TreeHeight": float(heights[k]), "Centroid": (float(cent[k][0]), float(cent[k][1])),
            "Diameter": 2 * (areas[i] / np.pi) ** 0.5,
            "is_contained": bool(is_c[k]), "num_contained": int(num_c[k]),
            "coords": [(round(x * 1000) / 1000, round(y * 1000) / 1000) for (x, y) in ring],
        })
    return out, debug


# ============================================================================
# P1  tile cut + normalise   (prediction.py:159-176)
# ============================================================================


def resize_shortest_edge(h, w, short=800, max_size=1333):
    """detectron2 ResizeShortestEdge.get_output_shape (test-time: 800 / 1333)."""
    size = short * 1.0
    scale = size / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def tile_cut_normalize(image, window):
    """prediction.py:164-171 for one tile.  image (bands,H,W) uint8 / uint16, window
    (col_off,row_off,w,h) as rasterio.mask(crop=True) cuts it for a pixel-aligned box.
    Returns float32 CHW, or None where the reference raises (uint16 data whose green band
    does not exceed 255: torch cannot interpolate uint16) and skips the tile."""
    import torch
    import torch.nn.functional as F
    from PIL import Image

    c0, r0, w, h = window
    out_img = image[:, r0:r0 + h, c0:c0 + w]
    rgb = np.dstack((out_img[2], out_img[1], out_img[0]))
    rgb_rescaled = 255 * rgb / 65535 if np.max(out_img[1]) > 255 else rgb
    nh, nw = resize_shortest_edge(h, w)
    if rgb_rescaled.dtype == np.uint8:
        pil = Image.fromarray(np.ascontiguousarray(rgb_rescaled))
        res = np.asarray(pil.resize((nw, nh), Image.BILINEAR))
    elif rgb_rescaled.dtype == np.float64:
        t = torch.from_numpy(np.ascontiguousarray(rgb_rescaled)).permute(2, 0, 1)[None]
        res = F.interpolate(t, (nh, nw), mode="bilinear", align_corners=False)[0].permute(1, 2, 0).numpy()
    else:
        return None
    return res.astype("float32").transpose(2, 0, 1)


# ============================================================================
# P0a  seam strips   (merging.py:34-110, helpers.py:1023-1085)
# ============================================================================


def seam_crop(a, b, axis, strip_w, strip_h):
    """Mosaic of two edge-adjacent rasters on the same grid (rasterio.merge of
    non-overlapping neighbours = concatenation), centre-cropped as crop_image does.
    Returns (strip, (left, top)) -- the window offset gives the strip's transform."""
    mosaic = np.concatenate([a, b], axis=2 if axis == 0 else 1)
    _, mh, mw = mosaic.shape
    left = max(mw // 2 - strip_w // 2, 0)
    top = max(mh // 2 - strip_h // 2, 0)
    return mosaic[:, top:top + strip_h, left:left + strip_w], (left, top)


# ============================================================================
# P10  forest-outline predicates   (helpers.py:795-811, preprocessing.py:67-96)
# ============================================================================
# shapely/GEOS are absent: the predicates are restated for single-ring polygons against the
# union of single-ring forest polygons (see treedetection_b200/csrc/forest_core.cuh for the
# definition and its two documented simplifications).  parity with GEOS: unpinned.


def locate_in_ring(p, ring):
    """1 inside, 0 boundary, -1 outside (GEOS RayCrossingCounter)."""
    crossings = 0
    for k in range(len(ring) - 1):
        p1, p2 = ring[k], ring[k + 1]
        if p1[0] < p[0] and p2[0] < p[0]:
            continue
        if p[0] == p2[0] and p[1] == p2[1]:
            return 0
        if p1[1] == p[1] and p2[1] == p[1]:
            if min(p1[0], p2[0]) <= p[0] <= max(p1[0], p2[0]):
                return 0
            continue
        if (p1[1] > p[1] and p2[1] <= p[1]) or (p2[1] > p[1] and p1[1] <= p[1]):
            sign = geom.orientation(p1[0], p1[1], p2[0], p2[1], p[0], p[1])
            if sign == 0:
                return 0
            if p2[1] < p1[1]:
                sign = -sign
            if sign > 0:
                crossings += 1
    return 1 if crossings & 1 else -1


def segments_touch(p1, p2, q1, q2):
    if not geom._env_intersects_seg(p1, p2, q1, q2):
        return False
    o1 = geom.orientation(p1[0], p1[1], p2[0], p2[1], q1[0], q1[1])
    o2 = geom.orientation(p1[0], p1[1], p2[0], p2[1], q2[0], q2[1])
    if (o1 > 0 and o2 > 0) or (o1 < 0 and o2 < 0):
        return False
    o3 = geom.orientation(q1[0], q1[1], q2[0], q2[1], p1[0], p1[1])
    o4 = geom.orientation(q1[0], q1[1], q2[0], q2[1], p2[0], p2[1])
    if (o3 > 0 and o4 > 0) or (o3 < 0 and o4 < 0):
        return False
    return True


def ring_intersects_ring(A, F):
    for i in range(len(A) - 1):
        for j in range(len(F) - 1):
            if segments_touch(A[i], A[i + 1], F[j], F[j + 1]):
                return True
    if A and locate_in_ring(A[0], F) >= 0:
        return True
    if F and locate_in_ring(F[0], A) >= 0:
        return True
    return False


def _split_params(a0, a1, q1, q2):
    if not segments_touch(a0, a1, q1, q2):
        return []
    dx = a1[0] - a0[0]; dy = a1[1] - a0[1]
    ex = q2[0] - q1[0]; ey = q2[1] - q1[1]
    den = dx * ey - dy * ex
    if den != 0.0:
        return [((q1[0] - a0[0]) * ey - (q1[1] - a0[1]) * ex) / den]
    len2 = dx * dx + dy * dy
    if len2 == 0.0:
        return []
    return [((q1[0] - a0[0]) * dx + (q1[1] - a0[1]) * dy) / len2,
            ((q2[0] - a0[0]) * dx + (q2[1] - a0[1]) * dy) / len2]


def _as_polygon(F):
    """a forest polygon is a list of rings [shell, hole, ...]; a bare ring is a polygon without holes"""
    if len(F) and isinstance(F[0][0], (int, float, np.floating, np.integer)):
        return [F]
    return F


def locate_in_polygon(p, poly):
    """1 interior, 0 boundary (shell or hole), -1 exterior (outside the shell or strictly inside a hole)."""
    s = locate_in_ring(p, poly[0])
    if s <= 0:
        return s
    for h in poly[1:]:
        k = locate_in_ring(p, h)
        if k == 1:
            return -1
        if k == 0:
            return 0
    return 1


def ring_intersects_polygon(A, poly):
    """GEOS ``intersects`` of the areal ring A with a polygon with holes (closed sets, touching counts)."""
    for R in poly:
        for i in range(len(A) - 1):
            for j in range(len(R) - 1):
                if segments_touch(A[i], A[i + 1], R[j], R[j + 1]):
                    return True
    if len(A) and locate_in_polygon(A[0], poly) >= 0:
        return True
    if len(poly[0]) and locate_in_ring(poly[0][0], A) >= 0:
        return True
    return False


def ring_within_union(A, forest):
    """GEOS ``within`` against the union of the forest polygons without building the union: every piece of A's
    boundary (edges split at their crossings with shells and holes) has its midpoint in some closed polygon,
    and no hole has a vertex strictly inside A."""
    forest = [_as_polygon(F) for F in forest]
    if len(A) < 2 or not forest:
        return False
    for i in range(len(A) - 1):
        a0, a1 = A[i], A[i + 1]
        ts = [0.0, 1.0]
        for P in forest:
            for R in P:
                for j in range(len(R) - 1):
                    for t in _split_params(a0, a1, R[j], R[j + 1]):
                        if 0.0 < t < 1.0:
                            ts.append(t)
        ts.sort()
        for k in range(len(ts) - 1):
            if not ts[k + 1] > ts[k]:
                continue
            tm = (ts[k] + ts[k + 1]) / 2.0
            m = (a0[0] + (a1[0] - a0[0]) * tm, a0[1] + (a1[1] - a0[1]) * tm)
            if not any(locate_in_polygon(m, P) >= 0 for P in forest):
                return False
    for P in forest:
        for h in P[1:]:
            if any(locate_in_ring(v, A) == 1 for v in h[:-1]):
                return False
    return True


def _bounds(ring):
    xs = [p[0] for p in ring]; ys = [p[1] for p in ring]
    return min(xs), min(ys), max(xs), max(ys)


def forest_predicates(rings, forest):
    """(intersects, within) of every ring against the union of the forest polygons (rings or [shell, holes...]
    lists); polygons are pre-filtered by bounding-box overlap of their shells exactly as the kernel does."""
    forest = [_as_polygon(F) for F in forest]
    fb = [_bounds(P[0]) for P in forest]
    inter, within = [], []
    for A in rings:
        ab = _bounds(A)
        cand = [P for P, b in zip(forest, fb) if not (ab[0] > b[2] or ab[2] < b[0] or ab[1] > b[3] or ab[3] < b[1])]
        hit = any(ring_intersects_polygon(A, P) for P in cand)
        inter.append(hit)
        within.append(bool(hit and ring_within_union(A, cand)))
    return np.array(inter), np.array(within)


def tile_flags(tile_bounds, buffered_box_ring, forest):
    """preprocessing.py:67-96 for one tile: (only_forest, only_urban).  ``tile_bounds`` =
    un-buffered tile box (bbox prefilter, strict inequalities), ``buffered_box_ring`` = the
    buffered tile box as a ring (the geometry tests)."""
    forest = [_as_polygon(F) for F in forest]
    minx, miny, maxx, maxy = tile_bounds
    cand = [P for P in forest
            if (_bounds(P[0])[2] > minx and _bounds(P[0])[0] < maxx and _bounds(P[0])[3] > miny and
                _bounds(P[0])[1] < maxy)]
    if not cand:
        return False, True
    hit = [P for P in cand if ring_intersects_polygon(buffered_box_ring, P)]
    if not hit:
        return False, True
    return bool(ring_within_union(buffered_box_ring, cand)), False


def fuse(urban, forest_crowns, forest):
    """helpers.py:795-811: indices of the forest-model crowns that intersect the forest union
    and of the urban-model crowns that are NOT within it (forest first in the output)."""
    fi, _ = forest_predicates(forest_crowns, forest)
    _, uw = forest_predicates(urban, forest)
    return np.nonzero(fi)[0], np.nonzero(~uw)[0]
