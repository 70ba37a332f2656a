"""CPU oracle for the detection hot path -- TEST INFRASTRUCTURE ONLY.

A ctypes front-end to ``oracle/yolo_oracle.c`` (a plain-C restatement of
``models/yolo_loss.py``, ``utils/box.py``, ``utils/iou.py`` and the arithmetic of
``torchvision.ops.nms``; every C function cites the reference file:line it
follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; the product
(``mobilenet_yolo_pytorch_b200``) never does.

Parity status: **pinned** -- ``tests/test_oracle_golden.py`` checks it against
fixtures in ``tests/golden/`` generated from the reference's own Python by
``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = os.path.join(_HERE, "yolo_oracle.c")
_lock = threading.Lock()
_lib = None

NMS_IOU_THRESHOLD = 0.45  # utils/box.py:28


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc, a second or two).  Returns the .so path."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
    base = [cc, "-O2", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-fvisibility=hidden", "-shared", "-o", _SO, _SRC]
    try:
        subprocess.check_call(base[:4] + ["-fopenmp"] + base[4:] + ["-lm"], stderr=subprocess.DEVNULL)
    except (subprocess.CalledProcessError, OSError):
        subprocess.check_call(base + ["-lm"])  # scalar build if libgomp is missing
    return _SO


def lib() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            _lib = C.CDLL(build())
            _lib.oracle_get_max_threads.restype = C.c_int
            _lib.oracle_target_loss.restype = C.c_int
    return _lib


def set_threads(n: int) -> None:
    lib().oracle_set_threads(C.c_int(int(n)))


def max_threads() -> int:
    return int(lib().oracle_get_max_threads())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def scaled_anchors(anchors, img_size) -> np.ndarray:
    """yolo_loss.py:214 -- python-double division, then rounded to fp32 (:67-68)."""
    return np.array([[aw / img_size[0], ah / img_size[1]] for aw, ah in anchors], dtype=np.float64).astype(np.float32)


def decode_head(head, anchor_wh, num_classes: int, conf_thr: float):
    """YOLOLoss.get_pred_boxes (yolo_loss.py:180-204).  ``anchor_wh`` is this
    head's (A,2) scaled anchors.  Returns (list of (n_b,7) arrays, list of id arrays)."""
    x = _f32(head)
    N, ch, H, W = x.shape
    A = ch // (5 + num_classes)
    cells = A * H * W
    rows = np.zeros((N, cells, 7), np.float32)
    count = np.zeros(N, np.int32)
    ids = np.zeros((N, cells), np.int32)
    aw = _f32(anchor_wh)
    lib().oracle_decode_head(_p(x), N, A, num_classes, H, W, _p(aw), C.c_float(np.float32(conf_thr)), _p(rows),
                             _p(count), _p(ids))
    return [rows[b, :count[b]].copy() for b in range(N)], [ids[b, :count[b]].copy() for b in range(N)]


def nms(cands, num_classes: int, iou_thr: float = NMS_IOU_THRESHOLD):
    """utils.box.nms (box.py:11-31) on per-image concatenated candidates
    (list of (n_b,7)).  Returns (list of (k_b,7), list of kept index arrays)."""
    N = len(cands)
    K = max([len(c) for c in cands] + [1])
    rows = np.zeros((N, K, 7), np.float32)
    count = np.zeros(N, np.int32)
    for b, c in enumerate(cands):
        rows[b, :len(c)] = c
        count[b] = len(c)
    out = np.zeros_like(rows)
    oc = np.zeros(N, np.int32)
    oi = np.zeros((N, K), np.int32)
    lib().oracle_nms(_p(rows), _p(count), N, K, num_classes, C.c_double(iou_thr), _p(out), _p(oc), _p(oi))
    return [out[b, :oc[b]].copy() for b in range(N)], [oi[b, :oc[b]].copy() for b in range(N)]


def decode_nms_padded(head0, head1, anchor_wh2, num_classes: int, conf_thr: float,
                      iou_thr: float = NMS_IOU_THRESHOLD, want_cand: bool = False):
    """Fixed-stride form used for timing and full-size checks: returns
    (out (N,K,7), out_count (N,), out_idx (N,K)[, cand, cand_count, cand_ids])."""
    h0, h1 = _f32(head0), _f32(head1)
    N, ch, H0, W0 = h0.shape
    _, _, H1, W1 = h1.shape
    A = ch // (5 + num_classes)
    K = A * H0 * W0 + A * H1 * W1
    out = np.zeros((N, K, 7), np.float32)
    oc = np.zeros(N, np.int32)
    oi = np.zeros((N, K), np.int32)
    cand = np.zeros((N, K, 7), np.float32) if want_cand else None
    cc = np.zeros(N, np.int32) if want_cand else None
    ci = np.zeros((N, K), np.int32) if want_cand else None
    aw = _f32(anchor_wh2).reshape(2, A, 2)
    lib().oracle_decode_nms(_p(h0), _p(h1), N, A, num_classes, H0, W0, H1, W1, _p(aw),
                            C.c_float(np.float32(conf_thr)), C.c_double(iou_thr), _p(out), _p(oc), _p(oi),
                            _p(cand), _p(cc), _p(ci))
    if want_cand:
        return out, oc, oi, cand, cc, ci
    return out, oc, oi


def decode_nms(head0, head1, anchor_wh2, num_classes: int, conf_thr: float, iou_thr: float = NMS_IOU_THRESHOLD):
    """mbv2_yolo.py:158-160 inference branch.  Returns (list of (k_b,7) detections,
    list of kept global candidate ids (head1 ids offset by A*H0*W0))."""
    out, oc, oi, cand, cc, ci = decode_nms_padded(head0, head1, anchor_wh2, num_classes, conf_thr, iou_thr, True)
    dets = [out[b, :oc[b]].copy() for b in range(len(oc))]
    ids = [ci[b][oi[b, :oc[b]]].copy() for b in range(len(oc))]
    return dets, ids


def pairwise(set_1, set_2, mode: str = "iou") -> np.ndarray:
    """utils/iou.py: 'inter' = find_intersection, 'union' = find_union, 'iou' = find_jaccard_overlap."""
    a, b = _f32(set_1).reshape(-1, 4), _f32(set_2).reshape(-1, 4)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    m = {"inter": 0, "union": 1, "iou": 2}[mode]
    if out.size:
        lib().oracle_pairwise(_p(a), a.shape[0], _p(b), b.shape[0], m, _p(out))
    return out


def box_ciou(box1, box2):
    """YOLOLoss.box_ciou (yolo_loss.py:257-293): returns (iou - term, iou)."""
    o = np.zeros(2, np.float32)
    lib().oracle_box_ciou(_p(_f32(box1)), _p(_f32(box2)), _p(o))
    return float(o[0]), float(o[1])


def box_giou(box1, box2):
    """YOLOLoss.box_giou (yolo_loss.py:295-317)."""
    o = np.zeros(2, np.float32)
    lib().oracle_box_giou(_p(_f32(box1)), _p(_f32(box2)), _p(o))
    return float(o[0]), float(o[1])


SCALAR_NAMES = ("loss", "recall", "avg_iou", "obj", "no_obj", "cls", "count_per_img", "l_dense", "l_iou", "sum_w",
                "n_assign", "sum_sq_w", "iou_num", "iou_wsum", "sum_conf", "n_recall")


def pack_targets(targets):
    """list[N] of (n_b,5) -> (G,5) fp32 rows + (N+1,) int32 offsets."""
    offs = np.zeros(len(targets) + 1, np.int32)
    rows = []
    for b, t in enumerate(targets):
        t = _f32(t).reshape(-1, 5)
        rows.append(t)
        offs[b + 1] = offs[b] + t.shape[0]
    gt = np.concatenate(rows, 0) if rows else np.zeros((0, 5), np.float32)
    if gt.shape[0] == 0:
        gt = np.zeros((1, 5), np.float32)  # keep a valid pointer
    return _f32(gt), offs


def target_loss(head, targets, anchors, mask, num_classes, img_size, ignore_threshold, iou_thresh, iou_weighting,
                want_dense: bool = False):
    """YOLOLoss.forward(input, targets) (yolo_loss.py:206-236) for one head.
    Returns dict(scalars..., assign (n,6) int32 rows (b,t,k,gj,gi,best_n), terms (n,3))."""
    x = _f32(head)
    N, ch, H, W = x.shape
    A = len(mask)
    sa = scaled_anchors(anchors, img_size)
    gt, offs = pack_targets(targets)
    G = int(offs[-1])
    max_assign = max(1, G * A)
    assign = np.zeros((max_assign, 6), np.int32)
    terms = np.zeros((max_assign, 3), np.float32)
    scal = np.zeros(16, np.float64)
    tg = np.zeros((N, A, H, W, num_classes + 1), np.float32) if want_dense else None
    wg = np.zeros((N, A, H, W, num_classes + 1), np.float32) if want_dense else None
    m = _i32(mask)
    n = lib().oracle_target_loss(_p(x), N, A, num_classes, H, W, _p(sa), sa.shape[0], _p(m), _p(gt), _p(offs),
                                 C.c_float(np.float32(ignore_threshold)), C.c_float(np.float32(iou_thresh)),
                                 C.c_float(np.float32(iou_weighting)), _p(scal), _p(assign), _p(terms), max_assign,
                                 _p(tg), _p(wg))
    if n < 0:
        raise IndexError("GT cell index out of range (the reference raises IndexError here, yolo_loss.py:149)")
    res = {k: float(v) for k, v in zip(SCALAR_NAMES, scal)}
    res["assign"] = assign[:n].copy()
    res["terms"] = terms[:n].copy()
    if want_dense:
        res["targets"], res["weights"] = tg, wg
    return res


def _ciou_value_and_grad(gt, pr):
    """v = iou - ciou_term of YOLOLoss.box_ciou(box1=gt, box2=pred) (yolo_loss.py:257-293) and dv/d(pred xyxy),
    float64, differentiating every operation of the reference's graph (alpha is NOT detached, :283;
    max/min/clamp pass the gradient to the selected operand)."""
    a1, b1, a2, b2 = [float(x) for x in gt]
    p1, q1, p2, q2 = [float(x) for x in pr]
    g = np.zeros
    # intersection (utils/iou.py:4-13)
    lox, loy, hix, hiy = max(a1, p1), max(b1, q1), min(a2, p2), min(b2, q2)
    iw_raw, ih_raw = hix - lox, hiy - loy
    iw, ih = max(iw_raw, 0.0), max(ih_raw, 0.0)
    d_iw = g(4)
    d_ih = g(4)
    if iw_raw >= 0:
        d_iw[2] = 1.0 if p2 < a2 else (0.5 if p2 == a2 else 0.0)
        d_iw[0] = -(1.0 if p1 > a1 else (0.5 if p1 == a1 else 0.0))
    if ih_raw >= 0:
        d_ih[3] = 1.0 if q2 < b2 else (0.5 if q2 == b2 else 0.0)
        d_ih[1] = -(1.0 if q1 > b1 else (0.5 if q1 == b1 else 0.0))
    inter = iw * ih
    d_inter = d_iw * ih + iw * d_ih
    area1 = (a2 - a1) * (b2 - b1)
    w2, h2 = p2 - p1, q2 - q1
    d_w2 = np.array([-1.0, 0, 1.0, 0])
    d_h2 = np.array([0, -1.0, 0, 1.0])
    area2 = w2 * h2
    d_area2 = d_w2 * h2 + w2 * d_h2
    union = area1 + area2 - inter
    d_union = d_area2 - d_inter
    iou = inter / union
    d_iou = (d_inter * union - inter * d_union) / (union * union)
    # enclosing box area (box_c :249-256, :264)
    cw, ch = max(a2, p2) - min(a1, p1), max(b2, q2) - min(b1, q1)
    d_cw = np.array([-(1.0 if p1 < a1 else (0.5 if p1 == a1 else 0.0)), 0, (1.0 if p2 > a2 else (0.5 if p2 == a2 else 0.0)), 0])
    d_ch = np.array([0, -(1.0 if q1 < b1 else (0.5 if q1 == b1 else 0.0)), 0, (1.0 if q2 > b2 else (0.5 if q2 == b2 else 0.0))])
    c = cw * ch
    d_c = d_cw * ch + cw * d_ch
    # centre distance (:269-272)
    dx, dy = (a2 + a1) / 2 - (p2 + p1) / 2, (b1 + b2) / 2 - (q1 + q2) / 2
    u = dx * dx + dy * dy
    d_u = np.array([-dx, -dy, -dx, -dy])
    dd = u / c
    d_dd = (d_u * c - u * d_c) / (c * c)
    # aspect-ratio term (:279-283); "ar_gt" is the PRED box's w/h in the reference
    w1, h1 = a2 - a1, b2 - b1
    k = 4.0 / (math.pi * math.pi)
    delta = math.atan(w2 / h2) - math.atan(w1 / h1)
    d_delta = (h2 * d_w2 - w2 * d_h2) / (h2 * h2 + w2 * w2)
    A = k * delta * delta
    d_A = 2.0 * k * delta * d_delta
    D = 1.0 - iou + A + 0.000001
    d_D = -d_iou + d_A
    f = A * A / D                      # alpha * ar_loss
    d_f = (2.0 * A * d_A * D - A * A * d_D) / (D * D)
    if c == 0.0:                       # :286-287 term = iou -> v = 0
        return 0.0, iou, np.zeros(4)
    return iou - (dd + f), iou, d_iou - d_dd - d_f


def target_loss_backward(head, targets, anchors, mask, num_classes, img_size, ignore_threshold, iou_thresh,
                         iou_weighting, grad_out: float = 1.0):
    """d loss / d input of YOLOLoss.forward(input, targets) (yolo_loss.py:206-236) as the reference's autograd
    graph defines it: the custom sigmoid (:15-32) passes gradients through unchanged, exp has its true
    derivative, only entries of `targets` overwritten with constants (weight 1) carry a gradient
    2 (o - t) w / sum(w) (:53-60), and the CIoU loss sum_i (v_i - 1)^2 / n_assign (:224, weights cancel)
    reaches tx, ty, tw, th of the assigned cells through the decoded box.  float64 maths, float32 result."""
    x = _f32(head)
    N, ch, H, W = x.shape
    A = len(mask)
    attrs = 5 + num_classes
    fwd = target_loss(x, targets, anchors, mask, num_classes, img_size, ignore_threshold, iou_thresh, iou_weighting,
                      want_dense=True)
    xv = x.reshape(N, A, attrs, H, W).astype(np.float64)
    grad = np.zeros_like(xv)
    out = 1.0 / (1.0 + np.exp(-xv[:, :, 4:]))                       # (N,A,C+1,H,W)
    tg = np.moveaxis(fwd["targets"].astype(np.float64), 4, 2)       # (N,A,H,W,C+1) -> (N,A,C+1,H,W)
    wg = np.moveaxis(fwd["weights"].astype(np.float64), 4, 2)
    sw = wg.sum()
    grad[:, :, 4:] = 2.0 * (out - tg) * wg / sw                     # identity through the custom sigmoid
    n = len(fwd["assign"])
    sa = scaled_anchors(anchors, img_size).astype(np.float64)
    for (b, t, k, gj, gi, _bn) in fwd["assign"]:
        tb = _f32(targets[b]).reshape(-1, 5)[t]
        cx, cy, w, h = [np.float32(v) for v in tb[1:5]]
        g1 = np.float32(cx - w / np.float32(2))
        g2 = np.float32(cy - h / np.float32(2))
        gtb = (g1, g2, np.float32(w + g1), np.float32(h + g2))      # :112-113 in fp32 like the reference
        tx, ty, tw, th = xv[b, k, 0:4, gj, gi]
        sx, sy = 1.0 / (1.0 + math.exp(-tx)), 1.0 / (1.0 + math.exp(-ty))
        bw, bh = math.exp(tw) * sa[mask[k]][0], math.exp(th) * sa[mask[k]][1]
        pcx, pcy = (sx + gi) / W, (sy + gj) / H
        p1, q1 = pcx - bw / 2, pcy - bh / 2
        pr = (p1, q1, bw + p1, bh + q1)
        v, _iou, dv = _ciou_value_and_grad(gtb, pr)
        gl = float(iou_weighting) * 2.0 * (v - 1.0) / n             # d(iou_weighting * sum (v-1)^2 / n) / dv
        # d pred / d (tx,ty,tw,th): sigmoid passes through (d sx/d tx = 1), exp is exact
        grad[b, k, 0, gj, gi] += gl * (dv[0] + dv[2]) / W
        grad[b, k, 1, gj, gi] += gl * (dv[1] + dv[3]) / H
        grad[b, k, 2, gj, gi] += gl * (dv[2] - dv[0]) * 0.5 * bw
        grad[b, k, 3, gj, gi] += gl * (dv[3] - dv[1]) * 0.5 * bh
    return (grad * float(grad_out)).reshape(N, ch, H, W).astype(np.float32)


# ----------------------------------------------------------------------------- mAP (utils/eval_mAP.py), SURVEY 8 f2
def _iou_f32(box, boxes):
    """find_jaccard_overlap of one box against (n,4) boxes in float32 (utils/iou.py:32-49)."""
    box = np.asarray(box, np.float32)
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    lo = np.maximum(box[:2], boxes[:, :2])
    hi = np.minimum(box[2:], boxes[:, 2:])
    d = hi - lo
    d = np.where(d < 0, np.float32(0), d).astype(np.float32)   # clamp(min=0) keeps NaN
    inter = (d[:, 0] * d[:, 1]).astype(np.float32)
    a1 = np.float32((box[2] - box[0]) * (box[3] - box[1]))
    a2 = ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])).astype(np.float32)
    union = ((a1 + a2).astype(np.float32) - inter).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(np.float32)


def map_match_image(det_boxes, det_labels, true_boxes, true_labels, true_difficulties, c, iou_thr=0.5):
    """eval_single_image_recall (utils/eval_mAP.py:8-63) for class ``c``: (tp, fp) flags of the image's class-c
    detections in the order given, and the number of non-difficult class-c objects."""
    tsel = np.asarray(true_labels) == c
    dsel = np.asarray(det_labels) == c
    tb = np.asarray(true_boxes, np.float32).reshape(-1, 4)[tsel]
    td = np.asarray(true_difficulties).reshape(-1)[tsel]
    n_easy = int((1 - td.astype(np.int64)).sum())                    # :17
    db = np.asarray(det_boxes, np.float32).reshape(-1, 4)[dsel]
    tp = np.zeros(len(db), np.float32)
    fp = np.zeros(len(db), np.float32)
    detected = np.zeros(len(tb), np.uint8)
    for d in range(len(db)):                                          # :32
        if len(tb) == 0:
            fp[d] = 1                                                 # :37-39
            continue
        ov = _iou_f32(db[d], tb)                                      # :41
        if np.isnan(ov).any():                                        # torch.max propagates NaN; NaN > 0.5 is False
            fp[d] = 1
            continue
        ind = int(np.argmax(ov))                                      # first maximum (:42)
        if ov[ind] > np.float32(iou_thr):                             # :51
            if td[ind] == 0:                                          # :53
                if detected[ind] == 0:
                    tp[d] = 1
                    detected[ind] = 1
                else:
                    fp[d] = 1
        else:
            fp[d] = 1                                                 # :61-62
    return tp, fp, n_easy


def calculate_map(det_boxes, det_labels, det_scores, true_boxes, true_labels, true_difficulties, n_classes, iou_thr=0.5):
    """calculate_mAP (utils/eval_mAP.py:134-188) with eval_class_ap (:65-132): lists over images; labels 1..n_classes-1
    (0 = background).  Returns (ap (n_classes-1,) float32, mAP, tp sums, fp sums).  Ties in score are ordered by
    (image, detection) index -- the reference's torch.sort leaves them unspecified."""
    aps = np.zeros(n_classes - 1, np.float32)
    tps = np.zeros(n_classes - 1, np.float32)
    fps = np.zeros(n_classes - 1, np.float32)
    rec_thr = np.arange(0, 1.1, 0.1, dtype=np.float32)               # torch.arange(0, 1.1, .1) (:120), compared as fp32
    for c in range(1, n_classes):
        tp_all, fp_all, sc_all, n_easy = [], [], [], 0
        for b in range(len(det_boxes)):
            tp, fp, ne = map_match_image(det_boxes[b], det_labels[b], true_boxes[b], true_labels[b], true_difficulties[b], c, iou_thr)
            tp_all.append(tp)
            fp_all.append(fp)
            sc_all.append(np.asarray(det_scores[b], np.float32).reshape(-1)[np.asarray(det_labels[b]) == c])
            n_easy += ne
        tp_all = np.concatenate(tp_all) if tp_all else np.zeros(0, np.float32)
        fp_all = np.concatenate(fp_all) if fp_all else np.zeros(0, np.float32)
        sc_all = np.concatenate(sc_all) if sc_all else np.zeros(0, np.float32)
        order = np.argsort(-sc_all, kind="stable")                   # :107
        tp_s, fp_s = tp_all[order], fp_all[order]
        ctp = np.cumsum(tp_s, dtype=np.float32)                       # :113-114
        cfp = np.cumsum(fp_s, dtype=np.float32)
        prec = (ctp / ((ctp + cfp).astype(np.float32) + np.float32(1e-10))).astype(np.float32)   # :115-116
        with np.errstate(divide="ignore", invalid="ignore"):
            rec = (ctp / np.float32(n_easy)).astype(np.float32)      # :117
        p11 = np.zeros(11, np.float32)
        for i, t in enumerate(rec_thr):                               # :121-126
            above = rec >= t
            if above.any():
                p11[i] = prec[above].max()
        aps[c - 1] = p11.mean(dtype=np.float32)
        tps[c - 1] = tp_s.sum()
        fps[c - 1] = fp_s.sum()
    return aps, float(aps.mean(dtype=np.float32)), tps, fps


# ----------------------------------------------------------------------------- seg head (models/seg_loss.py), SURVEY 8 f4
def seg_loss(inp, targets):
    """SegLoss.forward(input, targets) (models/seg_loss.py:51-76): input (N,C,H,W), targets (N,H,W,C).
    Returns (loss, obj_mean, no_obj_mean); fp32 elementwise maths, float64 sums."""
    x = _f32(inp)
    t = np.transpose(_f32(targets), (0, 3, 1, 2))                      # :54
    o = (np.float32(1.0) / (np.float32(1.0) + np.exp(-x, dtype=np.float32))).astype(np.float32)   # :19
    d = (o - t).astype(np.float32)
    sq = (d * d).astype(np.float32)
    loss = 0.05 * float(sq.sum(dtype=np.float64)) / x.size           # :40-45 with all-ones weights (:73), :76
    with np.errstate(divide="ignore", invalid="ignore"):
        obj = float(o[t >= 0.5].sum(dtype=np.float64) / np.float64((t >= 0.5).sum()))   # :65, torch.mean of empty = nan
        no_obj = float(o[t < 0.5].sum(dtype=np.float64) / np.float64((t < 0.5).sum()))  # :66
    return loss, obj, no_obj


def seg_loss_backward(inp, targets, grad_out: float = 1.0):
    """d loss / d input of SegLoss.forward(input, targets): the custom sigmoid passes gradients through (:23-31),
    so it is grad_out * 0.05 * 2 (o - t) / numel."""
    x = _f32(inp).astype(np.float64)
    t = np.transpose(_f32(targets), (0, 3, 1, 2)).astype(np.float64)
    o = 1.0 / (1.0 + np.exp(-x))
    return (grad_out * 0.05 * 2.0 * (o - t) / x.size).astype(np.float32)


def seg_sigmoid(inp):
    """SegLoss.forward(input) (:77-80): sigmoid(input)[0] as a numpy array."""
    x = _f32(inp)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x[0], dtype=np.float32))).astype(np.float32)
