#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own modules (models/yolo_loss.py, utils/box.py,
utils/iou.py -- unmodified, with in-memory stubs for the two unrelated missing
imports matplotlib/progress), runs them on seeded synthetic
inputs on CPU and stores inputs + outputs as small .npz files.  The fixtures pin
the C oracle (tests/test_oracle_golden.py) and, on the GPU box, the CUDA path.

Index-level outputs (NMS kept ids, assignment tuples) come from index-tracking
mirrors of utils/box.py:16-30 and models/yolo_loss.py:127-145 that are asserted
row-for-row against what the real reference functions return.
"""
import math
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

VOC = dict(
    anchors=[[143, 265], [153, 121], [280, 279], [20, 37], [49, 94], [73, 201]],  # models/voc/config.yaml:20-26
    mask=[[0, 1, 2], [3, 4, 5]],
    num_classes=20, img_size=[352, 352],
    ignore_thresh=[0.6076333316652263, 0.5623606200028424], iou_thresh=0.5497280113447018,
    iou_weighting=0.021830872589525777,
)
BDD = dict(
    anchors=[[34, 47], [66, 93], [122, 182], [6, 11], [11, 43], [16, 22]],  # models/bdd100k/config.yaml:17-23
    mask=[[0, 1, 2], [3, 4, 5]],
    num_classes=10, img_size=[640, 384],  # BASELINE.json config 3 (W x H), 10 classes
    ignore_thresh=[0.6, 0.55], iou_thresh=0.6, iou_weighting=0.02,
)


def load_reference():
    sys.path.insert(0, REF)
    for n in ["matplotlib", "matplotlib.pyplot", "progress", "progress.bar"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["progress.bar"].Bar = object
    sys.modules["progress.bar"].IncrementalBar = object
    from models.yolo_loss import YOLOLoss  # noqa
    from utils.box import nms  # noqa
    from utils import iou as iou_mod  # noqa
    return YOLOLoss, nms, iou_mod


YOLOLoss, ref_nms, ref_iou = load_reference()


def fixed_pre_maps(self, bs, is_cuda, anchors, in_w, in_h):
    """Quirk Q1: the reference's pre_maps (yolo_loss.py:71-72) only
    works for square grids.  This is the intended meshgrid; on square grids it is
    torch.equal to the original (asserted in main())."""
    this = torch.FloatTensor(np.array(anchors)[self.mask])
    A = self.num_mask
    anchor_wh = this.view(1, A, 1, 1, 2).expand(bs, A, in_h, in_w, 2).contiguous()
    gx = torch.linspace(0, in_w - 1, in_w).view(1, 1, 1, in_w, 1).expand(bs, A, in_h, in_w, 1)
    gy = torch.linspace(0, in_h - 1, in_h).view(1, 1, in_h, 1, 1).expand(bs, A, in_h, in_w, 1)
    return torch.cat((gx, gy), 4).contiguous(), anchor_wh


def make_losses(cfg, val_conf, patch_nonsquare=False):
    ls = []
    for i in range(2):
        l = YOLOLoss(cfg["anchors"], cfg["mask"][i], cfg["num_classes"], cfg["img_size"], cfg["ignore_thresh"][i],
                     cfg["iou_thresh"], val_conf=val_conf, iou_weighting=cfg["iou_weighting"])
        if patch_nonsquare:
            l.pre_maps = types.MethodType(fixed_pre_maps, l)
        ls.append(l)
    return ls


def heads(cfg, N, seed, grid_div=(32, 16), conf_shift=0.0):
    g = torch.Generator().manual_seed(seed)
    A = 3
    attrs = 5 + cfg["num_classes"]
    W, H = cfg["img_size"]
    hs = []
    for d in grid_div:
        h = torch.randn(N, A * attrs, H // d, W // d, generator=g)
        if conf_shift:
            h.view(N, A, attrs, H // d, W // d)[:, :, 4] += conf_shift
        hs.append(h.contiguous())
    return hs


def nms_with_ids(preds, num_classes):
    """Index-tracking mirror of utils/box.py:16-30."""
    import torchvision
    out_rows, out_ids = [], []
    for b in range(len(preds[0])):
        per = torch.cat((preds[0][b], preds[1][b]), 0)
        rows = torch.zeros(0, 7)
        ids = torch.zeros(0, dtype=torch.long)
        if per.size(0):
            for i in range(num_classes):
                m = per[..., 6] == i
                idx_in = m.nonzero().flatten()
                sub = per[m]
                if sub.size(0):
                    keep = torchvision.ops.nms(sub[..., :4], sub[..., 5] * sub[..., 4], 0.45)
                    rows = torch.cat((rows, sub[keep]), 0)
                    ids = torch.cat((ids, idx_in[keep]), 0)
        out_rows.append(rows)
        out_ids.append(ids)
    return out_rows, out_ids


def pack_ragged(prefix, lst, d):
    """store list of (n_b, k) arrays as one concatenated array + counts"""
    lst = [np.asarray(a) for a in lst]
    d[prefix + "_count"] = np.array([len(a) for a in lst], np.int32)
    if len(lst) and sum(len(a) for a in lst):
        d[prefix] = np.concatenate([a for a in lst if len(a)], 0)
    else:
        d[prefix] = np.zeros((0,) + lst[0].shape[1:], lst[0].dtype if len(lst) else np.float32)


def case_decode_nms(name, cfg, N, seed, val_conf, conf_shift=0.0, nonsquare=False):
    losses = make_losses(cfg, val_conf, patch_nonsquare=nonsquare)
    h0, h1 = heads(cfg, N, seed, conf_shift=conf_shift)
    with torch.no_grad():
        p0 = losses[0](h0)
        p1 = losses[1](h1)
        det = ref_nms((p0, p1), cfg["num_classes"])
        rows, ids = nms_with_ids((p0, p1), cfg["num_classes"])
    for a, b in zip(det, rows):
        assert torch.equal(a, b), "index-tracking NMS mirror diverged from utils.box.nms"
    d = dict(head0=h0.numpy(), head1=h1.numpy(), val_conf=np.float64(val_conf), num_classes=np.int32(cfg["num_classes"]),
             anchors=np.array(cfg["anchors"], np.float64), mask=np.array(cfg["mask"], np.int32),
             img_size=np.array(cfg["img_size"], np.float64))
    pack_ragged("p0", [p.numpy() for p in p0], d)
    pack_ragged("p1", [p.numpy() for p in p1], d)
    pack_ragged("det", [p.numpy() for p in det], d)
    pack_ragged("det_idx", [i.numpy().astype(np.int32).reshape(-1, 1) for i in ids], d)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "cand", [len(a) + len(b) for a, b in zip(p0, p1)], "kept", [len(x) for x in det])


def case_nms_ties():
    """Hand-made candidates pinning tie-breaking, zero-area/NaN IoU, negative-size
    boxes and an unused class, through the real utils.box.nms."""
    C = 4
    rows0 = [
        # x1   y1   x2   y2   conf  cls_score cls
        [0.10, 0.10, 0.50, 0.50, 0.9, 0.8, 0.0],
        [0.10, 0.10, 0.50, 0.50, 0.9, 0.8, 0.0],   # exact duplicate: same score, suppressed by the earlier one
        [0.12, 0.10, 0.52, 0.50, 0.8, 0.9, 0.0],   # same product 0.72 as the two above (0.9*0.8) up to rounding
        [0.60, 0.60, 0.60, 0.60, 0.7, 0.7, 0.0],   # zero area
        [0.60, 0.60, 0.60, 0.60, 0.7, 0.7, 0.0],   # zero area twin -> IoU NaN -> never suppressed
        [0.90, 0.90, 0.80, 0.80, 0.6, 0.6, 1.0],   # negative size
        [0.85, 0.85, 0.95, 0.95, 0.6, 0.6, 1.0],
        [0.00, 0.00, 1.00, 1.00, 0.5, 0.5, 3.0],
    ]
    rows1 = [
        [0.00, 0.00, 0.45, 1.00, 0.5, 0.5, 3.0],   # IoU with the unit box = 0.45 exactly-ish (not > 0.45 in f32?)
        [0.00, 0.00, 0.46, 1.00, 0.4, 0.5, 3.0],
        [0.30, 0.30, 0.70, 0.70, 0.95, 0.99, 0.0],
        [0.31, 0.30, 0.71, 0.70, 0.95, 0.99, 0.0],
        [0.31, 0.30, 0.71, 0.70, 0.96, 0.99, 0.0],
    ]
    p0 = [torch.tensor(rows0, dtype=torch.float32), torch.zeros(0, 7)]
    p1 = [torch.tensor(rows1, dtype=torch.float32), torch.zeros(0, 7)]
    det = ref_nms((p0, p1), C)
    rows, ids = nms_with_ids((p0, p1), C)
    for a, b in zip(det, rows):
        assert torch.equal(a, b)
    d = dict(num_classes=np.int32(C))
    pack_ragged("p0", [p.numpy() for p in p0], d)
    pack_ragged("p1", [p.numpy() for p in p1], d)
    pack_ragged("det", [p.numpy() for p in det], d)
    pack_ragged("det_idx", [i.numpy().astype(np.int32).reshape(-1, 1) for i in ids], d)
    np.savez_compressed(os.path.join(OUT, "nms_ties.npz"), **d)
    print("nms_ties kept ids", [i.tolist() for i in ids])


def case_iou(seed=3):
    g = torch.Generator().manual_seed(seed)

    def boxes(n):
        c = torch.rand(n, 2, generator=g)
        wh = torch.rand(n, 2, generator=g) * 0.5
        return torch.cat((c - wh / 2, c + wh / 2), 1)

    a, b = boxes(37), boxes(53)
    a[3] = torch.tensor([0.2, 0.2, 0.2, 0.2])  # zero area
    b[5] = torch.tensor([0.2, 0.2, 0.2, 0.2])
    b[6] = a[7]                                # identical boxes -> IoU 1
    d = dict(a=a.numpy(), b=b.numpy(), inter=ref_iou.find_intersection(a, b).numpy(),
             union=ref_iou.find_union(a, b).numpy(), iou=ref_iou.find_jaccard_overlap(a, b).numpy())
    # box_ciou / box_giou (yolo_loss.py:257-317) on matched pairs
    l = YOLOLoss(VOC["anchors"], VOC["mask"][0], 20, [352, 352], 0.6, 0.5)
    n = 37
    ci = np.zeros((n, 2), np.float32)
    gi = np.zeros((n, 2), np.float32)
    for k in range(n):
        v, i = l.box_ciou(a[k:k + 1], b[k:k + 1])
        ci[k] = [v.item(), i.item()]
        v, i = l.box_giou(a[k:k + 1], b[k:k + 1])
        gi[k] = [v.item(), i.item()]
    d["ciou"] = ci
    d["giou"] = gi
    np.savez_compressed(os.path.join(OUT, "iou.npz"), **d)
    print("iou ok; nan in iou:", int(np.isnan(d["iou"]).sum()))


def synth_targets(N, counts, C, seed, dup=True):
    """SURVEY 8d config-4 generator: cls~U{1..C}, w,h~U(0.02,0.47), cx~U(w/2,1-w/2)."""
    g = torch.Generator().manual_seed(seed)
    ts = []
    for b in range(N):
        n = counts[b]
        if n == 0:
            ts.append(torch.zeros(0, 5))
            continue
        cls = torch.randint(1, C + 1, (n, 1), generator=g).float()
        wh = torch.rand(n, 2, generator=g) * 0.45 + 0.02
        c = wh / 2 + torch.rand(n, 2, generator=g) * (1 - wh)
        t = torch.cat((cls, c, wh), 1)
        if dup and n >= 3:
            t[2, 1:] = t[0, 1:]            # same box twice -> same cell, same anchors (duplicate assignment)
            t[2, 0] = (t[0, 0] % C) + 1    # ... with a different class
        ts.append(t)
    return ts


def assignment_mirror(loss, targets, in_w, in_h):
    """Index mirror of yolo_loss.py:127-145 -> rows (b,t,k,gj,gi,best_n)."""
    scaled = [(aw / loss.img_size[0], ah / loss.img_size[1]) for aw, ah in loss.anchors]
    anchor_shapes = torch.FloatTensor(np.concatenate((np.zeros((loss.num_anchors, 2)), np.array(scaled)), 1))
    in_dim = torch.Tensor([in_w, in_h])
    rows = []
    for b, tb in enumerate(targets):
        if len(tb) == 0:
            continue
        gt = tb.clone().detach()
        gxgy = gt[..., 1:3] * in_dim
        gt[..., 1:3] = 0
        gt_box = gt[..., 1:]
        anch_ious = ref_iou.find_jaccard_overlap(gt_box, anchor_shapes)
        best_n = torch.argmax(anch_ious, 1)
        for t in range(len(tb)):
            gi, gj = int(gxgy[t, 0]), int(gxgy[t, 1])
            lst = (anch_ious[t][loss.mask] > loss.iou_thresh).tolist()
            bn = loss.num_anchors + 1
            if best_n[t] in loss.mask:
                bn = loss.mask.index(best_n[t])
            for k in range(loss.num_mask):
                if k == bn or lst[k]:
                    rows.append((b, t, k, gj, gi, int(best_n[t])))
    return np.array(rows, np.int32).reshape(-1, 6)


def case_loss(name, cfg, N, counts, seed, nonsquare=False):
    losses = make_losses(cfg, 0.1, patch_nonsquare=nonsquare)
    hs = heads(cfg, N, seed)
    targets = synth_targets(N, counts, cfg["num_classes"], seed + 100)
    d = dict(head0=hs[0].numpy(), head1=hs[1].numpy(), num_classes=np.int32(cfg["num_classes"]),
             anchors=np.array(cfg["anchors"], np.float64), mask=np.array(cfg["mask"], np.int32),
             img_size=np.array(cfg["img_size"], np.float64), ignore_thresh=np.array(cfg["ignore_thresh"], np.float64),
             iou_thresh=np.float64(cfg["iou_thresh"]), iou_weighting=np.float64(cfg["iou_weighting"]))
    pack_ragged("targets", [t.numpy() for t in targets], d)
    for i, (l, h) in enumerate(zip(losses, hs)):
        with torch.no_grad():
            tup = l(h, targets)
            in_h, in_w = h.size(2), h.size(3)
            scaled = [(aw / l.img_size[0], ah / l.img_size[1]) for aw, ah in l.anchors]
            tg, wg, outp, recall, avg_iou, obj, no_obj, cls, cnt, ious, iouw = l.get_target(
                targets, h, scaled, in_w, in_h, l.ignore_threshold, l.iou_thresh)
        mirror = assignment_mirror(l, targets, in_w, in_h)
        assert len(mirror) == round(tup[6] * N), (len(mirror), tup[6] * N)
        # every mirrored cell must be an objectness-1 target in the reference
        for (b, t, k, gj, gi, bn) in mirror:
            assert tg[b, k, gj, gi, 0].item() == 1.0 and wg[b, k, gj, gi, 0].item() == 1.0
        assert int((tg[..., 0] == 1.0).sum()) == len({(b, k, gj, gi) for (b, t, k, gj, gi, bn) in mirror})
        d[f"tuple{i}"] = np.array([float(tup[0]), float(tup[1]), float(tup[2]), float(tup[3]), float(tup[4]),
                                   float(tup[5]), float(tup[6])], np.float64)
        d[f"assign{i}"] = mirror
        d[f"targets{i}"] = tg.numpy()
        d[f"weights{i}"] = wg.numpy()
        d[f"ciou_terms{i}"] = ious.numpy().reshape(-1)
        d[f"ciou_weights{i}"] = iouw.numpy().reshape(-1)
        # d loss / d input by the reference's own autograd graph (custom pass-through sigmoid :15-32, exp,
        # the in-place wh_to_x2y2, box_ciou with alpha NOT detached, batch-global normalisers)
        hg = h.clone().requires_grad_(True)
        l(hg, targets)[0].backward()
        d[f"grad{i}"] = hg.grad.numpy().copy()
        print(name, "head", i, "tuple", d[f"tuple{i}"], "assign", len(mirror))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def check_pre_maps_patch():
    l = make_losses(VOC, 0.3)[0]
    a = [(aw / 352, ah / 352) for aw, ah in VOC["anchors"]]
    g0, a0 = l.pre_maps(2, False, a, 11, 11)
    g1, a1 = fixed_pre_maps(l, 2, False, a, 11, 11)
    assert torch.equal(g0, g1) and torch.equal(a0, a1), "pre_maps patch is not identical on square grids"


def case_map(name, N, C, seed):
    """utils/eval_mAP.py::calculate_mAP on synthetic detections / ground truth: detections are
    jittered copies of GT boxes (some with the wrong label) plus random boxes, random scores; some GT boxes are
    'difficult'; some images have no GT or no detections; the last class has no GT at all."""
    from utils.eval_mAP import calculate_mAP
    r = np.random.RandomState(seed)
    lists = {k: [] for k in ("det_boxes", "det_labels", "det_scores", "true_boxes", "true_labels", "true_difficulties")}
    for b in range(N):
        ng = 0 if b % 7 == 3 else r.randint(1, 7)
        nd = 0 if b % 5 == 4 else r.randint(1, 18)
        g = np.sort(r.rand(ng, 2, 2), axis=1).reshape(ng, 4).astype(np.float32)
        gl = r.randint(1, C - 1, ng)                      # labels 1..C-2: class C-1 never has ground truth
        d, dl = [], []
        for _ in range(nd):
            if ng and r.rand() < 0.6:
                j = r.randint(ng)
                d.append(g[j] + r.randn(4).astype(np.float32) * 0.03)
                dl.append(gl[j] if r.rand() < 0.8 else r.randint(1, C))
            else:
                d.append(np.sort(r.rand(2, 2), axis=0).reshape(4).astype(np.float32))
                dl.append(r.randint(1, C))
        lists["det_boxes"].append(np.array(d, np.float32).reshape(-1, 4))
        lists["det_labels"].append(np.array(dl, np.int64).reshape(-1))
        lists["det_scores"].append(r.rand(nd).astype(np.float32))
        lists["true_boxes"].append(g.reshape(-1, 4))
        lists["true_labels"].append(gl.astype(np.int64).reshape(-1))
        lists["true_difficulties"].append((r.rand(ng) < 0.2).astype(np.uint8))
    names = ["background"] + [f"c{i}" for i in range(1, C)]
    t = {k: [torch.from_numpy(x) for x in v] for k, v in lists.items()}
    aps, m, tp, fp = calculate_mAP(t["det_boxes"], t["det_labels"], t["det_scores"], t["true_boxes"], t["true_labels"],
                                   t["true_difficulties"], names)
    d = dict(n_classes=np.int32(C), ap=np.array(list(aps.values()), np.float32), mAP=np.float64(m),
             tp=np.array(list(tp.values()), np.float32), fp=np.array(list(fp.values()), np.float32))
    for k, v in lists.items():
        pack_ragged(k, [x.astype(np.float32) if x.dtype != np.uint8 and x.dtype != np.int64 else x for x in v], d)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "mAP", m, "AP", d["ap"])


def case_seg(name, N, C, H, W, seed):
    """models/seg_loss.py::SegLoss on a synthetic drivable-area head: forward with targets (+ the
    gradient the reference's autograd produces) and the eval forward."""
    import models.seg_loss as sl
    sl.device = torch.device("cpu")
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, C, H, W, generator=g)
    t = (torch.rand(N, H, W, C, generator=g) < 0.3).float() * torch.rand(N, H, W, C, generator=g).clamp(min=0.5)
    t[0, :2] = 0.49999                                          # values right at the 0.5 split
    m = sl.SegLoss(C)
    xg = x.clone().requires_grad_(True)
    loss, obj, no_obj = m(xg, t)
    loss.backward()
    ev = m(x)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), input=x.numpy(), targets=t.numpy(),
                        out=np.array([loss.item(), obj, no_obj], np.float64), grad=xg.grad.numpy(), eval=np.asarray(ev, np.float32))
    print(name, loss.item(), obj, no_obj)


def main():
    torch.set_num_threads(1)
    check_pre_maps_patch()
    case_decode_nms("voc_n2_conf03", VOC, 2, 0, 0.3)
    case_decode_nms("voc_sparse_n3", VOC, 3, 1, 0.3, conf_shift=-2.6)
    case_decode_nms("voc_none_n2", VOC, 2, 2, 0.999999)
    case_decode_nms("bdd_nonsquare_n2", BDD, 2, 4, 0.3, nonsquare=True)
    # 832x832 input: 10 140 cells per image (the "~10k boxes" reading of config 5) -> the large-image kernels
    case_decode_nms("voc832_sparse_n1", dict(VOC, img_size=[832, 832]), 1, 7, 0.3, conf_shift=-1.5)
    case_nms_ties()
    case_iou()
    case_loss("loss_voc_n3", VOC, 3, [6, 0, 14], 5)
    case_loss("loss_bdd_nonsquare_n2", BDD, 2, [9, 4], 6, nonsquare=True)
    # 100 / 60 GT boxes on the 11x11 and 22x22 grids: several GT boxes share a cell and an anchor (duplicate chains)
    case_loss("loss_voc_dense_n2", VOC, 2, [100, 60], 8)
    case_map("map_n40_c6", 40, 6, 11)
    case_seg("seg_n3_c2", 3, 2, 24, 40, 13)


if __name__ == "__main__":
    main()
