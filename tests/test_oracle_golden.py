"""Pins the C oracle (oracle/yolo_oracle.c) against fixtures produced by the
reference's own Python (tests/golden/make_golden.py).  CPU only.

Tolerances: floats 1e-5 relative (BASELINE.json north_star); candidate sets,
NMS keep indices and anchor assignments bit-exact.
"""
import numpy as np
import pytest

import oracle
from conftest import load_golden, unpack_ragged

RTOL = 1e-5
ATOL = 1e-6  # boxes are normalised to [0,1]; values near 0 need an absolute floor

DECODE_CASES = ["voc_n2_conf03", "voc_sparse_n3", "voc_none_n2", "bdd_nonsquare_n2", "voc832_sparse_n1"]


def head_anchor_wh(d, i):
    sa = oracle.scaled_anchors(d["anchors"].tolist(), d["img_size"].tolist())
    return sa[d["mask"][i]]


@pytest.mark.parametrize("case", DECODE_CASES)
def test_decode_matches_reference(case):
    d = load_golden(case)
    C = int(d["num_classes"])
    for i in range(2):
        rows, _ = oracle.decode_head(d[f"head{i}"], head_anchor_wh(d, i), C, float(d["val_conf"]))
        ref = unpack_ragged(d, f"p{i}")
        assert [len(r) for r in rows] == [len(r) for r in ref]
        for a, b in zip(rows, ref):
            np.testing.assert_allclose(a[:, :6], b[:, :6], rtol=RTOL, atol=ATOL)
            assert np.array_equal(a[:, 6], b[:, 6])  # class ids exact


@pytest.mark.parametrize("case", DECODE_CASES + ["nms_ties"])
def test_nms_matches_reference_on_reference_candidates(case):
    """NMS fed the reference's OWN candidate rows: keep indices bit-exact, rows bit-exact."""
    d = load_golden(case)
    C = int(d["num_classes"])
    p0, p1 = unpack_ragged(d, "p0"), unpack_ragged(d, "p1")
    cands = [np.concatenate((a.reshape(-1, 7), b.reshape(-1, 7)), 0) for a, b in zip(p0, p1)]
    dets, idx = oracle.nms(cands, C)
    ref_det, ref_idx = unpack_ragged(d, "det"), unpack_ragged(d, "det_idx")
    for a, b, ia, ib in zip(dets, ref_det, idx, ref_idx):
        assert np.array_equal(ia, ib.reshape(-1))
        assert np.array_equal(a, b.reshape(-1, 7))


@pytest.mark.parametrize("case", DECODE_CASES)
def test_decode_nms_end_to_end(case):
    d = load_golden(case)
    C = int(d["num_classes"])
    aw2 = np.stack([head_anchor_wh(d, 0), head_anchor_wh(d, 1)])
    dets, _ = oracle.decode_nms(d["head0"], d["head1"], aw2, C, float(d["val_conf"]))
    ref_det = unpack_ragged(d, "det")
    for a, b in zip(dets, ref_det):
        assert a.shape == b.reshape(-1, 7).shape
        np.testing.assert_allclose(a[:, :6], b.reshape(-1, 7)[:, :6], rtol=RTOL, atol=ATOL)
        assert np.array_equal(a[:, 6], b.reshape(-1, 7)[:, 6])


def test_pairwise_iou_and_ciou():
    d = load_golden("iou")
    for mode in ("inter", "union", "iou"):
        got = oracle.pairwise(d["a"], d["b"], mode)
        np.testing.assert_allclose(got, d[mode], rtol=1e-6, atol=1e-7, equal_nan=True)
    for k in range(len(d["ciou"])):
        np.testing.assert_allclose(oracle.box_ciou(d["a"][k], d["b"][k]), d["ciou"][k], rtol=RTOL, atol=ATOL,
                                   equal_nan=True)
        np.testing.assert_allclose(oracle.box_giou(d["a"][k], d["b"][k]), d["giou"][k], rtol=RTOL, atol=ATOL,
                                   equal_nan=True)


@pytest.mark.parametrize("case", ["loss_voc_n3", "loss_bdd_nonsquare_n2", "loss_voc_dense_n2"])
def test_target_loss_matches_reference(case):
    d = load_golden(case)
    C = int(d["num_classes"])
    targets = unpack_ragged(d, "targets")
    for i in range(2):
        r = oracle.target_loss(d[f"head{i}"], targets, d["anchors"].tolist(), d["mask"][i].tolist(), C,
                               d["img_size"].tolist(), float(d["ignore_thresh"][i]), float(d["iou_thresh"]),
                               float(d["iou_weighting"]), want_dense=True)
        assert np.array_equal(r["assign"], d[f"assign{i}"])  # (b,t,k,gj,gi,best_n) bit-exact, reference order
        tup = d[f"tuple{i}"]
        got = np.array([r["loss"], r["recall"], r["avg_iou"], r["obj"], r["no_obj"], r["cls"], r["count_per_img"]])
        np.testing.assert_allclose(got, tup, rtol=RTOL, atol=1e-7)
        np.testing.assert_allclose(r["targets"], d[f"targets{i}"], rtol=RTOL, atol=ATOL)
        assert np.array_equal(r["weights"], d[f"weights{i}"])
        np.testing.assert_allclose(r["terms"][:, 0], d[f"ciou_terms{i}"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(r["terms"][:, 2], d[f"ciou_weights{i}"], rtol=RTOL, atol=ATOL)


def test_out_of_range_gt_raises():
    d = load_golden("loss_voc_n3")
    t = [np.array([[1.0, 1.0, 0.5, 0.1, 0.1]], np.float32)] + [np.zeros((0, 5), np.float32)] * 2  # cx == 1.0
    with pytest.raises(IndexError):
        oracle.target_loss(d["head0"], t, d["anchors"].tolist(), d["mask"][0].tolist(), 20, [352, 352], 0.6, 0.5, 0.02)


GRAD_TOL = 1e-5  # |d| <= GRAD_TOL * max|grad|: the reference's autograd runs in fp32


@pytest.mark.parametrize("case", ["loss_voc_n3", "loss_bdd_nonsquare_n2", "loss_voc_dense_n2"])
def test_target_loss_backward_matches_reference_autograd(case):
    """oracle.target_loss_backward (analytic, float64) vs input.grad after the reference's own
    loss.backward() (tests/golden/make_golden.py)."""
    d = load_golden(case)
    targets = unpack_ragged(d, "targets")
    for i in range(2):
        g = oracle.target_loss_backward(d[f"head{i}"], targets, d["anchors"].tolist(), d["mask"][i].tolist(),
                                        int(d["num_classes"]), d["img_size"].tolist(), float(d["ignore_thresh"][i]),
                                        float(d["iou_thresh"]), float(d["iou_weighting"]))
        ref = d[f"grad{i}"]
        assert np.array_equal(g != 0, ref != 0), "gradient support differs"
        assert np.abs(g - ref).max() <= GRAD_TOL * np.abs(ref).max()
        # box channels of assigned cells carry the CIoU gradient
        A, attrs = len(d["mask"][i]), 5 + int(d["num_classes"])
        box = ref.reshape(ref.shape[0], A, attrs, *ref.shape[2:])[:, :, :4]
        assert (box != 0).any()


MAP_KEYS = ("det_boxes", "det_labels", "det_scores", "true_boxes", "true_labels", "true_difficulties")


def test_map_matches_reference_calculate_mAP():
    """oracle.calculate_map vs utils/eval_mAP.py::calculate_mAP run by make_golden.py (difficult objects, empty
    images, a class without ground truth)."""
    d = load_golden("map_n40_c6")
    L = [unpack_ragged(d, k) for k in MAP_KEYS]
    ap, m, tp, fp = oracle.calculate_map(*L, int(d["n_classes"]))
    np.testing.assert_allclose(ap, d["ap"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(m, float(d["mAP"]), rtol=1e-6)
    assert np.array_equal(tp, d["tp"]) and np.array_equal(fp, d["fp"])


def test_seg_loss_matches_reference():
    """oracle.seg_loss / seg_loss_backward / seg_sigmoid vs models/seg_loss.py::SegLoss and its autograd."""
    d = load_golden("seg_n3_c2")
    np.testing.assert_allclose(oracle.seg_loss(d["input"], d["targets"]), d["out"], rtol=1e-5)
    g = oracle.seg_loss_backward(d["input"], d["targets"])
    assert np.abs(g - d["grad"]).max() <= 1e-5 * np.abs(d["grad"]).max()
    np.testing.assert_allclose(oracle.seg_sigmoid(d["input"]), d["eval"], rtol=1e-6, atol=1e-7)
