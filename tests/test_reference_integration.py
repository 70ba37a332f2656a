"""BASELINE config 1 / SURVEY row A17: the reference's REAL caller -- models/mbv2_yolo.py `yolo.forward`
(:137-173, call site :158-160) driven the way inference.py:109-126 drives it -- with this package patched in.

The unmodified reference files come from oracle/_ref (oracle/snapshot_reference.py places them there at build();
in the build container /root/reference itself is used).  The reference fixes its device when it is imported
(quirk Q4), so its CPU run happens in a subprocess with CUDA_VISIBLE_DEVICES=""; its CUDA run and the patched run share
this process (same cuDNN kernels for the backbone, so the head tensors are bit-identical and the comparison isolates the
replaced hot path).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

RTOL, ATOL = 1e-5, 1e-6   # north star: decoded floats within 1e-5 relative


def _need_reference():
    if ref_loader.reference_root() is None:
        pytest.skip("no reference snapshot (oracle/_ref) and no /root/reference on this box")


def test_snapshot_recipe_matches_the_checkout():
    """(CPU) where the checkout exists, the snapshot is a byte-for-byte copy of the listed files."""
    from oracle import snapshot_reference as sr
    if not os.path.isdir(sr.REF):
        pytest.skip("no reference checkout here")
    assert sr.snapshot()
    man = json.load(open(os.path.join(sr.DEST, "MANIFEST.json")))["sha256"]
    for rel in sr.FILES:
        assert man[rel] == sr.sha256(os.path.join(sr.REF, rel)), rel


def test_reference_config1_runs_on_cpu(tmp_path):
    """(CPU) the harness itself: the unmodified model builds random-init and produces detections at batch 1."""
    _need_reference()
    out = tmp_path / "cfg1_cpu.npz"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_config1.py"), "--out", str(out), "--reps", "1"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = np.load(out)
    assert d["out0"].shape == (1, 75, 11, 11) and d["out1"].shape == (1, 75, 22, 22)
    assert d["dets"].shape[1] == 7 and d["dets"].shape[0] > 0
    # class-ascending blocks, score-descending inside (utils/box.py:29-30)
    cls = d["dets"][:, 6]
    assert np.all(np.diff(cls) >= 0)


def _compare_rows(got, want, what):
    assert got.shape == want.shape, f"{what}: {got.shape[0]} rows, the reference has {want.shape[0]}"
    assert np.array_equal(got[:, 6], want[:, 6]), f"{what}: class columns differ"
    np.testing.assert_allclose(got[:, :6], want[:, :6], rtol=RTOL, atol=ATOL, err_msg=what)


def _patch(b200):
    def patch(ns, m):
        b200.patch_reference(models_yolo_loss=ns.yolo_loss, utils_box=ns.box, mbv2_yolo=m, fuse_inference=True)
    return patch


class _Restore:
    """undo patch_reference on the reference's modules (the process is shared with other tests)"""

    def __enter__(self):
        self.ns = ref_loader.load()
        import models.mbv2_yolo as m
        self.m = m
        self.saved = (self.ns.yolo_loss.YOLOLoss, self.ns.box.nms, m.YOLOLoss, m.nms)
        return self

    def __exit__(self, *exc):
        import mobilenet_yolo_pytorch_b200 as b200
        self.ns.yolo_loss.YOLOLoss, self.ns.box.nms, self.m.YOLOLoss, self.m.nms = self.saved
        b200.YOLOLoss.lazy_eval = False


@pytest.mark.gpu
def test_config1_inference_patched_vs_reference(tmp_path):
    _need_reference()
    import mobilenet_yolo_pytorch_b200 as b200
    from oracle import ref_config1
    torch.backends.cudnn.benchmark = False
    # the reference, unmodified, on this GPU (torchvision's CUDA nms per (image, class)) ...
    ref = ref_config1.run("cuda", reps=1)
    # ... and on the CPU (separate process: quirk Q4)
    out = tmp_path / "cfg1_cpu.npz"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_config1.py"), "--out", str(out), "--reps", "1"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    cpu = np.load(out)
    with _Restore():
        got = ref_config1.run("cuda", reps=1, patch=_patch(b200))
    # same weights, same cuDNN kernels: the patched model sees the same head tensors
    np.testing.assert_allclose(got["out0"], ref["out0"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(got["out1"], ref["out1"], rtol=1e-6, atol=1e-7)
    # (a) the whole patched model against the unpatched one, both on this GPU
    _compare_rows(got["dets"], ref["dets"], "patched yolo.forward vs the reference on CUDA")
    # (b) the replaced hot path alone on the reference's own head tensors: CUDA heads vs its CUDA detections, CPU heads
    #     vs its CPU detections
    dev = torch.device("cuda", 0)
    losses = [b200.YOLOLoss(b200_anchors(), MASK[i], 20, [352, 352], 0.6, 0.55, val_conf=0.3) for i in range(2)]
    for label, src in (("CUDA", ref), ("CPU", cpu)):
        dets = b200.decode_nms(torch.from_numpy(src["out0"]).to(dev), torch.from_numpy(src["out1"]).to(dev), losses, 20)
        _compare_rows(dets[0].cpu().numpy(), src["dets"], f"decode_nms on the reference's {label} heads vs its {label} detections")
    # every random-init cell passes val_conf 0.3 (conf ~ 0.5): the reference's 849 detections at survey time
    assert ref["dets"].shape[0] > 100


MASK = [[0, 1, 2], [3, 4, 5]]


def b200_anchors():
    return [[143, 265], [153, 121], [280, 279], [20, 37], [49, 94], [73, 201]]   # models/voc/config.yaml:20-26


@pytest.mark.gpu
def test_config1_training_tuple_and_backward_through_the_real_model():
    """yolo.forward(x, targets) (mbv2_yolo.py:158, train.py:258-283): the two 7-tuples and d loss / d weights of the
    patched model against the unpatched reference, same weights, same batch."""
    _need_reference()
    import mobilenet_yolo_pytorch_b200 as b200
    from oracle import ref_config1
    torch.backends.cudnn.benchmark = False
    dev = torch.device("cuda", 0)
    x = ref_config1.make_image(2).to(dev)
    r = np.random.RandomState(3)
    targets = []
    for n in (6, 3):
        wh = r.rand(n, 2) * 0.4 + 0.05
        c = wh / 2 + r.rand(n, 2) * (1 - wh)
        targets.append(torch.from_numpy(np.concatenate((r.randint(1, 21, (n, 1)), c, wh), 1).astype(np.float32)))

    def run(patch):
        model, _, _ = ref_config1.build_model(patch)
        model = model.to(dev)   # eval(): BatchNorm uses its running statistics, the comparison needs no batch statistics
        out = model(x, targets)
        loss = out[0][0] + out[1][0]
        loss.backward()
        w = model.yolo_headS16[-1].weight.grad.detach().cpu().numpy().copy()
        w0 = model.yolo_headS32[-1].weight.grad.detach().cpu().numpy().copy()
        tup = [[float(v) for v in t] for t in out]
        return tup, w0, w

    with _Restore():
        want, g0, g1 = run(None)
        got, h0, h1 = run(_patch(b200))
    for a, b in zip(got, want):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-7)
    for g, h in ((g0, h0), (g1, h1)):
        assert np.abs(g).max() > 0
        assert np.abs(g - h).max() <= 1e-4 * np.abs(g).max()


@pytest.mark.gpu
@pytest.mark.parametrize("conf_shift", [0.0, -2.6])
def test_exact_decode_is_bit_equal_to_the_reference_on_cuda(conf_shift):
    """SURVEY section 7's first-choice test: the unmodified reference on device='cuda' (ATen kernels + torchvision's CUDA
    nms, driven per (image, class) by utils/box.py:16-29) against this library on the same head tensors -- torch.equal
    on the decoded rows of both heads AND on the final detections -- with the exact decode switched on
    (ops.set_exact_decode: the reference's own IEEE operations).  The default decode is within 1e-5 of the same."""
    _need_reference()
    import mobilenet_yolo_pytorch_b200 as b200
    from mobilenet_yolo_pytorch_b200 import ops
    ref = ref_loader.load()
    dev = torch.device("cuda", 0)
    N, C = 48, 20
    g = torch.Generator().manual_seed(5)
    h0 = torch.randn(N, 75, 11, 11, generator=g)
    h1 = torch.randn(N, 75, 22, 22, generator=g)
    if conf_shift:
        h0.view(N, 3, 25, 11, 11)[:, :, 4] += conf_shift
        h1.view(N, 3, 25, 22, 22)[:, :, 4] += conf_shift
    h0, h1 = h0.to(dev), h1.to(dev)
    r_losses = [ref.YOLOLoss(b200_anchors(), MASK[i], C, [352, 352], 0.6, 0.55, val_conf=0.3) for i in range(2)]
    with torch.no_grad():
        preds = [r_losses[0](h0), r_losses[1](h1)]
        want = ref.nms(preds, C)
    losses = [b200.YOLOLoss(b200_anchors(), MASK[i], C, [352, 352], 0.6, 0.55, val_conf=0.3) for i in range(2)]
    ops.set_exact_decode(True)
    try:
        got_preds = [losses[0](h0), losses[1](h1)]
        dets = b200.decode_nms(h0, h1, losses, C)
        sep = b200.nms(got_preds, C)
    finally:
        ops.set_exact_decode(False)
    for i in range(2):
        for b in range(N):
            assert torch.equal(got_preds[i][b], preds[i][b]), f"head {i} image {b}: decoded rows are not bit-equal"
    for b in range(N):
        assert torch.equal(dets[b], want[b]), f"image {b}: fused detections are not bit-equal to the reference on CUDA"
        assert torch.equal(sep[b], want[b])
    fast = b200.decode_nms(h0, h1, losses, C)
    for b in range(N):
        assert fast[b].shape == want[b].shape and torch.equal(fast[b][:, 6], want[b][:, 6])
        assert torch.allclose(fast[b], want[b], rtol=RTOL, atol=ATOL)
