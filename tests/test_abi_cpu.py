"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol
that include/b200yolo.h declares, validates arguments without a GPU, and the
host-side helpers (anchor scaling, shard bounds) behave like the reference."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT

from mobilenet_yolo_pytorch_b200 import _lib, build as b200_build, dist as b200_dist, ops


@pytest.fixture(scope="module")
def lib():
    b200_build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "b200yolo.h")).read()
    declared = set(re.findall(r"\b(b200yolo_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.b200yolo_version() == 100


def test_argument_validation_needs_no_gpu(lib):
    aw = np.ones((2, 3, 2), np.float32)
    rc = lib.b200yolo_decode_nms(None, None, 1, 3, 20, 11, 11, 22, 22, aw.ctypes.data, 0.3, 0.45, None, None, None, None)
    assert rc == -1 and b"null" in lib.b200yolo_last_error()
    rc = lib.b200yolo_pairwise(None, -1, None, 2, 2, None, None)
    assert rc == -1
    assert lib.b200yolo_target_loss_workspace_bytes(4) == 4 * 8 * 16 * 8  # N x (CTAs per image <= 8) x 16 doubles
    # round-2 entry points: replayable plans and multicast memory
    import ctypes as C
    assert lib.b200yolo_plan_launch(None, None) == -1 and b"null plan" in lib.b200yolo_last_error()
    assert lib.b200yolo_plan_destroy(None) == 0
    h = C.c_void_p()
    assert lib.b200yolo_plan_create(None, 0, 1, 3, 20, 11, 11, 22, 22, aw.ctypes.data, 0.3, 0.45, C.byref(h)) == -1 and not h.value
    assert lib.b200yolo_mc_free(None) == 0
    assert lib.b200yolo_mc_add_device(None) == -1
    fd = C.c_int(-1)
    assert lib.b200yolo_mc_create(0, 8, C.byref(fd), C.byref(h)) != 0 and not h.value   # (no driver here, or zero bytes)
    assert lib.b200yolo_mc_supported(0) in (0, 1)


def test_loss_finalize_matches_reference_formulas(lib):
    s = np.zeros(16)
    s[_lib.S_SQW], s[_lib.S_W] = 12.0, 48.0
    s[_lib.S_IOU_SQ], s[_lib.S_NASSIGN] = 3.0, 6.0
    s[_lib.S_OBJ], s[_lib.S_CONF_ALL], s[_lib.S_CLS], s[_lib.S_IOU], s[_lib.S_RECALL] = 3.0, 100.0, 2.4, 1.8, 2.0
    s[_lib.S_NCELLS], s[_lib.S_NIMG] = 206.0, 3.0
    r = ops.loss_finalize(s, 0.5)
    w = float(np.float32(0.5))
    np.testing.assert_allclose(r, [12 / 48 + (3 / 6) * w, 2 / 6, 1.8 / 6, 3 / 6, (100 - 3) / (206 - 6), 2.4 / 6, 2.0])
    s[_lib.S_NASSIGN] = 0
    r = ops.loss_finalize(s, 0.5)
    np.testing.assert_allclose(r, [0.25, 0, 0, 0, 0, 0, 0])  # yolo_loss.py:176-177


def test_scaled_anchors_round_like_the_reference():
    import torch
    anchors = [[143, 265], [153, 121], [280, 279], [20, 37], [49, 94], [73, 201]]
    got = ops.scaled_anchors(anchors, [352, 352])
    ref = torch.FloatTensor(np.array([(a / 352, b / 352) for a, b in anchors])).numpy()  # yolo_loss.py:214,67
    assert np.array_equal(got, ref)


def test_shard_bounds_partition_the_batch():
    for n in (0, 1, 7, 256, 1023):
        for w in (1, 2, 4, 8):
            spans = [b200_dist.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_no_cpu_fallback():
    import torch
    import mobilenet_yolo_pytorch_b200 as m
    l = m.YOLOLoss([[10, 10]] * 6, [0, 1, 2], 2, [64, 64], 0.5, 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        l(torch.zeros(1, 21, 2, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.find_jaccard_overlap(torch.zeros(1, 4), torch.zeros(1, 4))


def test_adjust_confidence_like_the_reference():
    """train.py:434-440"""
    import torch
    import mobilenet_yolo_pytorch_b200 as m
    assert m.adjust_confidence(10, 31, 0.3) == pytest.approx(0.31)
    assert m.adjust_confidence(10, 19, 0.3) == pytest.approx(0.29)
    assert m.adjust_confidence(10, 25, 0.3) == 0.3
    assert m.adjust_confidence(10, 0, 0.01) == 0.01          # never below 0.01
    assert m.adjust_confidence(10, torch.tensor([20, 15], dtype=torch.int32), 0.3) == pytest.approx(0.31)


def test_bench_reference_arm_runs_without_gpu():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the CUDA arm) needs no CUDA device and prints
    the contract's JSON line."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "decode+NMS images/sec" and line["value"] > 0
    assert line["unit"] == "images/s" and line["higher_is_better"] is True
    # the unmodified reference Python where its snapshot (oracle/_ref) or checkout exists, else the C port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["name"] == "cfg2" and line["config"]["batch_per_gpu"] == 256   # same config dict as the CUDA arm


def test_bench_config_dict_is_shared_by_both_arms():
    """the driver compares the two arms' `config`: it is built by one function from the workload alone"""
    import bench
    for world in (1, 2, 8):
        c = bench.workload_config("cfg2", bench.WORKLOADS["cfg2"], world)
        assert c["global_batch"] == 256 * world and c["cells_per_image"] == 1815
    assert bench.workload_config("cfg3", bench.WORKLOADS["cfg3"], 8)["batch_per_gpu"] == 128
    assert bench.local_batch(bench.WORKLOADS["cfg5"], 8, 3) == 512
