"""Size-independent properties of the CUDA path at BASELINE.json's FULL per-GPU sizes, where the CPU oracle would take
minutes (the small-size parity tests are in test_gpu_parity.py).  Everything is checked on the device with torch ops.

decode + NMS (yolo_loss.py:180-204, utils/box.py:11-31, torchvision nms):
  * every kept row is a candidate (conf > val_conf) and equals the stand-alone decode of its cell, kept cell ids unique;
  * output order: classes ascending, scores (col5*col4) non-increasing inside a class (box.py:29-30);
  * fixed point: NMS of the kept rows keeps all of them, in the same order (greedy NMS is idempotent);
  * no two kept rows of one class overlap by more than the threshold, and every dropped candidate is overlapped by a
    kept row of its class with a score at least as high (sampled images, utils/iou.py as the IoU);
  * images are independent: permuting the batch permutes the output.
target assignment + loss (yolo_loss.py:77-178, 206-236): the 16 partial sums are additive over any split of the batch
(the 'checksum of checksums': they are what the data-parallel all-reduce adds up), and the number of assignments /
normalisers are integers consistent with the cell counts."""
import numpy as np
import pytest
import torch

import mobilenet_yolo_pytorch_b200 as b200
from mobilenet_yolo_pytorch_b200 import _lib, ops
from test_gpu_parity import BDD_ANCHORS, MASK, VOC_ANCHORS, anchor_tables, make_heads

pytestmark = pytest.mark.gpu

NMS_THR = 0.45


def _check_decode_nms_properties(N, C, grids, anchors, img, thr, shift, dev, sample=6):
    h0, h1 = make_heads(N, C, grids, seed=123, conf_shift=shift)
    tables = anchor_tables(anchors, img)
    d0, d1 = h0.to(dev), h1.to(dev)
    out, cnt, idx = ops.decode_nms_padded(d0, d1, tables, C, thr, want_idx=True)
    K = out.shape[1]
    cells0 = 3 * grids[0][0] * grids[0][1]
    r0, c0, i0 = ops.decode_head_padded(d0, tables[0], C, thr, want_ids=True)
    r1, c1, i1 = ops.decode_head_padded(d1, tables[1], C, thr, want_ids=True)
    assert int(cnt.min()) >= 0 and int(cnt.max()) <= K
    assert torch.all(cnt <= c0 + c1)
    ar = torch.arange(K, device=dev)[None, :]
    valid = ar < cnt[:, None]                                     # (N, K) kept rows

    # candidates by cell id: a dense (N, K, 7) table of the stand-alone decode, NaN where the cell did not pass
    table = torch.full((N, K, 7), float("nan"), device=dev)
    for rows, count, ids, off in ((r0, c0, i0, 0), (r1, c1, i1, cells0)):
        v = torch.arange(rows.shape[1], device=dev)[None, :] < count[:, None]
        bi = torch.arange(N, device=dev)[:, None].expand_as(ids)[v]
        table[bi, ids[v].long() + off] = rows[v]
    bi = torch.arange(N, device=dev)[:, None].expand(N, K)[valid]
    kept_ids = idx[valid].long()
    assert torch.equal(table[bi, kept_ids], out[valid]), "a kept row differs from the decode of its cell"
    # unique cell ids per image
    flat = bi * K + kept_ids
    assert flat.unique().numel() == flat.numel()

    # order: class ascending, score non-increasing inside a class
    cls, score = out[..., 6], out[..., 5] * out[..., 4]
    nxt = valid[:, 1:]
    assert torch.all((cls[:, 1:] >= cls[:, :-1])[nxt])
    same = nxt & (cls[:, 1:] == cls[:, :-1])
    assert torch.all((score[:, 1:] <= score[:, :-1])[same])
    assert torch.all((out[..., 4] > np.float32(thr))[valid])
    assert torch.all(((cls >= 0) & (cls < C) & (cls == cls.round()))[valid])

    # fixed point: NMS(kept rows) == kept rows
    out2, cnt2 = ops.nms_padded(out, cnt, None, None, C)
    assert torch.equal(cnt2, cnt)
    assert torch.equal(out2[valid], out[valid])

    # images are independent
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(5)).to(dev)
    outp, cntp = ops.decode_nms_padded(d0[perm], d1[perm], tables, C, thr)
    assert torch.equal(cntp, cnt[perm])
    assert torch.equal(outp[valid[perm]], out[perm][valid[perm]])

    # sampled images: no kept pair of a class above the threshold; every dropped candidate is covered
    for b in np.linspace(0, N - 1, num=min(sample, N)).astype(int):
        k = int(cnt[b])
        kept = out[b, :k]
        cand = table[b][~torch.isnan(table[b, :, 4])]
        if k == 0:
            assert cand.shape[0] == 0 or torch.all(torch.isnan(cand[:, :4]).any(1))
            continue
        iou = b200.find_jaccard_overlap(kept[:, :4].contiguous(), kept[:, :4].contiguous())
        same_cls = kept[:, 6][:, None] == kept[:, 6][None, :]
        off_diag = ~torch.eye(k, dtype=torch.bool, device=dev)
        assert not torch.any((iou.double() > NMS_THR) & same_cls & off_diag)
        kept_mask = torch.zeros(K, dtype=torch.bool, device=dev)
        kept_mask[idx[b, :k].long()] = True
        dropped = table[b][(~torch.isnan(table[b, :, 4])) & ~kept_mask]
        if dropped.shape[0]:
            iou_d = b200.find_jaccard_overlap(dropped[:, :4].contiguous(), kept[:, :4].contiguous()).double()
            s_d, s_k = dropped[:, 5] * dropped[:, 4], kept[:, 5] * kept[:, 4]
            cover = (iou_d > NMS_THR) & (dropped[:, 6][:, None] == kept[:, 6][None, :]) & (s_k[None, :] >= s_d[:, None])
            assert torch.all(cover.any(1)), "a dropped candidate is not suppressed by any kept row"
    return int(cnt.sum())


@pytest.mark.parametrize("name,N,C,grids,anchors,img,thr,shift", [
    ("cfg2_voc_b256", 256, 20, [(11, 11), (22, 22)], VOC_ANCHORS, [352, 352], 0.3, 0.0),
    ("cfg2_sparse_b256", 256, 20, [(11, 11), (22, 22)], VOC_ANCHORS, [352, 352], 0.3, -2.6),
    ("cfg3_bdd_b1024", 1024, 10, [(12, 20), (24, 40)], BDD_ANCHORS, [640, 384], 0.3, 0.0),
    ("cfg5_416_b4096", 4096, 20, [(13, 13), (26, 26)], VOC_ANCHORS, [416, 416], 0.001, 0.0),
])
def test_decode_nms_properties_at_full_size(name, N, C, grids, anchors, img, thr, shift, cuda_device):
    assert _check_decode_nms_properties(N, C, grids, anchors, img, thr, shift, cuda_device) > 0


def test_decode_nms_properties_large_images(cuda_device):
    """the 832x832 variant (10 140 cells per image): same properties through the large-image kernels"""
    assert _check_decode_nms_properties(24, 20, [(26, 26), (52, 52)], VOC_ANCHORS, [832, 832], 0.001, 0.0, cuda_device, sample=3) > 0


def test_loss_sums_are_additive_at_full_size(cuda_device):
    """config 4: N=512, 100 GT boxes per image.  sums(batch) == sums(first part) + sums(second part) for uneven splits
    (fp64 partial sums: equal to 1e-12 relative), counts are integers, weights count cells."""
    N, G, C = 512, 100, 20
    r = np.random.RandomState(4)
    h0, h1 = make_heads(N, C, [(11, 11), (22, 22)], seed=77)
    targets = []
    for _ in range(N):
        wh = r.uniform(0.02, 0.47, (G, 2))
        c = wh / 2 + r.rand(G, 2) * (1 - wh)
        targets.append(torch.from_numpy(np.concatenate((r.randint(1, C + 1, (G, 1)), c, wh), 1).astype(np.float32)))
    sa = np.array([[aw / 352, ah / 352] for aw, ah in VOC_ANCHORS], np.float64).astype(np.float32)
    for k, h in enumerate((h0, h1)):
        hd = h.to(cuda_device)
        cells = hd.shape[1] // (5 + C) * hd.shape[2] * hd.shape[3]

        def sums_of(lo, hi):
            gt, gt_off, Gt, _ = ops.pack_targets(targets[lo:hi], cuda_device)
            s, _ = ops.target_loss_sums(hd[lo:hi].contiguous(), gt, gt_off, Gt, sa, MASK[k], C, 0.6, 0.55, max_gt=G)
            return s.double().cpu().numpy().copy()

        whole = sums_of(0, N)
        for cut in (1, 200, 511):
            parts = sums_of(0, cut) + sums_of(cut, N)
            np.testing.assert_allclose(parts, whole, rtol=1e-12, atol=1e-9)
        n_assign = whole[_lib.S_NASSIGN]
        assert n_assign == round(n_assign) and 0 < n_assign <= N * G * 3
        w = whole[_lib.S_W]
        # weight 1 on every objectness cell that is not ignored + the C class channels of every distinct assigned cell
        assert w == round(w) and 0 < w <= N * cells * (1 + C)
        assert whole[_lib.S_NCELLS] == N * cells and whole[_lib.S_NIMG] == N
