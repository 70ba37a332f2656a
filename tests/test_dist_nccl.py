"""Two-GPU NCCL check of the data-parallel path (skipped unless two CUDA devices are visible; the CPU/gloo
counterpart is tests/test_dist_gloo.py).  Each rank owns half of the batch: the all-gathered detections must equal
the single-GPU result on the whole batch, and loss / gradient of the sharded YOLOLoss (16 partial sums all-reduced
before the batch-global division, yolo_loss.py:55,224) must equal the unsharded ones."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu

VOC_ANCHORS = [[143, 265], [153, 121], [280, 279], [20, 37], [49, 94], [73, 201]]
MASKS = [[0, 1, 2], [3, 4, 5]]
C = 20


def _inputs(N):
    g = torch.Generator().manual_seed(21)
    h0 = torch.randn(N, 75, 11, 11, generator=g)
    h1 = torch.randn(N, 75, 22, 22, generator=g)
    r = np.random.RandomState(9)
    targets = []
    for b in range(N):
        n = [4, 0, 25, 2, 60, 9, 1, 13][b % 8]
        wh = r.rand(n, 2) * 0.45 + 0.02
        c = wh / 2 + r.rand(n, 2) * (1 - wh)
        targets.append(torch.from_numpy(np.concatenate((r.randint(1, C + 1, (n, 1)), c, wh), 1).astype(np.float32)))
    return h0, h1, targets


def _worker(rank, world, port, N, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import mobilenet_yolo_pytorch_b200 as b200
        h0, h1, targets = _inputs(N)
        lo, hi = b200.dist.shard_bounds(N, world, rank)
        losses = [b200.YOLOLoss(VOC_ANCHORS, MASKS[i], C, [352, 352], 0.6, 0.55, val_conf=0.3, iou_weighting=0.02,
                                process_group=dist.group.WORLD) for i in range(2)]
        dets, cnt = b200.decode_nms_padded(h0[lo:hi].to(dev), h1[lo:hi].to(dev), losses)
        g_dets, g_cnt = b200.dist.all_gather_detections(dets, cnt)
        c_rows, c_cnt = b200.dist.all_gather_detections_compact(dets, cnt)
        # the same gather fused into the kernel: kept rows stored into every rank's buffer over NVLink peer mappings
        pg = b200.dist.PeerGather(hi - lo, dets.shape[1])
        tables = b200.fused.head_anchor_table(losses)
        for it in range(4):   # repeated launches reuse the buffers; both fences (flag kernels, NCCL all-reduce)
            pg.decode_nms(h0[lo:hi].to(dev), h1[lo:hi].to(dev), tables, C, 0.3)
            pg.fence(collective=(it == 1))
        torch.cuda.synchronize()
        pg.check()
        p_dets, p_cnt = pg.dets.cpu().numpy().copy(), pg.counts.cpu().numpy().copy()
        pg.close()
        # ... and through NVSwitch multicast memory (one store per row, replicated by the switch), where the box has it:
        # single steps with fences, then the one-call form
        from mobilenet_yolo_pytorch_b200 import _lib
        m_dets = m_cnt = None
        votes = [None] * world
        dist.all_gather_object(votes, bool(_lib.load().b200yolo_mc_supported(rank)))
        if all(votes):
            pm = b200.dist.PeerGather(hi - lo, dets.shape[1], multicast=True)
            assert pm.multicast
            for it in range(3):
                pm.decode_nms(h0[lo:hi].to(dev), h1[lo:hi].to(dev), tables, C, 0.3)
                pm.fence()
            hs = [(h0[lo:hi].to(dev).contiguous(), h1[lo:hi].to(dev).contiguous()) for _ in range(4)]
            pm.run_steps(hs, tables, C, 0.3)
            torch.cuda.synchronize()
            pm.check()
            m_dets, m_cnt = pm.dets.cpu().numpy().copy(), pm.counts.cpu().numpy().copy()
            pm.close()
        x = h1[lo:hi].to(dev).requires_grad_(True)
        tup = losses[1](x, targets[lo:hi])
        tup[0].backward()
        q.put((rank, m_dets, m_cnt, p_dets, p_cnt, c_rows.cpu().numpy(), c_cnt.cpu().numpy(), g_dets.cpu().numpy(), g_cnt.cpu().numpy(), float(tup[0].detach()), [float(v) for v in tup[1:4]] + [float(tup[4]), tup[5], tup[6]],
               x.grad.cpu().numpy(), lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_shards_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (run with gpurun --gpus 2)")
    import mobilenet_yolo_pytorch_b200 as b200
    N, world = 16, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    h0, h1, targets = _inputs(N)
    losses = [b200.YOLOLoss(VOC_ANCHORS, MASKS[i], C, [352, 352], 0.6, 0.55, val_conf=0.3, iou_weighting=0.02) for i in range(2)]
    dets, cnt = b200.decode_nms_padded(h0.to(dev), h1.to(dev), losses)
    dets, cnt = dets.cpu().numpy(), cnt.cpu().numpy()
    x = h1.to(dev).requires_grad_(True)
    tup = losses[1](x, targets)
    tup[0].backward()
    grad = x.grad.cpu().numpy()
    want_rows = np.concatenate([dets[b, :cnt[b]] for b in range(N)], 0)
    for rank, m_dets, m_cnt, p_dets, p_cnt, c_rows, c_cnt, g_dets, g_cnt, loss, stats, g, lo, hi in got:
        assert np.array_equal(g_cnt, cnt) and np.array_equal(c_cnt, cnt) and np.array_equal(p_cnt, cnt)
        if m_dets is not None:                           # multicast gather: the same, on every rank
            assert np.array_equal(m_cnt, cnt)
            for b in range(N):
                assert np.array_equal(m_dets[b, :cnt[b]], dets[b, :cnt[b]])
        else:
            print("NVSwitch multicast not available here: multicast gather not exercised")
        for b in range(N):                               # fused peer gather: every rank holds the whole batch
            assert np.array_equal(p_dets[b, :cnt[b]], dets[b, :cnt[b]])
        assert np.array_equal(c_rows, want_rows)            # compact gather: kept rows only, rank-then-image order
        for b in range(N):
            assert np.array_equal(g_dets[b, :cnt[b]], dets[b, :cnt[b]])
        np.testing.assert_allclose(loss, float(tup[0].detach()), rtol=1e-6)
        np.testing.assert_allclose(stats[:5], [tup[1], tup[2], tup[3], float(tup[4]), tup[5]], rtol=1e-6)
        assert stats[5] == pytest.approx(tup[6])       # count / batch size: sums and image counts are global
        # the shard's gradient carries the batch-global normalisers: it equals the slice of the full gradient
        assert np.abs(g - grad[lo:hi]).max() <= 1e-6 * np.abs(grad).max()
