"""The fused all-gather (dist.PeerGather: b200yolo_decode_nms_gather + b200yolo_peer_fence over CUDA-IPC mappings) on ONE
GPU: two processes share cuda:0, rendezvous over gloo, and map each other's gather buffers through CUDA IPC exactly
as two ranks on two GPUs do (the 2-GPU NCCL variant is tests/test_dist_nccl.py).  Several steps with different
inputs, a consumer that reads the gathered buffers between the fence of a step and the launch of the next one
(the back-pressure), the one-call multi-step entry, and the sharded YOLOLoss with its all-reduce of the 16
partial sums (loss, statistics and gradient slices equal the unsharded ones).
"""
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
STEPS = 5


def _heads(N, step):
    g = torch.Generator().manual_seed(100 + step)
    h0 = torch.randn(N, 75, 11, 11, generator=g)
    h1 = torch.randn(N, 75, 22, 22, generator=g)
    if step % 2:   # alternate dense and sparse steps: very different row counts land in the same buffer slots
        h0.view(N, 3, 25, 11, 11)[:, :, 4] -= 2.6
        h1.view(N, 3, 25, 22, 22)[:, :, 4] -= 2.6
    return h0, h1


def _targets(N):
    r = np.random.RandomState(9)
    out = []
    for b in range(N):
        n = [4, 0, 25, 2, 60, 9, 1, 13][b % 8]
        wh = r.rand(n, 2) * 0.45 + 0.02
        c = wh / 2 + r.rand(n, 2) * (1 - wh)
        out.append(torch.from_numpy(np.concatenate((r.randint(1, C + 1, (n, 1)), c, wh), 1).astype(np.float32)))
    return out


def _digest(dets, counts):
    """what a consumer would read: per image the count and a checksum of its kept rows"""
    K = dets.shape[1]
    mask = (torch.arange(K, device=dets.device)[None, :] < counts[:, None]).to(dets.dtype)
    return counts.clone(), (dets.double() * mask[..., None].double()).sum(dim=(1, 2))


def _worker(rank, world, port, N, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        import mobilenet_yolo_pytorch_b200 as b200
        losses = [b200.YOLOLoss(VOC_ANCHORS, MASKS[i], C, [352, 352], 0.6, 0.55, val_conf=0.3) for i in range(2)]
        tables = b200.fused.head_anchor_table(losses)
        lo, hi = b200.dist.shard_bounds(N, world, rank)
        pg = b200.dist.PeerGather(hi - lo, 1815, device=dev)
        digests = []
        for step in range(STEPS):
            h0, h1 = _heads(N, step)
            pg.decode_nms(h0[lo:hi].to(dev), h1[lo:hi].to(dev), tables, C, 0.3)
            pg.fence(timeout_s=30.0)
            d, c = pg.current()
            cc, ck = _digest(d, c)            # the consumer: ordinary kernels on the same stream, before the next step
            digests.append((cc, ck))
        # the one-call form: three more steps issued from C, results of the last one read afterwards
        hs = [tuple(t[lo:hi].to(dev).contiguous() for t in _heads(N, STEPS + k)) for k in range(3)]
        pg.run_steps(hs, tables, C, 0.3)
        d, c = pg.current()
        digests.append(_digest(d, c))
        torch.cuda.synchronize()
        pg.check()
        out = [(a.cpu().numpy(), b.cpu().numpy()) for a, b in digests]
        rows = d.cpu().numpy().copy()
        pg.close()
        # the other exchange step of the path: YOLOLoss on a shard with the 16 partial sums all-reduced before the
        # batch-global division (here over gloo, which stages the CUDA tensor through the host; NCCL on real ranks)
        lh0, lh1 = _heads(N, 50)
        targets = _targets(N)
        ll = b200.YOLOLoss(VOC_ANCHORS, MASKS[1], C, [352, 352], 0.6, 0.55, iou_weighting=0.02, process_group=dist.group.WORLD)
        x = lh1[lo:hi].to(dev).requires_grad_(True)
        tup = ll(x, targets[lo:hi])
        tup[0].backward()
        loss_out = ([float(tup[0].detach()), tup[1], tup[2], tup[3], float(tup[4]), tup[5], tup[6]], x.grad.cpu().numpy(), lo, hi)
        # the same shard through the no-synchronisation path (lazy_stats + packed device targets): same seven values
        ll.lazy_stats = True
        x2 = lh1[lo:hi].to(dev).requires_grad_(True)
        lz = ll(x2, b200.ops.PackedTargets.from_list(targets[lo:hi], dev))
        lz[0].backward()
        ll.check()
        assert all(v.is_cuda for v in lz)
        np.testing.assert_allclose([float(v) for v in lz], loss_out[0], rtol=2e-7, atol=1e-9)
        assert torch.equal(x2.grad, x.grad)
        q.put((rank, out, rows, loss_out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_peer_gather_two_processes_one_gpu():
    import mobilenet_yolo_pytorch_b200 as b200
    N, world = 8, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = []
    try:
        got = [q.get(timeout=400) for _ in range(world)]
    finally:
        for p in procs:
            p.join(60)
            if p.is_alive():
                p.kill()
    for p in procs:
        assert p.exitcode == 0
    # the single-process answer for every step
    dev = torch.device("cuda", 0)
    losses = [b200.YOLOLoss(VOC_ANCHORS, MASKS[i], C, [352, 352], 0.6, 0.55, val_conf=0.3) for i in range(2)]
    want = []
    last_rows = None
    for step in list(range(STEPS)) + [STEPS + 2]:
        h0, h1 = _heads(N, step)
        dets, cnt = b200.decode_nms_padded(h0.to(dev), h1.to(dev), losses)
        a, b = _digest(dets, cnt)
        want.append((a.cpu().numpy(), b.cpu().numpy()))
        last_rows = (dets.cpu().numpy(), cnt.cpu().numpy())
    # the unsharded loss and gradient
    _, lh1 = _heads(N, 50)
    targets = _targets(N)
    ll = b200.YOLOLoss(VOC_ANCHORS, MASKS[1], C, [352, 352], 0.6, 0.55, iou_weighting=0.02)
    x = lh1.to(dev).requires_grad_(True)
    tup = ll(x, targets)
    tup[0].backward()
    full = [float(tup[0].detach()), tup[1], tup[2], tup[3], float(tup[4]), tup[5], tup[6]]
    grad = x.grad.cpu().numpy()
    for rank, digests, rows, (lt, g, lo, hi) in got:
        # reduce-then-normalise: every shard reports the loss and statistics of the WHOLE batch, and its gradient is the
        # slice of the full-batch gradient (the backward kernel reads the all-reduced sums)
        np.testing.assert_allclose(lt, full, rtol=1e-6, atol=1e-9)
        assert np.abs(g - grad[lo:hi]).max() <= 1e-6 * np.abs(grad).max()
        assert len(digests) == len(want)
        for k, ((c_got, s_got), (c_want, s_want)) in enumerate(zip(digests, want)):
            assert np.array_equal(c_got, c_want), f"rank {rank}, step {k}: counts differ"
            np.testing.assert_allclose(s_got, s_want, rtol=1e-12, err_msg=f"rank {rank}, step {k}: gathered rows differ")
        for b in range(N):   # last step, row for row
            assert np.array_equal(rows[b, :last_rows[1][b]], last_rows[0][b, :last_rows[1][b]])
