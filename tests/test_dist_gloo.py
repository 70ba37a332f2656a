"""World-size-2 checks of the N>1 host logic on CPU (gloo): the batch shards by image,
detections are all-gathered in ONE message (counts ride along), and the 16 loss
partial sums are all-reduced BEFORE the batch-global normalisation
(yolo_loss.py:55,224,170-178).  The per-shard numbers fed to the collectives come from
the CPU oracle (there is no GPU here); what is under test is the sharding, the packing
and the reduce-then-finalize order, against the oracle run on the unsharded batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

VOC_ANCHORS = [[143, 265], [153, 121], [280, 279], [20, 37], [49, 94], [73, 201]]
MASKS = [[0, 1, 2], [3, 4, 5]]
C = 20
IMG = [352, 352]
IGN, IOU_T, IOU_W = 0.5623606200028424, 0.5497280113447018, 0.021830872589525777


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(N):
    g = torch.Generator().manual_seed(11)
    h0 = torch.randn(N, 75, 11, 11, generator=g).numpy()
    h1 = torch.randn(N, 75, 22, 22, generator=g).numpy()
    r = np.random.RandomState(5)
    targets = []
    for b in range(N):
        n = [0, 3, 17, 1, 40, 9, 2][b % 7]
        wh = r.rand(n, 2) * 0.45 + 0.02
        c = wh / 2 + r.rand(n, 2) * (1 - wh)
        targets.append(np.concatenate((r.randint(1, C + 1, (n, 1)), c, wh), 1).astype(np.float32))
    return h0, h1, targets


def _sums_from_oracle(o, n_img, cells):
    """The oracle's scalars in the layout of b200yolo_target_loss's partial sums."""
    from mobilenet_yolo_pytorch_b200 import _lib
    s = np.zeros(_lib.S_COUNT)
    s[_lib.S_SQW], s[_lib.S_W] = o["sum_sq_w"], o["sum_w"]
    s[_lib.S_IOU_SQ] = o["l_iou"] * o["n_assign"]
    s[_lib.S_IOU_W] = o["iou_wsum"]
    s[_lib.S_NASSIGN] = o["n_assign"]
    s[_lib.S_OBJ] = o["obj"] * o["n_assign"]
    s[_lib.S_CONF_ALL] = o["sum_conf"]
    s[_lib.S_CLS] = o["cls"] * o["n_assign"]
    s[_lib.S_IOU] = o["avg_iou"] * o["n_assign"]
    s[_lib.S_RECALL] = o["n_recall"]
    s[_lib.S_NCELLS], s[_lib.S_NIMG] = n_img * cells, n_img
    return s


def _worker(rank, world, port, N, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from mobilenet_yolo_pytorch_b200 import dist as b2dist, ops
        oracle.set_threads(1)
        h0, h1, targets = _inputs(N)
        lo, hi = b2dist.shard_bounds(N, world, rank)
        # ---- decode + NMS: shard by image, ONE all-gather
        sa = oracle.scaled_anchors(VOC_ANCHORS, IMG)
        tables = np.stack([sa[MASKS[0]], sa[MASKS[1]]])
        out, oc, _ = oracle.decode_nms_padded(h0[lo:hi], h1[lo:hi], tables, C, 0.3)
        g_out, g_cnt = b2dist.all_gather_detections(torch.from_numpy(out), torch.from_numpy(oc))
        # ---- loss: shard targets the same way, all-reduce the sums, THEN normalise
        my_targets = b2dist.split_targets(targets, world, rank)
        assert len(my_targets) == hi - lo
        o = oracle.target_loss(h1[lo:hi], my_targets, VOC_ANCHORS, MASKS[1], C, IMG, IGN, IOU_T, IOU_W)
        sums = torch.from_numpy(_sums_from_oracle(o, hi - lo, 3 * 22 * 22))
        b2dist.all_reduce_loss_sums(sums)
        res = ops.loss_finalize(sums.numpy(), IOU_W)
        # ---- the rendezvous of the multicast set-up: rank 0's file descriptor reaches the other rank (SCM_RIGHTS over an
        # abstract unix socket whose name travels through the process group); here the descriptor is a pipe
        fd_ok = True
        if rank == 0:
            rd, wr = os.pipe()
            got = b2dist._pass_fd(wr, None, rank, world)
            assert got == wr
            dist.barrier()                                   # the peer has written through its copy of the descriptor
            fd_ok = os.read(rd, 16) == b"from rank 1"
            os.close(rd)
            os.close(wr)
        else:
            got = b2dist._pass_fd(None, None, rank, world)
            os.write(got, b"from rank 1")
            os.close(got)
            dist.barrier()
        q.put((rank, g_out.numpy(), g_cnt.numpy(), res, fd_ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo_shards_equal_unsharded():
    import oracle
    N, world = 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    h0, h1, targets = _inputs(N)
    sa = oracle.scaled_anchors(VOC_ANCHORS, IMG)
    tables = np.stack([sa[MASKS[0]], sa[MASKS[1]]])
    out, oc, _ = oracle.decode_nms_padded(h0, h1, tables, C, 0.3)
    o = oracle.target_loss(h1, targets, VOC_ANCHORS, MASKS[1], C, IMG, IGN, IOU_T, IOU_W)
    want = np.array([o["loss"], o["recall"], o["avg_iou"], o["obj"], o["no_obj"], o["cls"], o["count_per_img"]])
    for rank, g_out, g_cnt, res, fd_ok in got:
        assert fd_ok, "the file descriptor handed over by rank 0 did not work on the other rank"
        assert np.array_equal(g_cnt, oc), f"rank {rank}: gathered counts differ from the unsharded run"
        for b in range(N):
            assert np.array_equal(g_out[b, :oc[b]], out[b, :oc[b]]), f"rank {rank}: image {b} rows differ"
        np.testing.assert_allclose(res, want, rtol=1e-6, atol=1e-9)


def test_per_shard_normalisation_would_be_wrong():
    """Why the sums are reduced before dividing: averaging per-shard losses differs from
    the reference's batch-global normalisation when shards have different weights."""
    import oracle
    from mobilenet_yolo_pytorch_b200 import ops
    N = 8
    _, h1, targets = _inputs(N)
    full = oracle.target_loss(h1, targets, VOC_ANCHORS, MASKS[1], C, IMG, IGN, IOU_T, IOU_W)
    halves = [oracle.target_loss(h1[s], targets[s], VOC_ANCHORS, MASKS[1], C, IMG, IGN, IOU_T, IOU_W)
              for s in (slice(0, 4), slice(4, 8))]
    naive = 0.5 * (halves[0]["loss"] + halves[1]["loss"])
    assert abs(naive - full["loss"]) > 1e-6 * abs(full["loss"])
    s = sum(_sums_from_oracle(h, 4, 3 * 22 * 22) for h in halves)
    assert abs(ops.loss_finalize(s, IOU_W)[0] - full["loss"]) <= 1e-6 * abs(full["loss"])
