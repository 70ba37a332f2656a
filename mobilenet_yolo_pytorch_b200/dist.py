"""Data-parallel sharding of the hot path over the GPUs of one box (one process
per GPU, torch.distributed; NCCL over NVLink on the B200s, gloo in CPU tests).

The batch shards by image (images are independent in decode + NMS,
utils/box.py:16, yolo_loss.py:202-203).  The only exchange steps are
  * an all-gather of the fixed-stride detections + counts, and
  * an all-reduce(SUM) of the 16 loss partial sums (the normalisers of
    yolo_loss.py:55,224,170-178 are batch-global).
No collective sits inside the data path.  ``PeerGather`` fuses the first exchange into the kernel: the output phase
of decode + NMS stores the kept rows into every rank's gather buffer over NVLink peer mappings."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced chunks: the first n % world ranks get one extra image."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_detections(dets: torch.Tensor, counts: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """dets (n_local, K, 7) fp32 and counts (n_local,) int32 of this rank's shard ->
    the same for the whole batch in rank order.  Shards must have equal n_local
    (pad the batch to a multiple of the world size).  The counts travel inside the
    same message as one extra row per image, so this is ONE collective."""
    world = dist.get_world_size(group)
    n, K, _ = dets.shape
    packed = torch.empty((n, K + 1, 7), dtype=torch.float32, device=dets.device)
    packed[:, :K] = dets
    packed[:, K, 0] = counts.to(torch.float32)  # exact for counts < 2^24
    out = torch.empty((world * n, K + 1, 7), dtype=torch.float32, device=dets.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out[:, :K], out[:, K, 0].to(torch.int32)


def all_gather_detections_compact(dets: torch.Tensor, counts: torch.Tensor, group=None):
    """Like ``all_gather_detections`` but only kept rows travel: the shard's rows are packed back to back on the device
    (``b200yolo_compact_rows``), the per-rank totals and per-image counts are gathered (tiny), and ONE all-gather moves
    ``max total`` rows per rank -- 0.5 MB instead of 13 MB per rank for trained-like heads at 256 images.
    Returns (rows (sum of totals, 7) in rank-then-image order, counts (world*n_local,) int32)."""
    from . import ops
    world = dist.get_world_size(group)
    n = dets.shape[0]
    packed, offsets = ops.compact_rows(dets, counts)
    totals = torch.empty((world,), dtype=torch.int32, device=dets.device)
    dist.all_gather_into_tensor(totals, offsets[n:n + 1].contiguous(), group=group)
    all_counts = torch.empty((world * n,), dtype=torch.int32, device=dets.device)
    dist.all_gather_into_tensor(all_counts, counts.contiguous(), group=group)
    tot = totals.cpu().tolist()                      # the one host sync: buffer sizes
    m = max(max(tot), 1)
    recv = torch.empty((world, m, 7), dtype=torch.float32, device=dets.device)
    dist.all_gather_into_tensor(recv, packed[:m].contiguous(), group=group)
    rows = torch.cat([recv[r, :tot[r]] for r in range(world)], 0)
    return rows, all_counts


def all_reduce_loss_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def split_targets(targets: List, world_size: int, rank: int) -> List:
    lo, hi = shard_bounds(len(targets), world_size, rank)
    return targets[lo:hi]


def _pass_fd(fd: Optional[int], group, rank: int, world: int) -> int:
    """Hand rank 0's file descriptor to every rank of the (single-node) group: SCM_RIGHTS over an abstract unix socket
    whose name travels through the process group."""
    import os
    import socket
    src = 0 if group is None else dist.get_global_rank(group, 0)
    obj = [None]
    srv = None
    if rank == 0:
        _pass_fd.counter = getattr(_pass_fd, "counter", 0) + 1
        name = "\0b200yolo-mc-%d-%d" % (os.getpid(), _pass_fd.counter)
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(name)
        srv.listen(world)
        obj = [name]
    dist.broadcast_object_list(obj, src=src, group=group)
    if rank == 0:
        srv.settimeout(60.0)
        for _ in range(world - 1):
            conn, _ = srv.accept()
            socket.send_fds(conn, [b"f"], [fd])
            conn.close()
        srv.close()
        return fd
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    c.settimeout(60.0)
    c.connect(obj[0])
    _, fds, _, _ = socket.recv_fds(c, 16, 1)
    c.close()
    if not fds:
        raise RuntimeError("PeerGather: no file descriptor received from rank 0")
    return fds[0]


class _RawCuda:
    """zero-copy view of raw device memory for ``torch.as_tensor`` (CUDA array interface v2)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerGather:
    """Decode + NMS fused with the all-gather of the detections (``b200yolo_decode_nms_gather``).

    Every rank owns THREE gather buffers (buffer = step % 3) -- ``dets`` (world*n_local, K, 7) fp32 and ``counts``
    (world*n_local,) int32 each -- in memory the other ranks of the node map through CUDA IPC.  A step post-processes
    this rank's shard; its output phase stores the kept rows into the buffers of ALL ranks (its own and, over NVLink,
    the peers'), so the transfer overlaps the kernel instead of following it as a separate NCCL all-gather.
    ``fence`` raises this rank's arrival flag in every rank's flag array once the step's kernel has completed and
    waits until all flags of this rank's own array show the step: every rank's rows are then in this rank's buffer.
    Inside ``run_steps`` a step is ONE kernel launch -- the arrival signal of step s rides on the launch of step s+1
    -- so consecutive steps overlap exactly as they do on one GPU.

    Back-pressure without a release signal.  The kernel of step s stores into buffer s % 3, last used by step s-3.
    ``run_steps`` makes it wait, inside the kernel and before its first store, until EVERY rank has completed step
    s-1; with the per-step API the fence of step s-1 precedes it in the stream.  A rank that has completed step s-1
    has -- in stream order -- moved past whatever it enqueued on the buffer of step s-3.  So: consume the result of
    step s -- ``current()`` -- after its fence with ordinary kernel launches on the same stream BEFORE launching step
    s+1, and no peer can overwrite what a consumer still reads.  One process per GPU on one NVSwitch box, at most 8
    ranks; ``close`` releases the mappings.  ``group`` may be a gloo group (the rendezvous only moves IPC handles).

    ``multicast`` (None: the environment variable B200YOLO_MULTICAST=1 switches it on; True: required; False: off): the
    gather buffers live in memory bound to one NVSwitch multicast object (``b200yolo_mc_*``) and the kernel stores every
    row ONCE, through the multicast view: the switch replicates it into all ranks' buffers, so a GPU sends 1/(R-1) of
    the bytes the peer-mapping path sends.  The arrival flags stay on peer mappings.  Needs one GPU per rank."""

    NBUF = 3

    def __init__(self, n_local: int, cells_per_image: int, group=None, device: Optional[torch.device] = None,
                 multicast: Optional[bool] = None):
        import ctypes as C
        import os
        from . import _lib
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("PeerGather: at most 8 ranks (one NVSwitch box)")
        self.n_local, self.K = int(n_local), int(cells_per_image)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        total = self.world * self.n_local
        row_bytes = (total * self.K * 7 * 4 + 255) // 256 * 256
        cnt_bytes = (total * 4 + 255) // 256 * 256
        self._par_bytes = row_bytes + cnt_bytes
        want_mc = (os.environ.get("B200YOLO_MULTICAST", "0") == "1") if multicast is None else bool(multicast)
        self._mc = None
        if want_mc and self.world >= 2:
            votes: List = [None] * self.world
            dist.all_gather_object(votes, bool(lib.b200yolo_mc_supported(self.device.index)), group=group)
            if all(votes):
                self._mc_setup(lib, self.NBUF * self._par_bytes)
            elif multicast:
                raise RuntimeError("PeerGather: NVSwitch multicast is not available on every rank")
        self.multicast = self._mc is not None
        # (multicast: rows and counts live in the multicast-bound memory, the peer-mapped allocation holds the flags only)
        self._flag_off = 0 if self.multicast else self.NBUF * self._par_bytes    # int[8] arrival flags + int timed_out
        nbytes = self._flag_off + 256
        with torch.cuda.device(self.device):
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(lib.b200yolo_peer_alloc(nbytes, C.byref(ptr), handle))
            self._own = ptr.value
            handles: List = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self._bases, self._opened = [], []
            for r in range(self.world):
                if r == self.rank:
                    self._bases.append(self._own)
                    continue
                p2 = C.c_void_p()
                _lib.check(lib.b200yolo_peer_open(handles[r], C.byref(p2)))
                self._bases.append(p2.value)
                self._opened.append(p2.value)
            g = _lib.Gather()
            g.R, g.rank = self.world, self.rank
            g.multicast = 1 if self.multicast else 0
            for par in range(self.NBUF):
                if self.multicast:
                    g.peer_out[par][0] = self._mc_view + par * self._par_bytes
                    g.peer_count[par][0] = self._mc_view + par * self._par_bytes + row_bytes
                    continue
                for r in range(self.world):
                    g.peer_out[par][r] = self._bases[r] + par * self._par_bytes
                    g.peer_count[par][r] = self._bases[r] + par * self._par_bytes + row_bytes
            for r in range(self.world):
                g.peer_flags[r] = self._bases[r] + self._flag_off
            g.timed_out = self._own + self._flag_off + 64
            g.timeout_s = 5.0
            self._g = g
            self._out_ptrs = [(C.c_void_p * self.world)(*[self._bases[r] + par * self._par_bytes for r in range(self.world)])
                              for par in range(self.NBUF)]
            self._cnt_ptrs = [(C.c_void_p * self.world)(*[self._bases[r] + par * self._par_bytes + row_bytes for r in range(self.world)])
                              for par in range(self.NBUF)]
            self._flag_ptrs = (C.c_void_p * self.world)(*[b + self._flag_off for b in self._bases])
            self._step = 0        # steps launched so far
            self._fenced = 0      # steps whose fence has been launched
            flags_raw = torch.as_tensor(_RawCuda(self._own, nbytes), device=self.device)
            raw = (torch.as_tensor(_RawCuda(self._mc_local, self.NBUF * self._par_bytes), device=self.device)
                   if self.multicast else flags_raw)
            self._dets = [raw[par * self._par_bytes:par * self._par_bytes + total * self.K * 7 * 4].view(torch.float32).view(total, self.K, 7)
                          for par in range(self.NBUF)]
            self._counts = [raw[par * self._par_bytes + row_bytes:par * self._par_bytes + row_bytes + total * 4].view(torch.int32)
                            for par in range(self.NBUF)]
            self._timed_out = flags_raw[self._flag_off + 64:self._flag_off + 68].view(torch.int32)
            self._flag = torch.zeros((1,), dtype=torch.float32, device=self.device)
        dist.barrier(group=group)  # every rank has mapped every buffer before anyone writes

    def _mc_setup(self, lib, nbytes: int) -> None:
        """the multicast object (rank 0 creates it, the others import its file descriptor), this rank's memory bound to
        it, and the two views"""
        import ctypes as C
        import os
        from . import _lib
        with torch.cuda.device(self.device):
            h = C.c_void_p()
            fd = C.c_int(-1)
            err = None
            if self.rank == 0:
                try:
                    _lib.check(lib.b200yolo_mc_create(nbytes, self.world, C.byref(fd), C.byref(h)))
                except Exception as e:  # noqa: BLE001  (the other ranks must not be left waiting for the descriptor)
                    err = str(e)
            errs: List = [None] * self.world
            dist.all_gather_object(errs, err, group=self.group)
            if errs[0] is not None:
                raise RuntimeError("PeerGather: multicast object: " + errs[0])
            got = _pass_fd(fd.value if self.rank == 0 else None, self.group, self.rank, self.world)
            try:
                if self.rank != 0:
                    _lib.check(lib.b200yolo_mc_import(nbytes, self.world, got, C.byref(h)))
            finally:
                os.close(got)               # (the driver keeps its own reference)
            self._mc = h
            _lib.check(lib.b200yolo_mc_add_device(h))
            dist.barrier(group=self.group)      # every device is in the group before memory is bound
            local, view = C.c_void_p(), C.c_void_p()
            _lib.check(lib.b200yolo_mc_bind(h, C.byref(local), C.byref(view)))
            self._mc_local, self._mc_view = local.value, view.value
            dist.barrier(group=self.group)

    # ---- results of the most recently launched step (valid on the stream after its fence)
    @property
    def dets(self) -> torch.Tensor:
        return self._dets[(self._step - 1) % self.NBUF]

    @property
    def counts(self) -> torch.Tensor:
        return self._counts[(self._step - 1) % self.NBUF]

    def current(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.dets, self.counts

    def _shape_args(self, head0, head1, anchor_wh2, num_classes):
        from . import ops
        ops._require_cuda(head0, "head0")
        ops._require_cuda(head1, "head1")
        N, ch, H0, W0 = head0.shape
        _, _, H1, W1 = head1.shape
        A = ch // (5 + num_classes)
        if N != self.n_local or A * (H0 * W0 + H1 * W1) != self.K:
            raise RuntimeError("PeerGather: shard shape differs from the buffer's")
        aw = ops._host_f32(anchor_wh2).reshape(2, A, 2)
        return N, A, H0, W0, H1, W1, aw

    def decode_nms(self, head0: torch.Tensor, head1: torch.Tensor, anchor_wh2, num_classes: int, conf_thr: float,
                   iou_thr: float = 0.45) -> None:
        """launch one step on the current stream; results are complete on every rank after ``fence()``"""
        import ctypes as C
        from . import _lib
        if self._fenced != self._step:
            raise RuntimeError("PeerGather: fence() the previous step before launching the next one")
        head0, head1 = head0.contiguous(), head1.contiguous()
        N, A, H0, W0, H1, W1, aw = self._shape_args(head0, head1, anchor_wh2, num_classes)
        one = (_lib.Batch * 1)(_lib.Batch(head0.data_ptr(), head1.data_ptr(), None, None, None))
        _lib.check(_lib.load().b200yolo_decode_nms_gather_steps(
            C.byref(self._g), one, 1, self._step, 0, N, A, num_classes, H0, W0, H1, W1, aw.ctypes.data,
            float(np.float32(conf_thr)), float(iou_thr), torch.cuda.current_stream(self.device).cuda_stream))
        self._step += 1

    def fence(self, collective: bool = False, timeout_s: float = 5.0) -> None:
        """Cross-rank fence for the step just launched: this rank's arrival is announced to every rank once the step's
        kernel has completed, and the stream then waits until every rank has announced the step -- all rows are in this
        rank's buffer (``check()`` tells whether a peer failed to arrive in time).  ``collective=True`` additionally
        runs a 1-element NCCL all-reduce (for comparison; the flags are still raised, they carry the back-pressure)."""
        from . import _lib
        lib = _lib.load()
        if self._fenced >= self._step:
            return
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.b200yolo_peer_fence(self._flag_ptrs, self._own + self._flag_off, self.world, self.rank, self._step,
                                           float(timeout_s), self._own + self._flag_off + 64, st))
        self._fenced = self._step
        if collective:
            dist.all_reduce(self._flag, group=self.group)

    def run_steps(self, heads, anchor_wh2, num_classes: int, conf_thr: float, iou_thr: float = 0.45, plan=None):
        """``len(heads)`` steps -- kernel + arrival signal + fence each -- issued from C in ONE call
        (``b200yolo_decode_nms_gather_steps``); ``heads`` is a sequence of (head0, head1).  The fence of a step is issued
        one step late (it then costs nothing) and the last one at the end of the call.  Returns a reusable plan (pass it
        back as ``plan`` to skip the set-up).  For pipelines without a per-step consumer (benchmarks, or consumers that
        read ``current()`` after the call)."""
        import ctypes as C
        from . import _lib
        if self._fenced != self._step:
            raise RuntimeError("PeerGather: fence() the previous step first")
        if plan is None:
            h0, h1 = heads[0]
            N, A, H0, W0, H1, W1, aw = self._shape_args(h0, h1, anchor_wh2, num_classes)
            arr = (_lib.Batch * len(heads))()
            for k, (a0, a1) in enumerate(heads):
                if a0.shape != h0.shape or a1.shape != h1.shape or not a0.is_contiguous() or not a1.is_contiguous():
                    raise RuntimeError("PeerGather.run_steps: NCHW-contiguous heads of one shape")
                arr[k] = _lib.Batch(a0.data_ptr(), a1.data_ptr(), None, None, None)
            plan = (arr, len(heads), list(heads), (N, A, num_classes, H0, W0, H1, W1, aw.ctypes.data, float(np.float32(conf_thr)),
                                                  float(iou_thr)), aw)
        arr, n = plan[0], plan[1]
        _lib.check(_lib.load().b200yolo_decode_nms_gather_steps(
            C.byref(self._g), arr, n, self._step, 1, *plan[3], torch.cuda.current_stream(self.device).cuda_stream))
        self._step += n
        self._fenced = self._step
        return plan

    def check(self) -> None:
        """host-synchronising: raise if a fence gave up waiting for a peer"""
        if int(self._timed_out.item()):
            raise RuntimeError("PeerGather: a peer did not reach the fence in time")

    def close(self) -> None:
        from . import _lib
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)     # nobody still writes into a buffer that is about to go away
        for p2 in self._opened:
            lib.b200yolo_peer_close(p2)
        self._opened = []
        if self._own:
            self._dets = self._counts = self._timed_out = None
            lib.b200yolo_peer_free(self._own)
            self._own = None
        if self._mc is not None:
            lib.b200yolo_mc_free(self._mc)
            self._mc = None
