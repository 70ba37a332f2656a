"""Thin tensor-level wrappers over the C ABI (device pointers + current stream).

PyTorch is plumbing here: it owns device memory and streams; every computation
is a kernel in libb200yolo.so.  Inputs must be CUDA tensors -- there is no CPU
path (north star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Sequence
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib

LARGE_MAX_CELLS = 16384   # b200yolo_decode_nms_large: sort keys of one image in shared memory
NMS_IOU_THRESHOLD = 0.45  # utils/box.py:28
_WORKSPACES = {}  # (device index, stream) -> cached target-loss workspace tensor
_HOST_WS = {}     # device index -> staging of decode_nms_host (b200yolo_decode_nms_host_ws)


def set_exact_decode(exact: bool) -> None:
    """b200yolo_set_exact_decode: decode with the reference's own operations (IEEE sigmoid / expf, true division) so that
    rows and detections are bit-identical to the reference on device='cuda'; ~3 % slower.  Process-wide, default off."""
    _lib.load().b200yolo_set_exact_decode(1 if exact else 0)


def scaled_anchors(anchors, img_size) -> np.ndarray:
    """yolo_loss.py:214: python-double division, rounded to fp32 when it enters a
    FloatTensor (pre_maps :67-68)."""
    return np.array([[aw / img_size[0], ah / img_size[1]] for aw, ah in anchors], dtype=np.float64).astype(np.float32)


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: the b200yolo kernels have no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what} must be float32 (got {t.dtype})")


def _host_f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def bind_host_to_gpu(device_index: int) -> dict:
    """Pin the calling thread to the CPUs closest to a GPU (NVML's affinity mask of the device), so that pinned host
    buffers allocated afterwards land on that GPU's NUMA node and its H2D / D2H copies do not cross the socket link.
    For one-process-per-GPU jobs that feed the host entry (``decode_nms_host``); call it before allocating the buffers.
    Returns {"bound": bool, "cpus": n, ...}; never raises (no NVML, no permission: {"bound": False, "why": ...})."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        try:
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:  # noqa: BLE001  (older torch without the PCI ids: NVML index = CUDA index when nothing is masked)
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return {"bound": False, "why": "no allowed CPU in the device's affinity mask"}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "cpus": len(cpus), "of": len(allowed), "first_cpu": min(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": str(e)[:120]}


class CandidateList(list):
    """list[N] of (n_b, 7) views, as YOLOLoss.forward(input) returns it
    (yolo_loss.py:202-204), that also remembers the padded device buffer it views
    so utils.box.nms can consume it without repacking."""
    padded: torch.Tensor  # (N, stride, 7)
    counts: torch.Tensor  # (N,) int32, device
    ids: Optional[torch.Tensor]  # (N, stride) int32 cell ids, device


class LazyCandidates(Sequence):
    """What ``YOLOLoss.forward(input)`` returns when ``lazy_eval`` is on: the head tensor and the decode parameters
    as they were at call time, with nothing launched yet.  ``utils.box.nms`` on two of these runs the FUSED kernel
    (one launch instead of two decodes + NMS); any other use -- ``[b]``, iteration, ``.padded`` -- decodes first and
    behaves like the eager list.  ``len()`` needs no work (one entry per image, yolo_loss.py:202-204)."""

    def __init__(self, head: torch.Tensor, anchor_wh: np.ndarray, num_classes: int, conf_thr: float):
        self.head, self.anchor_wh, self.num_classes, self.conf_thr = head, anchor_wh, int(num_classes), conf_thr
        self._list: Optional[CandidateList] = None

    @property
    def pending(self) -> bool:
        return self._list is None

    def materialise(self) -> "CandidateList":
        if self._list is None:
            rows, count, ids = decode_head_padded(self.head, self.anchor_wh, self.num_classes, self.conf_thr, want_ids=True)
            self._list = _as_list(rows, count, ids)
        return self._list

    def __len__(self) -> int:
        return int(self.head.shape[0])

    def __getitem__(self, i):
        return self.materialise()[i]

    def __iter__(self):
        return iter(self.materialise())

    padded = property(lambda self: self.materialise().padded)
    counts = property(lambda self: self.materialise().counts)
    ids = property(lambda self: self.materialise().ids)


def _as_list(padded: torch.Tensor, counts_dev: torch.Tensor, ids: Optional[torch.Tensor] = None) -> CandidateList:
    counts = counts_dev.cpu().tolist()  # the one D2H sync needed to build a ragged python list
    out = CandidateList(padded[b, :n] for b, n in enumerate(counts))
    out.padded, out.counts, out.ids = padded, counts_dev, ids
    out.host_counts = counts
    return out


def decode_head_padded(head: torch.Tensor, anchor_wh, num_classes: int, conf_thr: float, want_ids: bool = False):
    """b200yolo_decode_head: returns (rows (N,cells,7), count (N,) int32[, ids (N,cells) int32])."""
    _require_cuda(head, "head")
    head = head.contiguous()
    N, ch, H, W = head.shape
    attrs = 5 + num_classes
    if ch % attrs:
        raise RuntimeError(f"channel dim {ch} is not a multiple of 5+num_classes={attrs}")
    A = ch // attrs
    aw = _host_f32(anchor_wh).reshape(A, 2)
    cells = A * H * W
    with _on_device(head.device):
        rows = torch.empty((N, cells, 7), dtype=torch.float32, device=head.device)
        count = torch.empty((N,), dtype=torch.int32, device=head.device)
        ids = torch.empty((N, cells), dtype=torch.int32, device=head.device) if want_ids else None
        _lib.check(_lib.load().b200yolo_decode_head(
            head.data_ptr(), N, A, num_classes, H, W, aw.ctypes.data, float(np.float32(conf_thr)), rows.data_ptr(),
            count.data_ptr(), ids.data_ptr() if want_ids else None, _stream(head)))
    return (rows, count, ids) if want_ids else (rows, count)


def nms_padded(cand0: torch.Tensor, count0: torch.Tensor, cand1: Optional[torch.Tensor], count1: Optional[torch.Tensor],
               num_classes: int, iou_thr: float = NMS_IOU_THRESHOLD, want_idx: bool = False):
    """b200yolo_nms on fixed-stride candidates: returns (out (N,S,7), out_count (N,)[, out_idx (N,S)])."""
    _require_cuda(cand0, "cand0")
    cand0 = cand0.contiguous()
    N, s0 = cand0.shape[0], cand0.shape[1]
    s1 = 0
    if cand1 is not None:
        _require_cuda(cand1, "cand1")
        cand1 = cand1.contiguous()
        s1 = cand1.shape[1]
    S = max(s0 + s1, 1)
    with _on_device(cand0.device):
        out = torch.empty((N, S, 7), dtype=torch.float32, device=cand0.device)
        oc = torch.empty((N,), dtype=torch.int32, device=cand0.device)
        oi = torch.empty((N, S), dtype=torch.int32, device=cand0.device) if want_idx else None
        _lib.check(_lib.load().b200yolo_nms(
            cand0.data_ptr(), count0.data_ptr(), s0, cand1.data_ptr() if cand1 is not None else None,
            count1.data_ptr() if cand1 is not None else None, s1, N, num_classes, float(iou_thr), out.data_ptr(),
            oc.data_ptr(), oi.data_ptr() if want_idx else None, _stream(cand0)))
    return (out, oc, oi) if want_idx else (out, oc)


class _on_device:
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current (the context manager costs ~3 us)."""

    def __init__(self, dev: torch.device):
        self.ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def decode_nms_padded(head0: torch.Tensor, head1: torch.Tensor, anchor_wh2, num_classes: int, conf_thr: float,
                      iou_thr: float = NMS_IOU_THRESHOLD, want_idx: bool = False,
                      out: Optional[torch.Tensor] = None, out_count: Optional[torch.Tensor] = None,
                      out_idx: Optional[torch.Tensor] = None, force_large: bool = False):
    """b200yolo_decode_nms: ONE launch, no host sync.  Returns (out (N,K,7), out_count (N,)[, out_idx (N,K)]).
    Pre-allocated outputs may be passed (CUDA-graph capture, NCCL send buffers).  Images with more cells than one
    CTA can stage in shared memory (e.g. 832x832 inputs: 10 140 cells) go through b200yolo_decode_nms_large with a
    workspace from torch's allocator; `force_large` takes that path for any shape (tests)."""
    _require_cuda(head0, "head0")
    _require_cuda(head1, "head1")
    attrs = 5 + num_classes
    # channels-last heads (what cuDNN prefers) are consumed as they are; anything else becomes NCHW-contiguous
    nhwc = (attrs <= 32 and head0.dim() == 4 and not head0.is_contiguous() and not head1.is_contiguous()
            and head0.is_contiguous(memory_format=torch.channels_last)
            and head1.is_contiguous(memory_format=torch.channels_last))
    if force_large:
        nhwc = False
    if not nhwc:
        if not head0.is_contiguous():
            head0 = head0.contiguous()
        if not head1.is_contiguous():
            head1 = head1.contiguous()
    N, ch, H0, W0 = head0.shape
    N1, ch1, H1, W1 = head1.shape
    if N1 != N or ch1 != ch or ch % attrs:
        raise RuntimeError("head shapes do not match (N, A*(5+C), H, W) for both heads")
    A = ch // attrs
    K = A * H0 * W0 + A * H1 * W1
    aw = anchor_wh2 if (isinstance(anchor_wh2, np.ndarray) and anchor_wh2.dtype == np.float32
                        and anchor_wh2.flags.c_contiguous and anchor_wh2.size == 4 * A) else _host_f32(anchor_wh2).reshape(2, A, 2)
    dev = head0.device
    with _on_device(dev):
        if out is None:
            out = torch.empty((N, K, 7), dtype=torch.float32, device=dev)
        if out_count is None:
            out_count = torch.empty((N,), dtype=torch.int32, device=dev)
        if want_idx and out_idx is None:
            out_idx = torch.empty((N, K), dtype=torch.int32, device=dev)
        lib = _lib.load()
        args = (N, A, num_classes, H0, W0, H1, W1, aw.ctypes.data, float(np.float32(conf_thr)), float(iou_thr),
                out.data_ptr(), out_count.data_ptr(), out_idx.data_ptr() if want_idx else None,
                torch.cuda.current_stream(dev).cuda_stream)
        rc = -2
        if not force_large:
            rc = (lib.b200yolo_decode_nms_nhwc if nhwc else lib.b200yolo_decode_nms)(head0.data_ptr(), head1.data_ptr(), *args)
            if rc == -2 and nhwc:  # no shared memory left for the channels-last staging: NCHW copies, then the planar kernel
                head0, head1 = head0.contiguous(), head1.contiguous()
                rc = lib.b200yolo_decode_nms(head0.data_ptr(), head1.data_ptr(), *args)
        if rc == -2 and K <= LARGE_MAX_CELLS:  # too many cells for one CTA's shared memory: records in a workspace
            if nhwc:
                head0, head1 = head0.contiguous(), head1.contiguous()
            ws = torch.empty((int(lib.b200yolo_decode_nms_large_workspace_bytes(N, K)),), dtype=torch.uint8, device=dev)
            rc = lib.b200yolo_decode_nms_large(head0.data_ptr(), head1.data_ptr(), *args[:-1], ws.data_ptr(), ws.numel(), args[-1])
        if rc:
            _lib.check(rc)
    return (out, out_count, out_idx) if want_idx else (out, out_count)


class BatchPlan:
    """A list of batches of ONE shape for ``b200yolo_decode_nms_batches``: the descriptor array is built once, then
    ``run()`` issues every launch from C in one call (no Python between the launches; launch k > 0 overlaps the tail of
    launch k - 1, see include/b200yolo.h).  ``batches`` is a sequence of ``(head0, head1, out, out_count[, out_idx])``
    CUDA tensors (NCHW-contiguous heads of equal shape)."""

    def __init__(self, batches, anchor_wh2, num_classes: int, conf_thr: float, iou_thr: float = NMS_IOU_THRESHOLD):
        if len(batches) == 0:
            raise RuntimeError("BatchPlan needs at least one batch")
        h0, h1 = batches[0][0], batches[0][1]
        _require_cuda(h0, "head0")
        _require_cuda(h1, "head1")
        attrs = 5 + num_classes
        N, ch, H0, W0 = h0.shape
        _, _, H1, W1 = h1.shape
        if ch % attrs:
            raise RuntimeError("head shapes do not match (N, A*(5+C), H, W)")
        A = ch // attrs
        self.K = A * H0 * W0 + A * H1 * W1
        self.dev = h0.device
        self.keep = list(batches)   # the tensors must outlive the plan
        self.aw = _host_f32(anchor_wh2).reshape(2, A, 2)
        self.arr = (_lib.Batch * len(batches))()
        for k, bt in enumerate(batches):
            a0, a1, out, cnt = bt[0], bt[1], bt[2], bt[3]
            idx = bt[4] if len(bt) > 4 else None
            if a0.shape != h0.shape or a1.shape != h1.shape or not a0.is_contiguous() or not a1.is_contiguous():
                raise RuntimeError("BatchPlan: every batch needs NCHW-contiguous heads of the first batch's shape")
            if tuple(out.shape) != (N, self.K, 7) or cnt.numel() != N:
                raise RuntimeError("BatchPlan: out must be (N, K, 7) and out_count (N,)")
            self.arr[k] = _lib.Batch(a0.data_ptr(), a1.data_ptr(), out.data_ptr(), cnt.data_ptr(),
                                     idx.data_ptr() if idx is not None else None)
        self.args = (N, A, num_classes, H0, W0, H1, W1, self.aw.ctypes.data, float(np.float32(conf_thr)), float(iou_thr))
        self.n = len(batches)
        self._graph = None

    def capture(self) -> "BatchPlan":
        """b200yolo_plan_create: freeze the whole list into one CUDA graph; ``run()`` over the whole list then costs the
        host one graph launch instead of one kernel launch per batch."""
        if self._graph is None:
            h = C.c_void_p()
            with _on_device(self.dev):
                _lib.check(_lib.load().b200yolo_plan_create(C.cast(self.arr, C.c_void_p), self.n, *self.args, C.byref(h)))
            self._graph = h
        return self

    def close(self) -> None:
        if self._graph is not None:
            h, self._graph = self._graph, None
            _lib.check(_lib.load().b200yolo_plan_destroy(h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, first: int = 0, count: Optional[int] = None):
        """Launch batches [first, first + count) on the current stream of the plan's device."""
        count = self.n - first if count is None else count
        if first < 0 or count < 0 or first + count > self.n:
            raise RuntimeError("BatchPlan.run: range outside the plan")
        if self._graph is not None and first == 0 and count == self.n:
            with _on_device(self.dev):
                _lib.check(_lib.load().b200yolo_plan_launch(self._graph, torch.cuda.current_stream(self.dev).cuda_stream))
            return
        ptr = C.cast(C.byref(self.arr, first * C.sizeof(_lib.Batch)), C.c_void_p)
        with _on_device(self.dev):
            _lib.check(_lib.load().b200yolo_decode_nms_batches(ptr, count, *self.args,
                                                               torch.cuda.current_stream(self.dev).cuda_stream))


def decode_nms_host(head0: torch.Tensor, head1: torch.Tensor, anchor_wh2, num_classes: int, conf_thr: float,
                    iou_thr: float = NMS_IOU_THRESHOLD, device: int = 0, out: Optional[torch.Tensor] = None,
                    out_count: Optional[torch.Tensor] = None):
    """b200yolo_decode_nms_host_ws: HOST tensors in (pinned recommended), HOST padded detections out (rows past an
    image's count are not written); chunked H2D / kernel / D2H pipeline inside the library, device staging from
    torch's allocator.  Not for concurrent use from several threads on one device (one staging buffer per device)."""
    if head0.is_cuda or head1.is_cuda:
        raise RuntimeError("decode_nms_host takes host tensors")
    head0, head1 = head0.contiguous(), head1.contiguous()
    N, ch, H0, W0 = head0.shape
    _, _, H1, W1 = head1.shape
    attrs = 5 + num_classes
    A = ch // attrs
    K = A * H0 * W0 + A * H1 * W1
    aw = _host_f32(anchor_wh2).reshape(2, A, 2)
    if out is None:
        out = torch.empty((N, K, 7), dtype=torch.float32).pin_memory()
    if out_count is None:
        out_count = torch.empty((N,), dtype=torch.int32).pin_memory()
    lib = _lib.load()
    # device staging from torch's allocator (the library allocates nothing), kept per device and grown on demand
    need = int(lib.b200yolo_decode_nms_host_workspace_bytes(N, A, num_classes, H0, W0, H1, W1))
    ws = _HOST_WS.get(int(device))
    if ws is None or ws.numel() < need:
        ws = _HOST_WS[int(device)] = torch.empty((need,), dtype=torch.uint8, device=torch.device("cuda", int(device)))
    _lib.check(lib.b200yolo_decode_nms_host_ws(
        head0.data_ptr(), head1.data_ptr(), N, A, num_classes, H0, W0, H1, W1, aw.ctypes.data,
        float(np.float32(conf_thr)), float(iou_thr), out.data_ptr(), out_count.data_ptr(), ws.data_ptr(), ws.numel(), int(device)))
    return out, out_count


def compact_rows(dets: torch.Tensor, counts: torch.Tensor):
    """b200yolo_compact_rows: (N, K, 7) fixed-stride detections + (N,) counts -> (packed (N*K, 7) whose first
    offsets[N] rows are the kept rows back to back, offsets (N+1,) int32), both on the device, no host sync."""
    _require_cuda(dets, "dets")
    dets = dets.contiguous()
    N, K, _ = dets.shape
    with _on_device(dets.device):
        packed = torch.empty((max(N * K, 1), 7), dtype=torch.float32, device=dets.device)
        offsets = torch.empty((N + 1,), dtype=torch.int32, device=dets.device)
        _lib.check(_lib.load().b200yolo_compact_rows(dets.data_ptr(), counts.data_ptr(), N, K, packed.data_ptr(),
                                                     offsets.data_ptr(), _stream(dets)))
    return packed, offsets


def pairwise(set_1: torch.Tensor, set_2: torch.Tensor, mode: int) -> torch.Tensor:
    _require_cuda(set_1, "set_1")
    _require_cuda(set_2, "set_2")
    a = set_1.reshape(-1, 4).contiguous()
    b = set_2.reshape(-1, 4).contiguous()
    with _on_device(a.device):
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
        _lib.check(_lib.load().b200yolo_pairwise(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], mode,
                                                 out.data_ptr(), _stream(a)))
    return out


_PACK_CACHE = {}     # device -> [key, targets list, packed result]: ONE more use only (the second head of the same step)
_PIN_STAGING = {}    # device -> [pinned staging tensor for the packed rows, event of the last copy out of it]


def pack_targets(targets, device) -> Tuple[torch.Tensor, torch.Tensor, int, List[int]]:
    """list[N] of (n_b,5) tensors/arrays (reference: CPU tensors, train.py:246) -> one (G,5) device buffer + (N+1,)
    device offsets.  One concatenation, one staged H2D copy each.  The reference hands ONE `targets` list to both
    heads of a step (mbv2_yolo.py:158), so the packed copy serves exactly one more call with the same list object --
    same length, same first / last tensors at the same in-place version -- and is dropped then: a list that is
    mutated or refilled between steps is always packed again, and nothing is kept alive beyond the step."""
    def _ver(t):
        return (id(t), getattr(t, "_version", 0), t.data_ptr() if isinstance(t, torch.Tensor) else 0)
    key = (id(targets), len(targets)) + ((_ver(targets[0]) + _ver(targets[-1]) + _ver(targets[len(targets) // 2]))
                                         if len(targets) else ())
    hit = _PACK_CACHE.pop(device, None)
    if hit is not None and hit[0] == key and hit[1] is targets:
        return hit[2]
    try:  # the reference's case: CPU tensors of shape (n_b, 5) -- one torch.cat, no per-image Python work beyond the counts
        counts = [t.shape[0] for t in targets]
        flat = torch.cat(targets) if len(targets) else None
        if flat is not None and (flat.dim() != 2 or flat.shape[1] != 5):
            raise TypeError
    except (TypeError, AttributeError, RuntimeError, IndexError):  # arrays, lists, ragged / 1-D empties: the slow way
        counts = [int(t.shape[0]) if hasattr(t, "shape") and len(t.shape) else 0 for t in targets]
        rows = [torch.as_tensor(t).reshape(n, 5) for t, n in zip(targets, counts) if n]
        flat = torch.cat(rows) if rows else None
    G = int(sum(counts))
    offs = np.zeros(len(targets) + 1, np.int32)
    np.cumsum(counts, out=offs[1:])
    if G:
        flat = flat.to(dtype=torch.float32, device="cpu")
        entry = _PIN_STAGING.get(device)
        if entry is not None:
            entry[1].synchronize()   # the previous H2D copy out of the staging buffer has finished
        if entry is None or entry[0].shape[0] < G:
            entry = _PIN_STAGING[device] = [torch.empty((max(G, 1024), 5), dtype=torch.float32).pin_memory(),
                                            torch.cuda.Event()]
        stage, done = entry
        stage[:G].copy_(flat)
        gt = stage[:G].to(device, non_blocking=True)
        done.record(torch.cuda.current_stream(device))
    else:
        gt = torch.zeros((1, 5), dtype=torch.float32, device=device)
    off_d = torch.from_numpy(offs).to(device)
    res = (gt, off_d, G, counts)
    _PACK_CACHE[device] = [key, targets, res]
    return res


def target_loss_sums(head: torch.Tensor, gt: torch.Tensor, gt_off: torch.Tensor, G: int, anchors_all_scaled,
                     mask, num_classes: int, ignore_thr: float, iou_thr: float,
                     want_assign: bool = False, max_gt: int = 0, cell_state: Optional[torch.Tensor] = None):
    """b200yolo_target_loss: returns (sums (16,) float64 device, status (1,) int32 device[, assign, terms]).
    ``cell_state``: optional (N, A*H*W) uint8 output consumed by ``target_loss_backward``."""
    _require_cuda(head, "head")
    head = head.contiguous()
    N, ch, H, W = head.shape
    A = len(mask)
    if ch != A * (5 + num_classes):
        raise RuntimeError("head channel dim does not match len(mask)*(5+num_classes)")
    sa = _host_f32(anchors_all_scaled).reshape(-1, 2)
    m = np.ascontiguousarray(np.asarray(mask, dtype=np.int32))
    lib = _lib.load()
    with _on_device(head.device):
        # one allocation: 16 partial sums (f64) + the status word; the per-CTA workspace is cached per
        # (device, stream) and reused (calls on one stream are ordered, so reuse is safe)
        buf = torch.empty((_lib.S_COUNT + 1,), dtype=torch.float64, device=head.device)
        sums, status = buf[:_lib.S_COUNT], buf[_lib.S_COUNT:].view(torch.int32)[:1]
        ws_bytes = int(lib.b200yolo_target_loss_workspace_bytes(N))
        key = (head.device.index, torch.cuda.current_stream(head.device).cuda_stream)
        ent = _WORKSPACES.get(key)
        if ent is None or ent[0].numel() < ws_bytes:
            ent = _WORKSPACES[key] = (torch.empty((ws_bytes,), dtype=torch.uint8, device=head.device), N)
        ws = ent[0]
        assign = torch.empty((max(G, 1), A, 4), dtype=torch.int32, device=head.device) if want_assign else None
        terms = torch.empty((max(G, 1), A, 2), dtype=torch.float32, device=head.device) if want_assign else None
        _lib.check(lib.b200yolo_target_loss(
            head.data_ptr(), N, A, num_classes, H, W, sa.ctypes.data, sa.shape[0], m.ctypes.data, gt.data_ptr(),
            gt_off.data_ptr(), int(G), float(np.float32(ignore_thr)), float(np.float32(iou_thr)), int(max_gt), sums.data_ptr(),
            assign.data_ptr() if want_assign else None, terms.data_ptr() if want_assign else None,
            status.data_ptr(), cell_state.data_ptr() if cell_state is not None else None, ws.data_ptr(), ws.numel(),
            _stream(head)))
    return (sums, status, assign, terms) if want_assign else (sums, status)


def target_loss_lazy(head: torch.Tensor, gt: torch.Tensor, gt_off: torch.Tensor, G: int, sa: np.ndarray, m: np.ndarray,
                     num_classes: int, ignore_thr: float, iou_thr: float, iou_weighting: float, max_gt: int,
                     cell_state: Optional[torch.Tensor], between=None):
    """b200yolo_target_loss + b200yolo_loss_finalize_dev back to back with the host work of ONE call (one stream lookup,
    two allocations, prepared anchor / mask arrays, thresholds already rounded to fp32): the ``lazy_stats`` path of
    ``YOLOLoss.forward``.  ``between(sums, status)`` runs between the two launches.  Returns (sums (16,) f64, status (1,) i32, result (7,) f32), all on the device."""
    _require_cuda(head, "head")
    head = head.contiguous()
    N, ch, H, W = head.shape
    A = m.shape[0]
    if ch != A * (5 + num_classes):
        raise RuntimeError("head channel dim does not match len(mask)*(5+num_classes)")
    dev = head.device
    lib = _lib.load()
    with _on_device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        buf = torch.empty((_lib.S_COUNT + 1,), dtype=torch.float64, device=dev)
        res = torch.empty((7,), dtype=torch.float32, device=dev)
        sums, status = buf[:_lib.S_COUNT], buf[_lib.S_COUNT:].view(torch.int32)[:1]
        key = (dev.index, st)
        ws = _WORKSPACES.get(key)
        if ws is None or ws[1] != N:
            ws_bytes = int(lib.b200yolo_target_loss_workspace_bytes(N))
            ws = _WORKSPACES.get(key)
            t = ws[0] if ws is not None and ws[0].numel() >= ws_bytes else torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            ws = _WORKSPACES[key] = (t, N)
        _lib.check(lib.b200yolo_target_loss(
            head.data_ptr(), N, A, num_classes, H, W, sa.ctypes.data, sa.shape[0], m.ctypes.data, gt.data_ptr(),
            gt_off.data_ptr(), int(G), ignore_thr, iou_thr, int(max_gt), buf.data_ptr(), None, None,
            buf.data_ptr() + 8 * _lib.S_COUNT, cell_state.data_ptr() if cell_state is not None else None,
            ws[0].data_ptr(), ws[0].numel(), st))
        if between is not None:
            between(sums, status)   # data-parallel shards: the all-reduce of the partial sums (and of the status word)
        _lib.check(lib.b200yolo_loss_finalize_dev(buf.data_ptr(), iou_weighting, res.data_ptr(), st))
    return sums, status, res


def target_loss_backward(head: torch.Tensor, gt: torch.Tensor, gt_off: torch.Tensor, G: int, anchors_all_scaled,
                         mask, num_classes: int, iou_thr: float, cell_state: torch.Tensor,
                         sums: torch.Tensor, iou_weighting: float, grad_out: Optional[torch.Tensor] = None,
                         max_gt: int = 0) -> torch.Tensor:
    """b200yolo_target_loss_backward: d loss / d head, same shape as ``head``.  ``sums`` is the (all-reduced)
    device vector of the forward call, ``grad_out`` an optional device scalar."""
    _require_cuda(head, "head")
    head = head.contiguous()
    N, ch, H, W = head.shape
    A = len(mask)
    sa = _host_f32(anchors_all_scaled).reshape(-1, 2)
    m = np.ascontiguousarray(np.asarray(mask, dtype=np.int32))
    with _on_device(head.device):
        grad = torch.empty_like(head)
        go = None
        if grad_out is not None:
            go = grad_out.detach().to(device=head.device, dtype=torch.float32).reshape(1).contiguous()
        _lib.check(_lib.load().b200yolo_target_loss_backward(
            head.data_ptr(), N, A, num_classes, H, W, sa.ctypes.data, sa.shape[0], m.ctypes.data, gt.data_ptr(),
            gt_off.data_ptr(), int(G), float(np.float32(iou_thr)), int(max_gt), cell_state.data_ptr(), sums.data_ptr(),
            float(np.float32(iou_weighting)), go.data_ptr() if go is not None else None, grad.data_ptr(), _stream(head)))
    return grad


class PackedTargets:
    """Ground truth already on the device in the layout of ``b200yolo_target_loss``: ``gt`` (G, 5) fp32 rows
    [cls(1-based), cx, cy, w, h] of all images back to back, ``gt_off`` (N+1,) int32 first row of every image,
    ``max_gt`` an upper bound on the rows of one image.  ``YOLOLoss.forward(input, PackedTargets)`` skips the host-side
    packing of the reference's list of CPU tensors (a data loader can build this once per batch for both heads)."""

    def __init__(self, gt: torch.Tensor, gt_off: torch.Tensor, G: int, max_gt: int):
        self.gt, self.gt_off, self.G, self.max_gt = gt, gt_off, int(G), int(max_gt)

    @classmethod
    def from_list(cls, targets, device) -> "PackedTargets":
        gt, off, G, counts = pack_targets(targets, device)
        return cls(gt, off, G, max(counts + [1]))


def loss_finalize_dev(sums: torch.Tensor, iou_weighting: float) -> torch.Tensor:
    """b200yolo_loss_finalize_dev: (16,) float64 device sums -> (7,) float32 device [loss, recall, avg_iou, obj, no_obj,
    cls, count/N]; no host synchronisation."""
    with _on_device(sums.device):
        out = torch.empty((7,), dtype=torch.float32, device=sums.device)
        _lib.check(_lib.load().b200yolo_loss_finalize_dev(sums.data_ptr(), float(np.float32(iou_weighting)), out.data_ptr(),
                                                          _stream(sums)))
    return out


def loss_finalize(sums_host: np.ndarray, iou_weighting: float) -> np.ndarray:
    s = np.ascontiguousarray(np.asarray(sums_host, dtype=np.float64))
    r = np.zeros(7, np.float64)
    _lib.check(_lib.load().b200yolo_loss_finalize(s.ctypes.data, float(np.float32(iou_weighting)), r.ctypes.data))
    return r
