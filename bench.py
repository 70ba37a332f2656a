#!/usr/bin/env python
"""bench.py -- decode + NMS images/s of the MobileNet-YOLO detection hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg2|cfg2_sparse|cfg3|cfg5|cfg4_loss]   (cfg4_loss: the YOLOLoss target-assignment path)

A "step" is one pass of the hot path over one batch of synthetic head tensors
(BASELINE.json configs[1]: MobileNetV2-YOLO 352x352 VOC heads, batch 256 per GPU,
torch.randn heads, val_conf 0.3).  With N > 1 (launched under torchrun, one rank
per GPU) the batch shards by image: every rank owns 256 images (weak scaling);
there is no collective in the data path.  Rank 0 prints ONE JSON line.

Keys beyond the base contract: `roofline` (dominant kernel, algorithmic bytes /
CUDA-event time vs the measured HBM peak), `cpu_baseline` (the CPU oracle port on
this box's host cores), `e2e` (same metric through the host-buffer C-ABI call,
H2D + D2H inside the timed region), `extra` (sparse-head variant, NMS-only and
all-gather timings).

`--impl reference` times the reference's algorithm on the CPU: the reference is
pure Python that cannot travel to the GPU box (/root/reference is absent there)
so the arm runs the oracle port (oracle/yolo_oracle.c, OpenMP over all host
threads) -- see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

VOC_ANCHORS = [[143, 265], [153, 121], [280, 279], [20, 37], [49, 94], [73, 201]]   # models/voc/config.yaml:20-26
BDD_ANCHORS = [[34, 47], [66, 93], [122, 182], [6, 11], [11, 43], [16, 22]]         # models/bdd100k/config.yaml:17-23
MASK = [[0, 1, 2], [3, 4, 5]]

WORKLOADS = {
    # name: (batch per GPU, classes, grids (H,W) of head0/head1, anchors, img_size [W,H], val_conf, conf logit shift)
    "cfg2": dict(N=256, C=20, grids=[(11, 11), (22, 22)], anchors=VOC_ANCHORS, img=[352, 352], conf=0.3, shift=0.0,
                 desc="MobileNetV2-YOLO 352x352 VOC 20-class heads, batch 256 per GPU, randn heads, val_conf 0.3"),
    "cfg2_sparse": dict(N=256, C=20, grids=[(11, 11), (22, 22)], anchors=VOC_ANCHORS, img=[352, 352], conf=0.3,
                        shift=-2.6, desc="cfg2 with objectness logits shifted by -2.6 (~4% of cells pass, as trained heads do)"),
    "cfg3": dict(N=128, C=10, grids=[(12, 20), (24, 40)], anchors=BDD_ANCHORS, img=[640, 384], conf=0.3, shift=0.0,
                 desc="MobileNetV3-YOLO BDD100k 10-class 640x384 heads, batch 1024/8 = 128 per GPU"),
    "cfg5": dict(N=512, C=20, grids=[(13, 13), (26, 26)], anchors=VOC_ANCHORS, img=[416, 416], conf=0.001, shift=0.0,
                 desc="dense-candidate NMS stress: 416x416 heads, val_conf 0.001, batch 4096/8 = 512 per GPU"),
    "cfg5_832": dict(N=128, C=20, grids=[(26, 26), (52, 52)], anchors=VOC_ANCHORS, img=[832, 832], conf=0.001, shift=0.0,
                     desc="the ~10k-boxes-per-image reading of the stress configuration: 832x832 heads (10140 cells per "
                          "image, all pass val_conf 0.001), batch 128 per GPU, large-image path (b200yolo_decode_nms_large)"),
}
A = 3


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS) + ["cfg4_loss"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements")
    return ap.parse_args()


def make_heads(wl, N, seed, pin=False):
    g = torch.Generator().manual_seed(seed)
    C = wl["C"]
    hs = []
    for (H, W) in wl["grids"]:
        h = torch.randn(N, A * (5 + C), H, W, generator=g)
        if wl["shift"]:
            h.view(N, A, 5 + C, H, W)[:, :, 4] += wl["shift"]
        h = h.contiguous()
        hs.append(h.pin_memory() if pin else h)
    return hs


def anchor_tables(wl):
    sa = np.array([[aw / wl["img"][0], ah / wl["img"][1]] for aw, ah in wl["anchors"]], np.float64).astype(np.float32)
    return np.stack([sa[MASK[0]], sa[MASK[1]]])


def bytes_in_per_image(wl):
    return sum(4 * A * (5 + wl["C"]) * H * W for (H, W) in wl["grids"])


def cells_per_image(wl):
    return sum(A * H * W for (H, W) in wl["grids"])


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """NVML poll (the recipe's nvidia-smi clocks line, at a few ms instead of 200 ms
    so that sub-second timed regions are still seen)."""

    def __init__(self, index):
        self.samples = []
        self.reasons = set()
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None
            self.max_sm = None

    def _poll(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.003)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU (oracle) arm
def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm overrides it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_decode_nms_rate(wl, budget_s=12.0, max_reps=50):
    """images/s of the CPU oracle port on this box's cores, bounded sample."""
    import oracle
    oracle.set_threads(host_threads())
    N = wl["N"]
    h0, h1 = make_heads(wl, N, seed=0)
    h0, h1 = h0.numpy(), h1.numpy()
    tables = anchor_tables(wl)
    oracle.decode_nms_padded(h0, h1, tables, wl["C"], wl["conf"])  # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    reps = 0
    times = []
    while reps < max_reps and time.perf_counter() - t0 < budget_s:
        t = time.perf_counter()
        oracle.decode_nms_padded(h0, h1, tables, wl["C"], wl["conf"])
        times.append(time.perf_counter() - t)
        reps += 1
    med = statistics.median(times)
    return N / med, reps, med


def cpu_loss_rate(N, G):
    """images/s of the CPU oracle port for YOLOLoss.forward(input, targets), both VOC heads."""
    import oracle
    oracle.set_threads(host_threads())
    wl = WORKLOADS["cfg2"]
    h0, h1 = make_heads(wl, N, seed=100)
    targets = make_targets(N, G, wl["C"], 1)
    args = [(h0.numpy(), MASK[0], VOC_IGNORE[0]), (h1.numpy(), MASK[1], VOC_IGNORE[1])]
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        for h, m, ign in args:
            oracle.target_loss(h, targets, VOC_ANCHORS, m, wl["C"], [352, 352], ign, VOC_IOU_THRESH, VOC_IOU_WEIGHTING)
        best = min(best, time.perf_counter() - t)
    return N / best


def run_reference(args, wl):
    """--impl reference: the CPU path (oracle port), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.set_threads(host_threads())
    N = wl["N"]
    h0, h1 = make_heads(wl, N, seed=0)
    h0, h1 = h0.numpy(), h1.numpy()
    tables = anchor_tables(wl)
    for _ in range(max(1, min(args.warmup, 3))):
        oracle.decode_nms_padded(h0, h1, tables, wl["C"], wl["conf"])
    steps = max(1, min(args.steps, 40))  # each step = one pass over the same batch of N images; bounded
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.decode_nms_padded(h0, h1, tables, wl["C"], wl["conf"])
    dt = time.perf_counter() - t0
    val = N * steps / dt
    line = {
        "impl": "reference", "metric": "decode+NMS images/sec", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "batch": N},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": oracle.max_threads(), "kind": "port",
                         "sample": f"{steps} passes over the full {N}-image batch (oracle/yolo_oracle.c, OpenMP)"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python and /root/reference does not exist on the GPU box; this arm is the "
                "C restatement of its algorithm (pinned by tests/golden), all host threads",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- loss target assignment (config 4)
VOC_IGNORE = [0.6076333316652263, 0.5623606200028424]   # models/voc/config.yaml:27-29
VOC_IOU_THRESH = 0.5497280113447018                      # :31
VOC_IOU_WEIGHTING = 0.021830872589525777                 # :12


def make_targets(N, G, C, seed):
    """cls~U{1..C}, w,h~U(0.02,0.47), cx~U(w/2,1-w/2), cy likewise."""
    r = np.random.RandomState(seed)
    out = []
    for _ in range(N):
        wh = r.uniform(0.02, 0.47, (G, 2))
        c = wh / 2 + r.rand(G, 2) * (1 - wh)
        out.append(np.concatenate((r.randint(1, C + 1, (G, 1)), c, wh), 1).astype(np.float32))
    return out


def time_loss(dev, N, G, steps=100, seed=1):
    """YOLOLoss.forward(input, targets) partial sums for BOTH VOC-352 heads (2 launches of
    target_loss_kernel + 2 tiny reductions per step), heads resident in HBM."""
    from mobilenet_yolo_pytorch_b200 import ops
    wl = WORKLOADS["cfg2"]
    C = wl["C"]
    sa = np.array([[aw / 352, ah / 352] for aw, ah in VOC_ANCHORS], np.float64).astype(np.float32)
    targets = make_targets(N, G, C, seed)
    gt, gt_off, Gt, _ = ops.pack_targets([torch.from_numpy(t) for t in targets], dev)
    in_bytes = N * bytes_in_per_image(wl)
    R = max(3, int(np.ceil(300e6 / in_bytes)))
    sets = [tuple(h.to(dev) for h in make_heads(wl, N, seed=100 + r)) for r in range(R)]

    def step(i):
        h0, h1 = sets[i % R]
        ops.target_loss_sums(h0, gt, gt_off, Gt, sa, MASK[0], C, VOC_IGNORE[0], VOC_IOU_THRESH, max_gt=G)
        ops.target_loss_sums(h1, gt, gt_off, Gt, sa, MASK[1], C, VOC_IGNORE[1], VOC_IOU_THRESH, max_gt=G)

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    ms = time_loop(step, steps) / steps
    algo = in_bytes + 2 * 20 * N * G
    res = {"images_per_s_per_gpu": N / (ms * 1e-3), "ms_per_step": ms, "batch": N, "gt_per_image": G,
           "algorithmic_bytes_per_step": algo, "algorithmic_gbs": algo / (ms * 1e-3) / 1e9,
           "iou_pairs_per_step": N * G * cells_per_image(wl)}
    # forward + backward (grad_input of both heads; every element written once: + in_bytes of stores)
    states = [torch.empty((N, A * H * W), dtype=torch.uint8, device=dev) for (H, W) in wl["grids"]]

    def step_fb(i):
        for k, h in enumerate(sets[i % R]):
            sums, _ = ops.target_loss_sums(h, gt, gt_off, Gt, sa, MASK[k], C, VOC_IGNORE[k], VOC_IOU_THRESH, max_gt=G,
                                           cell_state=states[k])
            ops.target_loss_backward(h, gt, gt_off, Gt, sa, MASK[k], C, VOC_IOU_THRESH, states[k], sums, VOC_IOU_WEIGHTING,
                                     max_gt=G)

    for i in range(5):
        step_fb(i)
    torch.cuda.synchronize()
    ms_fb = time_loop(step_fb, steps) / steps
    res["fwd_bwd"] = {"ms_per_step": ms_fb, "images_per_s_per_gpu": N / (ms_fb * 1e-3),
                      "algorithmic_gbs": (algo + in_bytes) / (ms_fb * 1e-3) / 1e9}
    return res


# ----------------------------------------------------------------------------- B200 arm
def time_loop(fn, steps, stream=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)  # ms


def run_b200(args, wl):
    import torch.distributed as dist
    from mobilenet_yolo_pytorch_b200 import _lib, ops
    from mobilenet_yolo_pytorch_b200 import dist as b2dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, C, conf = wl["N"], wl["C"], wl["conf"]
    tables = anchor_tables(wl)
    K = cells_per_image(wl)

    # R rotating input sets so that successive steps never find their heads in the 126 MB L2
    in_bytes = N * bytes_in_per_image(wl)
    R = max(4, int(np.ceil(400e6 / in_bytes)))
    sets = []
    for r in range(R):
        h0, h1 = make_heads(wl, N, seed=1000 * rank + r)
        sets.append((h0.to(dev), h1.to(dev)))
    out = torch.empty((N, K, 7), dtype=torch.float32, device=dev)
    cnt = torch.empty((N,), dtype=torch.int32, device=dev)

    def step(i):
        h0, h1 = sets[i % R]
        ops.decode_nms_padded(h0, h1, tables, C, conf, out=out, out_count=cnt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    # kept rows per launch (for the algorithmic bytes), averaged over the rotating sets
    kept = []
    for r in range(R):
        step(r)
        kept.append(int(cnt.sum().item()))
    kept_per_launch = float(np.mean(kept))

    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    l0 = _lib.launch_count()
    ms = time_loop(step, args.steps)
    launches = _lib.launch_count() - l0
    barrier()
    # keep the same launch loop running ~0.4 s so the clock sampler sees the GPU under this load
    if sampler:
        t_end = time.perf_counter() + 0.4
        i = 0
        while time.perf_counter() < t_end:
            step(i)
            i += 1
            if i % 256 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * N / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C-ABI host call (H2D + kernel + D2H inside)
    hh0, hh1 = make_heads(wl, N, seed=7 + rank, pin=True)
    ho = torch.empty((N, K, 7), dtype=torch.float32).pin_memory()
    hc = torch.empty((N,), dtype=torch.int32).pin_memory()
    for _ in range(3):
        ops.decode_nms_host(hh0, hh1, tables, C, conf, device=local, out=ho, out_count=hc)
    e2e_steps = max(3, min(args.steps, 30))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ops.decode_nms_host(hh0, hh1, tables, C, conf, device=local, out=ho, out_count=hc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * N * e2e_steps / float(t.item())

    extra = {}
    if not args.no_extra:
        # sparse-head variant of the same workload (trained heads pass ~4% of cells)
        if wl["shift"] == 0.0:
            wls = dict(wl, shift=-2.6)
            ssets = []
            for r in range(R):
                h0, h1 = make_heads(wls, N, seed=5000 + 1000 * rank + r)
                ssets.append((h0.to(dev), h1.to(dev)))

            def sstep(i):
                h0, h1 = ssets[i % R]
                ops.decode_nms_padded(h0, h1, tables, C, conf, out=out, out_count=cnt)

            for i in range(5):
                sstep(i)
            skept = []
            for r in range(R):
                sstep(r)
                skept.append(int(cnt.sum().item()))
            barrier()
            sms = time_loop(sstep, args.steps) / args.steps
            sbytes = in_bytes + 28 * float(np.mean(skept)) + 4 * N
            extra["sparse_heads"] = {"images_per_s_per_gpu": N / (sms * 1e-3), "ms_per_step": sms,
                                     "kept_rows_per_image": float(np.mean(skept)) / N,
                                     "algorithmic_gbs": sbytes / (sms * 1e-3) / 1e9}
            # the same launches replayed from a CUDA graph (one graph = one pass over the R input sets): at ~12 us per
            # step the Python launch loop is the limiter, the graph shows what the GPU does
            try:
                gstream = torch.cuda.Stream(device=dev)
                with torch.cuda.stream(gstream):
                    for i in range(R):
                        sstep(i)
                gstream.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=gstream):
                    for i in range(R):
                        sstep(i)
                reps = max(2, args.steps // R)
                for _ in range(2):
                    graph.replay()
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(reps):
                    graph.replay()
                g1.record()
                g1.synchronize()
                gms = g0.elapsed_time(g1) / (reps * R)
                extra["sparse_heads"]["cuda_graph"] = {"ms_per_step": gms, "images_per_s_per_gpu": N / (gms * 1e-3),
                                                       "algorithmic_gbs": sbytes / (gms * 1e-3) / 1e9}
            except Exception as e:  # noqa: BLE001
                extra["sparse_heads"]["cuda_graph"] = {"error": repr(e)}
            del ssets
        # the same loop without programmatic dependent launch (debug flag 2): every launch waits for the previous
        # one to drain -- the serialized per-launch time, comparable with ncu's gpu__time_duration
        _lib.load().b200yolo_debug_set_flags(2)
        try:
            for i in range(5):
                step(i)
            barrier()
            ser_ms = time_loop(step, args.steps) / args.steps
        finally:
            _lib.load().b200yolo_debug_set_flags(0)
        extra["serialized_launches"] = {"ms_per_step": ser_ms, "images_per_s_per_gpu": N / (ser_ms * 1e-3),
                                        "note": "plain stream order (no programmatic dependent launch)"}
        # the same steps issued round-robin on two streams (separate output buffers): consecutive launches
        # overlap, so one launch's decode (memory phase) runs under the other's NMS (issue-bound phase)
        s2 = [torch.cuda.Stream(device=dev) for _ in range(2)]
        o2 = [torch.empty_like(out) for _ in range(2)]
        c2 = [torch.empty_like(cnt) for _ in range(2)]

        def pstep(i):
            h0, h1 = sets[i % R]
            with torch.cuda.stream(s2[i & 1]):
                ops.decode_nms_padded(h0, h1, tables, C, conf, out=o2[i & 1], out_count=c2[i & 1])

        for i in range(6):
            pstep(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            pstep(i)
        torch.cuda.synchronize()
        pms = (time.perf_counter() - t0) * 1e3 / args.steps
        extra["two_streams"] = {"images_per_s_per_gpu": N / (pms * 1e-3), "ms_per_step": pms,
                                "note": "throughput of overlapped launches (host wall clock); the headline value is single-stream"}
        # YOLOLoss target assignment + loss sums, BASELINE config 4 (both heads, 100 GT boxes per image)
        if rank == 0:
            try:
                lN = 512
                lres = time_loss(dev, lN, 100, steps=max(20, args.steps // 2))
                lres["hbm_frac"] = lres["algorithmic_gbs"] / float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) \
                    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else lres["algorithmic_gbs"] / 6650.0
                lres["bound"] = "issue (92.9 M box pairs x ~14 instructions), not HBM"
                lres["cpu_port_images_per_s"] = cpu_loss_rate(32, 100)
                extra["loss_cfg4"] = lres
            except Exception as e:  # noqa: BLE001
                extra["loss_cfg4"] = {"error": repr(e)}
        # all-gather of the fixed-stride detections (the only collective the path may need)
        if world > 1:
            def gstep(i):
                step(i)
                b2dist.all_gather_detections(out, cnt)
            for i in range(3):
                gstep(i)
            barrier()
            gms = time_loop(gstep, max(10, args.steps // 4)) / max(10, args.steps // 4)
            tg = torch.tensor([gms], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            extra["with_allgather"] = {"images_per_s": world * N / (float(tg.item()) * 1e-3),
                                       "ms_per_step": float(tg.item()),
                                       "bytes_gathered_per_rank": int(world * N * (K + 1) * 28)}

            # compact form: kept rows only (one host read of the per-rank totals to size the buffer)
            def cstep(i):
                step(i)
                b2dist.all_gather_detections_compact(out, cnt)
            for i in range(3):
                cstep(i)
            barrier()
            cms = time_loop(cstep, max(10, args.steps // 4)) / max(10, args.steps // 4)
            tc = torch.tensor([cms], dtype=torch.float64, device=dev)
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            extra["with_allgather_compact"] = {"images_per_s": world * N / (float(tc.item()) * 1e-3),
                                               "ms_per_step": float(tc.item()),
                                               "bytes_gathered_per_rank": int(world * kept_per_launch * 28)}

            # the same gather fused into the kernel: the output phase stores the kept rows into every rank's buffer over
            # NVLink peer mappings (b200yolo_decode_nms_gather); the fence is the cross-rank barrier a consumer needs
            try:
                pg = b2dist.PeerGather(N, K)

                def pstep2(i):
                    h0, h1 = sets[i % R]
                    pg.decode_nms(h0, h1, tables, C, conf)
                    pg.fence(timeout_s=1.0)
                for i in range(3):
                    pstep2(i)
                barrier()
                pg.check()   # a fence that timed out in the warm-up ends this measurement here (reported as an error)
                pgms = time_loop(pstep2, max(10, args.steps // 4)) / max(10, args.steps // 4)
                tp = torch.tensor([pgms], dtype=torch.float64, device=dev)
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
                same = bool(torch.equal(pg.counts[rank * N:(rank + 1) * N], cnt))
                pg.check()

                def pstep3(i):
                    h0, h1 = sets[i % R]
                    pg.decode_nms(h0, h1, tables, C, conf)
                    pg.fence(collective=True)
                for i in range(3):
                    pstep3(i)
                barrier()
                pcms = time_loop(pstep3, max(10, args.steps // 4)) / max(10, args.steps // 4)
                tpc = torch.tensor([pcms], dtype=torch.float64, device=dev)
                dist.all_reduce(tpc, op=dist.ReduceOp.MAX)
                extra["with_peer_gather"] = {"images_per_s": world * N / (float(tp.item()) * 1e-3),
                                             "ms_per_step": float(tp.item()),
                                             "ms_per_step_with_nccl_fence": float(tpc.item()),
                                             "bytes_stored_per_rank": int(world * kept_per_launch * 28),
                                             "counts_equal_plain_launch": same,
                                             "note": "decode + NMS + all-gather in ONE kernel (peer stores over NVLink) + the fence (arrival "
                                                     "flags in peer memory, two tiny kernels), per step"}
                pg.close()
            except Exception as e:  # noqa: BLE001
                extra["with_peer_gather"] = {"error": repr(e)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        algo_bytes = in_bytes + 28.0 * kept_per_launch + 4 * N
        achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "decode_nms_traffic.json")))
            traffic = prof.get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        cpu_val, cpu_reps, cpu_med = cpu_decode_nms_rate(wl)
        import oracle
        line = {
            "metric": "decode+NMS images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "batch_per_gpu": N, "global_batch": world * N,
                       "cells_per_image": K, "kept_rows_per_image": kept_per_launch / N,
                       "l2": f"{R} rotating input sets ({R * in_bytes / 1e6:.0f} MB) > 126 MB L2, so every step reads its heads from HBM",
                       "launch": ("one kernel per step on one stream; consecutive launches overlap through programmatic dependent "
                                  "launch (a launch starts on free SM slots while the previous one finishes, and waits for it before "
                                  "writing); extra.serialized_launches is the same loop in plain stream order")
                       if K <= _lib.load().b200yolo_max_cells(local) else "one kernel per step on one stream, plain stream order",
                       "parallelism": f"dp{world} by image, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": "decode_nms_kernel<MODE_FUSED>" if K <= _lib.load().b200yolo_max_cells(local)
                         else "decode_nms_large_kernel (more cells per image than one CTA stages in shared memory)",
                         "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src + ", of measured"},
            "cpu_baseline": {"value": cpu_val, "unit": "images/s", "cores": oracle.max_threads(), "kind": "port",
                             "sample": f"{cpu_reps} passes over the same {N}-image batch, median {cpu_med * 1e3:.1f} ms "
                                       "(oracle/yolo_oracle.c, OpenMP, all host threads)"},
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": int(in_bytes),
                    "d2h_bytes_per_step": int(N * K * 28 + 4 * N), "steps": e2e_steps,
                    "api": "b200yolo_decode_nms_host (pinned host heads -> host detections, 4 chunks on 3 streams: H2D, kernel and D2H overlap)",
                    "timer": "host wall clock around the synchronous calls, max over ranks"},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_loss(args):
    """--workload cfg4_loss: BASELINE config 4 -- YOLOLoss target assignment + loss (both VOC-352 heads), global batch
    512 with 100 synthetic GT boxes per image, sharded by image over the ranks (strong scaling); a step is
    YOLOLoss.forward(input, targets) x2 through the module API, i.e. it includes packing the host targets, their H2D
    copy, the all-reduce of the 16 partial sums (N > 1) and the D2H read of the sums."""
    import torch.distributed as dist
    import mobilenet_yolo_pytorch_b200 as b200
    from mobilenet_yolo_pytorch_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    GB, G = 512, 100
    wl = WORKLOADS["cfg2"]
    C = wl["C"]
    if args.impl == "reference":
        if rank == 0:
            n = 32
            val = cpu_loss_rate(n, G)
            print(json.dumps({"impl": "reference", "metric": "YOLOLoss target assignment images/sec", "value": val,
                              "unit": "images/s", "n_gpus": args.gpus, "steps": 3, "warmup": 0, "ms_per_step": 1e3 * n / val,
                              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic", "config": {"workload": "YOLOLoss target assignment, VOC 352 heads, 100 GT boxes per image", "batch": n},
                              "cpu_baseline": {"value": val, "unit": "images/s", "cores": host_threads(), "kind": "port",
                                               "sample": f"best of 3 passes over {n} images, both heads (oracle/yolo_oracle.c, OpenMP)"},
                              "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                              "gpu_launches": 0}), flush=True)
        return
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    lo, hi = b200.dist.shard_bounds(GB, world, rank)
    N = hi - lo
    targets = [torch.from_numpy(t) for t in make_targets(GB, G, C, 1)[lo:hi]]
    losses = [b200.YOLOLoss(VOC_ANCHORS, MASK[k], C, [352, 352], VOC_IGNORE[k], VOC_IOU_THRESH, iou_weighting=VOC_IOU_WEIGHTING,
                            process_group=group) for k in range(2)]
    in_bytes = N * bytes_in_per_image(wl)
    R = max(3, int(np.ceil(300e6 / max(in_bytes, 1))))
    sets = [tuple(h.to(dev) for h in make_heads(wl, N, seed=100 + 17 * rank + r)) for r in range(R)]

    def step(i):
        h0, h1 = sets[i % R]
        tl = list(targets)  # a new list object every step, like a data loader's: packed once, shared by both heads
        return losses[0](h0, tl)[0] + losses[1](h1, tl)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    l0 = _lib.launch_count()
    ms = time_loop(step, args.steps)
    launches = _lib.launch_count() - l0
    barrier()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    # kernel-only rate (device-resident packed targets, no host sync), for the roofline
    kres = time_loss(dev, N, G, steps=max(20, args.steps // 2)) if rank == 0 else None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        value = GB / (ms_per_step * 1e-3)
        print(json.dumps({
            "metric": "YOLOLoss target assignment images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "YOLOLoss.forward(input, targets) x2 (VOC 352 heads), global batch 512, 100 synthetic GT boxes per image",
                       "name": "cfg4_loss", "batch_per_gpu": N, "global_batch": GB,
                       "l2": f"{R} rotating head sets ({R * in_bytes / 1e6:.0f} MB) > 126 MB L2",
                       "parallelism": f"dp{world} by image; one all-reduce(SUM) of 16 doubles per head" if world > 1 else "dp1"},
            "roofline": {"bound": "hbm", "achieved": kres["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                         "frac": kres["algorithmic_gbs"] / peak, "traffic": None, "kernel": "target_loss_kernel (both heads, kernel-only loop)",
                         "algorithmic_bytes_per_launch": kres["algorithmic_bytes_per_step"],
                         "note": "the path is bound by the 92.9 M pred-vs-GT box pairs (~14 instructions each), not by HBM"},
            "cpu_baseline": {"value": cpu_loss_rate(32, G), "unit": "images/s", "cores": host_threads(), "kind": "port",
                             "sample": "best of 3 passes over 32 images, both heads (oracle/yolo_oracle.c, OpenMP)"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": int(2 * (N * G * 20 + 4 * (N + 1))),
                    "d2h_bytes_per_step": 2 * 17 * 8, "api": "YOLOLoss.forward(input, targets): host target lists in, python loss tuple out"},
            "gpu_launches": int(launches) * world, "clocks": clocks,
            "extra": {"kernel_only": kres},
        }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.workload == "cfg4_loss":
        run_loss(args)
        return
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
