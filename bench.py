#!/usr/bin/env python
"""bench.py -- decode + NMS images/s of the MobileNet-YOLO detection hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cfg2|cfg2_sparse|cfg3|cfg5|cfg5_832|cfg4_loss] [--no-extra]

A "step" is one pass of the hot path over one batch of synthetic head tensors
(BASELINE.json configs[1]: MobileNetV2-YOLO 352x352 VOC heads, batch 256 per GPU,
torch.randn heads, val_conf 0.3).  With N > 1 (launched under torchrun, one rank
per GPU) the batch shards by image: every rank owns 256 images (weak scaling) and
a step is kernel + all-gather of the detections + fence ("all-gather
included for > 1 GPU"): the gather is fused into the kernel's output phase (peer
stores over NVLink, dist.PeerGather).  Rank 0 prints ONE JSON line.

The K steps of the timed region are issued by ONE C call (b200yolo_decode_nms_batches /
b200yolo_decode_nms_gather_steps): no Python between the launches.

Keys beyond the base contract: `roofline` (dominant kernel, algorithmic bytes /
CUDA-event time vs the measured HBM peak), `cpu_baseline` (the CPU oracle port on
this box's host cores), `e2e` (same metric through the host-buffer C-ABI call,
H2D + D2H inside the timed region), `extra` (the reference's own Python on this box's
CPU and GPU, NMS alone against torchvision's CUDA kernel, sparse heads, the other
BASELINE configurations with their own roofline blocks, config-1 latency).

`--impl reference` times the UNMODIFIED reference Python (oracle/_ref snapshot, see
oracle/snapshot_reference.py) on the host cores in a subprocess with
CUDA_VISIBLE_DEVICES="" (cpu_baseline.kind "reference"); where the snapshot is
missing it falls back to the C port (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
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
    # N: batch per GPU when the workload is weak-scaled; GB: global batch when BASELINE.json fixes it (strong scaling)
    "cfg2": dict(N=256, GB=None, C=20, grids=[(11, 11), (22, 22)], anchors=VOC_ANCHORS, img=[352, 352], conf=0.3, shift=0.0,
                 desc="MobileNetV2-YOLO 352x352 VOC 20-class heads, batch 256 per GPU, randn heads, val_conf 0.3"),
    "cfg2_sparse": dict(N=256, GB=None, C=20, grids=[(11, 11), (22, 22)], anchors=VOC_ANCHORS, img=[352, 352], conf=0.3,
                        shift=-2.6, desc="cfg2 with objectness logits shifted by -2.6 (~4% of cells pass, as trained heads do)"),
    "cfg3": dict(N=128, GB=1024, C=10, grids=[(12, 20), (24, 40)], anchors=BDD_ANCHORS, img=[640, 384], conf=0.3, shift=0.0,
                 desc="MobileNetV3-YOLO BDD100k 10-class 640x384 heads, global batch 1024 sharded over the GPUs"),
    "cfg5": dict(N=512, GB=4096, C=20, grids=[(13, 13), (26, 26)], anchors=VOC_ANCHORS, img=[416, 416], conf=0.001, shift=0.0,
                 desc="dense-candidate NMS stress: 416x416 heads, val_conf 0.001 (all 2535 cells pass), global batch 4096 sharded over the GPUs"),
    "cfg5_832": dict(N=128, GB=None, C=20, grids=[(26, 26), (52, 52)], anchors=VOC_ANCHORS, img=[832, 832], conf=0.001, shift=0.0,
                     desc="the ~10k-boxes-per-image reading of the stress configuration: 832x832 heads (10140 cells per "
                          "image, all pass val_conf 0.001), batch 128 per GPU, large-image path (b200yolo_decode_nms_large)"),
}
A = 3
L2_NOTE = "inputs larger than L2: rotating input sets of >= 400 MB in total (126 MB L2), so every step reads its heads from HBM"


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


def local_batch(wl, world, rank=0):
    if wl["GB"] is None:
        return wl["N"]
    base, rem = divmod(wl["GB"], world)
    return base + (1 if rank < rem else 0)


def workload_config(name, wl, world):
    """Identical in both arms (the driver compares the dicts)."""
    n = local_batch(wl, world)
    return {"workload": wl["desc"], "name": name, "batch_per_gpu": n, "global_batch": wl["GB"] if wl["GB"] else world * n,
            "cells_per_image": cells_per_image(wl), "l2": L2_NOTE}


def load_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy), of measured"
    except Exception:  # noqa: BLE001
        pass
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md), of fallback"


# ----------------------------------------------------------------------------- clocks
_SAMPLER_SRC = r"""
import json, select, sys, time
idx = int(sys.argv[1])
samples, reasons, max_sm = [], set(), None
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(idx)
    max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
    names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
             "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
             "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
             "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
except Exception:
    nv = None
print("ready", flush=True)
while True:
    r, _, _ = select.select([sys.stdin], [], [], 0.003)
    if r:
        break
    if nv is not None:
        try:
            samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            t = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for k, bit in names.items():
                if t & bit:
                    reasons.add(k)
        except Exception:
            pass
samples.sort()
print(json.dumps({"sm_mhz": (samples[len(samples) // 2] if samples else None), "sm_max_mhz": max_sm,
                  "reasons": sorted(reasons), "samples": len(samples)}), flush=True)
"""


class ClockSampler:
    """NVML poll every ~3 ms (the recipe's nvidia-smi clocks line, faster) in a SEPARATE PROCESS, so that it never
    competes with the launching thread for the interpreter lock."""

    def __init__(self, index):
        self.proc = None
        try:
            phys = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(phys)], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, text=True)
            self.proc.stdout.readline()  # "ready"
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        try:
            self.proc.stdin.write("\n")
            self.proc.stdin.flush()
            line = self.proc.stdout.readline()
            self.proc.wait(timeout=5)
            return json.loads(line)
        except Exception:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


# ----------------------------------------------------------------------------- CPU arms (oracle port / reference python)
def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm overrides it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_port_rate(wl, N, budget_s=10.0, max_reps=30):
    """images/s of the CPU oracle port on this box's cores, bounded sample (median over the passes)."""
    import oracle
    oracle.set_threads(host_threads())
    h0, h1 = make_heads(wl, N, seed=0)
    h0, h1 = h0.numpy(), h1.numpy()
    tables = anchor_tables(wl)
    oracle.decode_nms_padded(h0, h1, tables, wl["C"], wl["conf"])  # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    times = []
    while len(times) < max_reps and (len(times) < 5 or time.perf_counter() - t0 < budget_s):
        t = time.perf_counter()
        oracle.decode_nms_padded(h0, h1, tables, wl["C"], wl["conf"])
        times.append(time.perf_counter() - t)
    med = statistics.median(times)
    return N / med, len(times), med


def cpu_loss_port_rate(N, G):
    """images/s of the CPU oracle port for YOLOLoss.forward(input, targets), both VOC heads."""
    import oracle
    oracle.set_threads(host_threads())
    wl = WORKLOADS["cfg2"]
    h0, h1 = make_heads(wl, N, seed=100)
    targets = make_targets(N, G, wl["C"], 1)
    args = [(h0.numpy(), MASK[0], VOC_IGNORE[0]), (h1.numpy(), MASK[1], VOC_IGNORE[1])]
    times = []
    for _ in range(5):
        t = time.perf_counter()
        for h, m, ign in args:
            oracle.target_loss(h, targets, VOC_ANCHORS, m, wl["C"], [352, 352], ign, VOC_IOU_THRESH, VOC_IOU_WEIGHTING)
        times.append(time.perf_counter() - t)
    return N / statistics.median(times)


def reference_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "models", "yolo_loss.py")) or \
        os.path.exists("/root/reference/models/yolo_loss.py")


def run_ref_bench(device, what, workload, n, reps, budget=40.0, timeout=600):
    """oracle/ref_bench.py in a subprocess (the reference fixes its device at import, quirk Q4). -> dict or {'error'}"""
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    if device == "cpu":
        env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_bench.py"), "--device", device, "--what", what, "--workload", workload,
           "--n", str(n), "--reps", str(reps), "--budget", str(budget)]
    if device == "cpu":
        cmd += ["--threads", str(host_threads())]
    try:
        res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
        lines = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
        if res.returncode != 0 or not lines:
            return {"error": (res.stderr or res.stdout)[-400:]}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def run_reference(args, name, wl):
    """--impl reference: the reference's own CPU path, rank 0 only.  Each step = one pass over a bounded sample of
    the workload (n_ref images of the same shapes / seeds); value = images/s from the median step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    steps, warm = max(1, args.steps), max(args.warmup, 3)
    cfg = workload_config(name, wl, world)
    if reference_available():
        n_ref = 32
        r = run_ref_bench("cpu", "decode_nms", name if name in WORKLOADS else "cfg2", n_ref, steps, budget=150.0, timeout=900)
        if "error" not in r:
            val = r["images_per_s"]
            line = {
                "impl": "reference", "metric": "decode+NMS images/sec", "value": val, "unit": "images/s", "n_gpus": world,
                "steps": r["reps"], "warmup": warm, "ms_per_step": 1e3 * r["median_s"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": val, "unit": "images/s", "cores": r["threads"], "kind": "reference",
                                 "sample": f"{r['reps']} passes over {n_ref} images of the workload (the batch is {cfg['batch_per_gpu']}): "
                                           "YOLOLoss.forward x2 + utils.box.nms of the unmodified reference (oracle/_ref), "
                                           f"torch {r['torch']} / torchvision {r['torchvision']} on the host cores, median pass"},
                "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "spread": {"min_s": r["min_s"], "max_s": r["max_s"], "median_s": r["median_s"]},
            }
            print(json.dumps(line), flush=True)
            return
    # no snapshot (or it failed): the C restatement of the algorithm
    import oracle
    N = local_batch(wl, world)
    val, reps, med = cpu_port_rate(wl, N, budget_s=20.0, max_reps=max(5, min(steps, 40)))
    line = {
        "impl": "reference", "metric": "decode+NMS images/sec", "value": val, "unit": "images/s", "n_gpus": world,
        "steps": reps, "warmup": warm, "ms_per_step": 1e3 * med, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": oracle.max_threads(), "kind": "port",
                         "sample": f"{reps} passes over the full {N}-image batch (oracle/yolo_oracle.c, OpenMP), median pass; "
                                   "the reference snapshot oracle/_ref is not present on this box"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
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


def time_loop(fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)  # ms


def time_loss(dev, N, G, steps=100, seed=1):
    """YOLOLoss.forward(input, targets) partial sums for BOTH VOC-352 heads (2 launches of
    target_loss_kernel + 2 tiny reductions per step), heads resident in HBM."""
    from mobilenet_yolo_pytorch_b200 import _lib as _lib_mod
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

    # the heads and targets are resident and no kernel in this stream produces them: consecutive calls may overlap
    # (include/b200yolo.h, b200yolo_set_inputs_ready); the module API (YOLOLoss.forward) never sets this
    _lib_mod.load().b200yolo_set_inputs_ready(1)
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
    _lib_mod.load().b200yolo_set_inputs_ready(0)
    res["fwd_bwd"] = {"ms_per_step": ms_fb, "images_per_s_per_gpu": N / (ms_fb * 1e-3),
                      "algorithmic_gbs": (algo + in_bytes) / (ms_fb * 1e-3) / 1e9}
    return res


# ----------------------------------------------------------------------------- B200 arm
class Dist:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max(self, v):
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v):
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


class DecodeNmsRun:
    """K steps of decode + NMS on this rank's shard, issued by one C call.  world > 1: kernel + fused all-gather
    (peer stores over NVLink) + fence per step."""

    def __init__(self, D, wl, n_local, steps, gather=True, seed0=0, graph=True, multicast=False):
        from mobilenet_yolo_pytorch_b200 import _lib, ops
        from mobilenet_yolo_pytorch_b200 import dist as b2dist
        self.D, self.wl, self.N, self.steps = D, wl, n_local, steps
        self.K = cells_per_image(wl)
        self.tables = anchor_tables(wl)
        self.in_bytes = n_local * bytes_in_per_image(wl)
        self.R = max(4, int(np.ceil(400e6 / max(self.in_bytes, 1))))
        self.sets = []
        for r in range(self.R):
            h0, h1 = make_heads(wl, n_local, seed=seed0 + 1000 * D.rank + r)
            self.sets.append((h0.to(D.dev), h1.to(D.dev)))
        self.large = self.K > _lib.load().b200yolo_max_cells(D.local)
        self.gather = gather and D.world > 1 and not self.large
        # two result buffers in rotation (what a consumer that works on step k while step k + 1 runs needs anyway):
        # consecutive launches then write disjoint outputs and are chained (no wait before the stores, include/b200yolo.h)
        # Every step of the list keeps its own result buffer (an evaluation loop that collects the detections of every
        # batch) while that stays below 2 GB, else two buffers in rotation: launches that write disjoint outputs do not
        # wait for their predecessor before they store (include/b200yolo.h, b200yolo_decode_nms_batches)
        out_bytes = n_local * self.K * 28
        n_out = steps if steps * out_bytes <= (2 << 30) else 2
        if gather and D.world > 1:
            n_out = 1
        self.outs = [(torch.empty((n_local, self.K, 7), dtype=torch.float32, device=D.dev),
                      torch.empty((n_local,), dtype=torch.int32, device=D.dev)) for _ in range(max(n_out, 1))]
        self.out, self.cnt = self.outs[0]
        self.ops = ops
        self.pg = None
        if self.large:
            self.plan = None
        elif self.gather:
            self.pg = b2dist.PeerGather(n_local, self.K, multicast=multicast)
            self.plan = self.pg.run_steps([self.sets[i % self.R] for i in range(steps)], self.tables, wl["C"], wl["conf"])
            torch.cuda.synchronize()
        else:
            self.plan = ops.BatchPlan([(self.sets[i % self.R][0], self.sets[i % self.R][1]) + self.outs[i % len(self.outs)] for i in range(steps)],
                                      self.tables, wl["C"], wl["conf"])
            if graph:
                self.plan.capture()   # b200yolo_plan_create: the K launches as one CUDA graph (K kernel nodes, programmatic edges)

    def run(self, count=None):
        count = self.steps if count is None else count
        if self.large:
            for i in range(count):
                h0, h1 = self.sets[i % self.R]
                self.ops.decode_nms_padded(h0, h1, self.tables, self.wl["C"], self.wl["conf"], out=self.out, out_count=self.cnt)
        elif self.gather:
            # (a shorter run re-uses the head of the plan)
            if count == self.steps:
                self.pg.run_steps(None, None, self.wl["C"], self.wl["conf"], plan=self.plan)
            else:
                sub = (self.plan[0], count) + self.plan[2:]
                self.pg.run_steps(None, None, self.wl["C"], self.wl["conf"], plan=sub)
        else:
            self.plan.run(0, count)

    def kept_per_launch(self):
        """kept rows of this rank's shard, averaged over the rotating sets"""
        kept = []
        for r in range(self.R):
            h0, h1 = self.sets[r]
            self.ops.decode_nms_padded(h0, h1, self.tables, self.wl["C"], self.wl["conf"], out=self.out, out_count=self.cnt)
            kept.append(int(self.cnt.sum().item()))
        return float(np.mean(kept))

    def time(self, warmup):
        D = self.D
        if getattr(self.plan, "_graph", None) is not None:
            self.run()                 # the plan replays as a whole: its K >= W steps are the warm-up
            if self.steps < max(warmup, 3):
                self.run()
        else:
            self.run(min(max(warmup, 3), self.steps))
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        from mobilenet_yolo_pytorch_b200 import _lib
        c0 = _lib.launch_count()
        e0.record()
        self.run()
        e1.record()
        e1.synchronize()
        self.timed_launches = _lib.launch_count() - c0   # kernels of this library inside the timed region
        ms = e0.elapsed_time(e1)
        D.barrier()
        if self.pg is not None:
            self.pg.check()
        return D.max(ms) / self.steps

    def close(self):
        if self.pg is not None:
            self.pg.close()
            self.pg = None
        self.sets = None


def measure_workload(D, name, steps, warmup, peak):
    """One BASELINE configuration at this world size -> dict with its own roofline block (rank 0 returns it)."""
    wl = WORKLOADS[name]
    n_local = local_batch(wl, D.world, D.rank)
    run = DecodeNmsRun(D, wl, n_local, steps, seed0=7000)
    kept = run.kept_per_launch()
    ms = run.time(warmup)
    total_images = D.sum(n_local)
    kept_all = D.sum(kept)
    in_all = total_images * bytes_in_per_image(wl)
    algo_rank = run.in_bytes + 28.0 * kept + 4 * n_local       # this rank's kernel
    gbs = algo_rank / (ms * 1e-3) / 1e9
    res = {"workload": wl["desc"], "n_gpus": D.world, "global_batch": int(total_images), "batch_per_gpu": n_local,
           "scaling": "strong" if wl["GB"] else "weak", "ms_per_step": ms, "images_per_s": total_images / (ms * 1e-3),
           "kept_rows_per_image": kept_all / total_images,
           "collective": ("fused all-gather of the kept rows (peer stores over NVLink) + fence, inside the timed step" if run.gather
                          else "none (1 GPU)" if D.world == 1 else "none: large-image path, shards stay on their rank"),
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                        "kernel": "decode_nms_large_kernel" if run.large else "decode_nms_kernel<MODE_FUSED>",
                        "algorithmic_bytes_per_launch": algo_rank,
                        "note": "per rank: this rank's algorithmic bytes / the step time (max over ranks)"}}
    if run.gather:
        res["nvlink_bytes_in_per_rank_per_step"] = 28.0 * (kept_all - kept)
        res["nvlink_in_gbs"] = 28.0 * (kept_all - kept) / (ms * 1e-3) / 1e9
    run.close()
    del in_all
    return res


def run_b200(args, name, wl):
    from mobilenet_yolo_pytorch_b200 import _lib, ops

    D = Dist()
    world, rank, local, dev = D.world, D.rank, D.local, D.dev
    peak, peak_src = load_peak()
    N, C, conf = local_batch(wl, world, rank), wl["C"], wl["conf"]
    K = cells_per_image(wl)
    tables = anchor_tables(wl)
    lib = _lib.load()

    # cpu_baseline: rank 0, and only while it has the host to itself (N = 1); at N > 1 the reference arm's line is the
    # CPU figure (a CPU number taken while 8 ranks share the host means nothing)
    cpu_base = None
    if rank == 0 and world == 1:
        import oracle
        cpu_val, cpu_reps, cpu_med = cpu_port_rate(wl, N)
        cpu_base = {"value": cpu_val, "unit": "images/s", "cores": oracle.max_threads(), "kind": "port",
                    "sample": f"{cpu_reps} passes over the same {N}-image batch, median {cpu_med * 1e3:.1f} ms "
                              "(oracle/yolo_oracle.c, OpenMP, all host threads); the unmodified reference Python is "
                              "timed in extra.reference_python and by --impl reference"}
    D.barrier()

    # ---- headline: K steps from one C call
    run = DecodeNmsRun(D, wl, N, args.steps)
    kept_per_launch = run.kept_per_launch()
    ms_per_step = run.time(args.warmup)
    launches_timed = run.timed_launches
    # clocks: the same launches for ~0.4 s under the NVML sampler (a 0.5 ms timed region is shorter than one NVML poll)
    clocks = None
    if rank == 0:
        sampler = ClockSampler(local)
    D.barrier()
    if world == 1:
        t_end = time.perf_counter() + 0.4
        while time.perf_counter() < t_end:
            run.run()
            torch.cuda.synchronize()
    else:
        # (the ranks must issue the same number of gather steps: a fixed number of passes, ~0.2-0.5 s)
        for _ in range(max(1, 4000 // max(args.steps, 1))):
            run.run()
        torch.cuda.synchronize()
    D.barrier()
    if rank == 0:
        clocks = sampler.stop()
    total_images = D.sum(N)
    value = total_images / (ms_per_step * 1e-3)
    kept_all = D.sum(kept_per_launch)
    gathered = run.gather

    # ---- e2e: host buffers through the C-ABI host call (H2D + kernel + D2H inside)
    # (the rank's thread moves to the CPUs next to its GPU first, so the pinned buffers land on that NUMA node)
    cpus_before = os.sched_getaffinity(0)
    numa = ops.bind_host_to_gpu(local) if os.environ.get("B200YOLO_NUMA", "1") == "1" else {"bound": False, "why": "B200YOLO_NUMA=0"}
    hh0, hh1 = make_heads(wl, N, seed=7 + rank, pin=True)
    ho = torch.empty((N, K, 7), dtype=torch.float32).pin_memory()
    hc = torch.empty((N,), dtype=torch.int32).pin_memory()
    for _ in range(3):
        ops.decode_nms_host(hh0, hh1, tables, C, conf, device=local, out=ho, out_count=hc)
    e2e_steps = max(3, min(args.steps, 30))
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ops.decode_nms_host(hh0, hh1, tables, C, conf, device=local, out=ho, out_count=hc)
    torch.cuda.synchronize()
    e2e_val = total_images * e2e_steps / D.max(time.perf_counter() - t0)
    d2h_bytes = int(lib.b200yolo_host_last_d2h_bytes()) if hasattr(lib, "b200yolo_host_last_d2h_bytes") else int(N * K * 28 + 4 * N)
    del hh0, hh1, ho, hc
    os.sched_setaffinity(0, cpus_before)   # (the CPU legs below use every host thread again)

    extra = {}
    if not args.no_extra:
        # plain stream order (no programmatic dependent launch): comparable with ncu's gpu__time_duration
        if not run.large:
            # (the graph froze the launch attributes: these two legs issue the K launches directly)
            held, held_graph = (run.plan, run.plan._graph) if getattr(run.plan, "_graph", None) is not None else (None, None)
            if held is not None:
                held._graph = None
                direct_ms = run.time(3)
                extra["direct_launches"] = {"ms_per_step": direct_ms, "images_per_s": total_images / (direct_ms * 1e-3),
                                            "note": "the same K overlapping launches issued one by one from one C call "
                                                    "(b200yolo_decode_nms_batches) instead of one graph launch"}
            lib.b200yolo_debug_set_flags(2)
            try:
                ser_ms = run.time(3)
            finally:
                lib.b200yolo_debug_set_flags(0)
                if held is not None:
                    held._graph = held_graph
            extra["serialized_launches"] = {"ms_per_step": ser_ms, "images_per_s": total_images / (ser_ms * 1e-3),
                                            "note": "the same steps in plain stream order (every launch waits for the previous one to drain)"}
        if gathered:
            # what the step costs without the exchange, and with NCCL's all-gather of the fixed-stride block instead
            from mobilenet_yolo_pytorch_b200 import dist as b2dist
            solo = DecodeNmsRun(D, wl, N, args.steps, gather=False)
            solo_ms = solo.time(args.warmup)
            extra["without_gather"] = {"ms_per_step": solo_ms, "images_per_s": total_images / (solo_ms * 1e-3),
                                       "note": "kernels only, every shard stays on its rank (round 1's headline protocol)"}

            def gstep(i):
                h0, h1 = solo.sets[i % solo.R]
                ops.decode_nms_padded(h0, h1, tables, C, conf, out=solo.out, out_count=solo.cnt)
                b2dist.all_gather_detections(solo.out, solo.cnt)
            for i in range(3):
                gstep(i)
            D.barrier()
            gn = max(10, args.steps // 4)
            gms = D.max(time_loop(gstep, gn) / gn)
            extra["with_nccl_allgather"] = {"ms_per_step": gms, "images_per_s": total_images / (gms * 1e-3),
                                            "bytes_gathered_per_rank": int(world * N * (K + 1) * 28),
                                            "note": "kernel, then ncclAllGather of the fixed-stride (N, K+1, 7) block"}
            solo.close()
            # the same step with the gather buffers in NVSwitch multicast memory (one store per row, replicated by the switch)
            mc_ok = [None] * world
            torch.distributed.all_gather_object(mc_ok, bool(lib.b200yolo_mc_supported(local)))
            if all(mc_ok):
                mrun = DecodeNmsRun(D, wl, N, args.steps, multicast=True)
                mc_ms = mrun.time(args.warmup)
                mrun.close()
                extra["with_multicast"] = {"ms_per_step": mc_ms, "images_per_s": total_images / (mc_ms * 1e-3),
                                           "note": "every row stored once through the switch instead of once per peer: egress drops "
                                                   "(R-1)-fold, but every rank still RECEIVES all rows -- plus its own copy back "
                                                   "through the switch -- and ingress is what bounds the exchange"}
            else:
                extra["with_multicast"] = {"unavailable": "cuMulticast* not supported on every rank"}
        run.close()
        # sparse-head variant of the same workload (trained heads pass ~4% of cells), same protocol
        if wl["shift"] == 0.0 and name == "cfg2":
            wls = dict(WORKLOADS["cfg2_sparse"])
            srun = DecodeNmsRun(D, wls, N, args.steps, seed0=5000)
            skept = srun.kept_per_launch()
            sms = srun.time(args.warmup)
            sbytes = srun.in_bytes + 28.0 * skept + 4 * N
            sg = sbytes / (sms * 1e-3) / 1e9
            straffic = None
            try:
                straffic = json.load(open(os.path.join(ROOT, "profiles", "decode_nms_traffic.json"))).get("cfg2_sparse", {}).get("dram_bytes_per_launch")
            except Exception:  # noqa: BLE001
                pass
            extra["sparse_heads"] = {"ms_per_step": sms, "images_per_s": total_images / (sms * 1e-3),
                                     "kept_rows_per_image": D.sum(skept) / total_images,
                                     "collective_in_step": srun.gather,
                                     "roofline": {"bound": "hbm", "achieved": sg, "peak": peak, "unit": "GB/s", "frac": sg / peak,
                                                  "traffic": straffic, "algorithmic_bytes_per_launch": sbytes}}
            srun.close()
            # what bit-equality with the reference on CUDA costs (b200yolo_set_exact_decode: IEEE sigmoid / expf, true division)
            if world == 1:
                ops.set_exact_decode(True)
                try:
                    xrun = DecodeNmsRun(D, wl, N, args.steps)
                    xms = xrun.time(args.warmup)
                    xrun.close()
                    extra["exact_decode"] = {"ms_per_step": xms, "images_per_s": total_images / (xms * 1e-3),
                                             "note": "decoded rows and detections bit-identical to the reference on device='cuda' "
                                                     "(tests/test_reference_integration.py); the headline uses the default decode (<= 1e-5)"}
                finally:
                    ops.set_exact_decode(False)
        # the other BASELINE configurations at this world size, each with its own roofline block
        if name == "cfg2":
            osteps = max(10, min(args.steps, 40))
            for other in ("cfg3", "cfg5", "cfg5_832"):
                try:
                    r = measure_workload(D, other, osteps if other != "cfg5_832" else max(5, osteps // 4), 3, peak)
                    if rank == 0:
                        extra[other] = r
                except Exception as e:  # noqa: BLE001
                    if rank == 0:
                        extra[other] = {"error": repr(e)}
                torch.cuda.empty_cache()
            try:
                r = measure_loss(D, max(10, min(args.steps, 50)), 3, peak)
                if rank == 0:
                    extra["cfg4_loss"] = r
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    extra["cfg4_loss"] = {"error": repr(e)}
        # the reference's own code on this box (rank 0, one GPU: the host and GPU 0 are otherwise idle now)
        if rank == 0 and world == 1 and name in ("cfg2", "cfg2_sparse") and reference_available():
            extra["reference_python"] = reference_legs(dev, wl, name, tables)
    else:
        run.close()

    if rank == 0:
        algo_bytes = run.in_bytes + 28.0 * kept_per_launch + 4 * N
        achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "decode_nms_traffic.json")))
            traffic = prof.get(name, {}).get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
        cfg = workload_config(name, wl, world)
        line = {
            "metric": "decode+NMS images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong" if wl["GB"] else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg,
            "run": {"kept_rows_per_image": kept_all / total_images,
                    "launch": (("the K steps are K kernel nodes of ONE CUDA graph (b200yolo_plan_create / _launch)" if not gathered else
                                "the K steps are issued by ONE C call (b200yolo_decode_nms_gather_steps)")
                               + "; consecutive launches overlap through programmatic dependent launch (a launch starts on free SM "
                               "slots while the previous one finishes and waits for it before writing); extra.serialized_launches is "
                               "plain stream order")
                    if not run.large else "one kernel per step on one stream, plain stream order",
                    "parallelism": (f"dp{world} by image; every step = kernel + fused all-gather of the kept rows (peer stores over NVLink) + fence"
                                    if gathered else f"dp{world} by image, no collective (1 GPU)" if world == 1
                                    else f"dp{world} by image, no data-path collective (large-image path)"),
                    "limiter": (f"NVLink ingress: every rank receives {28.0 * (kept_all - kept_per_launch) / 1e6:.1f} MB of kept rows per step "
                                f"({28.0 * (kept_all - kept_per_launch) / (ms_per_step * 1e-3) / 1e9:.0f} GB/s achieved of 900 GB/s nominal per direction, "
                                "770 GB/s measured peer copy); the kernel itself is issue-bound" if gathered
                                else "issue-bound after the heads are in (see DESIGN.md 4.1)")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "kernel": "decode_nms_kernel<MODE_FUSED>" if not run.large
                         else "decode_nms_large_kernel (more cells per image than one CTA stages in shared memory)",
                         "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src},
            "cpu_baseline": cpu_base if cpu_base is not None else {
                "value": None, "unit": "images/s", "cores": 0, "kind": "port",
                "sample": "not taken at N > 1 (the ranks share the host cores); see the N = 1 line and the reference arm"},
            "e2e": {"value": e2e_val, "unit": "images/s", "h2d_bytes_per_step": int(run.in_bytes),
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "api": "b200yolo_decode_nms_host (pinned host heads -> host detections, chunks on 3 streams: H2D, kernel and D2H overlap)",
                    "timer": "host wall clock around the synchronous calls, max over ranks",
                    "host_numa_binding": numa},
            "gpu_launches": int(launches_timed) * world,
            "clocks": clocks,
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    D.close()


def reference_legs(dev, wl, name, tables):
    """extra.reference_python: the UNMODIFIED reference (oracle/_ref) on this box -- its CPU path, its CUDA path
    (torchvision's sm_100 nms kernel per (image, class), utils/box.py:16-29) and NMS alone on identical candidates."""
    from mobilenet_yolo_pytorch_b200 import ops
    out = {}
    n_ref = 64
    out["cpu"] = run_ref_bench("cpu", "decode_nms", name, n_ref, 5)
    out["cuda"] = run_ref_bench("cuda", "decode_nms", name, n_ref, 5)
    out["cuda_nms_only"] = run_ref_bench("cuda", "nms_only", name, n_ref, 5)
    # b200yolo_nms alone on the candidates the reference's decode produces for the same heads (same seeds)
    try:
        C, conf = wl["C"], wl["conf"]
        h0, h1 = [h.to(dev) for h in make_heads(wl, n_ref, seed=0)]
        r0, c0 = ops.decode_head_padded(h0, tables[0], C, conf)
        r1, c1 = ops.decode_head_padded(h1, tables[1], C, conf)
        for _ in range(3):
            o, oc = ops.nms_padded(r0, c0, r1, c1, C)
        torch.cuda.synchronize()
        reps = 50
        ms = time_loop(lambda i: ops.nms_padded(r0, c0, r1, c1, C), reps) / reps
        out["b200yolo_nms_only"] = {"n": n_ref, "ms": ms, "images_per_s": n_ref / (ms * 1e-3),
                                    "kept_rows_per_image": float(oc.sum().item()) / n_ref,
                                    "note": "b200yolo_nms (one launch) on the decoded candidates of the same heads"}
        tv = out["cuda_nms_only"]
        if "error" not in tv:
            out["nms_only_speedup_vs_reference_cuda"] = (n_ref / (ms * 1e-3)) / tv["images_per_s"]
            if "torchvision_batched_nms" in tv:
                out["nms_only_speedup_vs_torchvision_batched_nms"] = (n_ref / (ms * 1e-3)) / tv["torchvision_batched_nms"]["images_per_s"]
    except Exception as e:  # noqa: BLE001
        out["b200yolo_nms_only"] = {"error": repr(e)}
    # BASELINE config 1: the real model, batch 1 (inference.py:120-124 prints this latency)
    try:
        out["config1_batch1"] = config1_latency()
    except Exception as e:  # noqa: BLE001
        out["config1_batch1"] = {"error": repr(e)}
    return out


def config1_latency():
    """MobileNetV2-YOLO 352x352, batch 1, random-init: model(x) latency of the reference on the CPU and on the GPU
    (subprocesses, quirk Q4) and of the same model with this package patched in (in this process)."""
    res = {}
    env_cpu = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    env_cpu.pop("OMP_NUM_THREADS", None)
    for label, env in (("reference_cpu", env_cpu), ("reference_cuda", dict(os.environ))):
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_config1.py"), "--reps", "10"], env=env,
                               capture_output=True, text=True, timeout=300)
            lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
            res[label] = json.loads(lines[-1]) if lines else {"error": r.stderr[-300:]}
        except Exception as e:  # noqa: BLE001
            res[label] = {"error": repr(e)}
    import mobilenet_yolo_pytorch_b200 as b200
    from oracle import ref_config1

    def patch(ns, m):
        b200.patch_reference(models_yolo_loss=ns.yolo_loss, utils_box=ns.box, mbv2_yolo=m, fuse_inference=True)
    r = ref_config1.run("cuda", reps=20, patch=patch)
    b200.YOLOLoss.lazy_eval = False
    res["b200_patched_cuda"] = {"latency_ms": r["latency_ms"], "detections": int(r["dets"].shape[0]),
                                "note": "the reference's models/mbv2_yolo.py with YOLOLoss / nms replaced by this package "
                                        "(patch_reference(..., fuse_inference=True)): backbone + heads in cuDNN, post-processing in one launch"}
    return res


def measure_loss(D, steps, warmup, peak):
    """BASELINE config 4: YOLOLoss.forward(input, targets) x2 (VOC 352 heads), global batch 512 with 100 synthetic GT
    boxes per image, sharded by image (strong scaling); N > 1: the all-reduce of the 16 partial sums per head is inside
    the step.  Module API: host target lists in, loss tensor out."""
    import mobilenet_yolo_pytorch_b200 as b200
    GB, G = 512, 100
    wl = WORKLOADS["cfg2"]
    C = wl["C"]
    group = D.dist.group.WORLD if D.world > 1 else None
    lo, hi = b200.dist.shard_bounds(GB, D.world, D.rank)
    N = hi - lo
    targets = [torch.from_numpy(t) for t in make_targets(GB, G, C, 1)[lo:hi]]
    losses = [b200.YOLOLoss(VOC_ANCHORS, MASK[k], C, [352, 352], VOC_IGNORE[k], VOC_IOU_THRESH, iou_weighting=VOC_IOU_WEIGHTING,
                            process_group=group) for k in range(2)]
    in_bytes = N * bytes_in_per_image(wl)
    R = max(3, int(np.ceil(300e6 / max(in_bytes, 1))))
    sets = [tuple(h.to(D.dev) for h in make_heads(wl, N, seed=100 + 17 * D.rank + r)) for r in range(R)]

    def step(i):
        h0, h1 = sets[i % R]
        tl = list(targets)  # a new list object every step, like a data loader's: packed once, shared by both heads
        return losses[0](h0, tl)[0] + losses[1](h1, tl)[0]

    for i in range(max(warmup, 3)):
        step(i)
    D.barrier()
    ms = D.max(time_loop(step, steps) / steps)
    D.barrier()
    # the same step without host round trips: ground truth packed on the device once per batch (ops.PackedTargets, what a
    # prefetching data loader hands over) and lazy_stats (loss and statistics stay device scalars; b200yolo_loss_finalize_dev)
    from mobilenet_yolo_pytorch_b200 import ops as _ops
    packed = _ops.PackedTargets.from_list(targets, D.dev)
    for l in losses:
        l.lazy_stats = True

    def step_lazy(i):
        h0, h1 = sets[i % R]
        return losses[0](h0, packed)[0] + losses[1](h1, packed)[0]

    for i in range(max(warmup, 3)):
        step_lazy(i)
    D.barrier()
    ms_lazy = D.max(time_loop(step_lazy, steps) / steps)
    for l in losses:
        l.check()
        l.lazy_stats = False
    D.barrier()
    kres = time_loss(D.dev, N, G, steps=max(20, steps)) if D.rank == 0 else None
    D.barrier()
    if D.rank != 0:
        return None
    return {"workload": "YOLOLoss.forward(input, targets) x2 (VOC 352 heads), global batch 512, 100 synthetic GT boxes per image",
            "n_gpus": D.world, "global_batch": GB, "batch_per_gpu": N, "scaling": "strong",
            "module_api": {"ms_per_step": ms, "images_per_s": GB / (ms * 1e-3),
                           "note": "host target lists in, loss tensor + python stats out"
                                   + ("; all-reduce(SUM) of 16 doubles per head inside the step" if D.world > 1 else "")},
            "module_api_device_targets_lazy_stats": {
                "ms_per_step": ms_lazy, "images_per_s": GB / (ms_lazy * 1e-3),
                "note": "YOLOLoss.forward(input, ops.PackedTargets) x2 with lazy_stats: no host synchronisation, loss tensor and "
                        "statistics stay on the device"},
            "kernel_only": kres,
            "roofline": {"bound": "hbm", "achieved": kres["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                         "frac": kres["algorithmic_gbs"] / peak, "traffic": None,
                         "kernel": "target_loss_kernel (both heads, device-resident packed targets)",
                         "algorithmic_bytes_per_launch": kres["algorithmic_bytes_per_step"],
                         "note": "bound by the pred-vs-GT box pairs, not by HBM"}}


def run_loss(args):
    """--workload cfg4_loss as the headline (same measurement as extra.cfg4_loss of the default run)."""
    rank = int(os.environ.get("RANK", "0"))
    GB, G = 512, 100
    cfg = {"workload": "YOLOLoss.forward(input, targets) x2 (VOC 352 heads), global batch 512, 100 synthetic GT boxes per image",
           "name": "cfg4_loss", "batch_per_gpu": GB // max(args.gpus, 1), "global_batch": GB, "cells_per_image": 1815, "l2": L2_NOTE}
    if args.impl == "reference":
        if rank != 0:
            return
        if reference_available():
            r = run_ref_bench("cpu", "loss", "cfg2", 8, max(3, min(args.steps, 10)), budget=120.0, timeout=900)
        else:
            r = {"error": "no snapshot"}
        if "error" in r:
            val, kind, sample, cores = cpu_loss_port_rate(32, G), "port", "median of 5 passes over 32 images, both heads (oracle/yolo_oracle.c, OpenMP)", host_threads()
            ms = 1e3 * 32 / val
        else:
            val, kind, cores = r["images_per_s"], "reference", r["threads"]
            sample = f"{r['reps']} passes over 8 images, both heads: the unmodified reference's YOLOLoss.forward(input, targets), median pass"
            ms = 1e3 * r["median_s"]
        print(json.dumps({"impl": "reference", "metric": "YOLOLoss target assignment images/sec", "value": val,
                          "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    from mobilenet_yolo_pytorch_b200 import _lib
    D = Dist()
    peak, _ = load_peak()
    sampler = ClockSampler(D.local) if D.rank == 0 else None
    l0 = _lib.launch_count()
    r = measure_loss(D, args.steps, args.warmup, peak)
    launches = _lib.launch_count() - l0
    if D.rank == 0:
        clocks = sampler.stop()
        value = r["module_api"]["images_per_s"]
        N = r["batch_per_gpu"]
        print(json.dumps({
            "metric": "YOLOLoss target assignment images/sec", "value": value, "unit": "images/s", "n_gpus": D.world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["module_api"]["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": r["roofline"],
            "cpu_baseline": {"value": cpu_loss_port_rate(32, G) if D.world == 1 else None, "unit": "images/s", "cores": host_threads(), "kind": "port",
                             "sample": "median of 5 passes over 32 images, both heads (oracle/yolo_oracle.c, OpenMP)"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": int(2 * (N * G * 20 + 4 * (N + 1))),
                    "d2h_bytes_per_step": 2 * 17 * 8, "api": "YOLOLoss.forward(input, targets): host target lists in, python loss tuple out"},
            "gpu_launches": int(launches) * D.world, "clocks": clocks,
            "extra": {"kernel_only": r["kernel_only"]},
        }), flush=True)
    D.close()


def main():
    args = parse_args()
    if args.workload == "cfg4_loss":
        run_loss(args)
        return
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    else:
        run_b200(args, args.workload, wl)


if __name__ == "__main__":
    main()
