"""Where the host time of YOLOLoss.forward(input, PackedTargets) with lazy_stats goes (config 4: N=512, 100 GT boxes per
image, both heads): wall time per step and a cProfile of 200 steps.  Run on a GPU box: python profiles/loss_module_profile.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mobilenet_yolo_pytorch_b200 as b200  # noqa: E402
from mobilenet_yolo_pytorch_b200 import ops  # noqa: E402

ANCH = [[143, 265], [153, 121], [280, 279], [20, 37], [49, 94], [73, 201]]
MASK = [[0, 1, 2], [3, 4, 5]]
dev = torch.device("cuda", 0)
N, C, G = 512, 20, 100
g = torch.Generator().manual_seed(1)
sets = [(torch.randn(N, 75, 11, 11, generator=g).to(dev), torch.randn(N, 75, 22, 22, generator=g).to(dev)) for _ in range(3)]
r = np.random.RandomState(2)
targets = []
for b in range(N):
    wh = r.rand(G, 2) * 0.4 + 0.03
    c = wh / 2 + r.rand(G, 2) * (1 - wh)
    targets.append(torch.from_numpy(np.concatenate((r.randint(1, C + 1, (G, 1)), c, wh), 1).astype(np.float32)))
packed = ops.PackedTargets.from_list(targets, dev)
losses = [b200.YOLOLoss(ANCH, MASK[i], C, [352, 352], 0.6, 0.55) for i in range(2)]
for l in losses:
    l.lazy_stats = True


def step(i):
    h0, h1 = sets[i % 3]
    return losses[0](h0, packed)[0] + losses[1](h1, packed)[0]


def timed(K):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(K):
        step(i)
    t_issue = time.perf_counter() - t
    torch.cuda.synchronize()
    return t_issue / K * 1e3, (time.perf_counter() - t) / K * 1e3


for i in range(10):
    step(i)
for K in (20, 200):
    a, b = timed(K)
    print(f"{K} steps: host issue {a:.4f} ms/step, with final synchronize {b:.4f} ms/step")
with torch.no_grad():
    a, b = timed(200)
    print(f"no_grad, 200 steps: host issue {a:.4f} ms/step, with final synchronize {b:.4f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(200):
    step(i)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
