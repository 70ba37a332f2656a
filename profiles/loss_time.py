#!/usr/bin/env python
"""Time the YOLOLoss target-assignment kernel (b200yolo_target_loss) on BASELINE config 4:
VOC 352 heads, N images, 100 synthetic GT boxes per image.  Usage: loss_time.py [N] [G]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mobilenet_yolo_pytorch_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
G = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda", 0)
res = bench.time_loss(dev, N, G, steps=100)
print(res)
