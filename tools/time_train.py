#!/usr/bin/env python
"""Training-step timing (bench.py next_rows.train_step) in isolation."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

d = bench.next_rows(torch.device("cuda:0"), 256, 120, {"hbm_gbs": 6536.7})
print(json.dumps(d["train_step"], indent=1))
