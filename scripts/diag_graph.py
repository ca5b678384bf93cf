#!/usr/bin/env python
"""Eager vs one-graph vs segmented-graph losses of the train step after 3 warm-up + 2 steps (diagnostic of tests/test_train_step.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from ffwm_b200.train_step import FFWMTrainer
from oracle.train_cpu import synthetic_batch
batches = [synthetic_batch(2, seed=900 + i) for i in range(3)]
def run(graph, segmented=False):
    torch.manual_seed(3)
    tr = FFWMTrainer("cuda:0", graph=graph)
    if graph: tr.enable_cuda_graph(batches[0], warmup=3, segmented=segmented)
    else:
        for _ in range(3): tr.step(batches[0])
    out = []
    for b in batches[1:]:
        tr.step(b); out.append(tr.get_current_losses())
    return out
res = {"eager1": run(False), "eager2": run(False), "graph": run(True), "segmented": run(True, True)}
for k in ("loss_G", "loss_D", "loss_adv", "loss_iden", "loss_l1"):
    print(k, {n: [round(s[k], 5) for s in v] for n, v in res.items()})
