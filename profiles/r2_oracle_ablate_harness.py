#!/usr/bin/env python
"""Scratch: time the TractOracle-Net forward under TTL_ORACLE_ABLATE masks (results are wrong by design)."""
import os
import sys
import json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from benchmarks.oracle_bench import make_streamlines
from tracktolearn_b200 import synthetic
from tracktolearn_b200.oracles.oracle import OracleSingleton

N = 131072
dev = torch.device('cuda:0')
ck = synthetic.oracle_checkpoint(n_head=4, n_layers=4, input_size=384, seed=2222)
model = OracleSingleton(ck, dev, batch_size=N, precision='fp16')
data, offsets = make_streamlines(N)
pts = torch.from_numpy(data).to(dev)
off = torch.from_numpy(offsets).to(dev)
out = {}
for mask in [int(a) for a in sys.argv[1:]] or [0]:
    os.environ['TTL_ORACLE_ABLATE'] = str(mask)
    for _ in range(2):
        model.predict_device(pts, off)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        model.predict_device(pts, off)
    e1.record()
    torch.cuda.synchronize()
    out[mask] = e0.elapsed_time(e1) / 3
    print(mask, out[mask], flush=True)
print(json.dumps(out))
