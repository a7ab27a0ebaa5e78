#!/usr/bin/env python
"""Small driver for ncu captures of the TractOracle-Net kernels: one warm-up pass and one measured
pass over --n streamlines (see profiles/README.md for the ncu command line)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=16384)
    ap.add_argument('--precision', default='fp16')
    a = ap.parse_args()
    import torch
    from benchmarks.oracle_bench import make_streamlines
    from tracktolearn_b200 import synthetic
    from tracktolearn_b200.oracles.oracle import OracleSingleton
    dev = torch.device('cuda:0')
    ck = synthetic.oracle_checkpoint(n_head=4, n_layers=4, input_size=384, seed=2222)
    model = OracleSingleton(ck, dev, batch_size=a.n, precision=a.precision)
    data, offsets = make_streamlines(a.n)
    pts = torch.from_numpy(data).to(dev)
    off = torch.from_numpy(offsets).to(dev)
    for _ in range(2):
        s = model.predict_device(pts, off)
    torch.cuda.synchronize()
    print(float(s.mean()))


if __name__ == '__main__':
    main()
