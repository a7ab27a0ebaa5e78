#!/usr/bin/env python
"""K sweep of the stand-alone tcgen05 dense layer (ttl_gemm_bf16) at the bench's row count: time vs K
tells a per-tile floor (epilogue / tile turnaround) from the per-k-block cost.  Back-to-back launches on
three rotating operand sets (inputs exceed L2), CUDA events on the launching stream."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tracktolearn_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device('cuda:0')
    sp = _lib.stream_ptr(dev)
    m, n = int(os.environ.get('GEMM_M', '50000')), 1024
    out = {}
    mode = 1
    for k in (128, 256, 512, 640, 768, 1024, 1536, 2048):
        sets = []
        for _ in range(3):
            A = torch.randn((m, k), device=dev).to(torch.bfloat16)
            W = (torch.randn((n, k), device=dev) * 0.03).to(torch.bfloat16)
            C = torch.zeros((m, n), device=dev, dtype=torch.bfloat16)
            sets.append((A, W, C))
        bias = torch.zeros((1024,), device=dev)
        m_dev = torch.tensor([m], dtype=torch.int32, device=dev)

        def run(i):
            A, W, C = sets[i % 3]
            _lib.check(lib.ttl_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(C), m, n, k, n, mode,
                                         _lib.ptr(m_dev), sp), 'gemm')
        for i in range(30):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 120
        e0.record()
        for i in range(reps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        out[k] = {'us': round(us, 2), 'tflops': round(2.0 * m * n * k / us / 1e6, 1)}
        del sets
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
