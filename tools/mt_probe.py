"""Time of nmrfit_ctx_mt19937 (numpy's legacy stream continued on the device) per call: python tools/mt_probe.py"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from nmrfit_b200 import _cabi
import torch
S, D = 100, 22
with _cabi.Context(1, 4096, 6) as ctx:
    for n in (1, 16, 32, 64, 128):
        np.random.seed(1)
        ctx.legacy_uniform_pairs(n, S * D)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            ctx.legacy_uniform_pairs(n, S * D)
            best = min(best, time.perf_counter() - t0)
        words = 4 * n * S * D
        print('pairs %4d  words %8d  %.3f ms  %.1f ns/step of 227 words' % (n, words, best * 1e3, best * 1e9 / (words / 227.0)))
