#!/usr/bin/env python
"""BASELINE config 5a: FitUtility.generate_result at scale=16 (24 peaks, 16,384 -> 262,144 points).

    python tools/bench_curves.py > gpurun_out/curves.json

Three numbers, one JSON line:
  kernel    generate_result_kernel alone on resident device buffers (CUDA events on the launching stream,
            L2 flushed between launches) against its HBM roofline: algorithmic bytes = (2P + 4) doubles
            written + 1 read per point (reference utils.py:262-295 builds exactly those arrays);
  api       wall time of nmrfit_b200.FitUtility.generate_result(scale=16) with host arrays (uploads the grid,
            downloads the (2P + 4) curves);
  cpu       the reference's way - scipy.integrate.quad per point per peak (equations.py:52-80) - through the
            oracle port on a bounded subsample of points, extrapolated to the full grid (stated as such).
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nmrfit_b200 import _cabi, synth, utils      # noqa: E402

P, N, SCALE = 24, 16384, 16


def main():
    import torch
    data, true = synth.multiplet(N, P, seed=5000)
    lo, up = data.generate_solution_bounds()
    fit = utils.FitUtility(data, lo, up, summary=False)
    fit.params = np.array(true, dtype=np.float64)
    ns = int(SCALE * N)
    w_up = np.linspace(data.w.min(), data.w.max(), ns)

    # ---- kernel alone
    dev = torch.device('cuda', _cabi.default_device())
    wd = torch.from_numpy(w_up).to(dev)
    real = torch.empty((P, ns), dtype=torch.float64, device=dev)
    imag = torch.empty_like(real)
    V, I, u, v = (torch.empty(ns, dtype=torch.float64, device=dev) for _ in range(4))
    params = np.ascontiguousarray(fit.params)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def launch():
        _cabi.check(_cabi.lib().nmrfit_generate_result(_cabi.ptr(params), P, _cabi.ptr(wd), ns, _cabi.ptr(real),
                                                       _cabi.ptr(imag), _cabi.ptr(V), _cabi.ptr(I), _cabi.ptr(u),
                                                       _cabi.ptr(v), ctypes.c_void_p(stream)))
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    times = []
    for _ in range(50):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); launch(); b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    # the event bracket also holds the library's cudaMallocAsync + 800-byte parameter upload; report the median
    k_ms = float(np.median(times))
    algo_bytes = ((2 * P + 4) + 1) * 8 * ns
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    hbm_peak = json.load(open(peaks_path)).get('hbm_gbs') if os.path.exists(peaks_path) else 6650.0

    # ---- public API, host arrays
    api = []
    for _ in range(4):
        t0 = time.perf_counter()
        fit.generate_result(scale=SCALE)
        api.append(time.perf_counter() - t0)
    assert fit.V.shape == (ns,) and len(fit.real_contribs) == P and np.all(np.isfinite(fit.I))

    # ---- CPU: the reference's quadrature on a subsample of points of ONE peak (the oracle restates
    # equations.py:9-80 with scipy.integrate.quad); this is the CPU-baseline leg, the only use of oracle/ here
    from oracle import nmrfit_oracle as orc
    sub = w_up[:: ns // 256][:256]
    r, yoff = fit.params[2], fit.params[3]
    width, loc, a = fit.params[4:7]
    t0 = time.perf_counter()
    ref = np.array([orc.kk_quad(x, r, yoff, width, loc, a) for x in sub])
    cpu_s = time.perf_counter() - t0
    per_point = cpu_s / len(sub)
    got = np.array(fit.imag_contribs[0])[:: ns // 256][:256]
    scale_ref = np.abs(ref).max()

    print(json.dumps({
        'workload': 'BASELINE config[4] C5a: generate_result, %d peaks, %d points x scale %d = %d points' % (P, N, SCALE, ns),
        'kernel': {'ms': k_ms, 'ms_min': float(np.min(times)), 'algorithmic_bytes': algo_bytes,
                   'achieved_gbs': algo_bytes / (k_ms * 1e-3) / 1e9, 'hbm_peak_gbs': hbm_peak,
                   'frac_of_hbm_peak': algo_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                   'peak_points_per_s': P * ns / (k_ms * 1e-3)},
        'api': {'ms_median': 1e3 * float(np.median(api[1:])), 'd2h_bytes': (2 * P + 4) * 8 * ns, 'h2d_bytes': 8 * ns},
        'cpu': {'kind': 'port', 'cores': 1, 'sample': '256 grid points of one peak by scipy.integrate.quad',
                'seconds_per_point_per_peak': per_point,
                'extrapolated_seconds_full_grid': per_point * ns * P,
                'max_abs_diff_vs_quad_over_scale': float(np.abs(got - ref).max() / scale_ref)},
    }))


if __name__ == '__main__':
    main()
