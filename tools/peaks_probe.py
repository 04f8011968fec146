"""Wall time of the batched AutoPeakSelector (utils.select_peaks_batch) on synthetic multiplets:
python tools/peaks_probe.py [out.json].  The reference's selector on the same 16,384-point spectrum took 6 minutes when
tests/golden/make_golden.py produced peaks_c3_16384x6 (scipy argrelmax over 1.6 M samples with order ~ 1e5)."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
from nmrfit_b200 import synth, utils, _cabi

rows = []
for n, P, B in ((4096, 6, 1), (16384, 6, 1), (16384, 6, 64), (16384, 6, 256)):
    ws, us = [], []
    for b in range(B):
        d, _ = synth.multiplet(n, P, seed=100 + b)
        ws.append(d.w); us.append(d.u)
    ws, us = np.stack(ws), np.stack(us)
    utils.select_peaks_batch(ws, us, thresh=0.01, window=0.02)            # warm-up (context, pinned buffers)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        res = utils.select_peaks_batch(ws, us, thresh=0.01, window=0.02)
        best = min(best, time.perf_counter() - t0)
    rows.append(dict(n_points=n, n_peaks=P, n_spectra=B, ms=best * 1e3, ms_per_spectrum=best * 1e3 / B,
                     peaks_found=[len(r) for r in res][:4]))
    print(rows[-1], flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], 'w'), indent=1)
