import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from nmrfit_b200 import synth, utils, equations
for (N, P, scale) in [(4096, 6, 1), (4096, 6, 4), (16384, 24, 16)]:
    data, true = synth.multiplet(N, P, seed=1000)
    lo, up = data.generate_solution_bounds()
    fit = utils.FitUtility(data, lo, up, summary=False)
    fit.params = np.array(true)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); fit.generate_result(scale=scale); ts.append(time.perf_counter() - t0)
    print('generate_result N=%d P=%d scale=%d: %.3f ms (median of last 6)' % (N, P, scale, 1e3 * np.median(ts[2:])))
w = np.linspace(3.2, 3.6, 4096)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); equations.voigt(w, 0.5, 0.0, 0.004, 3.4, 1.0); ts.append(time.perf_counter() - t0)
print('voigt 4096 pts: %.1f us' % (1e6 * np.median(ts[5:])))
