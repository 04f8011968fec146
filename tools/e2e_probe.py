import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from nmrfit_b200 import equations, synth
data, weights, lo, up, _ = bench.make_inputs('metric')
S, D = 65536, 22
xs = torch.empty((S, D), dtype=torch.float64).pin_memory().numpy()
xs[:] = synth.particles(lo, up, S, seed=7)
for _ in range(5): equations.objective_batch(xs, data.w, data.u, data.v, weights)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): f = equations.objective_batch(xs, data.w, data.u, data.v, weights)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 30
print('e2e ms per call %.4f  evals/s %.4g' % (dt * 1e3, S / dt))
