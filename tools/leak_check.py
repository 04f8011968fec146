import sys, io, contextlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import nmrfit_b200
from nmrfit_b200 import synth, _cabi, equations
def used():
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2**20
torch.cuda.init(); base = used()
rng = np.random.default_rng(0)
marks = []
for rep in range(60):
    N = int(rng.choice([1024, 2048, 4096, 5000, 8192])); P = int(rng.choice([6, 12]))
    data, true = synth.multiplet(N, P, seed=rep)
    lo, up = data.generate_solution_bounds()
    with contextlib.redirect_stdout(io.StringIO()):
        f = nmrfit_b200.fit(data, lo, up, summary=False, options={'swarmsize': int(rng.integers(20, 220)), 'maxiter': 20, 'rng': 'device'})
        f.generate_result(scale=2)
        equations.objective_batch(np.tile(f.params, (5, 1)), data.w, data.u, data.v, f.weights)
        if rep % 10 == 0:
            ds = [synth.multiplet(1024, 6, seed=100 + b)[0] for b in range(8)]
            bs = [d.generate_solution_bounds() for d in ds]
            nmrfit_b200.fit_batch(ds, [b[0] for b in bs], [b[1] for b in bs], options={'swarmsize': 30, 'maxiter': 10})
    if rep % 10 == 9:
        marks.append(round(used() - base, 1))
print('device MiB above baseline after every 10 fits:', marks)
_cabi.clear_pool()
for c in list(equations._ctx_cache.values()): c.close()
equations._ctx_cache.clear()
print('after clearing pools:', round(used() - base, 1))
