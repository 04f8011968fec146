import sys, os, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from nmrfit_b200 import _cabi, synth, utils
for (P, N, S, seed) in [(6, 4096, 65536, 1000), (6, 16384, 16384, 3000), (6, 2048, 65536, 7)]:
    data, _ = synth.multiplet(N, P, seed=seed)
    lo, up = (np.array(a) for a in data.generate_solution_bounds())
    wts = utils.compute_weights(data.w, data.peaks)
    xs = torch.from_numpy(synth.particles(lo, up, S, seed=7)).cuda()
    f = torch.empty(S, dtype=torch.float64, device='cuda')
    with _cabi.Context(1, N, P) as ctx:
        ctx.set_spectrum(0, data.w, data.u, data.v, wts)
        for th, r, sp in [(0, 0, 0), (128, 4, 0), (256, 4, 0), (128, 8, 0), (256, 8, 0), (128, 16, 0), (256, 16, 0), (256, 4, 16), (128, 4, 16)]:
            ctx.set_tuning(th, r, 0, sp)
            try:
                for _ in range(3): ctx.objective_device(xs, S, f)
                ctx.profile(True)
                for _ in range(10): ctx.objective_device(xs, S, f)
                ms, n = ctx.profile_read(); ctx.profile(False)
                print(P, N, S, (th, r, sp), ctx.get_tuning(S), 'ms %.4f' % (ms / n), 'pp/s %.3e' % (S * N * P / (ms / n * 1e-3)), flush=True)
            except Exception as e:
                print(P, N, (th, r, sp), 'ERR', str(e)[:100])
