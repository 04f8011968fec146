#!/usr/bin/env python
"""Condense ncu output into the small text/JSON summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
    python tools/ncu_summary.py full gpurun_out/prof_objective.ncu-rep profiles/r01_objective_full.md

`launches`: per-kernel totals and shares of a `--metrics gpu__time_duration.sum --csv` launch list.
`full`: the metrics of a `--set full` capture that DESIGN.md and bench.py's roofline quote
(durations, DRAM bytes, pipe utilisation, issue rate, stall reasons, occupancy), per launch.
Runs here without a GPU (ncu -i reads the report).
"""
import collections
import csv
import json
import subprocess
import sys

KEEP = (
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
    'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
)


def num(s):
    try:
        return float(s.replace(',', ''))
    except ValueError:
        return s


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 's': 1e6}
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        tot[r[ki]] += num(r[vi]) * scale[r[ui]]
        cnt[r[ki]] += 1
    total = sum(tot.values())
    with open(dst, 'w') as f:
        f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none): per-kernel totals\n\n')
        f.write('source: `%s`, %d launches, %.1f us in total (cold-cache, serialised: compare shares)\n\n' % (src, len(rows) - 1, total))
        f.write('| kernel | launches | total us | us / launch | share |\n|---|---:|---:|---:|---:|\n')
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write('| `%s` | %d | %.1f | %.2f | %.2f %% |\n' % (k[:110], cnt[k], v, v / cnt[k], 100 * v / total))


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    body = rows[2:]
    with open(dst, 'w') as f:
        f.write('# ncu --set full capture: selected metrics per profiled launch\n\n')
        f.write('source: `%s` (%d launches); kernels: %s\n\n' % (
            src, len(body), ', '.join('`%s`' % k for k in sorted({r[hdr.index('Kernel Name')] for r in body}))))
        f.write('| metric | unit | ' + ' | '.join('launch %d' % i for i in range(len(body))) + ' |\n')
        f.write('|---|---|' + '---:|' * len(body) + '\n')
        for i, h in enumerate(hdr):
            if h in KEEP:
                f.write('| %s | %s | %s |\n' % (h, units[i], ' | '.join(r[i] for r in body)))
        f.write('\n## warps stalled per issued instruction, by reason (>= 0.01)\n\n')
        f.write('| reason | ' + ' | '.join('launch %d' % i for i in range(len(body))) + ' |\n|---|' + '---:|' * len(body) + '\n')
        pre, post = 'smsp__average_warps_issue_stalled_', '_per_issue_active.ratio'
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(pre) and h.endswith(post):
                vals = [num(r[i]) for r in body]
                if all(isinstance(v, float) for v in vals) and max(vals) >= 0.01:
                    stalls.append((max(vals), h[len(pre):-len(post)], vals))
        for _, name, vals in sorted(stalls, reverse=True):
            f.write('| %s | %s |\n' % (name, ' | '.join('%.3f' % v for v in vals)))


def objective_json(src, dst, workload='c2', peak_points=4096 * 32768 * 12):
    """profiles/objective_ncu.json: what bench.py attaches to its roofline (traffic + executed-instruction view).
    peak_points = peak-points one launch processes (particles * points * peaks)."""
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[-1]

    def g(name):
        return num(r[hdr.index(name)])

    def to_bytes(name):
        v, u = g(name), units[hdr.index(name)]
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    peak_points = float(peak_points)
    cycles = g('sm__cycles_elapsed.avg')
    per_op = {k: g('smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed' % k) * cycles for k in ('dfma', 'dmul', 'dadd')}
    fp64 = sum(per_op.values())
    d = {
        'kernel': r[hdr.index('Kernel Name')], 'source': src,
        'duration_ms_under_ncu': g('gpu__time_duration.sum') * {'ms': 1, 'us': 1e-3, 'ns': 1e-6, 'usecond': 1e-3, 'msecond': 1}[units[hdr.index('gpu__time_duration.sum')]],
        'dram_bytes_per_launch': to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum'),
        'fp64_pipe_active_pct': g('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
        'issue_slots_busy_pct': g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
        'warp_inst_per_peak_point': g('smsp__inst_executed.sum') * 32 / peak_points,
        'fp64_arith_inst_per_peak_point': fp64 / peak_points,
        'fp64_flop_per_peak_point': (2 * per_op['dfma'] + per_op['dmul'] + per_op['dadd']) / peak_points,
        'peak_points_per_launch': peak_points,
        'registers_per_thread': g('launch__registers_per_thread'),
        'warps_active_pct': g('sm__warps_active.avg.pct_of_peak_sustained_active'),
    }
    try:
        out = json.load(open(dst))
    except (OSError, ValueError):
        out = {}
    out[workload] = d
    json.dump(out, open(dst, 'w'), indent=1)


if __name__ == '__main__':
    # objective_json <report> <dst.json> <workload> <peak-points per launch>
    {'launches': launches, 'full': full, 'objective_json': objective_json}[sys.argv[1]](*sys.argv[2:6])
