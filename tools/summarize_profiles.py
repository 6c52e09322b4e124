"""Turn the ncu exports in gpurun_out/ (tools/profile_all.sh) into the small tracked summaries under profiles/."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r01'


def launches(wl):
    rows = list(csv.reader(open(os.path.join(ROOT, 'gpurun_out', 'launches_%s.csv' % wl))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ni, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ni].split('(')[0][:80]
        v = float(r[vi].replace(',', ''))
        v = v / 1000 if r[ui] == 'ns' else v * 1000 if r[ui] == 'ms' else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ['# ncu launch list: gpu__time_duration.sum, --clock-control none, first 900 launches of `python tools/run_iters.py %s 4`' % wl,
           '# per-launch times are cold-cache and serialised: compare SHARES, not absolutes', 'kernel,launches,total_us,share']
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append('%s,%d,%.1f,%.4f' % (k, n, t, t / tot))
    open(os.path.join(ROOT, 'profiles', '%s_launches_%s.csv' % (TAG, wl)), 'w').write('\n'.join(out) + '\n')
    return out[:12]


WANT = collections.OrderedDict([
    ('gpu__time_duration.sum', 'duration'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs'), ('launch__occupancy_limit_shared_mem', 'occ_limit_smem_blocks'),
    ('launch__occupancy_limit_registers', 'occ_limit_regs_blocks'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved_occupancy_pct'),
    ('dram__bytes_read.sum', 'dram_read'), ('dram__bytes_write.sum', 'dram_write'),
    ('lts__t_bytes.sum', 'l2_bytes'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_pct'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma_pipe_pct'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64_pipe_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_throughput_pct'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_throughput_pct'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
    ('smsp__inst_executed.sum', 'warp_insts'),
])
UNIT = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1, 'us': 1, 'ms': 1e3, 'ns': 1e-3, 'usecond': 1, 'msecond': 1e3, 'nsecond': 1e-3}


def metrics(wl):
    rows = list(csv.reader(open(os.path.join(ROOT, 'gpurun_out', 'raw_%s.csv' % wl))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = collections.OrderedDict()
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('promp::', '')
        if name in out:
            continue
        d = collections.OrderedDict()
        for k, short in WANT.items():
            if k in idx and r[idx[k]] != '':
                try:
                    v = float(r[idx[k]].replace(',', ''))
                except ValueError:
                    continue
                u = units[idx[k]]
                if u in UNIT and ('bytes' in k or 'duration' in k):
                    v *= UNIT[u]
                d[short + ('_bytes' if 'bytes' in k and not short.endswith('bytes') else '_us' if 'duration' in k else '')] = v
        out[name] = d
    return out


if __name__ == '__main__':
    res = {}
    for wl in ('point', 'cheetah'):
        print('\n'.join(launches(wl)))
        res[wl] = metrics(wl)
    json.dump(res, open(os.path.join(ROOT, 'profiles', '%s_kernel_metrics.json' % TAG), 'w'), indent=1)
    for wl, ks in res.items():
        for k, d in ks.items():
            print(wl, k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in d.items()})
