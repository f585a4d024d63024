#!/usr/bin/env python
"""Text summary of an .ncu-rep (run HERE, where the report was pulled to): the raw metrics that matter for this repo's
kernels, per profiled launch, and the instructions that collected the most warp-stall samples in the first launch.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--kernel regex] [--top 30] > profiles/rNN_x_summary.txt
"""
import argparse
import csv
import io
import subprocess

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_shared_mem',
           'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']


def ncu(args):
    return subprocess.run(['ncu'] + args, check=True, capture_output=True, text=True).stdout


def to_int(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--kernel', default=None, help='regex on the kernel name')
    ap.add_argument('--top', type=int, default=30)
    a = ap.parse_args()
    sel = ['--kernel-name', 'regex:' + a.kernel] if a.kernel else []
    rows = list(csv.reader(io.StringIO(ncu(['-i', a.report, '--page', 'raw', '--csv'] + sel))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name = hdr.index('Kernel Name')
    print('report', a.report)
    print('launches:', [r[name][:60] for r in data])
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print('%-72s %-10s %s' % (m, units[i], [r[i] for r in data]))
    stall = [(h, i) for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
    for r in data[:1]:
        tot = sum(float(r[i] or 0) for _, i in stall)
        print('warp states of launch 0 (share of warp-cycles):')
        for h, i in sorted(stall, key=lambda s: -float(r[s[1]] or 0))[:8]:
            print('  %-28s %5.1f %%' % (h.split('issue_stalled_')[1].split('_per_issue')[0], 100 * float(r[i] or 0) / tot))
    src = list(csv.reader(io.StringIO(ncu(['-i', a.report, '--page', 'source', '--csv', '--launch-count', '1'] + sel))))
    sh = src[1]
    body = []
    for r in src[2:]:
        if len(r) > 1 and r[0] == 'Address':
            break                                                  # a second view of the same kernel follows
        if len(r) == len(sh):
            body.append(r)
    i_s, i_src, i_ex = sh.index('# Samples'), sh.index('Source'), sh.index('Instructions Executed')
    cols = [i for i, h in enumerate(sh) if h.startswith('stall_') and 'Not Issued' not in h]
    total = sum(to_int(r[i_s]) for r in body)
    print('SASS instructions %d, samples %d, warp instructions executed %d' % (len(body), total, sum(to_int(r[i_ex]) for r in body)))
    by_kind = {sh[i]: sum(to_int(r[i]) for r in body) for i in cols}
    for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1])[:8]:
        print('  %-24s %6d  %4.1f %%' % (k, v, 100.0 * v / max(total, 1)))
    top = sorted(range(len(body)), key=lambda k: -to_int(body[k][i_s]))[:a.top]
    for k in sorted(top):
        r = body[k]
        st = sorted(((to_int(r[i]), sh[i]) for i in cols), reverse=True)[:2]
        print('  %5d  %-70s %5s %s' % (k, r[i_src][:70], r[i_s], [(n, s) for n, s in st if n]))


if __name__ == '__main__':
    main()
