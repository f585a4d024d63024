"""Per-source-line executed warp instructions and stall samples from an .ncu-rep (needs -lineinfo).
usage: python profiles/ncu_lines.py REPORT.ncu-rep [min_instr_per_warp]"""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fpath, hdr, first_kernel, seen_fn = None, None, None, set()
lines = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        fn = r[1]
        if first_kernel is None:
            first_kernel = fn
        cur_fn = fn; continue
    if r[0] == 'Line No':
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr and r[0] and r[0].isdigit() and cur_fn == first_kernel:
        key = (fpath, int(r[0]))
        if key in lines:
            continue
        try:
            n = int(r[hdr['Instructions Executed']]); s = int(r[hdr['# Samples']])
        except ValueError:
            continue
        lines[key] = (n, s, r[1].strip()[:110])
nwarps = None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
g = int(float(rr[2][rr[0].index('launch__grid_size')])); b = int(float(rr[2][rr[0].index('launch__block_size')]))
nwarps = g * b // 32
tot = sum(v[0] for v in lines.values()); ts = sum(v[1] for v in lines.values())
print(f'{first_kernel}: {nwarps} warps, {tot / nwarps:.0f} warp-instr per warp attributed, {ts} samples')
for (f, l), (n, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][0]):
    if n / nwarps >= thr:
        print(f'{n / nwarps:7.1f} {100 * s / max(ts, 1):5.1f}%  {f}:{l:<4d} {src}')
