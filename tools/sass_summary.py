#!/usr/bin/env python
"""profiles/r02_sass_summary.txt: per kernel of libcrl_b200.so -- SASS instruction count and the mnemonics that prove
what the kernel is (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, ...); full listing of the headline
kernels beside it (gzip)."""
import collections, gzip, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'combinatorial_rl_tasks_b200', 'libcrl_b200.so')
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
KEY = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'HMMA', 'SYNCS', 'ACQBULK', 'ATOMG', 'REDG', 'CCTL',
       'MEMBAR', 'ERRBAR', 'LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL', 'FFMA', 'FMUL', 'FADD', 'DFMA', 'DMUL', 'DADD', 'SHFL', 'MUFU', 'BAR', 'CALL']
kern, cur, body = collections.OrderedDict(), None, []
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(.*', '', cur).replace('crl::', '').replace('void ', '')
        kern[cur] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        kern[cur]['_n'] += 1
        kern[cur][m.group(1)] += 1
arch = re.search(r'arch = (\S+)', out)
with open(os.path.join(ROOT, 'profiles', 'r02_sass_summary.txt'), 'w') as f:
    f.write(f'cuobjdump -sass combinatorial_rl_tasks_b200/libcrl_b200.so  ({arch.group(1) if arch else "?"}); instruction counts per kernel\n')
    for k, c in kern.items():
        f.write(f'{c["_n"]:6d}  {k}\n        ' + ' '.join(f'{m}={c[m]}' for m in KEY if c[m]) + '\n')
full = subprocess.run(['cuobjdump', '-sass', '-fun', '_ZN3crl11step_kernelILi0ELi15ELb0ELb0EEEvNS_7KParamsE', LIB], capture_output=True, text=True).stdout
with gzip.open(os.path.join(ROOT, 'profiles', 'r02_sass_step_kernel_tsp15.txt.gz'), 'wt') as f:
    f.write(full)
full = subprocess.run(['cuobjdump', '-sass', '-fun', '_ZN7crl_enc18zone_encode_kernelILb0EEEvNS_7EncArgsE', LIB], capture_output=True, text=True).stdout
with gzip.open(os.path.join(ROOT, 'profiles', 'r02_sass_zone_encode_kernel.txt.gz'), 'wt') as f:
    f.write(full)
print(open(os.path.join(ROOT, 'profiles', 'r02_sass_summary.txt')).read()[:3000])
