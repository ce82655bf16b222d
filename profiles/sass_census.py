"""SASS opcode census of libbatman_b200.so, per kernel (cuobjdump -sass; runs without a GPU).

    python profiles/sass_census.py > profiles/r02_sass_census.md

Counts the mnemonics that prove which hardware paths the kernels use: UTCIMMA (tcgen05.mma kind::i8), LDTM
(tcgen05.ld), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk), SYNCS (mbarrier), DMMA (FP64 tensor core), DFMA/DADD/DMUL
(FP64 pipe), MUFU, PRMT, LDS/STS, LDG/STG, BAR, and the total instruction count.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'batman_b200', 'libbatman_b200.so')
WATCH = ['UTCIMMA', 'UTCHMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTCCP', 'UBLKCP', 'UTMALDG', 'SYNCS', 'DMMA', 'DFMA', 'DADD',
         'DMUL', 'MUFU', 'PRMT', 'IMAD', 'LDS', 'STS', 'LDG', 'STG', 'BAR', 'SHFL', 'ATOM', 'RED']

out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and cur:
        op = m.group(1)
        kernels[cur]['total'] += 1
        for w in WATCH:
            if op == w or op.startswith(w + '.') or (w in ('UTCIMMA', 'UTCHMMA', 'UTCBAR', 'UTCCP') and op.startswith(w)):
                kernels[cur][w] += 1
                break


def demangle(names):
    try:
        r = subprocess.run(['cu++filt'] + names, stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
        return dict(zip(names, r))
    except Exception:
        return {n: n for n in names}


names = demangle(list(kernels))
agg = collections.OrderedDict()
for k, c in kernels.items():
    short = re.sub(r'\(.*', '', re.sub(r'\((int|bool)\)', '', names[k])).replace('void bk::', '').replace('bk::', '')
    base = re.sub(r'<.*', '', short)
    key = short if base in ('k_var_ozaki', 'k_var_trmm', 'k_pod_sobol', 'k_pod_inverse') or '<' not in short else base
    agg.setdefault(key, []).append((short, c))

print("# SASS census of libbatman_b200.so (sm_100a), round 2\n")
print("`cuobjdump -sass batman_b200/libbatman_b200.so`, opcode counts per kernel; template families are summed over "
      "their instantiations (count in brackets), the dominant instantiation of the bench is listed separately.\n")
cols = [w for w in WATCH if any(c[w] for lst in agg.values() for _, c in lst)]
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
focus = ['k_predict<1, 8, 3>', 'k_predict<1, 8, 0>', 'k_sobol_static<1, 8, true>']
for key, lst in agg.items():
    tot = collections.Counter()
    for _, c in lst:
        tot.update(c)
    label = key if len(lst) == 1 else "%s [%d inst.]" % (key, len(lst))
    print("| `%s` | %d | " % (label, tot['total']) + " | ".join(str(tot[w]) for w in cols) + " |")
    for short, c in lst:
        if any(short.startswith(f) for f in focus):
            print("| &nbsp;&nbsp;`%s` | %d | " % (short, c['total']) + " | ".join(str(c[w]) for w in cols) + " |")
