#!/usr/bin/env python
"""
Attribute an ncu SASS-level source page to the inlined device functions of this repo.

    nvdisasm -gi -c <cubin>  > kernel.dis          (same binary that was profiled)
    ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/sass_profile.py src.csv kernel.dis <kernel-name-substring>

Prints, per device function (innermost inlined frame and inclusive), the executed warp
instructions, FP64-pipe instructions and stall samples, plus the opcode mix.  Used to
write profiles/*.md; it reads files only.
"""
import collections
import csv
import os
import re
import sys

FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')


def load_functions(path):
    """line number -> enclosing function name for a .cuh / .cu file (heuristic)."""
    names = {}
    cur = None
    pat = re.compile(r'^(?:template\s*<[^>]*>\s*)?(?:PM_HD(?:_NOINLINE)?|__device__|__global__|static|inline)[^;=]*?\b([A-Za-z_]\w*)\s*\(')
    try:
        lines = open(path).read().split('\n')
    except OSError:
        return names
    for i, l in enumerate(lines, 1):
        m = pat.match(l.strip())
        if m and not l.strip().endswith(';'):
            cur = m.group(1)
        names[i] = cur
    return names


def parse_dis(path, kernel):
    out = {}
    cur = []
    fresh = False   # True while reading a run of annotation lines
    active = False
    for l in open(path):
        if l.startswith('.text.'):
            active = kernel in l
            cur, fresh = [], False
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if not fresh:
                cur = []
                fresh = True
            cur.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/\s+', l)
        if m:
            fresh = False
            out[int(m.group(1), 16)] = list(cur)
    return out


def main():
    src_csv, dis, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
    locs = parse_dis(dis, kernel)
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    base = int(data[0][ix['Address']], 16)
    fn_cache = {}

    def fn_of(f, n):
        if f not in fn_cache:
            fn_cache[f] = load_functions(f)
        return fn_cache[f].get(n) or os.path.basename(f)

    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    inner = collections.defaultdict(lambda: collections.Counter())
    incl = collections.defaultdict(lambda: collections.Counter())
    ops = collections.Counter()
    tot = collections.Counter()
    for r in data:
        off = int(r[ix['Address']], 16) - base
        ex = int(r[ix['Instructions Executed']])
        smp = int(r[ix['# Samples']])
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[ix['Source']].strip())
        op = m.group(2) if m else '?'
        is64 = op in FP64
        frames = locs.get(off, [])
        names = [fn_of(f, n) for f, n in frames] or ['?']
        rec = {'inst': ex, 'fp64': ex if is64 else 0, 'samples': smp}
        for c in stall_cols:
            rec[c] = int(r[ix[c]])
        for k, v in rec.items():
            inner[names[0]][k] += v
            tot[k] += v
        for nm in set(names):
            for k, v in rec.items():
                incl[nm][k] += v
        ops[op] += ex

    def table(d, title):
        print(f'\n== {title} ==')
        print(f'{"function":28s} {"inst %":>7s} {"fp64 %":>7s} {"fp64/inst":>9s} {"samples %":>9s}  top stalls')
        for nm, c in sorted(d.items(), key=lambda kv: -kv[1]['samples'])[:28]:
            st = sorted(((c[s], s[6:]) for s in stall_cols), reverse=True)[:3]
            sts = ', '.join(f'{s} {100 * v / max(c["samples"], 1):.0f}%' for v, s in st if v)
            print(f'{nm:28s} {100 * c["inst"] / tot["inst"]:7.1f} {100 * c["fp64"] / max(tot["fp64"], 1):7.1f} '
                  f'{c["fp64"] / max(c["inst"], 1):9.2f} {100 * c["samples"] / tot["samples"]:9.1f}  {sts}')

    print(f'total warp instructions {tot["inst"]}, FP64-pipe {tot["fp64"]} ({100 * tot["fp64"] / tot["inst"]:.1f} %), '
          f'samples {tot["samples"]}')
    print('stall totals:', ', '.join(f'{s[6:]} {100 * tot[s] / tot["samples"]:.1f}%' for s in
                                     sorted(stall_cols, key=lambda s: -tot[s])[:10]))
    table(inner, 'innermost inlined function')
    table(incl, 'inclusive')
    print('\n== opcode mix ==')
    for k, v in ops.most_common(24):
        print(f'{k:10s} {100 * v / tot["inst"]:5.1f}%')


if __name__ == '__main__':
    main()
