#!/usr/bin/env python
"""Development aid: attribute an ncu capture's executed instructions and stall samples to source lines.
usage: ncu_lines.py file.ncu-rep kernel_substring [lib.so] [top_n]
Joins `ncu --page source --print-source sass` (per-instruction counts) with `nvdisasm -g` line info of the
cubin inside the library (instruction offsets are identical)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table, active, cur = {}, False, ('?', 0)
    for ln in out.splitlines():
        if ln.startswith('.text.'):
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, 'gym_so100_c_b200', 'libso100_b200.so')
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(lib, kernel)
    # NCU_PICK="-k regex:phase_solve_light -c 1": pick one result of a multi-kernel report
    pick = os.environ.get('NCU_PICK', '').split()
    out = subprocess.run(['ncu', '-i', rep] + pick + ['--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    allrows = list(csv.reader(out.splitlines()))
    # a report with several results prints one block per kernel ("Kernel Name" row, header row, instruction rows): keep the one asked for
    want = os.environ.get('NCU_KERNEL', kernel.split('ILj')[0].replace('_ZN5so100', '').lstrip('0123456789'))
    rows, keep = [], False
    for r in allrows:
        if r and r[0] == 'Kernel Name':
            if rows:
                break
            keep = want in r[1]
            if keep:
                rows.append(r)
            continue
        if keep:
            rows.append(r)
    hdr = rows[1]
    ia, ii, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stall_cols = [(k, h) for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    base = None
    per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_i = tot_s = 0
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        addr = int(r[ia], 16)
        base = addr if base is None else base
        key, _ = table.get(addr - base, (('?', 0), ''))
        e = per_line[key]
        n, s = int(r[ii]), int(r[isamp])
        e[0] += n; e[1] += s
        tot_i += n; tot_s += s
        for k, h in stall_cols:
            v = int(r[k])
            if v:
                e[2][h[6:]] += v
    print(f'{kernel}: {len(table)} SASS instructions, {tot_i} warp-instructions executed, {tot_s} samples')
    by_file = collections.Counter()
    for (f, l), e in per_line.items():
        by_file[f] += e[0]
    print('by file:', ', '.join(f'{f} {100 * n / tot_i:.1f}%' for f, n in by_file.most_common(8)))
    print(f'{"line":28s} {"inst%":>6s} {"samp%":>6s}  top stalls')
    for (f, l), e in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        st = ' '.join(f'{k}:{v}' for k, v in e[2].most_common(4))
        print(f'{f + ":" + str(l):28s} {100 * e[0] / tot_i:6.2f} {100 * e[1] / max(tot_s, 1):6.2f}  {st}')


if __name__ == '__main__':
    main()
