"""Condense vbx_b200/csrc/ptxas.log (written by every build, `-Xptxas -v`) into one line per kernel:
registers, stack / spill bytes, static shared memory.  Usage: python tools/ptxas_summary.py > profiles/<round>_ptxas_summary.txt"""
import os
import re
import subprocess
import sys

LOG = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'vbx_b200', 'csrc', 'ptxas.log')


def demangle(names):
    try:
        out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True, timeout=60).stdout.splitlines()
        return [re.sub(r'\(.*', '', o.replace('(anonymous namespace)::', '').replace('void ', '')) for o in out]
    except Exception:
        return names


def main():
    rows, cur = [], None
    for line in open(LOG):
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            cur = dict(name=m.group(1), stack=0, st=0, ld=0, regs=0, smem=0)
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur['stack'], cur['st'], cur['ld'] = map(int, m.groups())
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur['regs'] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", line)
            cur['smem'] = int(s.group(1)) if s else 0
    names = demangle([r['name'] for r in rows])
    print(f'{len(rows)} kernels; {sum(1 for r in rows if r["st"] or r["ld"])} with spills')
    print(f'{"kernel":78s} {"regs":>5s} {"stack":>6s} {"spill st/ld":>12s} {"smem":>7s}')
    for r, n in sorted(zip(rows, names), key=lambda x: x[1]):
        print(f'{n[:78]:78s} {r["regs"]:5d} {r["stack"]:6d} {r["st"]:6d}/{r["ld"]:<5d} {r["smem"]:7d}')


if __name__ == '__main__':
    sys.exit(main())
