"""Static SASS size of a kernel's persistent tile loop (no GPU needed): cuobjdump -sass on an object / library, the
outermost backward branch before EXIT is the tile loop, the largest innermost loop inside it the view loop (two
trips for V = 8).  Prints the opcode mix of one trip through the tile loop (view loop counted twice).

    python profiles/sass_loop_count.py <file.o|.so> <mangled kernel name> [view-loop trips, default 2]

Static counts include the ragged-tile and fallback paths that a full tile never executes, so they sit ~15 % above what
ncu's smsp__inst_executed reports; they are for comparing builds of the same kernel, not for rooflines.
"""
import collections
import re
import subprocess
import sys


def load(path, fun):
    out = subprocess.run(['cuobjdump', '-sass', '-fun', fun, path], stdout=subprocess.PIPE, text=True, check=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opname(text):
    return re.sub(r'^@!?U?P\d+\s+', '', text).split()[0].split('.')[0]


def main():
    path, fun = sys.argv[1], sys.argv[2]
    trips = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ins = load(path, fun)
    first_exit = next(a for a, t in ins if re.search(r'\bEXIT\b', t))
    back = []
    for a, t in ins:
        m = re.search(r'\bBRA(?:\.U)?\b.*\b0x([0-9a-f]+)\s*$', t)
        if m and int(m.group(1), 16) < a and a < first_exit:
            back.append((int(m.group(1), 16), a))
    lo, hi = max(back, key=lambda b: b[1] - b[0])
    # the view loop: the largest loop inside the tile loop that contains no loop itself (the pass loop around it runs once
    # for ordinary joints)
    inner = [(b0, b1) for b0, b1 in back if lo < b0 and b1 < hi
             and not any(b0 <= c0 and c1 <= b1 and (c0, c1) != (b0, b1) for c0, c1 in back)]
    ilo, ihi = max(inner, key=lambda b: b[1] - b[0]) if inner else (0, -1)
    mix = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi:
            mix[opname(t)] += trips if ilo <= a <= ihi else 1
    n_loop = sum(1 for a, _ in ins if lo <= a <= hi)
    n_inner = sum(1 for a, _ in ins if ilo <= a <= ihi)
    print(f'{fun}\n  tile loop 0x{lo:x}-0x{hi:x}: {n_loop} instructions, view loop 0x{ilo:x}-0x{ihi:x}: {n_inner} x {trips} trips')
    print(f'  static instructions per thread and tile: {sum(mix.values())}')
    print('  ' + ', '.join(f'{k} {v}' for k, v in mix.most_common(24)))


if __name__ == '__main__':
    main()
