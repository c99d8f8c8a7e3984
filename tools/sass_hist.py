#!/usr/bin/env python
"""Static opcode histogram of the SASS of one or more kernels, with the pipe-cycle model measured on B200
(profiles/r02_microbench_pipes.txt): IMAD.WIDE / IMAD.HI occupy the integer-multiply (fmaheavy) pipe for 4 cycles
per warp, every other IMAD* for 2; IADD3 / LOP3 / SEL / ISETP / MOV / SHF occupy the ALU pipe for 2.

    python tools/sass_hist.py <file.so|file.cubin> <function-name regex> [butterflies per thread]
"""
import collections
import re
import subprocess
import sys


def histogram(path, pattern):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, res = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1) if re.search(pattern, m.group(1)) else None
            if fn:
                res[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        res[fn][op] += 1
    return res


def pipe_cycles(c):
    fma = alu = 0
    for op, n in c.items():
        if re.match(r"IMAD\.(WIDE|HI)", op):
            fma += 4 * n
        elif op.startswith("IMAD"):
            fma += 2 * n
        elif re.match(r"(IADD3|LOP3|SEL|ISETP|MOV|SHF|LEA|PRMT|VIADD|CS2R|IABS|VIMNMX|PLOP3)", op):
            alu += 2 * n
    return fma, alu


def main():
    path, pattern = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 0
    for fn, c in histogram(path, pattern).items():
        fma, alu = pipe_cycles(c)
        tot = sum(c.values())
        print(f"== {fn}\n   instructions {tot}  multiply-pipe cycles/warp {fma}  ALU-pipe cycles/warp {alu}")
        if per:
            print(f"   per butterfly ({per:g}/thread): {tot / per:.1f} instr, {fma / per:.1f} multiply-pipe cycles, {alu / per:.1f} ALU cycles")
        for op, n in c.most_common():
            if n >= max(3, tot // 400):
                print(f"   {n:6d} {op}")


if __name__ == "__main__":
    main()
