#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libmishagpu.so (cuobjdump -sass): the evidence of which hardware paths a kernel uses
(UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDL / STL = local memory) and a static instruction count.

    python tools/sass_summary.py [pattern] [--top N] [--dump DIR]      -> markdown on stdout
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "misha_b200", "libmishagpu.so")


def kernels():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name is not None:
            body.append(line)
    if name:
        yield name, body


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    pat = re.compile(args[0]) if args else None
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    print("| kernel | instr | UBLKCP | SYNCS | LDL+STL | LDS | LDG | top opcodes |")
    print("|---|---|---|---|---|---|---|---|")
    for name, body in kernels():
        dn = demangle(name)
        if pat and not pat.search(dn):
            continue
        ops = collections.Counter()
        for line in body:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
        if dump:
            os.makedirs(dump, exist_ok=True)
            open(os.path.join(dump, re.sub(r"[^A-Za-z0-9_]+", "_", dn)[:120] + ".sass"), "w").write("\n".join(body))
        tot = sum(ops.values())
        print("| `%s` | %d | %d | %d | %d | %d | %d | %s |" % (dn.replace("|", "\\|"), tot, ops["UBLKCP"], ops["SYNCS"], ops["LDL"] + ops["STL"], ops["LDS"], ops["LDG"],
                                                    " ".join("%s:%d" % kv for kv in ops.most_common(top))))


if __name__ == "__main__":
    main()
