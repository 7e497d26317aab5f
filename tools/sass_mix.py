#!/usr/bin/env python
"""SASS opcode histogram of one kernel of libelliptic_b200.so, split at its out-of-line sub-routines
(`cuobjdump -sass -fun <mangled>` + the `$kernel$callee` symbols of `cuobjdump -elf`).
usage: python tools/sass_mix.py k256_verify_kernel [lib] > profiles/rNN_sass_mix_<kernel>.txt"""
import collections
import re
import subprocess
import sys


def main():
    pat = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else "elliptic_b200/libelliptic_b200.so"
    elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
    kern = None
    for l in elf.splitlines():
        m = re.search(r"\.text\.(_Z\w*%s\w*)" % re.escape(pat), l)
        if m and "PROGBITS" in l:
            kern = m.group(1)
            break
    if not kern:
        sys.exit("kernel not found: " + pat)
    subs = []
    for l in elf.splitlines():
        f = l.split()
        if len(f) >= 7 and f[0].startswith("0x") and f[-1].startswith("$" + kern + "$"):
            subs.append((int(f[1], 16), int(f[2], 16), f[-1].split("$")[-1]))
    subs.sort()
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", kern, lib], capture_output=True, text=True).stdout
    ins = []
    for l in sass.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;?\s*/\*", l)
        if m:
            body = m.group(2).split()
            op = body[1] if body[0].startswith("@") else body[0]
            ins.append((int(m.group(1), 16), op.rstrip(";")))
    regions = [("main body", 0, subs[0][0] if subs else 1 << 30)] + [(n, a, a + sz) for a, sz, n in subs]
    print("# SASS opcode mix of %s (sm_100a), %s" % (kern, lib))
    print("whole kernel: %d instructions" % len(ins))
    tot = collections.Counter(op for _, op in ins)
    print("  " + ", ".join("%s %d" % kv for kv in tot.most_common(16)))
    calls = collections.Counter()
    for name, lo, hi in regions:
        c = collections.Counter(op for a, op in ins if lo <= a < hi)
        n = sum(c.values())
        print("%s @0x%x: %d instructions" % (name, lo, n))
        print("  " + ", ".join("%s %d" % kv for kv in c.most_common(14)))
    mac = sum(v for k, v in tot.items() if k.startswith("IMAD.WIDE"))
    mov = sum(v for k, v in tot.items() if k.startswith("IMAD.MOV") or k == "MOV")
    add = sum(v for k, v in tot.items() if k.startswith("IADD3"))
    print("static totals: IMAD.WIDE.U32(.X) %d, IMAD.MOV(.U32)+MOV %d (%.1f %% of the kernel), IADD3(.X) %d"
          % (mac, mov, 100.0 * mov / len(ins), add))


if __name__ == "__main__":
    main()
