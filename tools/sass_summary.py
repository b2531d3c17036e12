#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libvvb200.so, the counts of the instructions that identify the
techniques DESIGN.md claims (bulk async copies = TMA 1-D, mbarriers, packed fp32, dot products, byte permutes,
warp shuffles / votes, programmatic dependent launch) - and the absence of tensor-core instructions, by design.
Usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "videovanish_b200", "csrc", "libvvb200.so")
PAT = collections.OrderedDict([
    ("UBLKCP (cp.async.bulk g<->s)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("FFMA2/FMUL2/FADD2 (f32x2)", r"\bF(FMA|MUL|ADD)2\b"),
    ("IDP (dp2a/dp4a)", r"\bIDP"), ("PRMT", r"\bPRMT"), ("SHFL", r"\bSHFL"), ("VOTE/MATCH/REDUX", r"\b(VOTE|MATCH|REDUX)"),
    ("LDG.E.128", r"LDG\.E\.128"), ("STG.E.128", r"STG\.E(\.[A-Z0-9]+)*\.128"), ("ATOM/RED", r"\b(ATOM|RED|ATOMS|ATOMG)\b"),
    ("ACQBULK/griddepcontrol (PDL)", r"\b(ACQBULK|PREEXIT)"), ("LD/ST .SYS (halo flags)", r"\.(STRONG\.SYS|SYS)\b"),
    ("HMMA/IMMA/UTCMMA (tensor cores)", r"\b(HMMA|IMMA|QGMMA|UTC[A-Z]*MMA|UTMALDG)"),
])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    print("cuobjdump -sass %s  (sm_100a only; instruction counts per kernel)" % os.path.relpath(LIB, ROOT))
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print("architectures in the fat binary:", ", ".join(arch))
    kernels = re.split(r"\n\s*Function : ", sass)[1:]
    names = list(PAT)
    print("%-58s %6s " % ("kernel", "insts") + " ".join("%8s" % n.split(" ")[0][:8] for n in names))
    total = collections.Counter()
    for k in kernels:
        mangled, body = k.split("\n", 1)
        demangled = subprocess.run(["c++filt", mangled.strip()], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", demangled).replace("void ", "").replace("vv::", "")
        lines = [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        counts = [sum(1 for l in lines if re.search(p, l)) for p in PAT.values()]
        for n, c in zip(names, counts):
            total[n] += c
        print("%-58s %6d " % (short[:58], len(lines)) + " ".join("%8d" % c for c in counts))
    print()
    for n in names:
        print("%-40s %d" % (n, total[n]))


if __name__ == "__main__":
    main()
