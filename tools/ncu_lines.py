#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` capture: instructions executed, stall samples (with the
dominant stall reasons) and excessive shared-memory wavefronts, for the lines that matter most.
Usage: python tools/ncu_lines.py report.ncu-rep [kernel-index] [top-n]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, hdr, kernels, kidx = None, None, [], -1
    agg = collections.OrderedDict()
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            if r[1] not in kernels:
                kernels.append(r[1])
            kidx = kernels.index(r[1])
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and kidx == which:
            g = lambda name: int(r[hdr.index(name)] or 0) if name in hdr and r[hdr.index(name)].replace(".", "").isdigit() else 0
            e = agg.setdefault((cur, int(r[0])), dict(src=r[1], inst=0, samp=0, exc=0, stalls=collections.Counter()))
            e["inst"] += g("Instructions Executed")
            e["samp"] += g("Warp Stall Sampling (All Samples)")
            e["exc"] += g("L1 Wavefronts Shared Excessive")
            for name in hdr:
                if name.startswith("stall_") and "Not Issued" not in name:
                    e["stalls"][name[6:]] += g(name)
    ti = sum(e["inst"] for e in agg.values()) or 1
    ts = sum(e["samp"] for e in agg.values()) or 1
    print("kernel: %s" % kernels[which][:90])
    print("total warp instructions %.3f M, stall samples %d" % (ti / 1e6, ts))
    allst = collections.Counter()
    for e in agg.values():
        allst.update(e["stalls"])
    print("stall mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(sum(allst.values()), 1)) for k, v in allst.most_common(8)))
    top = sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:topn]
    for (f, ln), e in sorted(top, key=lambda kv: (kv[0][0] != "k3_composite.cu", kv[0][0], kv[0][1])):
        st = ",".join("%s:%d" % (k, 100.0 * v / max(sum(e["stalls"].values()), 1)) for k, v in e["stalls"].most_common(3))
        print("%-16s %5d  samp %5.2f%%  inst %5.2f%%  exc_smem %8d  [%s] | %s" %
              (f[:16], ln, 100.0 * e["samp"] / ts, 100.0 * e["inst"] / ti, e["exc"], st, e["src"].strip()[:90]))


if __name__ == "__main__":
    main()
