#!/usr/bin/env python
"""Per-phase warp-instruction breakdown of a K3 capture: `ncu --set full --import-source on` stores the CUDA source
with per-line 'Instructions Executed'; this sums them per kernel phase, using marker comments found in the source
text embedded in the report itself (so old reports stay readable after the file changes).
Usage: python tools/ncu_source_breakdown.py gpurun_out/prof_k3_fast_s3.ncu-rep [...]"""
import collections
import csv
import subprocess
import sys

# (bucket, marker that starts it) in file order, per kernel generation
# Only source lines that own SASS appear in the report, so the markers are code lines (alternatives per bucket).
FAST = [("prologue", ["extern __shared__", "const int h = gm.h", "const int strip_words"]),
        ("phase 1 (mask -> bit rows)", ["for (int i = warp; i < rows_s", "const uint32_t last_valid = (W0 & 31)", "if (BITS) {"]),
        ("alpha LUT / setup", ["if (warp == 0) {", "if (threadIdx.x < 16) {"]),
        ("worker: quad up-scale + blend", ["const int xq = item.x & 0xffff", "const Tap *taps_s = reinterpret_cast<const Tap *>",
                                           "const int xq = (int)(item & 0x3ffu) << 2"]),
        ("phase 2: classification", ["const int G = gm.G", "const int G = W0 >> 4;", "int qcount = 0, it = 0, j = 0;",
                                     "for (int step = 0; step <= n_steps", "const bool drain = step == n_steps;"]),
        ("compaction (queue push)", ["if (__ballot_sync(0xffffffffu, need != 0))", "const uint32_t bal_b = __ballot_sync"]),
        ("worker loop control", ["if (qcount >= 32 || (drain", "if (qcount >= 32) {"]),
        ("epilogue (bulk store)", ["fence_proxy_async();"])]
OLD = [("prologue", ["extern __shared__", "const int nthreads = TMA"]),
       ("phase 1 (mask -> bit rows)", ["uint16_t *b16 = reinterpret_cast<uint16_t *>(bits);"]),
       ("phase 2: classification", ["const int G = (W0 + 15) >> 4;"]),
       ("compaction (queue push)", ["uint32_t any_need = 0, qn = 0;", "any_need |= need[k];", "qn += ((need[k]"]),
       ("worker: quad up-scale + blend", ["if (qcount < 32 && !(drain"]),
       ("epilogue (bulk store)", ["fence_proxy_async();"])]


def main():
    for rep in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                             capture_output=True, text=True).stdout
        cur, hdr, kernel = None, None, ""
        lines = collections.OrderedDict()          # (file, line) -> [source, count]
        for r in csv.reader(txt.splitlines()):
            if not r:
                continue
            if r[0] == "File Path":
                cur = r[1].split("/")[-1]
            elif r[0] == "Function Name":
                kernel = r[1].split("(")[0]
            elif r[0] == "Line No":
                hdr = r
            elif hdr and r[0].isdigit():
                try:
                    n = int(r[hdr.index("Instructions Executed")])
                except ValueError:
                    continue
                ent = lines.setdefault((cur, int(r[0])), [r[1], 0])
                ent[1] += n
        main_file = "k3_composite.cu"
        src = {ln: s for (f, ln), (s, _) in lines.items() if f == main_file}
        fast = "k3_fast" in kernel
        marks = []
        for name, pats in (FAST if fast else OLD):
            hit = [ln for ln, s in sorted(src.items()) if any(p in s for p in pats) and (not marks or ln > marks[-1][1])]
            if hit:
                marks.append((name, hit[0]))
        first = marks[0][1] if marks else 0
        buckets = collections.Counter()
        for (f, ln), (s, n) in lines.items():
            if f == main_file:
                name = "helpers above the kernel (bit_window, x2/x4 fetch / assemble + hpass + vpass)" if ln < first else \
                    [nm for nm, start in marks if start <= ln][-1]
            elif f == "common.cuh":
                name = "common.cuh (u8<->f32, nonzero_bits16, TMA wrappers)"
            else:
                name = "intrinsics headers (shfl, syncwarp ...)"
            buckets[name] += n
        total = sum(buckets.values())
        print("== %s :: %s   total %.3f G warp instructions" % (rep.split("/")[-1], kernel, total / 1e9))
        for name, n in buckets.most_common():
            print("   %-70s %12d %5.1f %%" % (name, n, 100.0 * n / max(total, 1)))


if __name__ == "__main__":
    main()
