#!/usr/bin/env python
"""Attribute the hot-loop SASS instructions of a kernel to source lines (needs -lineinfo; uses
nvdisasm -g on build/kernels.cubin produced by scripts/sass_stats.py).
Usage: python scripts/sass_lines.py <kernel-substring> [loop-index]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUBIN = os.path.join(ROOT, "build", "kernels.cubin")


def helpers_end():
    src = open(os.path.join(ROOT, "svbrdf_estimation_b200", "csrc", "shading.cuh")).read().split("\n")
    for i, l in enumerate(src):
        if l.startswith("template <typename T> struct LaneTraits;"):
            return i + 1
    return 0


HELPERS_END = helpers_end()


def main():
    key = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    txt = subprocess.run(["nvdisasm", "-gi", "-c", CUBIN], stdout=subprocess.PIPE, text=True).stdout
    secs = re.split(r"\n(?=\.text\.)", txt)
    sec = [s for s in secs if s.startswith(".text.") and key in s.split("\n", 1)[0]][0]
    cur, rows, labels, chain, fresh = None, [], {}, [], True
    for l in sec.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                chain, fresh = [], False
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            # attribute to the innermost frame that is not one of the lane-type helper one-liners
            body = [c for c in chain if not (c[0] == "shading.cuh" and c[1] < HELPERS_END)]
            cur = body[0] if body else chain[-1]
            continue
        fresh = True
        m = re.match(r"^(\.L_x_\d+):", l.strip())
        if m:
            labels[m.group(1)] = len(rows)
            continue
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*([^;]*);", l)
        if m:
            rows.append((m.group(2), m.group(3), cur))
    loops = []
    for i, (op, args, _) in enumerate(rows):
        if op.startswith("BRA"):
            m = re.search(r"(\.L_x_\d+)", args)
            if m and m.group(1) in labels and labels[m.group(1)] < i:
                loops.append((labels[m.group(1)], i))
    s, e = loops[which]
    print("loop %d of %d: instructions %d..%d (%d)" % (which, len(loops), s, e, e - s + 1))
    agg = collections.OrderedDict()
    for op, args, cur in rows[s:e + 1]:
        agg.setdefault(cur, collections.Counter())[op.split(".")[0]] += 1
    src = {f: open(os.path.join(ROOT, "svbrdf_estimation_b200", "csrc", f)).read().split("\n") for f in ("shading.cuh", "kernels.cu")}
    for cur, cnt in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
        text = src[cur[0]][cur[1] - 1].strip()[:100] if cur and cur[0] in src else ""
        print("%-12s %4d  n=%2d %-46s | %s" % (cur[0] if cur else "-", cur[1] if cur else 0, sum(cnt.values()),
                                               " ".join("%s:%d" % kv for kv in cnt.most_common()), text))


if __name__ == "__main__":
    main()
