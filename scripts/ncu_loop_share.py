#!/usr/bin/env python
"""In-loop vs out-of-loop shares of a kernel from an ncu source-page dump (needs --import-source on at capture):
    ncu -i X.ncu-rep --page source --csv > src.csv; python scripts/ncu_loop_share.py src.csv [iterations=9]
Instructions executed `iterations` x as often as the once-per-warp ones are the record loop."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    hdr, data = rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    executed = lambda r: int(r[idx["Instructions Executed"]]) if r[idx["Instructions Executed"]].isdigit() else 0
    samples = lambda rs: sum(int(r[idx["# Samples"]]) for r in rs if r[idx["# Samples"]].isdigit())
    once = collections.Counter(executed(r) for r in data if executed(r) > 0).most_common(1)[0][0]
    loop = [r for r in data if executed(r) == once * iters]
    out = [r for r in data if 0 < executed(r) <= once]
    total = samples(data)
    n_out = sum(executed(r) for r in out) / once
    print("kernel:", rows[0][1][:90])
    print("record loop: %d instructions x %d iterations = %d per warp; outside the loop: %.0f per warp (%.1f %%)"
          % (len(loop), iters, len(loop) * iters, n_out, 100 * n_out / (len(loop) * iters + n_out)))
    print("warp-residency samples: loop %.1f %%, outside %.1f %%" % (100 * samples(loop) / total, 100 * samples(out) / total))
    for name, rs in (("in-loop", loop), ("outside", out)):
        agg = {s: sum(int(r[idx[s]]) for r in rs) for s in stalls}
        print("%-8s stall mix: %s" % (name, ", ".join("%s %.1f %%" % (s[6:], 100 * v / max(1, samples(rs)))
                                                   for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7])))


if __name__ == "__main__":
    main()
