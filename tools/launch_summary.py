#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (tools/one_step.py): per-kernel totals and
shares, split into the UNet steps and the VAE decode.  No GPU needed.   python tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys


def short(name):
    m = re.search(r"(\w+)<([^>]*)>\(", name) or re.search(r"(\w+)\(", name)
    if not m:
        return name[:50]
    base = m.group(1)
    targs = m.group(2) if m.lastindex and m.lastindex >= 2 else ""
    targs = targs.replace("__nv_bfloat16", "bf16").replace("(int)", "").replace(" ", "")
    return f"{base}<{targs}>" if targs else base


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    launches = [(short(r[ik]), float(r[iv].replace(",", "")) / 1e3, r[ig]) for r in rows[1:] if len(r) == len(hdr)]
    # weight upload (repack_kernel) precedes the first step; the decode starts at the first pointwise_in_kernel<bf16>
    first = next(i for i, l in enumerate(launches) if not l[0].startswith("repack"))
    body = [l for l in launches[first:] if not l[0].startswith("repack")]
    dec0 = next((i for i, l in enumerate(body) if l[0].startswith("pointwise_in_kernel<bf16")), len(body))
    for title, part in (("UNet steps", body[:dec0]), ("VAE decode", body[dec0:])):
        agg = collections.OrderedDict()
        for k, us, g in part:
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1; a[1] += us
        tot = sum(a[1] for a in agg.values()) or 1.0
        print(f"== {title}: {len(part)} launches, {tot:.1f} us (serialised, cold-cache ncu times)")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"  {k:58s} {a[0]:5d} launches {a[1]:10.1f} us {100 * a[1] / tot:5.1f}%  avg {a[1] / a[0]:7.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
