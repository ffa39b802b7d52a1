#!/usr/bin/env python
"""Segment an `ncu --page source --csv` export into straight-line runs with equal execution counts:
which part of the kernel the issue slots go to.  Usage: ncu_source_segments.py <src.csv> <units (e.g. tiles)>"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]; data = rows[2:]
isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); ith = hdr.index('Avg. Threads Executed'); ismp = hdr.index('# Samples')
iw = hdr.index('L1 Wavefronts Shared'); iwi = hdr.index('L1 Wavefronts Shared Ideal')
tot = sum(int(r[iex]) for r in data)
print("total warp instructions", tot, " per unit %.1f" % (tot / units))
prev = None; start = 0; segs = []
for i, r in enumerate(data):
    e = int(r[iex])
    if prev is None:
        prev = e; start = i
    elif abs(e - prev) > 0.02 * max(prev, 1):
        segs.append((start, i - 1, prev)); prev = e; start = i
segs.append((start, len(data) - 1, prev))
for s, e, c in segs:
    n = e - s + 1
    if c * n > tot * 0.004:
        thr = sum(float(data[k][ith]) for k in range(s, e + 1)) / n
        smp = sum(int(data[k][ismp]) for k in range(s, e + 1))
        print("lines %4d-%4d n=%3d exec=%10d (%.2f/unit) share=%5.1f%% threads=%4.1f samples=%5d  %s" %
              (s, e, n, c, c / units, 100.0 * c * n / tot, thr, smp, data[s][isrc].strip()[:48]))
wf = sum(int(r[iw]) for r in data); wfi = sum(int(r[iwi]) for r in data)
print("shared wavefronts", wf, "ideal", wfi)
for r in data:
    if wf and int(r[iw]) > wf * 0.02:
        print("   ", r[isrc].strip()[:60], "exec", r[iex], "wavefronts", r[iw], "ideal", r[iwi], "threads", r[ith])
