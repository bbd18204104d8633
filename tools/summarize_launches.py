#!/usr/bin/env python3
"""Per-kernel share of one bench step from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: summarize_launches.py launches.csv steps_in_list > shares.csv   (per-launch times under ncu are serialised and
cold-cache: compare SHARES with bench.py's live CUDA-event numbers, not absolutes)"""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
    tot[name] += float(r[iv]) / 1e3
    cnt[name] += 1
total = sum(tot.values())
print("kernel,launches_per_step,total_us_per_step,share_pct")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k},{cnt[k] / steps:.1f},{tot[k] / steps:.1f},{100 * tot[k] / total:.1f}")
print(f"TOTAL,{sum(cnt.values()) / steps:.1f},{total / steps:.1f},100.0")
