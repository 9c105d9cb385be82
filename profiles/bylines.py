#!/usr/bin/env python
"""Warp instructions and stall samples per CUDA source line, from an ncu source page exported with
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass ... (needs --import-source on at capture)
Falls back to SASS-address buckets if no source column is present.   python profiles/bylines.py page.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = rows[1]
print(h[:12])
