#!/usr/bin/env python
"""Aggregate the `--page source --print-source cuda` view of an ncu report by source line: instructions executed and
stall samples per file:line (needs -lineinfo at compile time)."""
import csv
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur_file = "?"
    rows = []
    head = None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            head = None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            head = r
            continue
        if head is None:
            continue
        if not r[0].strip():
            continue      # SASS row
        d = {}
        for k, v in zip(head, r):
            d.setdefault(k, v)   # "Source" appears twice (CUDA, SASS): keep the first
        try:
            inst = int(d.get("Instructions Executed", "0") or 0)
            samp = int(d.get("# Samples", "0") or 0)
        except ValueError:
            continue
        if inst or samp:
            rows.append((inst, samp, cur_file, d["Line No"], d["Source"].strip()[:110]))
    ti = sum(r[0] for r in rows) or 1
    ts = sum(r[1] for r in rows) or 1
    print(f"total warp instructions {ti}, stall samples {ts}")
    for inst, samp, f, ln, src in sorted(rows, reverse=True)[:top]:
        print(f"{100 * inst / ti:5.1f}% inst {100 * samp / ts:5.1f}% samp  {f}:{ln:>4s}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
