#!/usr/bin/env python
"""Instruction counts / stall samples per CUDA source line from an .ncu-rep captured with --import-source on."""
import csv, io, subprocess, sys

def main(path, top=32):
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                         capture_output=True, text=True).stdout
    out = []
    hd = None
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == 'Line No':
            hd = r; continue
        if hd is None or not r[0].isdigit():
            continue
        try:
            ie, sm, th = int(r[7]), int(r[6]), int(r[8])
        except ValueError:
            continue
        out.append((ie, sm, th, r[1].strip()[:100], r[0]))
    tot = sum(o[0] for o in out) or 1
    ts = sum(o[1] for o in out) or 1
    print('warp instructions', tot, 'samples', ts)
    for ie, sm, th, s, ln in sorted(out, reverse=True)[:top]:
        print(f"{100*ie/tot:5.1f}% inst {100*sm/ts:5.1f}% smp  act {th/max(ie,1):4.1f}  L{ln}: {s}")

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 32)
