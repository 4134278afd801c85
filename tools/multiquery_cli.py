#!/usr/bin/env python
"""BASELINE.json configs[4] through the drop-in binary: N synthetic lncRNAs (1000 + z % 9001 nt, seeds 4001..) against a synthetic
chromosome (seed 1002), `fasim --queries --devices all`.  Prints the wall time of the scan phase and the aggregate GCUPS.
Usage: multiquery_cli.py [queries] [Mbp] [devices]"""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import splitmix_bases, splitmix_first

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mbp = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
devs = sys.argv[3] if len(sys.argv) > 3 else "all"
d = tempfile.mkdtemp(prefix="mq_")
n = int(mbp * 1e6)
with open(os.path.join(d, "chr.fa"), "wb") as f:
    f.write(b">syn|chr1|1-%d\n" % n); f.write(splitmix_bases(1002, n).tobytes()); f.write(b"\n")
with open(os.path.join(d, "rnas.fa"), "wb") as f:
    for k in range(nq):
        m = 1000 + splitmix_first(4001 + k) % 9001
        f.write(b">synRNA%d\n" % k); f.write(splitmix_bases(4001 + k, m).tobytes()); f.write(b"\n")
os.mkdir(os.path.join(d, "out"))
env = dict(os.environ, LTG_TIMING="1")
t0 = time.perf_counter()
r = subprocess.run([os.path.join(ROOT, "fasim-longtarget_b200", "fasim"), "-f1", "chr.fa", "-f2", "rnas.fa", "-O", "out/", "--queries", "--devices", devs],
                   cwd=d, env=env, capture_output=True, text=True)
wall = time.perf_counter() - t0
assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
cells = sum(float(x) for x in re.findall(r"scan_cells=([0-9.e+]+)", r.stdout))
laps = dict((m.group(1), float(m.group(2))) for m in re.finditer(r"\[fasim timing\] (\S+)\s+at ([0-9.]+) s", r.stderr))
scan_s = laps.get("scan+write", wall) - laps.get("read", 0.0)
rows = sum(len(open(os.path.join(d, "out", f)).read().splitlines()) - 1 for f in os.listdir(os.path.join(d, "out")) if f.endswith("TFOsorted"))
# per-device breakdown of the one-process run: context creation, query switches, time inside the scan calls, destruction
print("\n".join(l for l in r.stderr.splitlines() if l.startswith("[fasim timing]")))
print({"queries": nq, "mbp": mbp, "devices": devs, "wall_s": round(wall, 2), "read_s": laps.get("read"), "scan_and_write_s": round(scan_s, 2),
       "scan_cells": cells, "gcups_scan_phase": round(cells / scan_s / 1e9, 1), "gcups_wall": round(cells / wall / 1e9, 1),
       "output_files": len(os.listdir(os.path.join(d, "out"))), "rows": rows})
