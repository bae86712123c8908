"""Per-op roofline table from a tools/profile_ops.py listing: for every (op, shape) the measured time of one call next to
its floor = max(FLOPs / sustained tensor peak, algorithmic bytes / HBM copy peak), the fraction of the floor reached and
the time the op would give back at its floor, sorted by that gap. Peaks from MEASURED_PEAKS.json.
Usage: python tools/roofline_table.py profiles/r01_per_op_unet_forward.txt"""
import ast
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
TF, GBS = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
rows = []
for line in open(sys.argv[1]):
    m = re.match(r"(\(.*?\))\s+(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", line)
    if not m:
        continue
    key, n, ms, us, tfs, gbs = ast.literal_eval(m.group(1)), int(m.group(2)), *map(float, m.groups()[2:])
    flops, nbytes = tfs * 1e12 * us * 1e-6, gbs * 1e9 * us * 1e-6          # per call
    floor_us = max(flops / (TF * 1e12), nbytes / (GBS * 1e9)) * 1e6
    bound = "tensor" if flops / (TF * 1e12) >= nbytes / (GBS * 1e9) else "hbm"
    rows.append((n * (us - floor_us) / 1e3, key, n, us, floor_us, bound, ms))
tot = sum(r[6] for r in rows)
tot_floor = sum(r[2] * r[4] for r in rows) / 1e3
print(f"peaks: {TF:.0f} TFLOP/s sustained, {GBS:.0f} GB/s (MEASURED_PEAKS.json); sum of calls {tot:.2f} ms, "
      f"sum of floors {tot_floor:.2f} ms ({100 * tot_floor / tot:.0f} %)")
print(f"{'op / shape':58s} {'n':>3s} {'us':>7s} {'floor':>7s} {'frac':>5s} {'bound':>6s} {'gap ms':>7s}")
for gap, key, n, us, fl, bound, ms in sorted(rows, key=lambda r: -r[0]):
    print(f"{str(key)[:58]:58s} {n:3d} {us:7.1f} {fl:7.1f} {fl / us:5.2f} {bound:>6s} {gap:7.3f}")
