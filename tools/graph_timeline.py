"""In-graph device timeline of one UNet3D forward: what each kernel costs INSIDE the captured CUDA graph (warm L2,
programmatic dependent launch overlapping prologues), which is what a DDIM step pays - unlike ncu (serialised, cold
cache) or the event-per-call profile of tools/profile_ops.py (launch gaps included).

Method: (1) one eager forward with insv2v_b200.ops.Prof on, recording for every C-ABI call its op key and the range of
launch indices it produced; (2) torch.profiler (CUPTI kernel activity records) around replays of the captured graph;
the i-th ivv:: kernel of a replay is the i-th launch of the plan. Prints per-(op, shape) in-graph time, the span of the
forward, and the idle / overlap time between consecutive kernels.  Usage: python tools/graph_timeline.py [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from insv2v_b200 import lib, ops  # noqa: E402

dev = torch.device("cuda")
unet, _ = bench.build_models(dev)
f, h, w = [int(v) for v in os.environ.get("SHAPE", "16,32,48").split(",")]
x = torch.randn(3, 8, f, h, w, device=dev)
ctx = torch.randn(3, 77, 768, device=dev)
t = torch.full((3,), 981.0, device=dev)

# ---- (1) launch index -> op key, from one eager pass ----
unet.use_cuda_graph = False
unet(x, t, encoder_hidden_states=ctx)
torch.cuda.synchronize()
keys = {}
orig_begin, orig_end = ops.Prof.begin, ops.Prof.end
marks = []


def begin():
    marks.append(lib.LAUNCH_COUNT)
    return True


def end(e0, key, flops=0.0, nbytes=0.0):
    if e0 is None:
        return
    n0 = marks.pop()
    n1 = lib.LAUNCH_COUNT + (2 if key[0] == "groupnorm" else 1)  # Prof.end runs before the wrapper counts its launches
    for i in range(n0, n1):
        keys[i] = (key, flops, nbytes)


ops.Prof.begin, ops.Prof.end = staticmethod(begin), staticmethod(end)
base = lib.LAUNCH_COUNT
unet(x, t, encoder_hidden_states=ctx)
n_eager = lib.LAUNCH_COUNT - base
torch.cuda.synchronize()
ops.Prof.begin, ops.Prof.end = orig_begin, orig_end
keys = {i - base: v for i, v in keys.items()}

# ---- (2) CUPTI timeline of graph replays ----
unet.use_cuda_graph = True
for _ in range(3):
    unet(x, t, encoder_hidden_states=ctx)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
REPS = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(REPS):
        unet(x, t, encoder_hidden_states=ctx)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if "ivv::" in e.name and e.time_range is not None]
evs.sort(key=lambda e: e.time_range.start)
print(f"eager plan: {n_eager} launches; profiler saw {len(evs)} ivv kernels over {REPS} replays")
N_CTX_KV = 16  # the context K/V GEMMs leave the graph (projected once per context)
if len(evs) < REPS * (n_eager - N_CTX_KV):
    print("CUPTI did not report the graph's kernels individually; falling back to the eager timeline")
    unet.use_cuda_graph = False
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(REPS):
            unet(x, t, encoder_hidden_states=ctx)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if "ivv::" in e.name and e.time_range is not None]
    evs.sort(key=lambda e: e.time_range.start)
per = len(evs) // REPS
last = evs[-per:]
# the context K/V GEMMs run outside the graph only when the context changes: a replay has n_eager - 16 kernels
off = n_eager - per
span = last[-1].time_range.end - last[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in last)
gap = sum(max(0.0, b.time_range.start - a.time_range.end) for a, b in zip(last, last[1:]))
ovl = sum(max(0.0, a.time_range.end - b.time_range.start) for a, b in zip(last, last[1:]))
print(f"one replay: {per} kernels, span {span / 1e3:.3f} ms, sum of kernel durations {busy / 1e3:.3f} ms, idle between "
      f"kernels {gap / 1e3:.3f} ms, overlap (PDL) {ovl / 1e3:.3f} ms")
# With programmatic dependent launch kernel i+1 starts while kernel i is still running and waits in griddepcontrol.wait,
# so its recorded duration includes that wait. What a kernel ADDS to the step is end_i - max(end_{i-1}, start_i).
agg = {}
for r in range(REPS):
    seq = evs[r * per:(r + 1) * per]
    prev_end = None
    for i, e in enumerate(seq):
        cost = e.time_range.end - (e.time_range.start if prev_end is None else max(prev_end, e.time_range.start))
        prev_end = e.time_range.end if prev_end is None else max(prev_end, e.time_range.end)
        k = keys.get(i + off)
        name = e.name.split("(")[0].replace("void ivv::", "")[:70]
        key = (str(k[0]) if k else name)
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, name])
        a[0] += 1
        a[1] += cost
        if k:
            nk = 2 if k[0][0] == "groupnorm" else 1
            a[2] += k[1] / nk
            a[3] += k[2] / nk
pk = bench.peaks()
rows = []
for key, (n, us, fl, nb, name) in agg.items():
    n1, us1 = n / REPS, us / REPS
    floor = max(fl / REPS / (pk["tf_sustained"] * 1e12), nb / REPS / (pk["hbm"] * 1e9)) * 1e6
    rows.append((us1, key, n1, floor, name))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"{'op / shape (in-graph)':66s} {'n':>4s} {'ms':>7s} {'avg us':>7s} {'floor':>7s} {'frac':>5s} {'share':>6s}")
for us1, key, n1, floor, name in rows[:70]:
    print(f"{key[:66]:66s} {n1:4.0f} {us1 / 1e3:7.3f} {us1 / n1:7.1f} {floor / n1:7.1f} {floor / us1 if us1 else 0:5.2f} "
          f"{100 * us1 / tot:5.1f}%")
if len(sys.argv) > 1:
    json.dump({"span_ms": span / 1e3, "busy_ms": busy / 1e3, "idle_ms": gap / 1e3, "overlap_ms": ovl / 1e3,
               "kernels": per, "rows": [dict(key=k, n=n, us=u, floor_us=fl, kernel=nm) for u, k, n, fl, nm in rows]},
              open(sys.argv[1], "w"), indent=1)
