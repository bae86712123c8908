"""Per-kernel SASS evidence: counts of the Blackwell-native mnemonics (B200_PROFILING.md table) in every kernel of
libivv_b200.so: UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor loads / stores), LDTM / STTM (tcgen05.ld / st),
HMMA (mma.sync, the legacy tensor path), FHADD (mixed fp32+fp16 add), LDGSTS (cp.async).
Usage: python tools/sass_counts.py > profiles/r02_sass_counts.txt   (runs without a GPU: cuobjdump on the built library)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "insv2v_b200", "libivv_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
pats = [("UTC*MMA", r"\bUTC\w*MMA\b"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"),
        ("STTM", r"\bSTTM"), ("HMMA", r"\bHMMA"), ("FHADD", r"\bFHADD"), ("LDGSTS", r"\bLDGSTS"), ("MUFU", r"\bMUFU")]
rows, cur, body = [], None, []


def flush():
    if cur is not None:
        txt = "\n".join(body)
        rows.append((cur, [len(re.findall(p, txt)) for _, p in pats], sum(1 for l in body if re.match(r"\s+/\*[0-9a-f]{4}\*/", l))))


for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        cur, body = m.group(1), []
    else:
        body.append(line)
flush()
tot = [sum(r[1][i] for r in rows) for i in range(len(pats))]
print(f"# {os.path.relpath(lib, ROOT)}: {len(rows)} kernels; totals: " + ", ".join(f"{n} {t}" for (n, _), t in zip(pats, tot)))
print(f"# arch: " + ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))))
print(f"{'kernel':110s} {'instr':>6s} " + " ".join(f"{n:>7s}" for n, _ in pats))
for name, counts, n in sorted(rows, key=lambda r: -r[1][0]):
    d = demangle(name)
    d = re.sub(r"\(.*$", "", d).replace("void ", "").replace("ivv::", "")
    print(f"{d[:110]:110s} {n:6d} " + " ".join(f"{c:7d}" for c in counts))
