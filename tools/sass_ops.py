"""Counts the SASS opcodes that prove the kernels are hand-written sm_100a code:
TMA bulk copies (UBLKCP) and their mbarriers (SYNCS), warp matches (MATCH), global
atomics / reductions (ATOMG, REDG, ATOM, RED), shared-memory atomics (ATOMS), 128-bit
no-allocate stores (STG.E.NA.128 / STG.E.EF.128), 128-bit loads, cluster barriers
(UCGABAR / CGABAR), and what must NOT appear (HMMA / tensor-core ops).

    python tools/sass_ops.py > profiles/sass_ops.txt

Needs only cuobjdump (no GPU).  One line per kernel of liblyftvoxel_b200.so.
"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lyft-3d-object-detection_b200", "liblyftvoxel_b200.so")
CSRC = os.path.join(ROOT, "lyft-3d-object-detection_b200", "csrc")

GROUPS = [
    ("UBLKCP", re.compile(r"\bUBLKCP")),
    ("SYNCS", re.compile(r"\bSYNCS")),
    ("MATCH", re.compile(r"\bMATCH")),
    ("ATOMG", re.compile(r"\bATOMG|\bATOM\b|\bATOM\.")),
    ("REDG", re.compile(r"\bREDG|\bRED\b|\bRED\.")),
    ("ATOMS", re.compile(r"\bATOMS")),
    ("STG.128", re.compile(r"\bSTG\.[A-Z.]*128")),
    ("STG.NA/EF", re.compile(r"\bSTG\.E\.(NA|EF)")),
    ("LDG.128", re.compile(r"\bLDG\.[A-Z.]*128")),
    ("CGABAR", re.compile(r"CGABAR")),
    ("SHFL", re.compile(r"\bSHFL")),
    ("VOTE", re.compile(r"\bVOTE")),
    ("MMA", re.compile(r"\b(HMMA|IMMA|DMMA|UTCMMA|UTCHMMA|QMMA)")),
]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return [o.strip() for o in out[: len(names)]]


def short(name):
    m = re.match(r"(?:void )?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def main():
    if not os.path.exists(LIB):
        sys.exit("build the library first (python -c 'import __graft_entry__ as g; g.build()')")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for g, rx in GROUPS:
            if rx.search(op):
                kernels[cur][g] += 1
    names = demangle(list(kernels.keys()))
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"sm_\d+a?", arch)))
    print("# cuobjdump -sass %s" % os.path.relpath(LIB, ROOT))
    print("# produced by tools/sass_ops.py; sources sha256[:16] = %s; ELF targets: %s" % (h.hexdigest()[:16], ",".join(archs)))
    cols = [g for g, _ in GROUPS]
    print("%-46s %7s " % ("kernel", "instrs") + " ".join("%9s" % c for c in cols))
    tot = collections.Counter()
    for (mangled, cnt), nm in zip(kernels.items(), names):
        print("%-46s %7d " % (short(nm)[:46], cnt["_total"]) + " ".join("%9d" % cnt[c] for c in cols))
        tot.update(cnt)
    print("%-46s %7d " % ("TOTAL (%d kernels)" % len(kernels), tot["_total"]) + " ".join("%9d" % tot[c] for c in cols))
    assert tot["MMA"] == 0, "tensor-core instructions found on a path that has no contraction"


if __name__ == "__main__":
    main()
