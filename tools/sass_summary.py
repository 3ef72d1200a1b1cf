"""Per-kernel counts of the Blackwell-native SASS opcodes in the built library (cuobjdump -sass; runs without a GPU).

  python tools/sass_summary.py > profiles/r02_sass_summary.txt
UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG / UTMAPF = TMA load / store / reduce-add / L2 prefetch, LDTM / STTM =
tcgen05.ld / st, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync (must be absent from the product kernels)."""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "arp_b200" / "_lib" / "libarp_b200.so"
OPS = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCBAR", "LDTM", "STTM", "HMMA", "IMMA", "MUFU.EX2", "SYNCS")
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
kernels = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), {"n": 0, **{o: 0 for o in OPS}, "2CTA": 0})
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if not m:
        continue
    op = m.group(1)
    cur["n"] += 1
    for o in OPS:
        if op.startswith(o):
            cur[o] += 1
    if ".2CTA" in op:
        cur["2CTA"] += 1
print(f"# {lib.name}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
print(f"{'kernel':86s} {'instr':>6s} " + " ".join(f"{o:>8s}" for o in OPS) + f" {'.2CTA':>6s}")
for name, c in kernels.items():
    d = re.sub(r"\(.*", "", demangle(name)).replace("void ", "").replace("arp::", "")
    print(f"{d[:86]:86s} {c['n']:6d} " + " ".join(f"{c[o]:8d}" for o in OPS) + f" {c['2CTA']:6d}")
tot = {o: sum(c[o] for c in kernels.values()) for o in OPS}
print("# totals: " + ", ".join(f"{o} {v}" for o, v in tot.items()))
