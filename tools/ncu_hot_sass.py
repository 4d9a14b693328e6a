"""Hot SASS instructions of one kernel of an .ncu-rep captured with --import-source on (read here, no GPU):
warp-sample share per instruction with its two main stall reasons, and the share per instruction class.
    python tools/ncu_hot_sass.py gpurun_out/prof.ncu-rep <kernel name substring, e.g. "gemm_tc_kernel<224, 0"> [top_n]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, sel = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
base = re.split(r"[<(]", sel)[0]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{base}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
norm = lambda t: t.replace("(int)", "").replace(" ", "")
# one section per kernel and view: "Kernel Name" row, header row, one row per SASS instruction
starts = [i for i, r in enumerate(rows) if len(r) >= 2 and r[0] == "Kernel Name"]
pick = next((i for i in starts if norm(sel) in norm(rows[i][1])), None)
if pick is None:
    sys.exit("no source page for " + sel + "; kernels: " + "; ".join(sorted({rows[i][1][:80] for i in starts})))
h = rows[pick + 1]
end = next((i for i in starts if i > pick), len(rows))
data = [r for r in rows[pick + 2:end] if len(r) == len(h) and r[h.index("# Samples")].isdigit()]
hdr_at = [pick + 1]
iS, iI, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
total = sum(int(r[iS]) for r in data) or 1
print(f"kernel {rows[hdr_at[0] - 1][1][:110]}\nsamples {total}, SASS instructions {len(data)}")
agg = defaultdict(int)
for r in data:
    for i in stall:
        agg[h[i][6:]] += int(r[i])
print("stall reasons: " + ", ".join(f"{k} {100 * v / total:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))


def klass(k):
    """Class of the instruction a warp is stalled AT; a branch right behind a SYNCS try-wait is the mbarrier spin."""
    op = data[k][iSrc].split()
    op = [t for t in op if not t.startswith("@")]
    m = op[0] if op else ""
    if m.startswith("BRA") and k > 0 and "SYNCS" in data[k - 1][iSrc]:
        return "mbarrier wait (spin)"
    for pat, c in (("SYNCS", "mbarrier wait (spin)"), ("UTC|UTMA|UBLK", "tcgen05 / TMA issue"), ("LDTM|STTM", "tcgen05.ld/st"),
                   ("LDG|LD\\.", "global load"), ("STG|ST\\.|RED|ATOM", "global store"), ("LDS|STS", "shared ld/st"),
                   ("MUFU", "MUFU"), ("BAR|WARPSYNC", "bar / warpsync"), ("F2F|F2I|I2F|PRMT", "convert / permute"),
                   ("FFMA|FMUL|FADD|HFMA|HMUL|HADD", "fp math"), ("BRA|EXIT|CALL|RET|BSSY|BSYNC", "branch")):
        if re.match(pat, m):
            return c
    return "other"


by = defaultdict(int)
for k, r in enumerate(data):
    by[klass(k)] += int(r[iS])
print("samples by instruction class: " + ", ".join(f"{k} {100 * v / total:.1f}%" for k, v in sorted(by.items(), key=lambda kv: -kv[1])))
print(f"\n{'#':>5} {'share':>6} {'executed':>9}  instruction  [stalls]")
for k in sorted(sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:top_n]):
    r = data[k]
    st = sorted(((int(r[i]), h[i][6:]) for i in stall), reverse=True)[:2]
    print(f"{k:5d} {100 * int(r[iS]) / total:5.1f}% {r[iI]:>9}  {r[iSrc].strip()[:70]:70s} " + " ".join(f"{n}:{c}" for c, n in st if c))
