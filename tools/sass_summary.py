"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG / UBLKCP, mbarrier -> SYNCS, cluster barrier -> UCGABAR, multimem.ld_reduce -> LDGMC, multimem.st -> STG.E.128.STRONG.SYS (a
plain system-scope store to the multicast address; the switch replicates it),
legacy mma.sync -> HMMA (must be 0). usage: python tools/sass_summary.py [lib.so] > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "egogen_b200/libegogen_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = collections.OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA"), ("UTCQMMA/other UTC*MMA", r"\bUTC(?!HMMA|BAR|ATOMSWS)[A-Z]*MMA"), ("UTCBAR", r"\bUTCBAR"),
    ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"),
    ("UCGABAR", r"\bUCGABAR"), ("LDGMC (multimem.ld_reduce)", r"MULTIMEM|\bLDGMC|\bSTGMC|\bREDG?MC"),
    ("STG.SYS (multimem.st)", r"\bSTG\.E\.128\.STRONG\.SYS"), ("LDGSTS", r"\bLDGSTS"), ("HMMA (legacy)", r"\bHMMA"),
    ("FFMA", r"\bFFMA"), ("DFMA", r"\bDFMA")])
cur, rows = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    for k, p in pats.items():
        if re.search(p, line):
            rows[cur][k] += 1
            break


def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    r = re.sub(r"\(.*", "", r)
    return r[:100]


print(f"# SASS evidence, {lib} (cuobjdump -sass, sm_100a). Columns = instruction counts per kernel.")
keys = list(pats)
print("kernel | " + " | ".join(keys))
tot = collections.Counter()
for fn, c in rows.items():
    tot.update(c)
    if not any(c[k] for k in keys[:11]):      # list every kernel that touches tensor cores / TMEM / TMA / mbarriers / clusters / multimem
        continue
    print(demangle(fn) + " | " + " | ".join(str(c[k]) for k in keys))
print("TOTAL (all %d kernels) | " % len(rows) + " | ".join(str(tot[k]) for k in keys))
