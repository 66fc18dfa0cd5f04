"""Per-kernel SASS instruction census of libbattgp_b200.so (cuobjdump -sass): which kernels carry tcgen05 (UTCIMMA, LDTM,
UTCBAR...), bulk-async copies (UBLKCP), FP64 tensor (DMMA) instructions.  Output: profiles/sass_rNN_summary.txt."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "battgp_b200/lib/libbattgp_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UBLKPF", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL",
         "LDG", "STG", "LDS", "STS", "LDGSTS", "BAR", "ACQBULK", "ELECT", "CCTL", "MUFU", "REDUX", "ATOM", "RED"]
arch = None
cur = None
counts = collections.OrderedDict()
for ln in txt.splitlines():
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        counts[cur]["arch"] = arch
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(\.[\w.]+)?", ln)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        if op in WATCH:
            counts[cur][op] += 1
        if op == "UBLKCP" and m.group(2) and "MULTICAST" in m.group(2):
            counts[cur]["UBLKCP.MULTICAST"] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n


print(f"# cuobjdump -sass {lib}: {len(counts)} kernels; instruction counts per kernel (static SASS)")
tot = collections.Counter()
for k, c in counts.items():
    name = demangle(k)
    parts = [f"{op}={c[op]}" for op in WATCH + ["UBLKCP.MULTICAST"] if c[op]]
    print(f"{name:70s} {c['arch']} total={c['total']:6d}  " + " ".join(parts))
    for op in WATCH + ["UBLKCP.MULTICAST"]:
        tot[op] += c[op]
print("# library totals: " + " ".join(f"{op}={tot[op]}" for op in WATCH + ["UBLKCP.MULTICAST"] if tot[op]))
