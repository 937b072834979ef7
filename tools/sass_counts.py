"""SASS mnemonic counts per kernel of the shipped gabby_b200/libb2l.so -> profiles/r02_sass_counts.txt
usage: python tools/sass_counts.py [LIB] > profiles/r02_sass_counts.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "gabby_b200/libb2l.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
cols = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UBLKPF", "HMMA", "LDSM", "SYNCS", "USETMAXREG", "LDL/STL"]
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
        order.append(cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for c in cols[:-1]:
            if op.startswith(c) and not (c == "HMMA" and op.startswith("UTCHMMA")):
                counts[cur][c] += 1
        if op.startswith("LDL") or op.startswith("STL"):
            counts[cur]["LDL/STL"] += 1
print("SASS mnemonic counts per kernel of the shipped gabby_b200/libb2l.so (tools/sass_counts.py: cuobjdump -sass), round-2 final build.")
print("UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk (TMA bulk),")
print("UBLKPF = cp.async.bulk.prefetch.L2, HMMA = mma.sync, LDSM = ldmatrix, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, LDL/STL = local memory.\n")
print(f"{'kernel':100s}" + "".join(f"{c:>11s}" for c in cols))
for k in order:
    if any(counts[k][c] for c in cols):
        print(f"{k[:100]:100s}" + "".join(f"{counts[k][c]:11d}" for c in cols))
