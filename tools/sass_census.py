#!/usr/bin/env python3
"""Per-kernel SASS census of libnddwt_b200.so (run here, no GPU needed): static instruction counts of the
mnemonics that show what engine a kernel uses -- FFMA2 (packed complex-single arithmetic), UBLKCP / UTMALDG (TMA bulk
and tensor copies), SYNCS (mbarrier), LDS/STS (shared memory), LDG/STG, MUFU (the rsqrt of the fused shrink),
LDL/STL (spills) -- grouped by kernel family; the headline instantiations are listed one by one.
usage: tools/sass_census.py > profiles/r02_sass_census.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "non-decimated_wavelets_b200", "libnddwt_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["FFMA2", "FFMA", "DFMA", "UBLKCP", "UTMALDG", "SYNCS", "LDS", "STS", "LDG", "STG", "RED", "MUFU", "BAR", "LDL", "STL"]
per = {}
name = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("nddwt::", "").split("(")[0].replace("void ", "")
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        per[name]["total"] += 1
        for k in MN:
            if op == k:
                per[name][k] += 1
fam = collections.defaultdict(collections.Counter)
for n, c in per.items():
    f = n.split("<")[0]
    fam[f]["kernels"] += 1
    for k, v in c.items():
        fam[f][k] += v
print("# SASS census of libnddwt_b200.so (sm_100a), static instruction counts\n")
archs = set(re.findall(r"arch = (sm_\w+)", sass))
print("cubins: %d, arch: %s\n" % (len(re.findall(r"Fatbin elf code", sass)), ", ".join(sorted(archs))))
hdr = ["kernel family", "instantiations", "total"] + MN
print("| " + " | ".join(hdr) + " |")
print("|" + "---|" * len(hdr))
for f, c in sorted(fam.items(), key=lambda kv: -kv[1]["total"]):
    print("| " + " | ".join([f, str(c["kernels"]), str(c["total"])] + [str(c[k]) for k in MN]) + " |")
tot = collections.Counter()
for c in fam.values():
    tot.update(c)
print("| " + " | ".join(["ALL", str(tot["kernels"]), str(tot["total"])] + [str(tot[k]) for k in MN]) + " |")
print("\n## Headline instantiations (complex single, db4)\n")
print("| " + " | ".join(["kernel", "total"] + MN) + " |")
print("|" + "---|" * (len(MN) + 2))
for n, c in sorted(per.items()):
    if "float2, 8" in n and any(k in n for k in ("k_dec3_fused", "k_rec3_rows", "k_rec3_bulk", "k_dec_last", "k_rec_last")):
        print("| " + " | ".join(["`%s`" % n, str(c["total"])] + [str(c[k]) for k in MN]) + " |")
