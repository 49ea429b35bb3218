"""SASS evidence for libtmla.so: which kernels contain tcgen05 / TMEM / bulk-TMA / mbarrier / peer-memory instructions.
    python profiles/sass_summary.py > profiles/r2_sass_tcgen05.txt        (no GPU needed: cuobjdump on the built library)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "three-mlagents_b200", "lib", "libtmla.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "LDGSTS", "FFMA2", "MUFU.TANH",
      "DFMA", "REDG", "LDG.STRONG.SYS", "STG.STRONG.SYS", "MEMBAR.SYS")
cur, per, first = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); first[cur] = {}
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line) if cur else None
    if not m:
        continue
    ins = m.group(1).strip()
    op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
    per[cur]["total"] += 1
    for k in MN:
        hit = op == k or op.startswith(k + ".")
        if k.endswith("STRONG.SYS"):                       # LDG.E.128.STRONG.SYS etc.: width sits between the mnemonic parts
            hit = op.startswith(k[:3] + ".") and op.endswith("STRONG.SYS")
        if k == "MEMBAR.SYS":
            hit = op.startswith("MEMBAR.") and op.endswith(".SYS")
        if hit:
            per[cur][k] += 1
            first[cur].setdefault(k, ins)
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for c in per.values():
    tot.update({m: c[m] for m in MN})
print("# SASS evidence for libtmla.so (sm_100a): `cuobjdump -sass three-mlagents_b200/lib/libtmla.so` summarised by profiles/sass_summary.py")
print("# tcgen05.mma -> UTCHMMA | tcgen05.ld / st -> LDTM / STTM | tcgen05.commit -> UTCBAR | cp.async.bulk -> UBLKCP | mbarrier -> SYNCS |")
print("# cp.async -> LDGSTS | ld/st .sys on IPC-mapped peer memory -> LDG/STG.E[.128].STRONG.SYS.  Legacy mma.sync would show as HMMA: none.")
print("TOTAL over the library: " + " ".join(f"{m}={tot[m]}" for m in MN))
print()
for (k, c), name in zip(per.items(), names):
    if not any(c[m] for m in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "LDG.STRONG.SYS", "STG.STRONG.SYS")):
        continue
    print(name)
    print(f"    instructions={c['total']} " + " ".join(f"{m}={c[m]}" for m in MN if c[m]))
    for m in ("UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "LDG.STRONG.SYS", "STG.STRONG.SYS"):
        if m in first[k]:
            print(f"      e.g. {first[k][m]}")
