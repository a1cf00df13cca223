"""Per-kernel counts of the Blackwell-native SASS instructions in libb200rec.so (no GPU needed):
UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), UTMALDG (TMA tensor loads), LDTM / STTM (tcgen05.ld / .st), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), plus registers per thread.   python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "real-time-recommendation-system-with-feature-store_b200", "libb200rec.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", res):
    regs[m.group(1)] = int(m.group(2))
pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "STTM", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT", "REDUX", "ATOMG", "RED"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur is None: continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m: continue
    op = m.group(1); counts[cur]["_total"] += 1
    for p in pats:
        if op.startswith(p):
            counts[cur][p] += 1
            break
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
print("libb200rec.so — SASS evidence per kernel (cuobjdump -sass, sm_100a)\n")
print(f"{'kernel':90s} {'regs':>5s} {'instr':>6s} " + " ".join(f"{p:>12s}" for p in pats[:7]))
tot = collections.Counter()
for fn, c in counts.items():
    if not any(c[p] for p in pats[:6]): continue
    name = demangle(fn)[:90]
    print(f"{name:90s} {regs.get(fn, 0):5d} {c['_total']:6d} " + " ".join(f"{c[p]:12d}" for p in pats[:7]))
    tot.update({p: c[p] for p in pats})
print("\nlibrary totals: " + ", ".join(f"{p} {tot[p]}" for p in pats if tot[p]))
print(f"kernels in the library: {len(counts)}; kernels using tcgen05 / TMA / TMEM: {sum(1 for c in counts.values() if any(c[p] for p in pats[:6]))}")
