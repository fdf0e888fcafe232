"""Per-kernel SASS evidence: instruction count and the mnemonics that prove what a kernel uses
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA / bulk copies -> UTMALDG / UBLKCP,
mbarrier -> SYNCS, cluster barrier -> UCGABAR_*, 3-input min/max -> FMNMX3, warp match -> MATCH).
Runs without a GPU.  Usage: python tools/sass_summary.py [lib.so] > profiles/sass_rNN.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "3dvlp_b200", "lib3dvlp_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
PAT = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|SYNCS|FMNMX3|REDUX|MATCH|"
                 r"LDGSTS|ATOMS|ATOMG|RED|UCGABAR_ARV|UCGABAR_WAIT|LDS\.128|LDG\.E\.128|STG\.E\.128|"
                 r"LD\.E\.128|ST\.E\.128|STL|LDL)\b")
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        per[cur]["_n"] += 1
        per[cur].update(PAT.findall(line))
names = list(per)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass %s: %d kernels (sm_100a); columns: kernel, SASS instructions, marker mnemonics"
      % (os.path.relpath(so, ROOT), len(names)))
print("# STL/LDL = local-memory spills; none listed means none")
for n, d in zip(names, dem):
    d = d.replace("(anonymous namespace)::", "").replace("void ", "").replace("pn2::", "")
    d = re.sub(r"\((?!.*<).*$", "", d)          # drop the argument list, keep template arguments
    c = per[n]
    print("%-64s %6d  %s" % (d[:64], c["_n"], " ".join("%s:%d" % kv for kv in sorted(c.items()) if kv[0] != "_n")))
