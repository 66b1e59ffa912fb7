"""Opcode histogram of every kernel in libflope_b200.so (cuobjdump -sass), for profiles/.  No GPU needed.

usage: python tools/sass_hist.py [out.txt]
Per kernel: instruction count and the mnemonics that prove what the kernel is made of - UTCHMMA / UTCBAR / LDTM
(tcgen05 MMA, commit, TMEM load), UBLKCP (TMA bulk copy), SYNCS (mbarrier), IDP (DP2A), FFMA.RM - plus the top opcodes."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "flope_b200", "libflope_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "IDP", "FFMA.RM", "HMMA", "IMMA", "LDGSTS", "BAR", "ATOM", "RED", "LDS", "STS", "LDG", "STG")
out = []
for k, h in hist.items():
    base = collections.Counter()
    for op, c in h.items():
        base[op.split(".")[0]] += c
        if op.startswith("FFMA.RM"):
            base["FFMA.RM"] += c
    total = sum(h.values())
    out.append(f"{demangle(k)}\n  instructions {total}")
    out.append("  key:  " + "  ".join(f"{n} {base[n]}" for n in KEY if base[n]))
    top = collections.Counter({op.split('.')[0]: 0 for op in h})
    for op, c in h.items():
        top[op.split(".")[0]] += c
    out.append("  top:  " + "  ".join(f"{n} {c}" for n, c in top.most_common(12)))
text = "\n".join(out) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
else:
    print(text)
