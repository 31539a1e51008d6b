"""profiles/sass_summary_<tag>.txt: per-kernel SASS opcode counts of the shipped library (cuobjdump -sass), the mnemonics
that prove the Blackwell paths (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load,
UTCBAR = tcgen05.commit, SYNCS = mbarrier) next to the legacy ones (HMMA = mma.sync, FFMA).   python tools/sass_summary.py <tag>"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = ROOT / "splice_b200" / "libsplice_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA", "DFMA", "MUFU", "LDS", "STS", "LDG", "STG",
         "ATOM", "RED", "MEMBAR", "CCTL", "BAR"]
kern, counts, total = None, {}, {}
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("splice::", "")
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[kern][w] += 1
                break
out = [f"# SASS opcode summary of splice_b200/libsplice_b200.so ({lib.stat().st_size} bytes, sm_100a), cuobjdump -sass; product build",
       "# columns: kernel | instructions | watched opcodes (count)", ""]
for k in sorted(counts, key=lambda k: -total[k]):
    c = counts[k]
    out.append(f"{k[:100]:100s} {total[k]:7d}  " + " ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
out += ["", "library totals: " + " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w])]
(ROOT / "profiles" / f"sass_summary_{tag}.txt").write_text("\n".join(out) + "\n")
print("\n".join(out[:14]))
print(out[-1])
