"""SASS opcode census per kernel of libdeepj_sm100.so (evidence that the hot kernels are tcgen05 / TMA / TMEM):
  python tools/sass_census.py > profiles/r02_sass_census.md
Counts the mnemonics B200_PROFILING.md names: UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG
(TMA), UTCBAR (tcgen05.commit), FFMA2 (packed fp32 FMA), HMMA (legacy tensor path: must be absent)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "music-generator_b200", "libdeepj_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = {"UTC*MMA": r"\bUTC[A-Z]*MMA\b", "LDTM": r"\bLDTM\b", "STTM": r"\bSTTM\b", "UTMALDG": r"\bUTMALDG\b",
       "UTMASTG": r"\bUTMASTG\b", "UTCBAR": r"\bUTCBAR\b", "multicast": r"MULTICAST", "FFMA2": r"\bFFMA2\b",
       "HMMA": r"\bHMMA\b", "RED/ATOM": r"\b(RED|ATOMG?)\b", "LDL/STL": r"\b(LDL|STL)\b"}
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k, p in pat.items():
        if re.search(p, line):
            counts[cur][k] += 1
    if re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
        counts[cur]["instr"] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode census of libdeepj_sm100.so (cuobjdump -sass, sm_100a)\n")
print(f"{len(counts)} kernels.  Columns: static instruction counts per kernel.\n")
cols = list(pat)
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
tot = collections.Counter()
for (mangled, c), name in sorted(zip(counts.items(), names), key=lambda kv: -kv[0][1]["instr"]):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    print(f"| `{name[:80]}` | {c['instr']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
    tot.update(c)
print(f"| **total** | {tot['instr']} | " + " | ".join(str(tot[k]) for k in cols) + " |")
