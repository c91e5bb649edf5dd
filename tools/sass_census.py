"""SASS census of libtinyaudio_b200.so: per kernel, how many tcgen05 / TMEM / TMA / legacy-MMA instructions it carries
(`cuobjdump -sass`; the mnemonics are the ones /opt/skills/guides/B200_PROFILING.md lists as evidence of a Blackwell-native kernel).
usage: python tools/sass_census.py [lib.so] > profiles/sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tiny_audio_b200", "libtinyaudio_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = {}
PAT = {"UTCHMMA": r"\bUTCHMMA", "UTCHMMA.2CTA": r"UTCHMMA\.2CTA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
       "UTMAREDG": r"\bUTMAREDG", "UTMAPF": r"\bUTMAPF", "HMMA": r"\bHMMA", "MUFU": r"\bMUFU", "LDGSTS": r"\bLDGSTS", "SYNCS": r"\bSYNCS"}
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        counts[cur]["instructions"] = 0
        continue
    if cur is None:
        continue
    if re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
        counts[cur]["instructions"] += 1
        for k, p in PAT.items():
            if re.search(p, line):
                counts[cur][k] += 1
names = list(counts)
try:
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
except Exception:
    pass
cols = list(PAT)
tot = collections.Counter()
print(f"# {os.path.basename(lib)}: {len(names)} kernels (sm_100a)")
print(f"{'instr':>7s} " + " ".join(f"{c:>12s}" for c in cols) + "  kernel")
for n in names:
    c = counts[n]
    tot.update(c)
    short = re.sub(r"\((int|bool|unsigned int)\)", "", demangle.get(n, n).replace("<unnamed>::", "")).split("(")[0].replace("void ", "")
    print(f"{c['instructions']:7d} " + " ".join(f"{c[k]:12d}" for k in cols) + f"  {short[:110]}")
print(f"{tot['instructions']:7d} " + " ".join(f"{tot[k]:12d}" for k in cols) + "  TOTAL")
