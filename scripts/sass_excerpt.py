"""Write profiles/r02_sass_excerpt.md: which Blackwell-native instructions each kernel of libss_b200.so
contains (cuobjdump -sass; no GPU needed)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "smartstartcontinuous_b200", "libss_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
pat = re.compile(r"\b(UTC[A-Z0-9]*MMA(?:\.[A-Z0-9_]+)*|LDTM(?:\.[A-Za-z0-9_]+)*|STTM(?:\.[A-Za-z0-9_]+)*|UBLKCP(?:\.[A-Z0-9_]+)*|"
                 r"UTCBAR(?:\.[A-Z0-9_]+)*|UTCATOM[A-Z0-9_.]*|SYNCS\.ARRIVE\.TRANS64|HMMA[A-Z0-9_.]*|FFMA2|MUFU\.EX2)")
out = ["# SASS evidence (round 2): `cuobjdump -sass smartstartcontinuous_b200/libss_b200.so`, sm_100a",
       "",
       "PTX names never appear in SASS: `tcgen05.mma` -> `UTCHMMA` (`.2CTA` = cta_group::2), `tcgen05.ld` / `tcgen05.st` ->",
       "`LDTM` / `STTM`, `tcgen05.commit` -> `UTCBAR` (`.2CTA.MULTICAST` = the pair commit), `tcgen05.alloc` -> `UTCATOMSWS`,",
       "`cp.async.bulk` (TMA bulk copy) -> `UBLKCP.S.G`, `mbarrier.arrive.expect_tx` -> `SYNCS.ARRIVE.TRANS64`.  Kernels without",
       "any of these (FP32 SIMT rollout, scoring tail, trainer, geometry, value net) are omitted.  Regenerate with",
       "`python scripts/sass_excerpt.py`.", "",
       "| kernel | instructions (count) |", "|---|---|"]
excerpt = None
for f in funcs[1:]:
    mangled = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    m = re.search(r"(\w+_kernel(?:<[^>]*>)?)", dem)
    name = m.group(1) if m else dem[:60]
    c = collections.Counter(x.group(1) for x in pat.finditer(f))
    if not any(k.startswith(("UTC", "LDTM", "STTM", "UBLKCP")) for k in c):
        continue
    out.append("| `%s` | %s |" % (name, ", ".join("`%s` (%d)" % kv for kv in sorted(c.items()))))
    if excerpt is None and "mpc_rollout_tc_kernel<4, 3, 16>" in name:
        lines = [l for l in f.split("\n") if re.search(r"UTC[A-Z0-9]*MMA|LDTM|STTM|UBLKCP|UTCBAR|UTCATOM", l)]
        excerpt = (name, lines[:16])
out.append("")
if excerpt:
    out += ["First tcgen05 / TMA instructions of `%s`:" % excerpt[0], "```"]
    out += [re.sub(r"\s*/\*[0-9a-fx]+\*/\s*$", "", l.strip())[:160] for l in excerpt[1]]
    out.append("```")
out += ["", "Legacy tensor path (`HMMA` = mma.sync / wmma) occurrences in the whole library: **%d**." % len(re.findall(r"\bHMMA", txt))]
path = os.path.join(ROOT, "profiles", "r02_sass_excerpt.md")
with open(path, "w") as fh:
    fh.write("\n".join(out) + "\n")
print(path)
