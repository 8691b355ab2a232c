"""Counts of the SASS mnemonics that prove the Blackwell paths, per kernel of libasd_b200.so (read here, no GPU):
  python tools/sass_mnemonics.py > profiles/sass_mnemonics_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "adaptive-speculative-decoding_b200", "libasd_b200.so")
WANT = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "LDSM",
        "ACQBULK", "UCGABAR", "FFMA2", "FADD2", "FMUL2")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
kern, counts, sizes = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("asd::", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        sizes[kern] += 1
        op = m.group(1)
        for w in WANT:
            if op.startswith(w):
                key = op if w in ("UTCHMMA", "HMMA", "UCGABAR") else w
                if w == "SYNCS":
                    key = "SYNCS"
                counts[kern][key] += 1
                break
print("# cuobjdump -sass adaptive-speculative-decoding_b200/libasd_b200.so (sm_100a): counts of the mnemonics that prove the")
print("# Blackwell paths (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; UTMALDG = TMA tensor load; UBLKCP = 1-D bulk TMA;")
print("# LDTM / STTM = tcgen05.ld / tcgen05.st; UTCBAR = tcgen05.commit; UTCATOMSWS = TMEM alloc; SYNCS = mbarrier;")
print("# FFMA2/FADD2/FMUL2 = packed fp32x2; HMMA = legacy mma.sync; ACQBULK = griddepcontrol.wait; UCGABAR = barrier.cluster)")
for k in sorted(counts, key=lambda k: (-sum(v for n, v in counts[k].items() if n.startswith("UTC")), k)):
    print(f"{k:46s} " + "  ".join(f"{n} {v}" for n, v in sorted(counts[k].items())) + f"   [{sizes[k]} instr]")
