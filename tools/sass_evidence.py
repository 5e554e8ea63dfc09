"""Dev tool: SASS mnemonics that prove the tcgen05 / TMEM / TMA paths, per kernel of libtradeaid_b200.so (cuobjdump -sass).
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores, UTCBAR = tcgen05.commit,
UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier ops (B200_PROFILING.md, "What proves a Blackwell-native kernel").
usage: sass_evidence.py [path/to/lib.so]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "trade-aid-multimodal-transformer_b200", "libtradeaid_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]MMA|UTCBAR|UTCCP|UTCATOMSWS|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|SYNCS|HMMA|HGMMA)\b")
cur, cnt, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        total[cur] += 1
        mm = pat.search(line)
        if mm:
            cnt[cur][mm.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(total), capture_output=True, text=True).stdout.splitlines()
cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA"]
print(f"{'kernel':78s} {'SASS':>6s} " + " ".join(f"{c:>8s}" for c in cols))
for f, name in sorted(zip(total, names), key=lambda x: x[1]):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    print(f"{name[:78]:78s} {total[f]:6d} " + " ".join(f"{cnt[f].get(c, 0):8d}" for c in cols))
