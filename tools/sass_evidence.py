"""SASS evidence of the tcgen05 / TMEM / TMA path: per-kernel counts of the Blackwell opcodes in the built library.
usage: python tools/sass_evidence.py > profiles/rN_sass_evidence.txt   (needs cuobjdump, no GPU)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mirror_b200", "libmirror_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "HMMA", "MUFU.EX2"]
cnt = collections.defaultdict(collections.Counter)
samples = collections.defaultdict(dict)
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    for o in OPS:
        if re.search(r"\b" + re.escape(o) + r"(\b|\.)", line):
            cnt[cur][o] += 1
            samples[cur].setdefault(o, line.strip())


def dem(n):
    s = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    s = s.replace("mb::(anonymous namespace)::", "").replace("void ", "")
    return re.sub(r"\((CUtensorMap_st|mb::|float|void|__nv|unsigned|int|long|const).*", "", s)


print("# cuobjdump -sass mirror_b200/libmirror_b200.so (sm_100a): Blackwell opcodes per kernel.")
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG / UTMASTG = TMA tensor load / store,")
print("# UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc, HMMA = mma.sync (the three FIR kernels only)\n")
tot = collections.Counter()
for f, c in sorted(cnt.items(), key=lambda kv: (-kv[1].get("UTCHMMA", 0), -kv[1].get("HMMA", 0))):
    if not any(c.get(o, 0) for o in ("UTCHMMA", "LDTM", "UTMALDG", "HMMA", "STTM")):
        continue
    print(f"{dem(f)[:78]:78s} " + " ".join(f"{o}={c[o]}" for o in OPS if c.get(o)))
    tot.update(c)
print("\nTOTAL " + " ".join(f"{o}={tot[o]}" for o in OPS if tot.get(o)))
print("\n# first occurrence of each opcode in the flash / contrastive / dominant GEMM kernels:")
for f in samples:
    n = dem(f)
    if any(k in n for k in ("flash_fwd", "flash_bwd_kernel<true>", "contrastive_grad", "gemm_tcgen05_kernel<192, 0, 1, true, 2>")):
        print("## " + n)
        for o, l in samples[f].items():
            print("   " + re.sub(r"\s+", " ", l)[:140])
