"""Per-kernel SASS opcode evidence of the built library (no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_opcodes.txt
Counts, per device function of libunigeo_b200.so, the mnemonics that prove the Blackwell-native paths
(B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads /
stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync, LDGSTS = cp.async, MUFU.EX2 / MUFU.TANH
(softmax exponentials / the one-MUFU GELU of the GEMM epilogues), FFMA2 = packed fp32 math."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "unigeo_b200", "libunigeo_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "MUFU.EX2",
       "MUFU.TANH", "FFMA2", "UBLKCP"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = set(re.findall(r"arch = (sm_\w+)", out))
fn, counts, size = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = fn.replace("(anonymous namespace)::", "").replace("ug::", "")
        fn = re.sub(r"^void ", "", re.sub(r"\(.*", "", fn)) or m.group(1)
        counts[fn] = collections.Counter()
        size[fn] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if fn and m:
        size[fn] += 1
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[fn][o] += 1
print(f"# {os.path.basename(lib)}: cubin architectures {sorted(arch)}; instructions per device function (cuobjdump -sass)")
print(f"# {'function':58s} {'instrs':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
for fn, c in counts.items():
    if not any(c.values()) and size[fn] < 400:
        continue
    print(f"{fn[:60]:60s} {size[fn]:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
