"""Per-kernel SASS opcode counts of librnb.so — the evidence that the hot path is tcgen05 / TMEM / TMA code
(B200_PROFILING.md "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st,
UTMALDG / UTMASTG / UBLKCP = TMA, HMMA = legacy mma.sync (must be 0).

    python tools/sass_summary.py > profiles/sass_summary_r2.md      (no GPU needed: cuobjdump reads the .so)
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "resnet_c_b200" / "librnb.so"
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMALDG.IM2COL", "UTMASTG", "UBLKCP", "UTCBAR",
       "SYNCS", "HMMA", "FFMA", "IMAD"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        c = kernels[cur]
        c[base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            c["UTCHMMA.2CTA"] += 1
        if base == "UTMALDG" and "IM2COL" in op:
            c["UTMALDG.IM2COL"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle))
    print("# SASS summary of `resnet_c_b200/librnb.so` (round 2)\n")
    print("`cuobjdump -sass` of the in-tree library, counted by `tools/sass_summary.py`. `UTCHMMA` = `tcgen05.mma` "
          "(`.2CTA` = `cta_group::2`), `UTCQMMA` = `tcgen05.mma kind::f8f6f4`, `LDTM` / `STTM` = `tcgen05.ld` / `.st`, "
          "`UTMALDG` / `UTMASTG` / `UBLKCP` = TMA loads / stores / bulk copies, `HMMA` = legacy `mma.sync` (none).\n")
    print("| kernel | " + " | ".join(OPS) + " | instructions |")
    print("|---|" + "---|" * (len(OPS) + 1))
    total = Counter()
    for k, c in kernels.items():
        short = re.sub(r"\(anonymous namespace\)::", "", names[k])
        short = re.sub(r"\(.*", "", short).replace("void ", "").replace("rnb::", "")
        n = sum(v for kk, v in c.items() if "." not in kk)
        print(f"| `{short}` | " + " | ".join(str(c.get(o, 0)) for o in OPS) + f" | {n} |")
        for o in OPS:
            total[o] += c.get(o, 0)
    print(f"| **total ({len(kernels)} kernels)** | " + " | ".join(str(total[o]) for o in OPS) + " | |")
    ldd = subprocess.run(["ldd", str(LIB)], capture_output=True, text=True).stdout
    libs = sorted(set(re.findall(r"^\s*(\S+)", ldd, flags=re.M)))
    print("\n`ldd`: " + ", ".join(f"`{l}`" for l in libs) + " — no cuDNN / cuBLAS / NCCL / torch.")


if __name__ == "__main__":
    sys.exit(main())
