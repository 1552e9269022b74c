#!/usr/bin/env python
"""Static evidence that needs no GPU: per kernel of lbm_b200/libblbm.so, the register count / spill bytes /
shared memory that ptxas settled on (cuobjdump -res-usage) and the count of the memory and shuffle mnemonics
in the SASS (cuobjdump -sass) — 128-bit global accesses (LDG.E.128 / STG.E.128), cp.async staging (LDGSTS),
TMA (UTMALDG / UTMASTG), mbarrier waits (SYNCS), warp shuffles (SHFL), local-memory traffic (LDL / STL =
spills), packed fp32 adds (FADD2) and the forbidden packed multiply-adds (FFMA2 / FMUL2).

    python profiles/sass_summary.py [--all] > profiles/r1/sass_summary.txt

Without --all only the instantiations the default launch paths use are listed (vec4 block shape 4 rows).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lbm_b200", "libblbm.so")
CUOBJDUMP = "/usr/local/cuda/bin/cuobjdump"
FILT = "/usr/local/cuda/bin/cu++filt"

MNEMONICS = ("LDG.E.128", "LDG.E.64", "LDG.E", "LDG.E.U8", "LDG.E.U16", "STG.E.128", "STG.E.64", "STG.E",
             "LDGSTS", "UTMALDG", "UTMASTG", "SYNCS", "LDS", "STS", "SHFL", "LDL", "STL", "FADD2", "FFMA2", "FMUL2",
             "FFMA", "MUFU.RCP", "RED", "ATOM")


def demangle(names):
    out = subprocess.run([FILT] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def resource_usage():
    txt = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True).stdout
    res = {}
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            d = dict(re.findall(r"(\w+):(\d+)", line))
            res[cur] = d
            cur = None
    return res


def sass_counts():
    txt = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    ninstr = collections.Counter()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not (cur and m):
            continue
        op = m.group(1)
        ninstr[cur] += 1
        for key in MNEMONICS:
            if key in ("LDG.E", "STG.E"):
                # plain 32-bit access: the mnemonic without a width suffix (cache hints may follow)
                if re.fullmatch(key.replace(".", r"\.") + r"(\.(?!128|64|U8|U16|S8|S16)[A-Z0-9_]+)*", op):
                    counts[cur][key] += 1
            elif op == key or op.startswith(key + "."):
                counts[cur][key] += 1
    return counts, ninstr


def short(name):
    name = re.sub(r"blbmk::", "", name)
    name = re.sub(r">\((?!.*>\().*$", ">", name)  # drop the parameter list, keep the template arguments
    if "<" not in name:
        name = re.sub(r"\(.*\)$", "", name)
    name = name.replace("(StepMode)", "mode ").replace("unsigned int", "u32").replace("unsigned long", "u64")
    name = re.sub(r"^void ", "", name)
    return name


def main():
    show_all = "--all" in sys.argv
    res = resource_usage()
    counts, ninstr = sass_counts()
    names = demangle(sorted(res))
    print("# ptxas resource usage and SASS mnemonic counts per kernel of lbm_b200/libblbm.so (sm_100a)")
    print("# produced by profiles/sass_summary.py from cuobjdump -res-usage / -sass; no GPU involved")
    print("# step_vec4_kernel<MOMENTS, BLOCK_ROWS, FLAVOUR, PACKED, offset type>: FLAVOUR 0 sparse fix-up, 1 dense,")
    print("#   2 dense + cp.async staging (LDGSTS), 3 as 2 without the chunk-flag test; offset u32 / u64;\n#   STACK(B) > 0 = register spills (LDL/STL)")
    print()
    print(f"{'kernel':<58} {'REG':>4} {'STACK(B)':>8} {'SHARED':>7} {'INSTR':>6}  mnemonic=count (non-zero only)")
    for mangled in sorted(res, key=lambda n: names[n]):
        nm = short(names[mangled])
        if not show_all and nm.startswith("step_vec4_kernel") and not re.search(r"<\(bool\)[01], \(int\)4,", nm):
            continue
        r = res[mangled]
        c = counts[mangled]
        print(f"{nm:<58} {r.get('REG', '?'):>4} {r.get('STACK', '0'):>8} {r.get('SHARED', '0'):>7} "
              f"{ninstr[mangled]:>6}  " + " ".join(f"{k}={c[k]}" for k in MNEMONICS if c.get(k)))
    bad = [short(names[m]) for m in res if counts[m].get("FFMA2") or counts[m].get("FMUL2")]
    print()
    print("packed multiply / multiply-add anywhere (must be none, parity contract):", bad or "none")


if __name__ == "__main__":
    main()
