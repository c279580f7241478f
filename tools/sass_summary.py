#!/usr/bin/env python
"""Trimmed SASS evidence per kernel of libjolineedle_b200.so (runs without a GPU):

    python tools/sass_summary.py > profiles/r02/sass_summary.txt

For every kernel: registers / shared memory / spills as ptxas reported them (csrc/build.log, `-Xptxas -v`) and the
count of the instructions that prove which hardware path the kernel takes -- TMA tensor tiles (UTMALDG), TMA bulk
copies (UBLKCP), mbarrier traffic (SYNCS), warp reductions (REDUX), vector stores (STG.E.128 / .64), reductions to
memory (RED / REDG), votes, shuffles, the programmatic-dependent-launch pair (ACQBULK / griddepcontrol shows up as
`ACQBULK`+`PREEXIT` in SASS) -- from `cuobjdump -sass`.  Stamped with the source hash compiled into the library.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "jolineedle_b200", "libjolineedle_b200.so")
LOG = os.path.join(ROOT, "jolineedle_b200", "csrc", "build.log")
MNEMONICS = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "REDUX", "STG.E.128", "STG.E.64", "LDG.E.128", "REDG", "VOTE",
             "SHFL", "POPC", "PRMT", "ACQBULK", "PREEXIT", "UTC", "HMMA", "FFMA", "DADD", "DMUL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    from jolineedle_b200 import buildinfo

    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels, current = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = m.group(1)
            kernels[current] = []
            continue
        if current and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            kernels[current].append(line)
    res = {}
    if os.path.exists(LOG):
        text = open(LOG).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes "
                             r"spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?"
                             r"(?:, (\d+) bytes smem)?", text):
            res[m.group(1)] = {"stack": int(m.group(2)), "spill": int(m.group(3)) + int(m.group(4)),
                               "regs": int(m.group(5)), "smem": int(m.group(6) or 0)}
    names = demangle(list(kernels))
    print(f"# SASS summary of libjolineedle_b200.so -- source hash {buildinfo.library_source_hash()}, arch {', '.join(arch)}")
    print("# kernel | registers | static smem | spills | instructions | " + " ".join(MNEMONICS) + " (only non-zero shown)")
    for k in sorted(kernels, key=lambda n: names[n]):
        body = kernels[k]
        counts = {m: sum(1 for l in body if re.search(r"\b" + re.escape(m), l)) for m in MNEMONICS}
        r = res.get(k, {})
        shown = " ".join(f"{m}={c}" for m, c in counts.items() if c)
        short = names[k].replace("jnk::", "").split("(")[0].replace("void ", "")
        print(f"{short:58s} regs={r.get('regs', '?'):>3} smem={r.get('smem', '?'):>5} spill={r.get('spill', '?')} "
              f"insts={len(body):>5}  {shown}")


if __name__ == "__main__":
    main()
