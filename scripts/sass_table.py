"""Mnemonic counts per kernel of the built library (`cuobjdump -sass`): the evidence that the contractions run on
tcgen05 / TMEM / TMA and that every kernel carries the programmatic-dependent-launch pair. Writes profiles/sass_r02.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "open_knowledge_graph_embeddings_b200", "csrc", "libokge_b200.so")
WATCH = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "LDTM", "SYNCS", "ELECT", "MUFU.EX2", "MUFU.RCP",
         "MUFU.LG2", "MUFU.SQRT", "MUFU.RSQ", "F2FP", "REDG", "ATOMG", "ATOMS", "HMMA", "IMMA", "ACQBULK", "PREEXIT")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", re.sub(r"okge::\(anonymous namespace\)::|void ", "", n)).replace("(bool)", "").replace("(int)", "")
            for n in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["instr"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    names = demangle(list(kernels))
    lines = ["# SASS evidence, round 2 (final build): `python scripts/sass_table.py` = `cuobjdump -sass csrc/libokge_b200.so` (sm_100a), mnemonic",
             "# counts per kernel. UTCHMMA = tcgen05.mma (kind::f16 / kind::tf32), UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld",
             "# (TMEM -> registers), UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier ops, ELECT = elect.sync,",
             "# MUFU.* = special-function unit, F2FP = fp32 -> fp16x2 pack, REDG / ATOMG = global reductions / atomics, ACQBULK / PREEXIT =",
             "# griddepcontrol.wait / launch_dependents (programmatic dependent launch: every kernel has the pair).",
             "# Template arguments of okge_gemm_tc_kernel: <F16 operands, MODE (0 store, 1 BCE, 2 LSE, 3 softmax grad, 4 rank, 5 Adagrad,",
             "# 6 Adagrad with the deep operand ring), LIMIT (BCE: device-side column limit; Adagrad: dropout mask on the gradient), RANK slots>.",
             "# No HMMA / IMMA (mma.sync) in the library: every contraction goes through tcgen05.", ""]
    tot = collections.Counter()
    for name, c in zip(names, kernels.values()):
        tot.update(c)
        rest = ", ".join(f"{w} {c[w]}" for w in WATCH if c[w])
        lines.append(f"{name:72s} instr {c['instr']:5d}  {rest}")
    lines += ["", "library totals: " + ", ".join(f"{w} {tot[w]}" for w in WATCH)]
    out = os.path.join(ROOT, "profiles", "sass_r02.txt")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"{len(kernels)} kernels -> {out}")
    assert tot["HMMA"] == 0 and tot["IMMA"] == 0 and tot["UTCHMMA"] > 0, "the contractions must run on tcgen05"
    assert tot["ACQBULK"] >= len(kernels), "every kernel starts with griddepcontrol.wait"


if __name__ == "__main__":
    sys.exit(main())
