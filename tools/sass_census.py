"""Per-kernel census of the tensor-core / TMA instructions in libkit_b200.so (cuobjdump -sass; no GPU needed).
usage: python tools/sass_census.py > profiles/rNN_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "keypoints_interpolation_transformer_b200", "libkit_b200.so")
PAT = [("UTCHMMA", r"\bUTCHMMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
       ("UTMASTG/REDG", r"\bUTMA(STG|REDG)"), ("HMMA", r"\bHMMA"), ("MUFU.EX2", r"MUFU\.EX2"), ("MUFU.TANH", r"MUFU\.TANH")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    stats, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            stats[cur] = collections.Counter()
            continue
        if cur is None or not re.search(r"/\*[0-9a-f]{4,6}\*/", line):
            continue
        stats[cur]["n"] += 1
        for k, p in PAT:
            if re.search(p, line):
                stats[cur][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.splitlines()
    rows = sorted((re.sub(r"\(.*", "", n).replace("void kit::", "").replace("kit::", ""), c) for n, c in zip(names, stats.values()))
    print("# SASS instruction census of libkit_b200.so (cuobjdump -sass, sm_100a), per kernel\n")
    print("Static instruction counts.  `UTCHMMA` = tcgen05.mma, `UTCBAR` = tcgen05.commit, `LDTM` / `STTM` = tcgen05.ld / st, `UTMALDG` /\n"
          "`UTMASTG` / `UTMAREDG` = TMA bulk tensor load / store / reduce-add, `HMMA` = warp-level mma.sync.  Every kernel of the BASELINE\n"
          "configs[1] train step that contracts tensors -- `gemm_tcgen05_kernel`, `ffn_kernel`, `gemm_wgrad_group_kernel`, `attn64_fwd_kernel`,\n"
          "`attn64_bwd_kernel` -- and the long-sequence attention (`attn_fwd_tc_kernel`, `attn_bwd_tc_kernel`) issue tcgen05 instructions and no\n"
          "HMMA; HMMA remains in the fall-back attention kernels only (explicit additive mask tensors, head sizes the tcgen05 kernels do not\n"
          "take).  Kernels without tensor-core or TMA instructions (row kernels, loss, pre-pass, Adam) are not listed.\n")
    print("| kernel | instructions | " + " | ".join(k for k, _ in PAT) + " |")
    print("|---|---|" + "---|" * len(PAT))
    for name, c in rows:
        if not (c["UTCHMMA"] or c["HMMA"] or c["UTMALDG"]):
            continue
        print(f"| `{name[:90]}` | {c['n']} | " + " | ".join(str(c[k]) for k, _ in PAT) + " |")


if __name__ == "__main__":
    sys.exit(main())
