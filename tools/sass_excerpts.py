"""Regenerates profiles/r02_sass_excerpts.txt from the built library (cuobjdump -sass; no GPU needed).

    python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mfmg_b200", "csrc", "libmfmg_b200.so")

# (title, regex on the mangled function name)
KERNELS = [
    ("csr_tile_kernel<4,Jacobi,int32,GHOST=false> (the A-kernel of cfg1)", r"csr_tile_kernelILi4ELi2EiLb0E"),
    ("csr_tile_kernel<4,Jacobi,int32,GHOST=true> (fused compute + NVLink exchange form)", r"csr_tile_kernelILi4ELi2EiLb1E"),
    ("mf_q1_stencil3_kernel<Jacobi, arithmetic flags, ring 3, in place> (strip form, default)", r"mf_q1_stencil3_kernelILi2ELb1ELi3ELb1E"),
    ("mf_q1_stencil3_kernel<Spmv, arithmetic flags, ring 4>", r"mf_q1_stencil3_kernelILi0ELb1ELi4ELb1E"),
    ("mf_q1_stencil_kernel<Jacobi, arithmetic flags, 2 CTAs/SM> (one node per thread, form 1)", r"mf_q1_stencil_kernelILi2ELb1ELi2E"),
    ("csr_short_kernel<Sub, int32, 4 rows per thread> (prolongation)", r"csr_short_kernelILi3EiLi4E"),
    ("csr_tile_kernel<8,Jacobi,int32,GHOST=false> (Q2 rows: 16 gathers in flight per lane)", r"csr_tile_kernelILi8ELi2EiLb0E"),
    ("gemv_stream_kernel", r"gemv_stream_kernel"),
    ("halo_push_kernel", r"halo_push_kernel"),
    ("halo_wait_kernel", r"halo_wait_kernel"),
    ("peer_allreduce_kernel", r"peer_allreduce_kernel"),
    ("dd_rhs_allreduce_kernel<int32>", r"dd_rhs_allreduce_kernelIiE"),
    ("dd_w_rhs_allreduce_kernel<256> (W b_I + separator rhs + all-reduce, last CTA reduces)", r"dd_w_rhs_allreduce_kernelILi256E"),
]
INTERESTING = re.compile(r"\b(UBLKCP|UTMALDG|SYNCS|LDGSTS|MEMBAR|ERRBAR|NANOSLEEP|ATOMG|RED|CCTL)\b|\.SYS\b")
COUNTED = ("UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDGDEPBAR", "MEMBAR", "ERRBAR", "NANOSLEEP", "ATOMG", "RED", "CCTL",
           "BAR", "SHFL", "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "STS", "LD", "ST", "CS2R")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    funcs, name, body = {}, None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = body
            name, body = m.group(1), []
        elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
            body.append(line)
    if name:
        funcs[name] = body
    print("# SASS evidence (cuobjdump -sass mfmg_b200/csrc/libmfmg_b200.so, sm_100a), round 2 HEAD; tools/sass_excerpts.py")
    print("# per kernel: instruction count, histogram of the mnemonics that matter, and the first lines that carry them")
    print("# UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier, LDGSTS = cp.async, *.SYS = system-scope (peer) accesses,")
    print("# MEMBAR.SC.SYS = __threadfence_system, SHFL = warp shuffles, DFMA/DADD/DMUL = FP64 pipe")
    for title, pat in KERNELS:
        hits = [f for f in funcs if re.search(pat, f)]
        print(f"\n## {title}")
        if not hits:
            print("   (not found in this build)")
            continue
        f = sorted(hits, key=len)[0]
        body = funcs[f]
        hist = collections.Counter()
        for line in body:
            ins = line.split("*/", 1)[1].strip()
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            mn = ins.split()[0].rstrip(";")
            base = mn.split(".")[0]
            if base in COUNTED:
                hist[base] += 1
            if ".SYS" in mn:
                hist["*.SYS"] += 1
        print(f"   Function : {f}")
        print("   instructions: %d   %s" % (len(body), "  ".join(f"{k}={v}" for k, v in sorted(hist.items()))))
        shown = collections.Counter()
        for line in body:
            m = INTERESTING.search(line)
            if m:
                key = re.sub(r"^@!?U?P\d+\s+", "", line.split("*/", 1)[1].strip()).split()[0].rstrip(";")
                if shown[key] < 2:   # at most two lines per distinct mnemonic
                    print("   " + line.strip())
                shown[key] += 1


if __name__ == "__main__":
    sys.exit(main())
