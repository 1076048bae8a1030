"""SASS mnemonic counts per kernel of the shipped library (cuobjdump -sass; no GPU needed).

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt

Evidence that the contraction is tcgen05 / TMEM / TMA code and which kernels spill.
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "maximizingacquisitionfunctions_b200", "libmaf_b200.so")
CLASSES = [("UTCIMMA.2CTA", r"\bUTCIMMA\.2CTA"), ("UTCIMMA (1 CTA)", r"\bUTCIMMA(?!\.2CTA)"), ("UTCBAR", r"\bUTCBAR"),
           ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMAPF", r"\bUTMAPF"), ("DMMA", r"\bDMMA"), ("DFMA", r"\bDFMA"),
           ("SYNCS", r"\bSYNCS"), ("LDL/STL (spills)", r"\b(LDL|STL)\b"), ("ATOM/RED", r"\b(ATOM|ATOMG|RED)\b")]
SHOWN = ("ozaki_i8_gemm_v4_kernel<5, 6, 1>", "ozaki_i8_gemm_v4_kernel<6, 7, 1>", "ozaki_i8_gemm_v4_kernel<7, 8, 1>",
         "cross_fwd_planes_kernel<6, 1, 6, 3>", "cross_fwd_planes_kernel<6, 1, 6, 1>", "cross_fwd_planes_kernel<6, 1, 7, 2>",
         "cross_bwd_g_kernel<6, 1, 2, false>", "cross_bwd_g_kernel<6, 1, 2, true>", "cross_bwd_kernel<6, 1, false>",
         "posterior_reg_kernel<8, false>", "vbar_planes_kernel<8, false>", "mc_acq_kernel<8>", "gemm_nt_f64_kernel<64, 64, 16, 2, 2, 2, 4>",
         "chol_diag_kernel", "lbfgs_step_kernel")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    per, total, cur, k = {}, collections.Counter(), None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[k]).strip()
            k += 1
            per[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        for label, pat in CLASSES:
            if re.search(pat, line):
                per[cur][label] += 1
                total[label] += 1
    print("SASS mnemonic counts per kernel, cuobjdump -sass of the shipped libmaf_b200.so (sm_100a); scripts/sass_summary.py.")
    print("UTCIMMA = tcgen05.mma (kind::i8), .2CTA = cta_group::2; LDTM = tcgen05.ld (TMEM -> registers); UTMALDG = TMA tensor load;")
    print("UTCBAR = tcgen05.commit -> mbarrier; DMMA = FP64 tensor-core MMA (DMMA.8x8x4); SYNCS = mbarrier ops.\n")
    print("whole library (%d kernels): " % len(per) + ", ".join("%s %d" % (l, total[l]) for l, _ in CLASSES) + "\n")
    for want in SHOWN:
        for name, c in sorted(per.items()):
            if want in name:
                print("%-60s %s" % (name[:60], ", ".join("%s %d" % (l, c[l]) for l, _ in CLASSES if c[l])))


if __name__ == "__main__":
    sys.exit(main())
