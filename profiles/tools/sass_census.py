"""SASS census of the hot kernels: python profiles/tools/sass_census.py [out.txt]
Runs `cuobjdump -sass eegan_b200/libeegan_b200.so` and counts, per kernel, the mnemonics that prove the Blackwell paths:
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UTMALDG (cp.async.bulk.tensor), SYNCS (mbarrier),
LDGSTS (cp.async), FFMA2 (fma.rn.f32x2).  No GPU needed."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = os.path.join(ROOT, "eegan_b200", "libeegan_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDGSTS", "UBLKCP", "FFMA2", "MUFU.EX2"]
pick = ("h_gemm_kernel", "hf_fwd_kernel", "gag_", "h_du_kernel", "h_pack", "h_pre", "cos_lse", "bn_fwd_small", "bn_bwd_small", "ssa_",
        "tc_gemm", "ts_gemm", "pair_ce", "rprecision", "ae_")
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if not any(p in name for p in pick):
        continue
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    rows.append((dem, len(re.findall(r"/\*[0-9a-f]{4}\*/", f)), [len(re.findall(r"\b" + re.escape(k), f)) for k in keys]))
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
out.write("# SASS census of the hot kernels of eegan_b200/libeegan_b200.so (sm_100a): cuobjdump -sass, instruction count and the\n"
          "# mnemonics that prove tcgen05 (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit), TMA (UTMALDG =\n"
          "# cp.async.bulk.tensor), mbarrier (SYNCS), cp.async (LDGSTS), packed fp32 FMA (FFMA2).  profiles/tools/sass_census.py\n")
out.write("%-112s %6s %s\n" % ("kernel", "instrs", " ".join(keys)))
for dem, n, c in sorted(rows, key=lambda r: -r[1]):
    out.write("%-112s %6d %s\n" % (dem[:112], n, " ".join("%*d" % (len(k), v) for k, v in zip(keys, c))))
