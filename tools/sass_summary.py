"""SASS evidence (no GPU needed): mnemonic counts per object of the sm_100a build -- TMA bulk copies (UBLKCP) and mbarrier
waits (SYNCS), 128-bit loads / stores, shared-memory atomics, warp votes, fp64 and fp32 arithmetic, tensor-core opcodes.
    python -m semantic_depth_b200.build --force && python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import os, re, subprocess
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "semantic_depth_b200", "build")
pats = {"UBLKCP (TMA bulk copy)": r"\bUBLKCP", "SYNCS (mbarrier)": r"\bSYNCS", "LDG.E.128": r"LDG\.E\.128", "LDG.E.64": r"LDG\.E\.64",
        "LDS.128": r"LDS\.128", "STG.E.128": r"STG\.E\.128", "ATOMS": r"\bATOMS", "REDUX / VOTE / MATCH": r"\b(REDUX|VOTE|MATCH)",
        "SHFL": r"\bSHFL", "DFMA": r"\bDFMA", "DADD": r"\bDADD", "DMUL": r"\bDMUL", "FFMA": r"\bFFMA", "FMNMX / VIMNMX": r"\b(FMNMX|VIMNMX)",
        "HMMA / UTCMMA (tensor)": r"\b(HMMA|UTC.MMA|UTCHMMA|TCGEN)"}
print("SASS mnemonic counts per object (cuobjdump -sass of the sm_100a cubins in libsd_fusion.so; tools/sass_summary.py)")
print("TMA = UBLKCP + SYNCS (pixel_label_kernel stages its logits tile with cp.async.bulk + mbarrier); no tensor-core instruction anywhere")
print("(north_star: not GEMM-shaped).  DFMA with -fmad=false come from the fp64 division / sqrt expansions, FFMA from explicit fmaf")
print("in the k-NN key and the fp32 division expansions.\n")
for f in sorted(os.listdir(root)):
    if not f.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(root, f)], capture_output=True, text=True).stdout
    funcs = re.findall(r"Function : (\S+)", sass)
    counts = {k: len(re.findall(p, sass)) for k, p in pats.items()}
    print(f"{f}: {len(funcs)} kernels; " + ", ".join(f"{k}={v}" for k, v in counts.items() if v))
