"""Stage-level clock64 profile of the tensor-core pairwise kernels (block 0 only): where the expander / issuer warps
spend their time.  Needs the profiling build: make -C cuking_b200/csrc prof.   usage: umma_stage_profile.py [variant]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuking_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libcuking_b200_prof.so")
import cuking_b200 as ck, numpy as np
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = ck.Context(0, king_variant=variant)
n, s = 4096, 100000
pl = ctx.planes(ck.submatrix(n), s); pl.synthesize(42, 0.01); pl.finalize()
out = np.empty(1 << 20, dtype=ck.RESULT_DTYPE)
L = capi.load()
prof = L.ck_debug_screen_prof if variant == 5 else L.ck_debug_fp4_pair_prof if variant == 4 else L.ck_debug_fp4_prof if variant == 3 else L.ck_debug_umma_prof
buf = (ctypes.c_ulonglong * 32)()
for it in range(2):
    prof(buf)
    r = pl.king(0.0884, 1 << 20, out=out)
    print("variant", variant, "kernel ms", ctx.timings()["king_ms"], len(r))
prof(buf)
allv = list(buf)
if variant == 4:  # second half: the peer CTA of the first cluster
    v = allv[16:]
    ka, kb = max(v[3], 1), max(v[7], 1)
    print("PEER per item per warp: A expand %.0f wait %.0f store %.0f | B expand %.0f wait %.0f store %.0f" % (v[0]/ka, v[1]/ka, v[2]/ka, v[4]/kb, v[5]/kb, v[6]/kb))
v = allv[:16]
names = ["A expand", "A wait-empty", "A store+arrive", "A items(warps*k)", "B expand", "B wait-empty", "B store+arrive", "B subs", "issuer0 wait-full", "issuer1 wait-full", "issuer2 wait-full", "issuer1 issue+commit" if variant == 3 else "k steps", "mainloop clk", "steps", "epilogue clk"]
for nme, x in zip(names, v): print(f"{nme:20s} {x}")
ka, kb = max(v[3], 1), max(v[7], 1)
kk = max(v[13] if variant >= 3 else v[11], 1)
print("per item per warp: A expand %.0f wait %.0f store %.0f | B expand %.0f wait %.0f store %.0f | issuers wait/step %.0f %.0f %.0f" % (v[0]/ka, v[1]/ka, v[2]/ka, v[4]/kb, v[5]/kb, v[6]/kb, v[8]/kk, v[9]/kk, v[10]/kk))
if variant >= 3:
    print("mainloop clk/step %.1f (tensor floor 212; screen kernel 127), issuer1 issue+commit clk/step %.1f, epilogue clk %d" % (v[12] / kk, v[11] / kk, v[14]))
