import ctypes, sys, os
sys.path.insert(0, "/root/repo")
from cuking_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libcuking_b200_prof.so")
import cuking_b200 as ck, numpy as np
ctx = ck.Context(0, king_variant=2)
n, s = 4096, 100000
pl = ctx.planes(ck.submatrix(n), s); pl.synthesize(42, 0.01); pl.finalize()
out = np.empty(1 << 20, dtype=ck.RESULT_DTYPE)
L = capi.load()
buf = (ctypes.c_ulonglong * 16)()
for it in range(2):
    L.ck_debug_umma_prof(buf)
    r = pl.king(0.0884, 1 << 20, out=out)
    print("kernel ms", ctx.timings()["king_ms"], len(r))
L.ck_debug_umma_prof(buf)
v = list(buf)
names = ["A expand", "A wait-empty", "A store+arrive", "A steps(warps*k)", "B expand", "B wait-empty", "B store+arrive", "B steps", "issuer0 wait-full", "issuer1 wait-full", "issuer2 wait-full", "k steps"]
for nme, x in zip(names, v): print(f"{nme:20s} {x}")
ka, kb, kk = max(v[3],1), max(v[7],1), max(v[11],1)
print("per step per warp: A expand %.0f wait %.0f store %.0f | B expand %.0f wait %.0f store %.0f | issuers wait %.0f %.0f %.0f" % (v[0]/ka, v[1]/ka, v[2]/ka, v[4]/kb, v[5]/kb, v[6]/kb, v[8]/kk, v[9]/kk, v[10]/kk))
