"""Bring-up diagnostics for the tcgen05 Gram kernel (gram_tc.cuh): what TMA put into shared
memory, what the accumulator holds, and a sweep of descriptor variants."""
import sys, os, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pogs_b200 import _lib as L

BK, STAGE = 16, 12288
np.set_printoptions(linewidth=250, precision=1, suppress=True)

def gram_dbg(A):
    A = np.ascontiguousarray(A, np.float32); m, n = A.shape
    G = np.zeros((n, n), np.float32); st = np.full(STAGE, -7, np.float32); acc = np.full((128, 256), -7, np.float32)
    f = ctypes.c_float
    rc = L.lib.pogs_b200_gram_debug_s(m, n, L.ptr(A, f), L.ptr(G, f), L.ptr(st, f), L.ptr(acc, f))
    return rc, G, st, acc

def expected_stage(A, I=0, J=0):
    """hi part only (inputs are TF32-exact): [hi_I 4 blocks | lo_I 4 | hi_J 8 | lo_J 8], block = 16 rows x 128 B swizzled."""
    out = np.zeros(STAGE, np.float32)
    def put(base_f, colblock0, nblk):
        for b in range(nblk):
            for r in range(BK):
                for c in range(4):   # 32-byte chunks, XORed with the row index mod 4
                    src = A[r, (colblock0 + b) * 32 + c * 8:(colblock0 + b) * 32 + c * 8 + 8] if r < A.shape[0] else np.zeros(8)
                    src = np.pad(src, (0, 8 - len(src)))
                    off = base_f + (b * 2048 + r * 128 + ((c ^ (r % 4)) * 32)) // 4
                    out[off:off + 8] = src
    put(0, I * 4, 4); put(2 * 4 * 512, J * 8, 8)
    return out

n = 256
A = np.zeros((16, n), np.float32)
for r in range(16): A[r, :] = (r % 8) * 256 + np.arange(n)
rc, G, st, acc = gram_dbg(A)
exp = expected_stage(A)
print("rc", rc, L.last_error() if rc else "")
print("stage0: unwritten(-7):", (st == -7).sum(), " nonzero:", (st != 0).sum(), " matches expected swizzled layout:", np.array_equal(st, exp),
      " mismatches:", (st != exp).sum())
if not np.array_equal(st, exp):
    print(" first 40 floats of stage :", st[:40]); print(" expected                 :", exp[:40])
    print(" floats 32..72 (row 1)    :", st[32:72]); print(" expected                 :", exp[32:72])
    # is it the unswizzled layout?
    uns = np.zeros(STAGE, np.float32)
    for b in range(4):
        for r in range(16): uns[(b * 2048 + r * 128) // 4:(b * 2048 + r * 128) // 4 + 32] = A[r, b * 32:b * 32 + 32]
    print(" hi_I equals UNswizzled layout:", np.array_equal(st[:2048], uns[:2048]))
    for name, lo, hi in (("hi_I", 0, 2048), ("lo_I", 2048, 4096), ("hi_J", 4096, 8192), ("lo_J", 8192, 12288)):
        print(f"  region {name}: nonzero {np.count_nonzero(st[lo:hi])} sum {st[lo:hi].astype(np.float64).sum():.0f} expected sum {exp[lo:hi].astype(np.float64).sum():.0f}")
print("acc0: unwritten(-7):", (acc == -7).sum(), " nonzero:", np.count_nonzero(acc), " unique values:", np.unique(acc)[:10])

def run(name, A, env=None):
    for k in ("POGS_B200_GRAM_DESC", "POGS_B200_GRAM_IDESC"): os.environ.pop(k, None)
    if env: os.environ.update(env)
    rc, G, st, acc = gram_dbg(A)
    ref = A.astype(np.float64).T @ A.astype(np.float64)
    bad = (G.astype(np.float64) != ref)
    print(f"-- {name} {env or ''}: rc {rc} wrong {bad.sum()}/{bad.size} G nonzero {np.count_nonzero(G)} acc nonzero {np.count_nonzero(acc)} acc[0,:6] {acc[0,:6]} acc[1,:6] {acc[1,:6]} ref[0,:6] {ref[0,:6]} ref[1,:6] {ref[1,:6]}")
    return G, acc, ref

ones = np.ones((16, n), np.float32)
ramp = np.zeros((16, n), np.float32); ramp[3, :] = np.arange(1, n + 1) % 64
D = 0x10 | (2 << 7) | (2 << 10) | (32 << 17) | (8 << 24)
variants = [None,
            {"POGS_B200_GRAM_DESC": "2048,1024,1"},
            {"POGS_B200_GRAM_DESC": "512,2048,1"},
            {"POGS_B200_GRAM_DESC": "2048,256,1"},
            ]
for v in variants:
    run("ones", ones, v)
    G, acc, ref = run("ramp row3", ramp, v)
for k in ("POGS_B200_GRAM_DESC", "POGS_B200_GRAM_IDESC"): os.environ.pop(k, None)
rng = np.random.default_rng(0)
A = rng.integers(-3, 4, size=(200, 300)).astype(np.float32)
G, acc, ref = run("random ints 200x300", A)
if (G != ref).any():
    bad = G != ref; nb = (300 + 31) // 32
    fm = np.array([[bad[bi*32:(bi+1)*32, bj*32:(bj+1)*32].mean() * 9.99 for bj in range(nb)] for bi in range(nb)])
    print(np.floor(fm).astype(int))
A = rng.standard_normal((1000, 300)).astype(np.float32)
rc, G, st, acc = gram_dbg(A)
ref = A.astype(np.float64).T @ A.astype(np.float64); sc = np.abs(A).astype(np.float64).T @ np.abs(A).astype(np.float64)
print("gaussian 1000x300: max scaled err", np.max(np.abs(G - ref) / sc))
