"""Issue-rate microbenchmark of tcgen05.mma (see csrc/tc_gemm_test.cu: umma_rate_kernel)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtvc_b200
from rtvc_b200 import _native

lib = _native.load()
eng = C.c_void_p()
assert lib.wrnn_create(0, 9, _native.MODE_MOL, C.byref(eng)) == 0
iters = 2000
for mode in (0, 1, 5, 7, 6):
    for N in (16, 32, 64, 128, 256):
        if mode in (5, 7) and N > 64:
            continue
        a, b = C.c_int64(), C.c_int64()
        rc = lib.wrnn_debug_umma_rate(eng, N, iters, mode, C.byref(a), C.byref(b))
        print(f"mode {mode} N {N:3d} rc {rc}  issue {a.value / (4 * iters):7.1f} clk/mma   total {b.value / (4 * iters):7.1f} clk/mma")
lib.wrnn_destroy(eng)
