"""Small driver for ncu: runs the tensor-core front end once for a 60 s mel (kernel: gemm_tc_split_kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel
model, _ = make_model(seed=12, bits=9, mode="MOL")
mel = norm_mel(4800, 1)
for _ in range(2):
    model.conditioning_tc(mel)
print("ok")
