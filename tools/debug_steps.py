import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import wavernn_oracle as orc
from tests.util import make_model, norm_mel
model, sd = make_model(seed=11, bits=9, mode="RAW")
mel = norm_mel(24, 5)
for batched, tg, ov in ((False, 0, 0), (True, 1000, 200)):
    out = model.generate_debug(mel, batched, tg, ov, want_logits=True, seed=3, max_steps=6)
    S = 4800 if not batched else 1400
    forced = np.pad(out["samples"], ((0, 0), (0, S - out["samples"].shape[1])))
    _, tr = orc.generate(mel, sd, mode="RAW", batched=batched, target=tg, overlap=ov, seed=3, forced_samples=forced,
                         return_trace=True, max_steps=6)
    err = np.abs(out["logits"] - tr["logits"]).max(axis=2)
    print("batched", batched, "per (fold, step) max abs logit err:\n", err)
