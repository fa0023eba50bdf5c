"""RAW through the role-specialised loop (sampler CTAs): teacher-forced on the fp32 loop's samples, logits and draws against the fp32 loop.
   python tools/rs_check_raw.py [bits]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import make_model, norm_mel

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 9
model, _ = make_model(seed=11, bits=bits, mode="RAW")
for T, tg, ov, steps in [(26, 800, 200, 64), (166, 800, 200, 64), (646, 800, 200, 48), (1200, 800, 200, 32)]:
    mel = norm_mel(T, 5)
    a = model.generate_debug(mel, True, tg, ov, want_logits=True, seed=6, max_steps=steps)                       # fp32 loop
    forced = np.pad(a["samples"], ((0, 0), (0, tg + 2 * ov - steps)))
    b = model.generate_debug(mel, True, tg, ov, forced=forced, want_logits=True, seed=6, max_steps=steps, precision=1)
    kern = dict(model.last_timings)["loop_kernel"]
    err = float(np.abs(b["logits"] - a["logits"]).max() / np.abs(a["logits"]).max())
    agree = float((a["samples"] == b["samples"]).mean())
    print("%s %d-bit, %d folds x %d steps vs fp32 loop: logits rel err %.3e, identical draws %.5f" % (kern, bits, a["samples"].shape[0], steps, err, agree), flush=True)
