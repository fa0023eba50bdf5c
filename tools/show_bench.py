"""Summary of bench lines: python tools/show_bench.py file.json ..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    lp = d.get("loop", {})
    print("%s: %.1fx RT (e2e %.1fx), %.2f ms/step, %s %.2f us/loop step = %.2fx floor, roofline %s frac %.4f traffic %s, cpu %s %.2fx RT, clocks %s" % (
        f.split("/")[-1], d["value"] / 16000, d["e2e"]["value"] / 16000, d["ms_per_step"], lp.get("kernel"), lp.get("us_per_step", 0), lp.get("step_over_floor", 0),
        d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"]["traffic"], d["cpu_baseline"].get("kind"), d["cpu_baseline"]["value"] / 16000, d.get("clocks", {}).get("sm_mhz")))
    for k, v in d.get("plans", {}).items():
        print("    plan %s: %.1fx RT, %.2f us/step (%.2fx floor) on %s" % (k, v["x_realtime"], v["loop"]["us_per_step"], v["loop"]["step_over_floor"], v["loop"]["kernel"]))
    if d.get("parity_check"): print("    parity", d["parity_check"].get("logits_rel_err"), d["parity_check"].get("draw_agreement"))
    for k, v in d.get("sharded", {}).items():
        print("    sharded %s: %s" % (k, {a: b for a, b in v.items() if a in ("x_realtime", "value", "ms_per_call", "error")}))
