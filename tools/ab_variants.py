"""A/B of library builds on the bench: `python tools/ab_variants.py default,<name>[,...] [bench args]`.
<name> = a build made with sassy_b200.build.build_variant(name, defines) (selected through SASSY_B200_LIB);
prints step and kernel times per variant (used for the sweeps recorded under profiles/)."""
import json, os, subprocess, sys
variants = sys.argv[1].split(",")
extra = sys.argv[2:] 
for v in variants:
    env = dict(os.environ)
    if v != "default":
        env["SASSY_B200_LIB"] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sassy_b200", "lib", f"libsassy_b200_{v}.so")
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"), "--no-cpu", "--no-e2e", "--steps", "60"] + extra, env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        r = j["roofline"]
        print(v, extra, "ms/step %.4f" % j["ms_per_step"], "kernel_ms %.4f" % r["kernel_ms"], "verify %.4f" % r["verify_kernel_ms"], "frac %.3f" % r["frac"], "bps", j["config"]["blocks_per_sm"], "rows", j["config"]["rows"], flush=True)
    except Exception as e:
        print(v, "FAILED", out.stderr[-500:], flush=True)
