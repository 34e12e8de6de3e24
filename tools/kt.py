import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d["kernels"]
if len(sys.argv) > 1:
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["launches_per_step"]); print({n:round(v["ms_per_step"],4) for n,v in k.items()})
    print({n: round(v["frac"], 3) for n, v in d["roofline_all"].items()}); print(d["loss"])
else:
    print(d["ms_per_step"], {n:round(k[n]["ms_per_step"],4) for n in ("block_fwd","block_bwd_pre","block_wgrad","softmax_xent") if n in k})
