import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d["kernels"]
print(d["ms_per_step"], {n:round(k[n]["ms_per_step"],4) for n in ("block_fwd","block_bwd_pre","block_wgrad","softmax_xent") if n in k})
