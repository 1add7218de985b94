import json
from collections import defaultdict
d=json.load(open("gpurun_out/unet_ops.json"))
agg=defaultdict(float)
for o in d["ops"]: agg[o["kind"]]+=o["ms"]
print({k:round(v,3) for k,v in agg.items()}, round(d["total_ms"],3))
