"""Turns an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
of scripts/profile_unet.py --eager-only into (a) a per-kernel table of the LAST complete UNet step in the
capture and (b) profiles/r01_gemm_traffic.json, the measured DRAM traffic per ldx_gemm_kernel launch that
bench.py reports as roofline.traffic.  Usage: python scripts/summarize_launches.py <csv> <ops_per_step>"""
import csv, json, re, sys
from collections import defaultdict

path = sys.argv[1]
rows = list(csv.reader(open(path)))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ik, im, iv, iid, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID"), h.index("Metric Unit")
launches = {}
for r in rows[start:]:
    if len(r) <= iv:
        continue
    d = launches.setdefault(int(r[iid]), {"name": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("ldx::", "")})
    v = float(r[iv].replace(",", ""))
    unit = r[iu]
    if unit == "Mbyte": v *= 1e6
    elif unit == "Kbyte": v *= 1e3
    elif unit == "Gbyte": v *= 1e9
    elif unit in ("us", "usecond"): v *= 1e3
    elif unit in ("ms", "msecond"): v *= 1e6
    d[r[im]] = v
ids = sorted(launches)
# the eager run is: plan.run (step 1), prog.run (step 2), ... ; a step starts at the nchw_to_nhwc kernel
cand = [i for i in ids if "nchw_to_nhwc" in launches[i]["name"]]
starts = [c for k, c in enumerate(cand) if k == 0 or c - cand[k - 1] > 10]   # the step begins with two layout kernels
first, last = starts[-2], starts[-1]
step = [launches[i] for i in ids if first <= i < last]
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for l in step:
    a = agg[l["name"]]
    a[0] += 1
    a[1] += l.get("gpu__time_duration.sum", 0.0)
    a[2] += l.get("dram__bytes_read.sum", 0.0)
    a[3] += l.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"one UNet step (B=16, 64x64 latent): {len(step)} launches, {tot/1e6:.3f} ms summed kernel time under ncu "
      f"(serialised, cold clocks/cache)")
print(f"{'kernel':48s} {'launches':>8s} {'ms':>8s} {'share':>6s} {'DRAM rd MB':>10s} {'DRAM wr MB':>10s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:48]:48s} {a[0]:8d} {a[1]/1e6:8.3f} {100*a[1]/tot:5.1f}% {a[2]/1e6:10.1f} {a[3]/1e6:10.1f}")
FAMILY = ("ldx_gemm_kernel", "ldx_conv_halo_kernel", "ldx_ff_fused_kernel")
g = [a for n, a in agg.items() if any(f in n for f in FAMILY)]
n = sum(a[0] for a in g)
per = {}
for nm in FAMILY:
    gg = [a for n2, a in agg.items() if nm in n2]
    if gg:
        k = sum(a[0] for a in gg)
        per[nm] = {"launches": k, "ms": sum(a[1] for a in gg) / 1e6, "dram_bytes_per_launch": (sum(a[2] for a in gg) + sum(a[3] for a in gg)) / k}
out = {"kernel": " + ".join(per), "per_kernel": per, "launches_per_unet_step": n,
       "dram_bytes_read_per_launch": sum(a[2] for a in g) / n, "dram_bytes_write_per_launch": sum(a[3] for a in g) / n,
       "dram_bytes_per_launch": (sum(a[2] for a in g) + sum(a[3] for a in g)) / n,
       "share_of_step_kernel_time": sum(a[1] for a in g) / tot,
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, scripts/profile_unet.py --eager-only (per launch, cache flushed between kernels by ncu)"}
json.dump(out, open(sys.argv[2] if len(sys.argv) > 2 else "profiles/r01_gemm_traffic.json", "w"), indent=1)
print(json.dumps(out))
