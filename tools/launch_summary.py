"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) into the
per-step table of profiles/: one step = the launches between the last two `adam_dev_kernel` launches.
Usage: python tools/launch_summary.py launches.csv "<command line>" "<timed ms per step>" [traffic.json [title]] > profiles/x.md ;
the traffic json (stamped with the hash of the convolution sources) is what bench.py reports as roofline.traffic."""
import csv, sys, json, collections, re

path, cmdline, timed = sys.argv[1], sys.argv[2], sys.argv[3]
title = sys.argv[5] if len(sys.argv) > 5 else "one training step, per-kernel device time and DRAM traffic (ncu)"
rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(int(r[0]), {"name": r[4], "grid": r[8], "block": r[7]})
    v = float(r[14])
    unit = r[13]
    if r[12] == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    elif r[12] == "dram__bytes_read.sum":
        d["rd"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif r[12] == "dram__bytes_write.sum":
        d["wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
ids = sorted(launch)
adam = [i for i in ids if "adam_dev_kernel" in launch[i]["name"]]
step = [launch[i] for i in ids if adam[-2] < i <= adam[-1]]


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*\)$", "", n)
    n = n.replace("vxm::", "").replace("at::", "")
    return n[:96]


tot = sum(x["us"] for x in step)
agg = collections.OrderedDict()
for x in step:
    a = agg.setdefault(short(x["name"]), [0, 0.0, 0.0])
    a[0] += 1; a[1] += x["us"]; a[2] += x.get("rd", 0) + x.get("wr", 0)
conv = [x for x in step if re.search(r"conv_tc|wgrad|pack_weights", x["name"])]
conv_us = sum(x["us"] for x in conv)
conv_bytes = sum(x.get("rd", 0) + x.get("wr", 0) for x in conv if re.search(r"conv_tc\w*_kernel|wgrad\w*_kernel", x["name"]) and "reduce" not in x["name"])
print("# %s\n" % title)
print("Command (gpurun, 1x B200): `%s`\n" % cmdline)
print("bf16 tensor-core engine, eager launches (no CUDA graph) so that every kernel is visible.  ncu serialises launches and runs")
print("them cold: read SHARES, not absolute times (the timed, graph-replayed bench step is %s ms)." % timed)
print("One step = the launches between two `adam_dev_kernel` launches: %d launches, %.2f ms total.\n" % (len(step), tot / 1e3))
print("| kernel | launches | time (us) | share | DRAM bytes (MB) |\n|---|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, a[0], a[1], 100 * a[1] / tot, a[2] / 1e6))
print("\nConvolution family (all `tc*::` kernels incl. packing / reduce): %.2f ms = %.1f%% of the step; DRAM traffic of the conv + wgrad kernels: %.1f MB per step.\n"
      % (conv_us / 1e3, 100 * conv_us / tot, conv_bytes / 1e6))
print("## Every launch of that step, in order\n\n| # | kernel | grid | time (us) | DRAM MB |\n|---:|---|---|---:|---:|")
for i, x in enumerate(step):
    print("| %d | `%s` | %s | %.1f | %.1f |" % (i, short(x["name"])[:80], x["grid"], x["us"], (x.get("rd", 0) + x.get("wr", 0)) / 1e6))
if len(sys.argv) > 4:
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench                      # the library sources this pass was captured from: bench.py refuses a stale figure
    json.dump({"conv_dram_mbytes_per_step": conv_bytes / 1e6, "conv_source_sha": bench.conv_source_hash(),
               "source": "%s (ncu dram__bytes_read.sum + dram__bytes_write.sum over the conv_tc*/wgrad* kernels of one step)" % os.path.basename(path)},
              open(sys.argv[4], "w"), indent=1)
