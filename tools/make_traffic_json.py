#!/usr/bin/env python
"""profiles/kernel_traffic.json from ncu reports: dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels bench.py's roofline
objects name.  usage: make_traffic_json.py headline.ncu-rep config1.ncu-rep [config2.ncu-rep ...]   (reports of `bench.py --no-extras` and of `tools/prof_r2.py c2|c3|c4|headline`)"""
import csv, io, json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        unit = {h: u for h, u in zip(hdr, rows[1])}

        def mbytes(k):
            v = float(d[k].replace(",", ""))
            u = unit[k].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        out.append({"name": d["Kernel Name"], "grid": d.get("launch__grid_size"), "us": float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}[unit["gpu__time_duration.sum"].lower()],
                    "traffic": int(mbytes("dram__bytes_read.sum") + mbytes("dram__bytes_write.sum"))})
    return out


head = launches(sys.argv[1])
conf = [l for rep in sys.argv[2:] for l in launches(rep)]
res = {"_source": {"headline": os.path.basename(sys.argv[1]), "configs": [os.path.basename(r) for r in sys.argv[2:]],
                   "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none; the LAST captured launch of each kernel"}}


def last(ls, pred):
    m = [l for l in ls if pred(l["name"])]
    return m[-1] if m else None


for key, ls, pred in (("fast_sparse_kernel_kN9", head, lambda n: "fast_sparse_kernel" in n), ("select_kernel_headline", head, lambda n: "select_kernel" in n),
                      ("brief_kernel_headline", head, lambda n: "brief_kernel" in n)):
    l = last(ls, pred)
    if l:
        res[key] = l["traffic"]
        res[key + "_us_under_ncu"] = round(l["us"], 1)
# tools/prof_r2.py all: corner<1> at 1280x720 x 512 (the largest Shi-Tomasi launch), corner<0> at 3840x2160 x 64 (the largest Harris launch), lsd at 1080p x 256
for key, pred in (("corner_tma_kernel<1>_c2", lambda n: "corner_tma_kernel<1" in n), ("corner_tma_kernel<0>_c3", lambda n: "corner_tma_kernel<0" in n),
                  ("lsd_kernel_c4", lambda n: "lsd_kernel" in n), ("gather_tiles_kernel_c3", lambda n: "gather_tiles_kernel" in n), ("match_kernel_headline", lambda n: "match_kernel" in n)):
    m = [l for l in conf if pred(l["name"])]
    if m:
        big = max(m, key=lambda l: l["traffic"])
        res[key] = big["traffic"]
        res[key + "_us_under_ncu"] = round(big["us"], 1)
json.dump(res, open(os.path.join(ROOT, "profiles", "kernel_traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
