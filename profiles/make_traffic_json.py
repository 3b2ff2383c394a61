"""profiles/<tag>_locate_cfg3_metrics.csv (ncu --csv of the locate kernels at cfg3) -> profiles/<tag>_traffic_cfg3.json, the
per-launch DRAM bytes bench.py reports as roofline.traffic.  Usage: python profiles/make_traffic_json.py r02z"""
import csv
import json
import os
import sys

tag = sys.argv[1]
here = os.path.dirname(os.path.abspath(__file__))
src = os.path.join(here, f"{tag}_locate_cfg3_metrics.csv")
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e6, "usecond": 1e3, "nsecond": 1.0, "second": 1e9}
out = {}
for r in rows[1:]:
    name = r[ki].split("<")[0].split("(")[0].replace("void ", "").strip()
    name = name.split("::")[-1]
    k = out.setdefault(name, {})
    m, v = r[mi], float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    if m == "dram__bytes_read.sum":
        k["dram_bytes_read"] = v
    elif m == "dram__bytes_write.sum":
        k["dram_bytes_write"] = v
    elif m == "gpu__time_duration.sum":
        k["gpu_time_ns"] = v
    elif m == "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed":
        k["l1tex2xbar_req_cycles_active_pct"] = float(r[vi])
    elif m == "lts__t_sector_hit_rate.pct":
        k["l2_sector_hit_rate_pct"] = float(r[vi])
for k in out.values():
    if "dram_bytes_read" in k and "dram_bytes_write" in k:
        k["traffic"] = k["dram_bytes_read"] + k["dram_bytes_write"]
doc = {"source": f"profiles/{tag}_locate_cfg3_metrics.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... at cfg3, "
                 "one launch = 10^6 patterns; last captured launch of every kernel)",
       "workload": "cfg3", "kernels": out}
json.dump(doc, open(os.path.join(here, f"{tag}_traffic_cfg3.json"), "w"), indent=1)
print(json.dumps(doc, indent=1))
