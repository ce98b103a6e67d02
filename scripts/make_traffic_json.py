#!/usr/bin/env python
"""profiles/stage_kernel_traffic.json from an `ncu --set full` capture of the stage kernel: DRAM bytes read + written by
one launch, tagged with the hash of the kernel sources so that bench.py only quotes it for the build it was taken on.
usage: make_traffic_json.py prof.ncu-rep 'row_stage_kernel<4,roe>' [launch index in the report]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, kernel = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2 + which]


def val(name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


out = {"kernel": kernel, "kernel_name_in_report": r[hdr.index("Kernel Name")], "source_sha16": bench.kernel_source_sha16(),
       "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
       "gpu_time_us_under_ncu": val("gpu__time_duration.sum") if "gpu__time_duration.sum" in hdr else None,
       "report": os.path.basename(rep), "note": "one launch of RK stage > 0 (reads u and old_solution), L2 flushed by bench.py before the step"}
json.dump(out, open(os.path.join(ROOT, "profiles", "stage_kernel_traffic.json"), "w"), indent=1)
print(out)
