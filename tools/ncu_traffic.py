"""profiles/traffic.json from `ncu --set full` captures of THIS build: DRAM bytes (read + write) of one
launch of the dominant kernels, stamped with the hash of the kernel sources they were taken from
(bench.py reports `roofline.traffic` only while that hash still matches).  Run here, no GPU needed:
  python tools/ncu_traffic.py window_1m=gpurun_out/r2b/win_1m.ncu-rep window_64m=... tick_cascade_1m=..."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def one(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    def val(name):
        i = hdr.index(name)
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
    return {"kernel": r[hdr.index("Kernel Name")], "dram_read_bytes": val("dram__bytes_read.sum"),
            "dram_write_bytes": val("dram__bytes_write.sum"),
            "traffic_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
            "duration_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
            {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0),
            "report": os.path.relpath(path, ROOT)}


res = {"kernel_source_sha": bench.kernel_source_sha(), "how": "ncu --set full --clock-control none, one launch each"}
for arg in sys.argv[1:]:
    key, path = arg.split("=", 1)
    res[key] = one(path)
    res["traffic_" + key + "_bytes"] = res[key]["traffic_bytes"]
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps(res, indent=1))
