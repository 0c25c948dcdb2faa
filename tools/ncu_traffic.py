#!/usr/bin/env python3
"""Extract per-launch DRAM traffic of the bench kernels from `ncu --set full` reports taken at bench shard size.
usage: python tools/ncu_traffic.py NSUB stats.ncu-rep spec.ncu-rep [order.ncu-rep] > profiles/ncu_traffic.json
The output names the sha256 of the kernel sources in the tree (bench.csrc_sha256): bench.py uses the figures only for
those very sources."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys


def traffic(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def get(name):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
        return v * mult
    return {"kernel": vals[hdr.index("Kernel Name")], "dram_read": get("dram__bytes_read.sum"),
            "dram_write": get("dram__bytes_write.sum"), "gpu_time": float(vals[hdr.index("gpu__time_duration.sum")].replace(",", "")),
            "time_unit": units[hdr.index("gpu__time_duration.sum")]}


def main():
    nsub = int(sys.argv[1])
    st, sp = traffic(sys.argv[2]), traffic(sys.argv[3])
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    out = {"subjects_per_gpu": nsub, "csrc_sha256": bench.csrc_sha256(),
           "source": "ncu --set full --clock-control none, one launch each at bench shard size",
           "traffic_bytes_per_launch": {"window_stats": st["dram_read"] + st["dram_write"],
                                        "window_spectral": sp["dram_read"] + sp["dram_write"]},
           "detail": {"window_stats": st, "window_spectral": sp}}
    if len(sys.argv) > 4:
        od = traffic(sys.argv[4])
        out["traffic_bytes_per_launch"]["window_order"] = od["dram_read"] + od["dram_write"]
        out["detail"]["window_order"] = od
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
