#!/usr/bin/env python
"""Record the DRAM bytes per launch of a kernel from an `ncu --set full` report into profiles/traffic.json,
keyed by workload:batch and by the content hash of the kernel's sources (bench.py reports `roofline.traffic`
only while that hash still matches: a stale capture reads as null).

    python tools/make_traffic_json.py REPORT.ncu-rep cfg3:16384 tnq_ladder.cu tnq_ladder_core.cuh tnq_f2.cuh
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    rep, key, files = sys.argv[1], sys.argv[2], sys.argv[3:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    get = lambda name: (float(vals[hdr.index(name)]), units[hdr.index(name)])
    total = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = get(name)
        total += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    path = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[key] = {"dram_bytes_per_launch": total, "files": files, "csrc_sha256": ge.csrc_digest(files),
                  "kernel": vals[hdr.index("Kernel Name")], "report": os.path.basename(rep)}
    json.dump(table, open(path, "w"), indent=1)
    print(key, total)


if __name__ == "__main__":
    main()
