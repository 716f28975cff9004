#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report: DRAM bytes per launch (dram__bytes_read.sum +
dram__bytes_write.sum, averaged over the captured launches of each kernel) keyed by kernel name, stamped with the hash of
the library that was profiled.  bench.py reports `roofline.traffic` only when that hash equals the library it loaded.

    python scripts/ncu_traffic.py gpurun_out/TAG/prof_cp.ncu-rep [more.ncu-rep ...] --lib-hash $(cat gpurun_out/TAG/lib_hash.txt)

The hash is written on the GPU box next to the report (scripts/gpu_*.sh: `python -c "import bench; print(bench.lib_build_id())"`)
so that it names the binary that actually ran."""
import argparse
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--lib-hash", required=True)
    ap.add_argument("--note", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    args = ap.parse_args()
    acc = {}
    for rep in args.reports:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        for r in rows[2:]:
            name = re.sub(r"^void\s+", "", re.sub(r"<.*", "", r[ik]).split("::")[-1].split("(")[0].strip())
            b = float(r[ir].replace(",", "")) * UNIT_SCALE[units[ir]] + float(r[iw].replace(",", "")) * UNIT_SCALE[units[iw]]
            acc.setdefault(name, []).append(b)
    doc = {"build_id": args.lib_hash.strip(), "reports": [os.path.relpath(r, ROOT) for r in args.reports], "note": args.note,
           "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), mean over the captured launches",
           "kernels": {k: sum(v) / len(v) for k, v in acc.items()}, "launches": {k: len(v) for k, v in acc.items()}}
    json.dump(doc, open(args.out, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
