#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the CG kernel: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, stamped
with the hash of the CUDA sources it was measured on (bench.py reports roofline.traffic only when the stamp matches the code it runs).
usage: tools/update_traffic.py gpurun_out/r2_fused.ncu-rep C2_f32_cg_fused_bytes"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402

rep, key = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
get = lambda name: float(vals[hdr.index(name)].replace(",", "")) * scale[units[hdr.index(name)]]
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
p = os.path.join(ROOT, "profiles", "traffic.json")
d = json.load(open(p)) if os.path.exists(p) else {}
if d.get("source_sha") != kernel_source_sha():
    d = {}
d.update({"source_sha": kernel_source_sha(), key: rd + wr, key + "_read_write": [rd, wr], "kernel": vals[hdr.index("Kernel Name")][:120],
          "how": f"ncu --set full --clock-control none, one launch (cache flushed between replay passes), {os.path.basename(rep)}"})
json.dump(d, open(p, "w"), indent=1)
print(json.dumps(d, indent=1))
