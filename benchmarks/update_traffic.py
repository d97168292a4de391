#!/usr/bin/env python
"""profiles/roi_fwd_traffic.json from an `ncu --set full` capture of roi_fwd_kernel<RESIZE> at the bench shape (no GPU needed):
    python benchmarks/update_traffic.py gpurun_out/r02_roi_fwd_resize_c1.ncu-rep
The capture is keyed to the sha256 of csrc/roi.cu; bench.py reports `roofline.traffic` only while that still matches."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, row = rows[0], rows[1], rows[2]


def val(name):
    v, u = float(row[hdr.index(name)]), units[hdr.index(name)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
grid = row[hdr.index("Grid Size")]
images = int(grid.strip("()").split(",")[2])
src = open(os.path.join(ROOT, "faster_rcnn_b200", "csrc", "roi.cu"), "rb").read()
doc = {"kernel": "roi_fwd_kernel<RESIZE>", "images_per_launch": images, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
       "dram_bytes_per_launch": int(rd + wr),
       "algorithmic_bytes_per_launch": images * (4 * 38 * 63 * 1024 + 8 * 320 + 4 * 320 * 49 * 1024),
       "roi_cu_sha256": hashlib.sha256(src).hexdigest(),
       "source": "ncu --set full --clock-control none of %s (benchmarks/run_profiles_r02.sh); summary in profiles/%s_ncu_summary.txt"
                 % (os.path.basename(rep), os.path.basename(rep).replace(".ncu-rep", ""))}
json.dump(doc, open(os.path.join(ROOT, "profiles", "roi_fwd_traffic.json"), "w"), indent=1)
print(json.dumps(doc, indent=1))
