"""profiles/ncu_spmm_r02.json from the raw page of an `ncu --set full` capture of tools/spmm_probe.py --ncu
(launch 1 = the whole product): DRAM bytes, L2 -> SM bytes (tex read sectors x 32), duration, and the hash of the
kernel source the capture belongs to (bench.py refuses the numbers when csrc/spmm.cu changed since).

    ncu --set full --clock-control none -k regex:spmm_kernel -c 3 -o gpurun_out/ncu_spmm_r02 python tools/spmm_probe.py --ncu
    ncu -i gpurun_out/ncu_spmm_r02.ncu-rep --page raw --csv > profiles/ncu_spmm_r02_raw.csv
    python tools/ncu_spmm_json.py profiles/ncu_spmm_r02_raw.csv
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, first = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def val(name, want_unit):
    v, u = float(first[col[name]].replace(",", "")), units[col[name]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "sector": 1, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u]
    return v * scale


out = {"kernel": "spmm_kernel<64,4,4,0>", "workload": "synthetic",
       "spmm_cu_sha16": hashlib.sha256(open(os.path.join(ROOT, "recad_b200", "csrc", "spmm.cu"), "rb").read()).hexdigest()[:16],
       "dram_bytes_read": int(val("dram__bytes_read.sum", "byte")), "dram_bytes_write": int(val("dram__bytes_write.sum", "byte")),
       "l2_to_sm_bytes": int(val("lts__t_sectors_srcunit_tex_op_read.sum", "sector") * 32),
       "ncu_time_ms": round(val("gpu__time_duration.sum", "ms"), 5),
       "source": f"{os.path.relpath(sys.argv[1], ROOT)} (ncu --set full, launch 1 of 3)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_spmm_r02.json"), "w"), indent=1)
print(out)
