#!/usr/bin/env python
"""profiles/r02_playout_traffic.json from an ncu capture of k_playout<0,false> (tools/profile_playouts.py under
`ncu --set full`): warp instructions per playout and DRAM bytes per launch, stamped with the hash of the kernel sources.

  python tools/playout_traffic.py gpurun_out/<rep>.ncu-rep [games per launch = 65536] > profiles/r02_playout_traffic.json
"""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinforcementlearning_blokus_b200 import build  # noqa: E402

rep = sys.argv[1]
games = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if "k_playout<0, 0>" in d.get("Kernel Name", "") or "k_playout<0, false>" in d.get("Kernel Name", ""):
        inst = float(d["smsp__inst_executed.sum"].replace(",", ""))
        dram = float(d["dram__bytes_read.sum"].replace(",", "")) + float(d["dram__bytes_write.sum"].replace(",", ""))
        unit = dict(zip(hdr, rows[1]))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit.get("dram__bytes_read.sum", "byte"), 1)
        print(json.dumps({"kernel": "k_playout<0,false>", "games": games, "dram_bytes_per_launch": int(dram * scale),
                          "warp_instr_per_playout": int(round(inst / games)), "duration_ns": d.get("gpu__time_duration.sum"),
                          "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          "kernel_source_hash": build.playout_source_hash(), "source": os.path.basename(rep)}))
        break
else:
    sys.exit("no k_playout<0,false> launch in " + rep)
