#!/usr/bin/env python
"""Combine an `ncu --set full` capture of k_level launches with the plan's per-level algorithmic bytes.
The capture was taken with `-k regex:k_level -s 20 -c 3` on `bench.py --witnesses 256` (one tile), so the
captured launches are levels 20, 21, 22 of the first pass.  Writes profiles/traffic.json, which bench.py reads
for roofline.traffic (DRAM bytes per launch = measured traffic/algorithmic ratio x algorithmic bytes per launch).
usage: python profiles/make_traffic_json.py profiles/r01_k_level_ncu_full_summary.csv 20 256"""
import csv
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkb_loader  # noqa: E402


def main():
    summary, first_level, wt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    z = zkb_loader.load()
    circ = importlib.import_module("zkir_b200.circuits")
    p = circ.BLS12_381_FR
    c = circ.random_circuit(1 << 24, 1024, p, 0x5EED0003)
    b = z.GpuBackend(-1)
    b.set_field(p)
    b.push_gates(c.gates, c.const_pool)
    b.finalize(False)
    rows = list(csv.reader(open(summary)))
    hdr = rows[0]
    rd, wr, tm = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    out = {"capture": os.path.basename(summary), "tile_witnesses": wt, "launches": []}
    for k, r in enumerate(rows[2:]):
        info = b.level_info(first_level + k)
        dram = float(r[rd]) * unit[rows[1][rd]] + float(r[wr]) * unit[rows[1][wr]]
        algo = info["algo_bytes_per_witness"] * wt
        out["launches"].append({"level": first_level + k, "gates": info["gates"], "dram_bytes": dram, "algorithmic_bytes": algo,
                                "ratio": dram / algo, "ncu_duration_ms": float(r[tm])})
    out["traffic_over_algorithmic"] = sum(l["dram_bytes"] for l in out["launches"]) / sum(l["algorithmic_bytes"] for l in out["launches"])
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
