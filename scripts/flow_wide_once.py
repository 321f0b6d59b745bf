#!/usr/bin/env python
"""one-witness evaluation of a random circuit over a wide field, device timing (CUDA events inside libzkb): the barrier kernels
(ZKB_FLOW=0) against the dataflow launch.  usage: flow_wide_once.py <field> <log2_gates> <window>"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkb_loader
z = zkb_loader.load()
c = importlib.import_module("zkir_b200.circuits")
p = {"bls381": c.BLS12_381_FR, "bn254": c.BN254_FR, "goldilocks": c.GOLDILOCKS}[sys.argv[1]]
lg, window = int(sys.argv[2]), int(sys.argv[3])
circ = c.random_circuit(1 << lg, 1024, p, 0x5EED0002, window=window)
b = z.GpuBackend(0)
b.set_field(p)
b.push_gates(circ.gates, circ.const_pool)
b.finalize(False)
w = c.make_witnesses(circ, 1, seed=3)
v = b.evaluate(None, w, 1)
assert int(v[0]["ok"]) == 1
bad = c.make_witnesses(circ, 1, seed=3, corrupt={0: 1})
vb = b.evaluate(None, bad, 1)
assert int(vb[0]["ok"]) == 0 and int(vb[0]["first_fail_seq"]) == c.expected_first_fail(circ, 1, {0: 1})[0]
b.upload_inputs(None, w, 1)
tot, lv = [], []
for i in range(7):
    b.run()
    t = b.timing()
    if i >= 2:
        tot.append(t["total_ms"]); lv.append(t["levels_ms"])
st = b.stats()
print(json.dumps({"field": sys.argv[1], "gates": 1 << lg, "window": window, "flow": os.environ.get("ZKB_FLOW", "1"), "levels": st["n_levels"],
                  "ms": float(np.median(tot)), "levels_ms": float(np.median(lv)), "us_per_level": float(np.median(lv)) * 1e3 / st["n_levels"],
                  "launches": b.timing()["kernel_launches"]}))
