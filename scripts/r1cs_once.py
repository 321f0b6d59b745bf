#!/usr/bin/env python
"""one R1CS check (C4 shape, smaller) for ncu captures"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkb_loader
z = zkb_loader.load()
c = importlib.import_module("zkir_b200.circuits")
lr = int(sys.argv[1]) if len(sys.argv) > 1 else 20
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = c.BN254_FR
r = c.random_r1cs(1 << lr, 1 << (lr - 2), p, 4)
zz = c.assignment_bytes(c.r1cs_assignment(r, 1), p)
b = z.GpuBackend(0)
b.set_field(p)
b.r1cs_load(r.A, r.B, r.C, r.coef_table, r.n_vars)
b.r1cs_upload(np.broadcast_to(zz, (batch,) + zz.shape).copy())
for _ in range(3):
    v = b.r1cs_run()
print(v["ok"].all(), b.timing())
