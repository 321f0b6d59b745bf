#!/usr/bin/env python
"""Integer-pipe ceiling of the gate kernels: register-resident Montgomery products / modular additions per second
(no memory traffic), per field width.  One JSON line per field."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkb_loader  # noqa: E402

z = zkb_loader.load()
FIELDS = {"goldilocks (2 limbs, portable CIOS)": (1 << 64) - (1 << 32) + 1,
          "kat124 (4 limbs, PTX chains)": 16249742125730185677094195492597105093,
          "bn254 (8 limbs, PTX chains)": 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001,
          "bls12-381 Fr (8 limbs, PTX chains)": 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001}
for name, p in FIELDS.items():
    b = z.GpuBackend(0)
    b.set_field(p)
    mul = b.debug_field_throughput(1, 4000)
    add = b.debug_field_throughput(0, 4000)
    print(json.dumps({"field": name, "mont_mul_per_s": mul, "mod_add_per_s": add}), flush=True)
