#!/usr/bin/env python
"""Secondary configs of BASELINE.json (C1, C2, C4, C5, C3 as a `.sieve` statement): parity check + device timing (CUDA
events inside libzkb, inputs resident).  One JSON line per config.  The headline (C3) is bench.py.
Lives under tests/ because it uses the oracle — as workload generator (structured relations, FlatBuffers messages), as
the checker of every result it times and as the CPU figure printed beside them — never as the thing measured."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkb_loader  # noqa: E402

z = zkb_loader.load()
c = importlib.import_module("zkir_b200.circuits")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed_runs(fn, timing, k=5, warm=2):
    for _ in range(warm):
        fn()
    tot, lv = [], []
    for _ in range(k):
        fn()
        t = timing()
        tot.append(t["total_ms"])
        lv.append(t["levels_ms"])
    return float(np.median(tot)), float(np.median(lv))


def c1():
    from oracle import fixtures as fx, sieve_fbs as F
    msgs = [fx.example_instance(), fx.example_witness(), fx.example_relation()]
    t0 = time.perf_counter()
    ev = z.Evaluator.from_messages(z.Source.from_buffers([F.write_messages(msgs)]), device=0)
    v = ev.get_violations()
    dt = time.perf_counter() - t0
    bad = z.Evaluator.from_messages(z.Source.from_buffers([F.write_messages(
        [fx.example_instance(), fx.example_witness_incorrect(), fx.example_relation()])]), device=0).get_violations()
    st = ev.backend.stats()
    return {"config": "C1 zki_sieve example (p=101)", "violations": v, "incorrect_witness": bad, "wall_ms_incl_flatten": dt * 1e3,
            "callbacks": sum(st["callbacks"].values()), "levels": st["n_levels"], "device_ms": ev.backend.timing()["total_ms"]}


def c2(log2_gates, window):
    p = c.GOLDILOCKS
    circ = c.random_circuit(1 << log2_gates, 1024, p, 0x5EED0002, window=window)
    b = z.GpuBackend(0)
    b.set_field(p)
    t0 = time.perf_counter()
    b.push_gates(circ.gates, circ.const_pool)
    b.finalize(False)
    prep = time.perf_counter() - t0
    good = c.make_witnesses(circ, 1, seed=3)
    k = next(k for k in range(circ.n_tracked) if circ.first_fail_of_input[k] >= 0)   # a tracked input some assertion sees
    bad = c.make_witnesses(circ, 1, seed=3, corrupt={0: k})
    b.upload_inputs(None, bad, 1)
    vb = b.run()
    assert int(vb[0]["first_fail_seq"]) == int(circ.first_fail_of_input[k])
    b.upload_inputs(None, good, 1)
    assert int(b.run()[0]["ok"]) == 1
    tot, lv = timed_runs(b.run, b.timing)
    st = b.stats()
    algo = 8 * 3 * (circ.hist["add"] + circ.hist["mul"]) + 8 * circ.hist["assert_zero"] + 16 * st["n_device_ops"]
    # the latency floor of a one-launch evaluation is n_levels x the barrier between wavefronts,
    # measured here with kernels that only synchronise, at the CTA count the all-levels launch uses (launch_coop_n's rule)
    widest = max(b.level_info(l)["gates"] for l in range(st["n_levels"]))
    cluster = widest <= 8 * 512 * 2
    blocks = max(148, -(-(-(-widest // 256)) // 148) * 148)
    bar = {"cooperative_groups_grid_sync_us": b.debug_barrier_cost(0, blocks), "counter_barrier_us": b.debug_barrier_cost(1, blocks),
           "cluster_barrier_us": b.debug_barrier_cost(2), "ctas": 8 if cluster else blocks}
    used = bar["cluster_barrier_us"] if cluster else bar["counter_barrier_us"]
    return {"barrier": bar, "barrier_floor_ms": st["n_levels"] * used * 1e-3, "frac_of_barrier_floor": st["n_levels"] * used * 1e-3 / tot,
            "widest_level_gates": widest,
            "config": f"C2 2^{log2_gates}-gate Goldilocks, 1 witness, operands {'windowed ' + str(window) if window else 'global'}",
            "gates_per_s": circ.n_gates / (tot * 1e-3), "ms": tot, "levels_ms": lv, "levels": st["n_levels"], "us_per_level": tot * 1e3 / st["n_levels"],
            "us_per_level_levels_only": lv * 1e3 / st["n_levels"], "kernel_launches": b.timing()["kernel_launches"],
            "algo_GBps_incl_descriptors": algo / (tot * 1e-3) / 1e9, "frac_of_hbm_peak": algo / (tot * 1e-3) / 1e9 / PEAK,
            "prep_s": prep, "note": "L2-resident, launch/latency bound (one launch per wavefront)"}


def c4(log2_rows, log2_free, batch):
    p = c.BN254_FR
    r = c.random_r1cs(1 << log2_rows, 1 << log2_free, p, 0x5EED0004)
    t0 = time.perf_counter()
    zz = c.r1cs_assignment(r, 1)
    gen = time.perf_counter() - t0
    zb = c.assignment_bytes(zz, p)
    bad_row = (1 << log2_rows) // 3
    zbad = zb.copy()
    zbad[r.n_free + 1 + bad_row, 0] ^= 1
    b = z.GpuBackend(0)
    b.set_field(p)
    b.r1cs_load(r.A, r.B, r.C, r.coef_table, r.n_vars)
    v = b.r1cs_check(np.stack([zb, zbad]))
    assert int(v[0]["ok"]) == 1 and int(v[1]["first_fail_seq"]) == min(bad_row, c.r1cs_first_row_reading(r, r.n_free + 1 + bad_row))
    batch_z = np.broadcast_to(zb, (batch,) + zb.shape).copy() if batch > 1 else zb[None]
    b.r1cs_upload(batch_z)
    tot, lv = timed_runs(b.r1cs_run, b.timing)
    algo = c.r1cs_algorithmic_bytes(r) * batch
    ge = c.r1cs_gate_equivalent(r)
    # ceiling of the z look-ups: random 32-byte gathers from a table of z's size (one assignment), measured here
    gather = b.debug_gather_throughput(max(r.n_vars * 32, 1 << 20)) / 1e9 if batch == 1 else None
    z_bytes = 32 * r.nnz * batch
    extra = {}
    if gather:
        # the gathered sectors at the gather ceiling + the streamed descriptors at the copy peak
        floor_ms = (z_bytes / gather + (algo - z_bytes) / PEAK) / 1e9 * 1e3
        extra = {"random_gather_GBps": gather, "gather_bound_ms": floor_ms, "frac_of_gather_bound": floor_ms / lv}
    # the integer-pipe ceiling of the same work: N^2 = 64 wide multiply-adds per general term (plain product), 64 per linear
    # combination that holds one (its single Montgomery reduction), 128 for (A_r.z)(B_r.z); against the measured rate of
    # register-resident Montgomery products (128 wide multiply-adds each)
    wide_mads = 0
    for (rp, col, ci) in (r.A, r.B):
        general = (ci != 0)
        wide_mads += 64 * int(general.sum())
        wide_mads += 64 * int((np.add.reduceat(general.astype(np.int64), rp[:-1].astype(np.int64)) > 0).sum())
    wide_mads += 128 * r.n_rows
    mads_per_s = b.debug_field_throughput(1, 500) * 128
    imad_ms = wide_mads * batch / mads_per_s * 1e3
    extra.update({"wide_mads_per_assignment": wide_mads, "integer_pipe_bound_ms": imad_ms, "frac_of_integer_pipe_bound": imad_ms / lv})
    return {**extra, "config": f"C4 R1CS 2^{log2_rows} constraints, BN254, batch {batch}", "constraints_per_s": r.n_rows * batch / (lv * 1e-3),
            "gate_equivalent_per_s": ge * batch / (lv * 1e-3), "check_kernel_ms": lv, "total_ms_incl_z_conversion": tot, "nnz": r.nnz,
            "n_vars": r.n_vars, "algo_GBps": algo / (lv * 1e-3) / 1e9, "frac_of_hbm_peak": algo / (lv * 1e-3) / 1e9 / PEAK,
            "assignment_gen_s": gen}


def c4_rows_sharded(log2_rows, log2_free):
    """SURVEY 8(e), one assignment: the constraints sharded in contiguous row blocks over every visible GPU (z replicated), one host
    thread per device, first violated row = MIN over the blocks.  Reports the slowest device's check kernel and the wall time of
    the concurrent runs beside the one-device figure of the same run."""
    import threading
    import torch
    shard = importlib.import_module("zkir_b200.sharding")
    n_dev = torch.cuda.device_count()
    p = c.BN254_FR
    r = c.random_r1cs(1 << log2_rows, 1 << log2_free, p, 0x5EED0004)
    zz = c.r1cs_assignment(r, 1)
    zb = c.assignment_bytes(zz, p)
    bad_row = (1 << log2_rows) // 3
    zbad = zb.copy()
    zbad[r.n_free + 1 + bad_row, 0] ^= 1
    want_bad = min(bad_row, c.r1cs_first_row_reading(r, r.n_free + 1 + bad_row))
    out = {}
    for world in sorted({1, n_dev}):
        backs, row0 = [], []
        for rank in range(world):
            A, B, Cm, lo = shard.shard_r1cs_rows(r.A, r.B, r.C, rank, world)
            b = z.GpuBackend(rank)
            b.set_field(p)
            b.r1cs_load(A, B, Cm, r.coef_table, r.n_vars)
            backs.append(b)
            row0.append(lo)
        # parity: the satisfying assignment and the corrupted one, MIN over the blocks
        for zv, exp in ((zb, -1), (zbad, want_bad)):
            parts = [shard.global_first_row(b.r1cs_check(zv[None]), lo) for b, lo in zip(backs, row0)]
            ff = np.minimum.reduce(parts)
            assert (-1 if ff[0] == shard.NO_FAIL else int(ff[0])) == exp
        for b in backs:
            b.r1cs_upload(zb[None])
        def one(b):
            b.r1cs_run()
        for _ in range(2):
            for b in backs:
                one(b)
        walls, kern = [], []
        for _ in range(5):
            th = [threading.Thread(target=one, args=(b,)) for b in backs]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            walls.append((time.perf_counter() - t0) * 1e3)
            kern.append(max(b.timing()["levels_ms"] for b in backs))
        out[world] = {"slowest_check_kernel_ms": float(np.median(kern)), "wall_ms_concurrent_runs": float(np.median(walls))}
        for b in backs:
            b.close()
    res = {"config": f"C4 R1CS 2^{log2_rows} constraints, BN254, 1 assignment, rows sharded over {n_dev} GPU(s)", "devices": n_dev,
           "one_device": out[1], "all_devices": out[n_dev]}
    if n_dev > 1:
        res["kernel_speedup"] = out[1]["slowest_check_kernel_ms"] / out[n_dev]["slowest_check_kernel_ms"]
    return res


def c5(lo, li, batch):
    from oracle import ir, sieve_fbs as F, workloads as wl
    n_wit = 4096
    rel, n_leaf = wl.boolean_for_relation(lo, li, n_wit)
    buf = F.write_messages([ir.Witness(rel.header, [b"\0"] * n_wit), rel])
    b = z.GpuBackend(0)
    ev = z.Evaluator(b)
    t0 = time.perf_counter()
    ev.ingest_source(z.Source.from_buffers([buf]))
    flat_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    assert ev.get_violations() == []
    first_s = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    W = rng.integers(0, 2, size=(batch, n_wit, 1)).astype(np.uint8)
    v = b.evaluate(None, W, batch)
    block = 2 << li
    for j in range(min(batch, 4)):
        outs = wl.boolean_for_expected_outputs(W[j, :, 0], lo, li)
        x = 0
        for k in range(8):
            x ^= int(outs[-1][k])
        assert bool(v[j]["ok"]) == (x == 0)
        probe = [n_wit + k for k in (0, 1, 12345 % (block << lo), block * 3 + 7)]
        got = b.read_values(j, [ev.value_handle(w_) for w_ in probe], 4)
        assert got == [int(outs.reshape(-1)[w_ - n_wit]) for w_ in probe]
    # the groups' templates are being compiled into a specialised kernel in the background (group_jit.cpp): steady state =
    # after that has ended; ZKB_GROUP_JIT=0 times the interpreter kernel instead
    jit_state, jit_s = b.wait_group_jit()
    b.upload_inputs(None, W, batch)
    tot, lv = timed_runs(b.run, b.timing)
    st = b.stats()
    # unrolled, one 16-byte GateOp + 3 operands of 1/8 byte per gate = 16.4 B/gate at one witness.
    # Loop-structured (csrc/program.h CallGroup): a call moves its 3 inputs and 2 outputs (one 32-witness word each per tile
    # word) and nothing else — the 8-gate body lives in registers, the 96-byte descriptor stands for 2^li calls.
    words = max(1, st["tile_witnesses"] // 32)
    grouped_bytes = st["n_group_calls"] * (3 + 2) * 4 * words + 96 * st["n_call_groups"] + (16 + 12 * words) * st["n_device_ops"]
    return {"config": f"C5 Boolean 2^{lo + li + 3} leaf gates (nested For 2^{lo} x 2^{li}), {batch} witness(es) bit-sliced",
            "gates_per_s": n_leaf * batch / (tot * 1e-3), "ms": tot, "levels_ms": lv, "levels": st["n_levels"], "host_flatten_s": flat_s,
            "finalize_and_first_eval_s": first_s, "host_prep_s": flat_s + first_s, "values": st["n_values"],
            "call_groups": st["n_call_groups"], "group_calls": st["n_group_calls"], "device_ops": st["n_device_ops"],
            "kernel_launches": b.timing()["kernel_launches"],
            "group_kernel": "specialised at run time (NVRTC)" if st["group_jit_state"] == 3 else "interpreter (k_bool_groups)",
            "group_kernel_compile_s_background": jit_s,
            "descriptor_form_bytes_per_gate": 16 + 3 / 8, "grouped_bytes_per_gate": grouped_bytes / n_leaf,
            "device_GBps": grouped_bytes / (tot * 1e-3) / 1e9,
            "descriptor_form_GBps_equivalent": (16 + 3 / 8) * n_leaf / (tot * 1e-3) / 1e9}


def c3_sieve(log2_gates):
    """`zki_sieve evaluate` drop-in timing: a flat 2^k-gate BLS12-381 statement from `.sieve` files on disk to the verdict
    (Source -> FlatBuffers reader -> Evaluator mirror -> levelizer -> device), one witness; the CPU restatement
    (oracle/plaintext_flat.c, 1 thread, gate array already in memory: no parsing) timed beside it."""
    import tempfile
    from oracle import flat, ir, sieve_fbs as F
    p = c.BLS12_381_FR
    circ = c.random_circuit(1 << log2_gates, 1024, p, 0x5EED0003)
    w = c.make_witnesses(circ, 1, seed=11)
    host = z.GpuBackend(-1)
    rel = host.write_flat_relation(p, circ.gates, circ.const_pool)
    h = ir.Header(p.to_bytes(32, "little"))
    wit = F.write_message(ir.Witness(h, [bytes(w[0, i]) for i in range(w.shape[1])]))
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "001_witness.sieve"), "wb").write(wit)
        open(os.path.join(d, "002_relation.sieve"), "wb").write(rel)
        best = None
        for _ in range(3):
            tc = time.perf_counter()
            be = z.GpuBackend(0)                 # CUDA context creation: paid once per process, reported apart
            t0 = time.perf_counter()
            e = z.Evaluator(be)
            e.ingest_source(z.Source.from_directory(d))
            t1 = time.perf_counter()
            v = e.get_violations()
            t2 = time.perf_counter()
            assert v == []
            cur = {"cuda_context_s": t0 - tc, "ingest_s": t1 - t0, "levelize_upload_evaluate_s": t2 - t1, "total_s": t2 - t0,
                   "device_ms": e.backend.timing()["total_ms"]}
            if best is None or cur["total_s"] < best["total_s"]:
                best = cur
            e.close()
    t0 = time.perf_counter()
    ref = flat.eval_batch(circ.gates, circ.const_pool, p.to_bytes(32, "little"), None, w, 1, n_threads=1)
    cpu_s = time.perf_counter() - t0
    assert int(ref[0]["status"]) == flat.EV_TRUE
    return {"config": f"C3 statement 2^{log2_gates} gates from .sieve files, BLS12-381, 1 witness (zki_sieve evaluate drop-in)",
            "sieve_bytes": len(rel) + len(wit), "messages": len(F.split_messages(rel)) + 1, **best,
            "gates_per_s_end_to_end": circ.n_gates / best["total_s"],
            "cpu_restatement_1_thread_s_no_parsing": cpu_s, "cpu_gates_per_s": circ.n_gates / cpu_s,
            "parse_threads": int(os.environ.get("ZKB_PARSE_THREADS", "0")) or "default min(16, cores)"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c1,c2,c4,c5")
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    todo = a.only.split(",")
    if "c1" in todo:
        print(json.dumps(c1()), flush=True)
    if "c2" in todo:
        print(json.dumps(c2(16 if a.small else 20, 0)), flush=True)
        print(json.dumps(c2(16 if a.small else 20, 4096)), flush=True)
    if "c4" in todo:
        print(json.dumps(c4(14 if a.small else 22, 12 if a.small else 20, 1)), flush=True)
        print(json.dumps(c4(14 if a.small else 18, 12 if a.small else 16, 64)), flush=True)
    if "c3s" in todo:
        print(json.dumps(c3_sieve(14 if a.small else 22)), flush=True)
    if "c3s24" in todo:      # the headline relation as a statement: 168 messages of 100 000 gates
        print(json.dumps(c3_sieve(14 if a.small else 24)), flush=True)
    if "c4rows" in todo:     # explicit only: wants a multi-GPU box (gpurun --gpus N)
        print(json.dumps(c4_rows_sharded(14 if a.small else 22, 12 if a.small else 20)), flush=True)
    if "c5" in todo:
        print(json.dumps(c5(6 if a.small else 13, 6 if a.small else 10, 1)), flush=True)
        print(json.dumps(c5(6 if a.small else 13, 6 if a.small else 10, 64)), flush=True)
        print(json.dumps(c5(6 if a.small else 13, 6 if a.small else 10, 128 if a.small else 4096)), flush=True)
