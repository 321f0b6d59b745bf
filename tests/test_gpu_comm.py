"""Multi-GPU behind the C ABI (include/zkb.h section 7): one program levelized once, replicated by
zkb_comm_broadcast_program, the batch split into contiguous blocks, verdicts MIN-reduced.

A sharded run must give the CONCATENATION of what one context gives for the same batch — verdict, first failing assertion
and every wire value — and the oracle decides what that is.  Contexts sharing device 0 use the in-process transport (NCCL
refuses duplicate GPUs), so the replica path is covered on a one-GPU box; with two devices the same tests run over NCCL,
in one process (ncclCommInitAll) and as one process per device (ncclCommInitRank)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.util import FIELDS, ROOT, circuits, random_flat_program, zkb

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _oracle_batch(c, gates, pool, p, inst, wit, n):
    from oracle import flat
    eb = c.elem_bytes(p)
    return flat, flat.eval_batch(gates, pool, p.to_bytes(eb, "little"), inst, wit, n, n_threads=4)


def _compare_with_oracle(flat, v, ref, backend_for_wire):
    for j in range(len(v)):
        ok_ref = int(ref[j]["status"]) == flat.EV_TRUE
        assert bool(v[j]["ok"]) == ok_ref, (j, v[j], ref[j])
        if not ok_ref:
            assert int(v[j]["first_fail_seq"]) == int(ref[j]["fail_assert_seq"]), (j, v[j], ref[j])
            assert backend_for_wire.assert_wire(int(v[j]["first_fail_seq"])) == int(ref[j]["fail_wire"])


@pytest.mark.parametrize("devices", [[0, 0, 0], [0, 0], [0], "all"])
@pytest.mark.parametrize("name", ["goldilocks", "bls381"])
def test_sharded_equals_single_context_and_oracle(devices, name):
    z, c = zkb(), circuits()
    if devices == "all":
        if _n_devices() < 2:
            pytest.skip("needs two devices (NCCL transport)")
        devices = list(range(_n_devices()))
    p = FIELDS[name]
    eb = c.elem_bytes(p)
    circ = c.random_circuit(4000, 48, p, seed=31, n_tracked=6)
    n_batch = 61                                                # ragged blocks: 61 = 20 + 20 + 21
    corrupt = {0: 1, 19: 0, 20: 3, 40: 2, 60: 5}                # first / last witness of every block
    w = c.make_witnesses(circ, n_batch, seed=8, corrupt=corrupt)

    one = z.GpuBackend(0)
    one.set_field(p)
    one.push_gates(circ.gates, circ.const_pool)
    one.finalize(keep_all_values=True)
    v_one = one.evaluate(None, w, n_batch)

    sh = z.ShardedBackend(devices)
    sh.root.set_field(p)
    sh.root.push_gates(circ.gates, circ.const_pool)
    sh.root.finalize(keep_all_values=True)
    v = sh.evaluate(None, w, n_batch)
    info = sh.root.comm_info()
    assert info["n_ranks"] == len(devices)
    assert info["transport"] == ("single" if len(devices) == 1 else "in-process" if len(set(devices)) < len(devices) else "nccl")
    assert (v["ok"] == v_one["ok"]).all() and (v["first_fail_seq"] == v_one["first_fail_seq"]).all()
    flat, ref = _oracle_batch(c, circ.gates, circ.const_pool, p, None, w, n_batch)
    for b in sh.backends:                                       # replicas answer assert_info like the root
        _compare_with_oracle(flat, v, ref, b)
    assert (sh.run()["first_fail_seq"] == v["first_fail_seq"]).all()      # resident inputs, second pass

    # wire values on every rank, including the replicas: first and last witness of each block
    handles = [one.scope_lookup(i) for i in range(0, circ.n_wires, 7)]
    for r in range(len(devices)):
        lo, hi = sh.shard_of(r, n_batch)
        for j in (lo, hi - 1):
            if j in corrupt:
                continue
            b, local = sh.owner(j, n_batch)
            assert b is sh.backends[r]
            assert b.read_values(local, handles, eb) == one.read_values(j, handles, eb), (r, j)
    st0, st1 = sh.root.stats(), sh.backends[-1].stats()
    for k in ("n_values", "n_asserts", "n_witness", "n_consts", "ir_gates", "n_slots", "n_levels", "n_device_ops",
              "algo_bytes_per_witness", "nlimb", "callbacks"):
        assert st0[k] == st1[k], k
    # a replica cannot record, and says so
    if len(devices) > 1:
        with pytest.raises(z.ZkbError) as e:
            sh.backends[1].witness()
        assert "replica" in str(e.value)
    sh.close()
    one.close()


def test_sharded_every_gate_kind_with_instances_and_raw_inputs():
    """rare gates, shared and per-witness instances, unreduced inputs (the raw flags and raw constants travel too)"""
    z, c = zkb(), circuits()
    p = FIELDS["bn254"]
    eb = c.elem_bytes(p)
    gates, pool, n_wires = random_flat_program(p, 1500, 5, 7, seed=12, bool_ops=True)
    rng = np.random.default_rng(3)
    n_batch = 23
    wit = c.random_field_elements(rng, (n_batch, 7), p)
    wit[4, 2] = 0xFF                                            # a witness value >= p (trap 1)
    wit[22, 0] = 0xFF
    for inst in (c.random_field_elements(rng, (5,), p), c.random_field_elements(rng, (n_batch, 5), p)):
        sh = z.ShardedBackend([0, 0, 0, 0])
        sh.root.set_field(p)
        sh.root.push_gates(gates, pool)
        sh.root.finalize(keep_all_values=True)
        v = sh.evaluate(inst, wit, n_batch)
        flat, ref = _oracle_batch(c, gates, pool, p, inst, wit, n_batch)
        _compare_with_oracle(flat, v, ref, sh.backends[2])
        for j in (0, 4, 5, 22):
            if int(ref[j]["status"]) != flat.EV_TRUE:
                continue
            _, dump = flat.eval_dump(gates, pool, p.to_bytes(eb, "little"), inst if inst.ndim == 2 else inst[j], wit[j], n_wires, stride=eb)
            live = [i for i in range(n_wires) if not (dump[i] == 0xFF).all()]
            b, local = sh.owner(j, n_batch)
            got = b.read_values(local, [sh.root.scope_lookup(i) for i in live], 64)
            assert got == [int.from_bytes(dump[i].tobytes(), "little") for i in live], j
        sh.close()


@pytest.mark.parametrize("devices", [[0, 0, 0], "all"])
def test_call_groups_travel_with_the_program(devices):
    """a relation recorded as loop-structured call groups (C5's shape): the group descriptors, pre-decoded ops, hints and the
    implicit-value slots reach the replicas with the broadcast; verdicts = one context's, outputs readable on every rank"""
    from oracle import ir, sieve_fbs as F, workloads as wl
    z = zkb()
    if devices == "all":
        if _n_devices() < 2:
            pytest.skip("needs two devices (NCCL transport)")
        devices = list(range(_n_devices()))
    lo, li, n_wit = 4, 5, 256
    rel, _ = wl.boolean_for_relation(lo, li, n_wit)
    buf = F.write_messages([ir.Witness(rel.header, [b"\0"] * n_wit), rel])
    n_batch = 70
    rng = np.random.default_rng(9)
    W = rng.integers(0, 2, size=(n_batch, n_wit, 1)).astype(np.uint8)

    one = z.GpuBackend(0)
    e1 = z.Evaluator(one)
    e1.ingest_source(z.Source.from_buffers([buf]))
    one.finalize(keep_all_values=True)
    v_one = one.evaluate(None, W, n_batch)

    sh = z.ShardedBackend(devices)
    e = z.Evaluator(sh.root)
    e.ingest_source(z.Source.from_buffers([buf]))
    assert sh.root.stats()["n_call_groups"] == 1 << lo
    sh.root.finalize(keep_all_values=True)
    v = sh.evaluate(None, W, n_batch)
    assert (v["ok"] == v_one["ok"]).all() and (v["first_fail_seq"] == v_one["first_fail_seq"]).all()
    n_out = (1 << lo) * (2 << li)
    handles = [e.value_handle(n_wit + k) for k in range(0, n_out, 5)]
    for r in range(len(devices)):
        b_lo, b_hi = sh.shard_of(r, n_batch)
        for j in (b_lo, b_hi - 1):
            b, local = sh.owner(j, n_batch)
            outs = wl.boolean_for_expected_outputs(W[j, :, 0], lo, li).reshape(-1)
            assert b.read_values(local, handles, 4) == [int(outs[k]) for k in range(0, n_out, 5)], (r, j)
    st0, st1 = sh.root.stats(), sh.backends[-1].stats()
    for k in ("n_values", "n_call_groups", "n_group_calls", "n_group_launches", "n_slots", "n_device_ops"):
        assert st0[k] == st1[k], k
    sh.close()
    one.close()


def test_boolean_program_sharded():
    z, c = zkb(), circuits()
    gates, pool, n_wires = random_flat_program(2, 2000, 6, 10, seed=5, bool_ops=True)
    rng = np.random.default_rng(2)
    n_batch = 100
    inst = rng.integers(0, 2, size=(n_batch, 6, 1), dtype=np.uint8)
    wit = rng.integers(0, 2, size=(n_batch, 10, 1), dtype=np.uint8)
    sh = z.ShardedBackend([0, 0, 0])
    sh.root.set_field(2, is_boolean=True)
    sh.root.push_gates(gates, pool)
    sh.root.finalize()
    v = sh.evaluate(inst, wit, n_batch)
    flat, ref = _oracle_batch(c, gates, pool, 2, inst, wit, n_batch)
    _compare_with_oracle(flat, v, ref, sh.backends[1])
    sh.close()


def test_comm_argument_errors():
    z = zkb()
    sh = z.ShardedBackend([0, 0])
    with pytest.raises(z.ZkbError) as e:                        # nothing finalized on the root
        sh.evaluate(None, np.zeros((4, 1, 8), np.uint8), 4)
    assert e.value.code == z.ZKB_E_ARG
    sh.root.set_field(101)
    w0 = sh.root.witness()
    sh.root.assert_zero(w0, 0)
    sh.root.finalize()
    with pytest.raises(z.ZkbError):                             # fewer witnesses than ranks
        sh.evaluate(None, np.zeros((1, 1, 1), np.uint8), 1)
    v = sh.evaluate(None, np.array([[[0]], [[5]], [[0]]], dtype=np.uint8), 3)
    assert list(v["ok"]) == [1, 0, 1]
    with pytest.raises(z.ZkbError):                             # a context joins one communicator only
        sh.root._chk(z._lib.zkb_comm_init(sh._arr, 2))
    lone = z.GpuBackend(0)
    with pytest.raises(z.ZkbError) as e:
        lone.comm_run(0, 1)
    assert "communicator" in str(e.value)
    lone.close()
    sh.close()


_RANK_SCRIPT = r"""
import os, sys, numpy as np
sys.path.insert(0, {root!r})
import zkb_loader
z = zkb_loader.load()
import importlib
c = importlib.import_module("zkir_b200.circuits")
rank, world, idfile, outfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
import time
p = {p}
b = z.GpuBackend(rank % {ndev})
if rank == 0:
    cid = z.comm_unique_id()
    open(idfile + ".tmp", "wb").write(cid)
    os.replace(idfile + ".tmp", idfile)
else:
    t0 = time.time()
    while not os.path.exists(idfile):
        assert time.time() - t0 < 60
        time.sleep(0.05)
    cid = open(idfile, "rb").read()
b.comm_init_rank(cid, world, rank)
circ = c.random_circuit(3000, 32, p, seed=17, n_tracked=4)
n_total = 41
corrupt = {{0: 1, 20: 0, 40: 3}}
w = c.make_witnesses(circ, n_total, seed=4, corrupt=corrupt)
if rank == 0:                     # only the root ever sees the relation
    b.set_field(p)
    b.push_gates(circ.gates, circ.const_pool)
    b.finalize(keep_all_values=True)
b.comm_broadcast_program(0)
lo, hi = n_total * rank // world, n_total * (rank + 1) // world
v = b.comm_evaluate(None, w[lo:hi], hi - lo, lo, n_total)
v2 = b.comm_run(lo, n_total)
assert (v["first_fail_seq"] == v2["first_fail_seq"]).all()
np.save(outfile, v)
info = b.comm_info()
assert info["transport"] == "nccl" and info["n_ranks"] == world and info["nccl_version"] > 20000
b.close()
"""


@pytest.mark.parametrize("world", [1, 2])
def test_one_process_per_device(world, tmp_path):
    """zkb_comm_unique_id / zkb_comm_init_rank: the torchrun-style set-up, without torch"""
    ndev = _n_devices()
    if world > ndev:
        pytest.skip("needs two devices")
    z, c = zkb(), circuits()
    p = FIELDS["bls381"]
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT, p=p, ndev=ndev))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(tmp_path / "id"), str(tmp_path / f"v{r}.npy")],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(world)]
    outs = []
    for pr in procs:
        try:
            out, _ = pr.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out.decode("utf-8", "replace"))
    for r, pr in enumerate(procs):
        assert pr.returncode == 0, outs[r][-3000:]
    circ = c.random_circuit(3000, 32, p, seed=17, n_tracked=4)
    w = c.make_witnesses(circ, 41, seed=4, corrupt={0: 1, 20: 0, 40: 3})
    flat, ref = _oracle_batch(c, circ.gates, circ.const_pool, p, None, w, 41)
    one = z.GpuBackend(0)
    one.set_field(p)
    one.push_gates(circ.gates, circ.const_pool)
    one.finalize()
    for r in range(world):                                      # every rank holds the verdicts of the whole batch
        v = np.load(tmp_path / f"v{r}.npy")
        _compare_with_oracle(flat, v, ref, one)
    one.close()
