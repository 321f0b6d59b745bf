"""GPU parity of flat relations: CUDA path (through the C ABI) vs the C/Python oracle.
Bit-exact verdicts, first failing assertion and wire values."""
import os

import numpy as np
import pytest

from tests.util import FIELDS, circuits, random_flat_program, zkb

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import flat
    return flat


def _check(c, gates, pool, p, inst, wit, n_batch, n_wires, sample=(0,), keep_all=True):
    z = zkb()
    flat = _oracle()
    eb = c.elem_bytes(p)
    b = z.GpuBackend(0)
    b.set_field(p)
    b.push_gates(gates, pool)
    b.finalize(keep_all_values=keep_all)
    v = b.evaluate(inst, wit, n_batch)
    ref = flat.eval_batch(gates, pool, p.to_bytes(eb, "little"), inst, wit, n_batch, n_threads=4)
    for j in range(n_batch):
        ok_ref = int(ref[j]["status"]) == flat.EV_TRUE
        assert bool(v[j]["ok"]) == ok_ref, (j, v[j], ref[j])
        if not ok_ref:
            assert int(ref[j]["status"]) == flat.EV_ASSERT_FAILED
            assert int(v[j]["first_fail_seq"]) == int(ref[j]["fail_assert_seq"]), (j, v[j], ref[j])
            assert b.assert_wire(int(v[j]["first_fail_seq"])) == int(ref[j]["fail_wire"])
    for j in sample:
        if int(ref[j]["status"]) != flat.EV_TRUE:
            continue   # wires after the first failure are undefined in the reference
        _, dump = flat.eval_dump(gates, pool, p.to_bytes(eb, "little"), None if inst is None else (inst if inst.ndim == 2 else inst[j]),
                                 None if wit is None else wit[j], n_wires, stride=eb)
        live = [i for i in range(n_wires) if not (dump[i] == 0xFF).all() or p == (1 << 256) - 1]
        if keep_all:
            vals = b.read_values(j, [b.scope_lookup(i) for i in live], eb)
            for i, val in zip(live, vals):
                assert val == int.from_bytes(dump[i].tobytes(), "little"), (j, i)
    st = b.stats()
    b.close()
    return v, ref, st


@pytest.mark.parametrize("name", list(FIELDS))
@pytest.mark.parametrize("n_batch", [1, 37])
def test_random_circuit_all_fields(name, n_batch):
    c = circuits()
    p = FIELDS[name]
    circ = c.random_circuit(3000, 48, p, seed=11 + n_batch, n_tracked=6)
    corrupt = {0: 1} if n_batch == 1 else {2: 0, 5: 3, 36: 5}
    w = c.make_witnesses(circ, n_batch, seed=5, corrupt=corrupt)
    v, ref, _ = _check(c, circ.gates, circ.const_pool, p, None, w, n_batch, circ.n_wires,
                       sample=(0, 1) if n_batch > 1 else (0,))
    if p >= (1 << 31):   # over a tiny field a corrupted input can cancel by coincidence (1 / p per assertion); the oracle decides
        exp = c.expected_first_fail(circ, n_batch, corrupt)
        for j in range(n_batch):
            assert (int(v[j]["first_fail_seq"]) if not v[j]["ok"] else -1) == exp[j]


def test_true_single_witness_values():
    c = circuits()
    p = FIELDS["bls381"]
    circ = c.random_circuit(5000, 32, p, seed=3, n_tracked=4)
    w = c.make_witnesses(circ, 1, seed=9)
    v, _, st = _check(c, circ.gates, circ.const_pool, p, None, w, 1, circ.n_wires)
    assert v[0]["ok"] == 1
    assert st["ir_gates"] == 5000


@pytest.mark.parametrize("name", ["p101", "m31", "goldilocks", "p61", "p64full", "kat124", "bn254", "bls381", "p256full"])
@pytest.mark.parametrize("flow_min", ["1", "256"])
@pytest.mark.parametrize("window", [0, 64])
def test_dataflow_launch_single_witness(name, flow_min, window, monkeypatch):
    """one witness: every wavefront in ONE barrier-free launch (k_levels_flow for 1- / 2-limb fields: gates poll their
    operand words until the all-ones marker is gone; k_levels_flow_wide for 4- / 8-limb fields: a flag word per slot).  Every gate kind, every wire value and the first failing assertion
    against the oracle; ZKB_FLOW_MIN=1 sends even the narrowest program through it, a windowed circuit makes it deep
    (hundreds of wavefronts, producers and consumers in neighbouring warps); ZKB_FLOW=0 must give the same answers."""
    monkeypatch.setenv("ZKB_FLOW_MIN", flow_min)
    monkeypatch.setenv("ZKB_FLOW", "1")
    monkeypatch.setenv("ZKB_FLOW_WIDE", "1")      # the flag-word variant is opt-in (slower than the barrier kernels): still exact
    c = circuits()
    p = FIELDS[name]
    for seed, corrupt in ((41, {}), (42, {0: 2})):
        circ = c.random_circuit(6000, 48, p, seed=seed, n_tracked=6, window=window)
        w = c.make_witnesses(circ, 1, seed=seed + 1, corrupt=corrupt)
        v, ref, st = _check(c, circ.gates, circ.const_pool, p, None, w, 1, circ.n_wires)
        monkeypatch.setenv("ZKB_FLOW", "0")
        v0, _, st0 = _check(c, circ.gates, circ.const_pool, p, None, w, 1, circ.n_wires)
        monkeypatch.setenv("ZKB_FLOW", "1")
        assert (int(v[0]["ok"]), int(v[0]["first_fail_seq"])) == (int(v0[0]["ok"]), int(v0[0]["first_fail_seq"]))
    gates, pool, n_wires = random_flat_program(p, 1500, 5, 9, seed=7, bool_ops=True)   # And / Xor / Not, Free + id re-use
    rng = np.random.default_rng(3)
    inst = c.random_field_elements(rng, (5,), p)
    wit = c.random_field_elements(rng, (1, 9), p)
    _check(c, gates, pool, p, inst, wit, 1, n_wires)


@pytest.mark.parametrize("name", ["goldilocks", "bn254"])
def test_all_levels_launches_beyond_the_cached_wavefront_table(name, monkeypatch):
    """2493 wavefronts: more than the 2048 wavefront offsets the all-levels kernels keep in shared memory (the rest is read from
    global memory) — the dataflow launch (forced on, flag-word variant included) and the barrier kernels against the oracle"""
    c = circuits()
    p = FIELDS[name]
    circ = c.random_circuit(20000, 48, p, seed=41, n_tracked=6, window=3)
    w = c.make_witnesses(circ, 1, seed=5, corrupt={0: 3})
    for env in ({"ZKB_FLOW": "1", "ZKB_FLOW_MIN": "1", "ZKB_FLOW_WIDE": "1"}, {"ZKB_FLOW": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        v_, ref, st = _check(c, circ.gates, circ.const_pool, p, None, w, 1, circ.n_wires)
        assert st["n_levels"] > 2048


@pytest.mark.parametrize("tile_log2", ["0", "2", "5"])
def test_multi_tile_batches(tile_log2, monkeypatch):
    # force small tiles so a batch takes several passes, with a ragged last tile
    monkeypatch.setenv("ZKB_TILE_LOG2", tile_log2)
    c = circuits()
    p = FIELDS["bn254"]
    circ = c.random_circuit(1500, 40, p, seed=21, n_tracked=5)
    n_batch = 77
    corrupt = {0: 0, 31: 2, 32: 4, 76: 1}
    w = c.make_witnesses(circ, n_batch, seed=6, corrupt=corrupt)
    v, _, st = _check(c, circ.gates, circ.const_pool, p, None, w, n_batch, circ.n_wires, sample=(1, 33, 75))
    assert st["tile_witnesses"] == 1 << int(tile_log2)


@pytest.mark.parametrize("name", ["p101", "goldilocks", "bls381", "p256full"])
@pytest.mark.parametrize("bool_ops", [False, True])
def test_every_gate_kind_with_wire_reuse(name, bool_ops):
    c = circuits()
    p = FIELDS[name]
    eb = c.elem_bytes(p)
    gates, pool, n_wires = random_flat_program(p, 1200, 5, 7, seed=hash(name) % 1000 + bool_ops, bool_ops=bool_ops)
    rng = np.random.default_rng(4)
    n_batch = 9
    inst = c.random_field_elements(rng, (5,), p)            # shared instance
    wit = c.random_field_elements(rng, (n_batch, 7), p)
    _check(c, gates, pool, p, inst, wit, n_batch, n_wires, sample=(0, 8))
    inst2 = c.random_field_elements(rng, (n_batch, 5), p)   # per-witness instances
    _check(c, gates, pool, p, inst2, wit, n_batch, n_wires, sample=(3,))


def test_binary_field_flat():
    c = circuits()
    p = 2
    gates, pool, n_wires = random_flat_program(p, 2000, 6, 10, seed=77, bool_ops=True)
    rng = np.random.default_rng(1)
    n_batch = 70
    inst = rng.integers(0, 2, size=(n_batch, 6, 1), dtype=np.uint8)
    wit = rng.integers(0, 2, size=(n_batch, 10, 1), dtype=np.uint8)
    _check(c, gates, pool, p, inst, wit, n_batch, n_wires, sample=(0, 31, 32, 69))


@pytest.mark.parametrize("n_batch", [128, 200, 1000])
def test_binary_field_wide_tiles_four_words_per_thread(n_batch):
    """>= 128 witnesses: the Boolean kernel takes four 32-witness words per thread (16-byte vectors); ragged tiles, raw
    input values >= 2 (trap 1) in some witnesses, assertions failing in different words"""
    c = circuits()
    p = 2
    gates, pool, n_wires = random_flat_program(p, 3000, 6, 10, seed=78 + n_batch, bool_ops=True)
    rng = np.random.default_rng(n_batch)
    inst = rng.integers(0, 2, size=(n_batch, 6, 1), dtype=np.uint8)
    wit = rng.integers(0, 2, size=(n_batch, 10, 1), dtype=np.uint8)
    for j in (1, 33, 97, n_batch - 1):
        wit[j, 3, 0] = 2 + (j % 5)          # unreduced inputs: non-zero integers that are 0 or 1 mod 2
        inst[j, 1, 0] = 3
    _check(c, gates, pool, p, inst, wit, n_batch, n_wires, sample=(0, 1, 31, 32, 97, 127, n_batch - 1))


def test_structural_errors_match_reference_text():
    z = zkb()
    c = circuits()
    g = np.zeros(2, dtype=c.GATE_DTYPE)
    g["op"] = [c.G_WITNESS, c.G_ADD]
    g["out"] = [0, 1]
    g["a"] = [0, 0]
    g["b"] = [0, 7]
    b = z.GpuBackend(0)
    b.set_field(101)
    with pytest.raises(z.ZkbError) as e:
        b.push_gates(g)
    assert str(e.value) == "No value given for wire_7"
    assert b.pending_error() == "No value given for wire_7"
    b2 = z.GpuBackend(0)
    b2.set_field(101)
    g["b"] = [0, 0]
    g["out"] = [0, 0]
    with pytest.raises(z.ZkbError) as e:
        b2.push_gates(g)
    assert str(e.value) == "Wire_0 already has a value in this scope."


def test_unreduced_inputs_keep_raw_semantics():
    # SURVEY.md 8a trap 1: a witness equal to p fails AssertZero in the reference (raw integer != 0)
    z = zkb()
    c = circuits()
    flat = _oracle()
    p = 101
    g = np.zeros(5, dtype=c.GATE_DTYPE)
    g["op"] = [c.G_WITNESS, c.G_WITNESS, c.G_ADD, c.G_ASSERT_ZERO, c.G_ASSERT_ZERO]
    g["out"] = [0, 1, 2, 0, 0]
    g["a"] = [0, 0, 0, 2, 1]
    g["b"] = [0, 0, 1, 0, 0]
    wit = np.zeros((3, 2, 4), dtype=np.uint8)
    wit[0, 0, 0], wit[0, 1, 0] = 0, 0        # TRUE
    wit[1, 0, 0], wit[1, 1, 0] = 101, 0      # 101 + 0 = 0 mod p: first assert holds; TRUE
    wit[2, 0, 0], wit[2, 1, 0] = 0, 101      # second assert tests the RAW witness 101 != 0: FALSE at seq 1
    b = z.GpuBackend(0)
    b.set_field(p)
    b.push_gates(g)
    b.finalize(True)
    v = b.evaluate(None, wit, 3)
    ref = flat.eval_batch(g, None, bytes([101]), None, wit, 3)
    assert [int(r["status"]) for r in ref] == [0, 0, 1]
    assert [int(x["ok"]) for x in v] == [1, 1, 0]
    assert int(v[2]["first_fail_seq"]) == 1 == int(ref[2]["fail_assert_seq"])
    # read-back of an unreduced input returns the raw value, like Evaluator::get
    assert b.read_values(1, [b.scope_lookup(0)], 4) == [101]


def test_run_twice_device_resident():
    z = zkb()
    c = circuits()
    p = FIELDS["bls381"]
    circ = c.random_circuit(2000, 32, p, seed=5, n_tracked=4)
    w = c.make_witnesses(circ, 16, seed=1, corrupt={7: 1})
    b = z.GpuBackend(0)
    b.set_field(p)
    b.push_gates(circ.gates, circ.const_pool)
    b.finalize(False)
    b.upload_inputs(None, w, 16)
    v1 = b.run()
    v2 = b.run()
    assert (v1 == v2).all()
    assert circ.first_fail_of_input[1] >= 0 and int(v1[7]["first_fail_seq"]) == int(circ.first_fail_of_input[1])
    t = b.timing()
    assert t["level_launches"] > 0 and t["levels_ms"] > 0


def test_slot_reuse_keeps_results(monkeypatch):
    """liveness-based slot re-use (keep_all_values = 0): same verdicts and live wires with far fewer slots"""
    z = zkb()
    c = circuits()
    flat = _oracle()
    p = FIELDS["bn254"]
    circ = c.random_circuit(20000, 64, p, seed=31, n_tracked=8, window=256)
    # free most wires so they are not observable: everything below the last 300 wire ids
    g = np.zeros(1, dtype=c.GATE_DTYPE)
    g["op"], g["a"], g["b"] = c.G_FREE, 0, circ.n_wires - 300
    gates = np.concatenate([circ.gates, g])
    n_batch = 19
    corrupt = {4: 3, 18: 0}
    w = c.make_witnesses(circ, n_batch, seed=2, corrupt=corrupt)
    ref = flat.eval_batch(gates, circ.const_pool, p.to_bytes(32, "little"), None, w, n_batch, n_threads=4)
    slots = {}
    for reuse in ("0", "1"):
        monkeypatch.setenv("ZKB_SLOT_REUSE", reuse)
        b = z.GpuBackend(0)
        b.set_field(p)
        b.push_gates(gates, circ.const_pool)
        b.finalize(keep_all_values=False)
        v = b.evaluate(None, w, n_batch)
        slots[reuse] = b.stats()["n_slots"]
        for j in range(n_batch):
            assert bool(v[j]["ok"]) == (int(ref[j]["status"]) == flat.EV_TRUE)
            if not v[j]["ok"]:
                assert int(v[j]["first_fail_seq"]) == int(ref[j]["fail_assert_seq"])
        # live (not freed) wires stay readable and exact
        _, dump = flat.eval_dump(gates, circ.const_pool, p.to_bytes(32, "little"), None, w[0], circ.n_wires)
        live = list(range(circ.n_wires - 299, circ.n_wires))
        vals = b.read_values(0, [b.scope_lookup(i) for i in live], 32)
        assert vals == [int.from_bytes(dump[i].tobytes(), "little") for i in live]
        if reuse == "1":   # a freed wire's value is no longer kept on device
            with pytest.raises(z.ZkbError):
                b.read_values(0, [circ.n_inputs + 50], 32)
        b.close()
    assert slots["1"] < slots["0"] / 10, slots


@pytest.mark.parametrize("name,n_batch", [("goldilocks", 300), ("p61", 64), ("m31", 300), ("p101", 130), ("goldilocks", 63)])
def test_narrow_fields_packed_lanes(name, n_batch):
    """N = 1 / 2 limbs: tiles of >= 128 / 64 witnesses run 4 / 2 lanes per thread (16-byte vector loads); ragged batches,
    corrupted witnesses on both lanes of a pair, every gate kind incl. constants; smaller tiles take the unpacked kernel"""
    c = circuits()
    p = FIELDS[name]
    gates, pool, n_wires = random_flat_program(p, 1500, 5, 40, seed=n_batch)
    rng = np.random.default_rng(n_batch + 7)
    eb = c.elem_bytes(p)
    inst = c.random_field_elements(rng, (n_batch, 5), p)
    wit = c.random_field_elements(rng, (n_batch, 40), p)
    v, ref, st = _check(c, gates, pool, p, inst, wit, n_batch, n_wires, sample=(0, 1, n_batch - 1))
    if n_batch >= 128:
        assert st["tile_witnesses"] >= 128
    circ = c.random_circuit(4000, 64, p, seed=21, n_tracked=8)
    corrupt = {0: 1, 1: 4, 2: 0, 7: 7, n_batch - 1: 3, n_batch - 2: 2}
    w = c.make_witnesses(circ, n_batch, seed=6, corrupt=corrupt)
    v, ref, _ = _check(c, circ.gates, circ.const_pool, p, None, w, n_batch, circ.n_wires, sample=(0, 3, n_batch - 1))
    if p >= (1 << 31):   # over a tiny field a corrupted input can cancel by coincidence (1 / p per assertion); the oracle decides
        exp = c.expected_first_fail(circ, n_batch, corrupt)
        for j in range(n_batch):
            assert (int(v[j]["first_fail_seq"]) if not v[j]["ok"] else -1) == exp[j]


@pytest.mark.parametrize("tma", ["1", "0"])
@pytest.mark.parametrize("name,tile_log2", [("bls381", "8"), ("p256full", "8"), ("bls381", "9"), ("bn254", "10")])
def test_full_tiles_bulk_copy_ring_and_vector_loads(name, tile_log2, tma, monkeypatch):
    """Tiles of >= 256 lanes of an 8-limb field run k_level_tma (operand rows through cp.async.bulk + an mbarrier ring);
    ZKB_LEVEL_TMA=0 keeps them on k_level_pipe.  Full tiles and a ragged last one (lanes past the batch are masked), every
    arithmetic gate kind (constants are not bulk-copied), wire re-use, failing witnesses in each tile; wavefronts with more
    gates than the ring has stages and with fewer."""
    monkeypatch.setenv("ZKB_LEVEL_TMA", tma)
    monkeypatch.setenv("ZKB_TILE_LOG2", tile_log2)
    c = circuits()
    p = FIELDS[name]
    n_batch = 256 * 2 + 45
    circ = c.random_circuit(2500, 40, p, seed=77, n_tracked=5)
    corrupt = {0: 0, 255: 2, 256: 4, 511: 1, 512: 3, n_batch - 1: 0}
    w = c.make_witnesses(circ, n_batch, seed=8, corrupt=corrupt)
    v, _, st = _check(c, circ.gates, circ.const_pool, p, None, w, n_batch, circ.n_wires, sample=(1, 254, 257, 510, 513, n_batch - 2))
    assert st["tile_witnesses"] == 1 << int(tile_log2)
    gates, pool, n_wires = random_flat_program(p, 900, 4, 6, seed=5, bool_ops=False)
    rng = np.random.default_rng(12)
    inst = c.random_field_elements(rng, (4,), p)
    wit = c.random_field_elements(rng, (n_batch, 6), p)
    _check(c, gates, pool, p, inst, wit, n_batch, n_wires, sample=(0, 255, 256, n_batch - 1))
