"""flatten == evaluate on the device (consumers/flattening.rs:227-252): the statement our flatten writes evaluates
on the GPU to the verdict of the original, wire value for wire value (a flattened gate's output wire id is its
position in the oracle's callback trace)."""
import pytest

from oracle import evaluator as ev
from oracle import fixtures as fx
from oracle import sieve_fbs as F
from tests.test_host_evaluator import STATEMENTS
from tests.util import zkb

pytestmark = pytest.mark.gpu


def flatten(msgs):
    z = zkb()
    e = z.Evaluator(flatten=True)
    e.ingest_source(z.Source.from_buffers([F.write_messages(msgs)]))
    return e.flatten()


@pytest.mark.parametrize("name", list(STATEMENTS))
def test_flattened_statement_evaluates_like_the_original(name):
    z = zkb()
    msgs = STATEMENTS[name]()
    bufs = flatten(msgs)
    b = z.GpuBackend(0)
    e = z.Evaluator(b)
    e.ingest_source(z.Source.from_buffers(list(bufs)))
    assert e.get_violations() == ev.evaluate(msgs) == []
    # every wire of the flattened relation is still live in the top scope (no Free): value = the oracle's trace value
    tb = ev.TracingBackend()
    ev.Evaluator.from_messages(msgs, tb)
    for wid in range(0, len(tb.trace), max(1, len(tb.trace) // 64)):
        assert e.get(wid) == tb.trace[wid][2] % tb.m, wid


def test_flattened_incorrect_witness_fails_on_the_device():
    z = zkb()
    msgs = [fx.example_instance(), fx.example_witness_incorrect(), fx.example_relation()]
    bufs = flatten(msgs)
    e = z.Evaluator(z.GpuBackend(0))
    e.ingest_source(z.Source.from_buffers(list(bufs)))
    got = e.get_violations()
    want = ev.evaluate(F.read_messages(bufs[0]) + F.read_messages(bufs[1]) + F.read_messages(bufs[2]))
    assert got == want and len(got) == 1 and got[0].startswith("Wire_")
