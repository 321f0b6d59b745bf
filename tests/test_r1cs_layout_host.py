"""Host check of the sliced (SELL-32-sigma) device layout zkb_r1cs_load builds: every row's terms, in order
A, B, C, must be recoverable from the slices; zero coefficients are dropped, ones are tagged."""
import numpy as np
import pytest

from tests.util import FIELDS, circuits, zkb

T_PAD, T_ONE = 0xFFFFFFFF, 0xFFFFFFFE


@pytest.mark.parametrize("n_rows,sigma", [(1, None), (31, None), (32, None), (1000, "64"), (5000, None)])
def test_sell_layout_reproduces_the_csr(n_rows, sigma, monkeypatch):
    if sigma:
        monkeypatch.setenv("ZKB_R1CS_SIGMA", sigma)
    c = circuits()
    z = zkb()
    p = FIELDS["bn254"]
    r = c.random_r1cs(n_rows, 40, p, seed=n_rows)
    # a zero coefficient and a coefficient >= p that reduces to one
    table = np.concatenate([r.coef_table, c.le_bytes(0, 32)[None], c.le_bytes(p + 1, 32)[None]])
    zero_idx, one2_idx = len(r.coefs), len(r.coefs) + 1
    A = (r.A[0], r.A[1], r.A[2].copy())
    A[2][::7] = zero_idx
    A[2][3::11] = one2_idx
    b = z.GpuBackend(-1)
    b.set_field(p)
    b.r1cs_load(A, r.B, r.C, table, r.n_vars)
    slices, terms, rows = b.r1cs_layout()
    assert sorted(rows.tolist()) == list(range(n_rows))
    assert len(slices) == (n_rows + 31) // 32
    g = 0
    klass = lambda ci: T_PAD if ci == zero_idx else T_ONE if ci in (0, one2_idx) else ci
    for s, (g0, ka, kb, kc) in enumerate(slices.tolist()):
        assert g0 == g
        g += ka + kb + kc
        for i in range(32):
            pos = s * 32 + i
            col = terms[g0:g0 + ka + kb + kc, i]
            if pos >= n_rows:
                assert (col[:, 1] == T_PAD).all()
                continue
            row = int(rows[pos])
            off = 0
            for (rp, cc, ci), k in ((A, ka), (r.B, kb), (r.C, kc)):
                lo, hi = int(rp[row]), int(rp[row + 1])
                assert hi - lo <= k
                want = [(0 if klass(int(ci[e])) == T_PAD else int(cc[e]), klass(int(ci[e]))) for e in range(lo, hi)]
                got = [tuple(x) for x in col[off:off + hi - lo].tolist()]
                assert got == want, (row, got, want)
                assert (col[off + hi - lo:off + k, 1] == T_PAD).all()
                off += k
    assert g == len(terms)
    # no device in this context: evaluation fails loudly
    with pytest.raises(z.ZkbError) as e:
        b.r1cs_upload(np.zeros((1, r.n_vars, 32), dtype=np.uint8))
    assert e.value.code in (z.ZKB_E_CUDA, z.ZKB_E_FATAL)
