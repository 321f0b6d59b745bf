"""Host check of the sliced (SELL-32-sigma) device layout zkb_r1cs_load builds: every row's terms, in order
A, B, C and per matrix first the coefficient-one terms, then the general ones, must be recoverable from the slices;
zero coefficients are dropped."""
import numpy as np
import pytest

from tests.util import FIELDS, circuits, zkb

T_PAD, T_ONE = 0xFFFFFFFF, 0xFFFFFFFE


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("n_rows,sigma", [(1, None), (31, None), (32, None), (1000, "64"), (5000, None)])
def test_sell_layout_reproduces_the_csr(n_rows, sigma, kind, monkeypatch):
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
    slices, terms, rows = b.r1cs_layout(kind)
    assert sorted(rows.tolist()) == list(range(n_rows))
    assert len(slices) == (n_rows + 31) // 32
    g = 0
    klass = lambda ci: T_PAD if ci == zero_idx else T_ONE if ci in (0, one2_idx) else ci
    for s, (g0, pa, pb, pc) in enumerate(slices.tolist()):
        assert g0 == g
        k = [pa & 0xFFFF, pa >> 16, pb & 0xFFFF, pb >> 16, pc & 0xFFFF, pc >> 16]   # A ones, A general, B ..., C ...
        g += sum(k)
        for i in range(32):
            pos = s * 32 + i
            col = terms[g0:g0 + sum(k), i]
            if pos >= n_rows:
                assert (col[:, 1] == T_PAD).all()
                continue
            row = int(rows[pos])
            off = 0
            for m, (rp, cc, ci) in enumerate((A, r.B, r.C)):
                lo, hi = int(rp[row]), int(rp[row + 1])
                tagged = [(int(cc[e]), klass(int(ci[e]))) for e in range(lo, hi)]
                if kind == 0:
                    ones = [t for t in tagged if t[1] == T_ONE]
                    general = [t for t in tagged if t[1] not in (T_ONE, T_PAD)]  # zero coefficients are dropped
                else:                                                            # one class per matrix, ones tagged inside
                    ones, general = [], [t for t in tagged if t[1] != T_PAD]
                for want, kq in ((ones, k[2 * m]), (general, k[2 * m + 1])):
                    assert len(want) <= kq
                    got = [tuple(x) for x in col[off:off + len(want)].tolist()]
                    assert got == want, (row, m, got, want)
                    assert (col[off + len(want):off + kq, 1] == T_PAD).all()
                    off += kq
    assert g == len(terms)
    # no device in this context: evaluation fails loudly
    with pytest.raises(z.ZkbError) as e:
        b.r1cs_upload(np.zeros((1, r.n_vars, 32), dtype=np.uint8))
    assert e.value.code in (z.ZKB_E_CUDA, z.ZKB_E_FATAL)
