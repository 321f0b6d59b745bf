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


def test_r1cs_load_rejects_malformed_systems():
    """argument validation of zkb_r1cs_load (host-only context): the reference's own error text for an undefined variable
    (from_r1cs.rs:90), plain argument errors for the rest"""
    c = circuits()
    z = zkb()
    p = FIELDS["bn254"]
    r = c.random_r1cs(50, 10, p, seed=1)

    def load(A=None, B=None, C=None, table=None, n_vars=None, field=True):
        b = z.GpuBackend(-1)
        if field:
            b.set_field(p)
        b.r1cs_load(A or r.A, B or r.B, C or r.C, r.coef_table if table is None else table, r.n_vars if n_vars is None else n_vars)
        return b

    load()
    with pytest.raises(z.ZkbError) as e:                              # set_field first
        load(field=False)
    assert e.value.code == z.ZKB_E_ARG
    bad_col = (r.A[0], r.A[1].copy(), r.A[2])
    bad_col[1][3] = r.n_vars + 5
    with pytest.raises(z.ZkbError) as e:
        load(A=bad_col)
    assert e.value.code == z.ZKB_E_SEMANTIC and str(e.value) == f"The WireId {r.n_vars + 5} has not been defined yet."
    bad_ci = (r.B[0], r.B[1], r.B[2].copy())
    bad_ci[2][0] = len(r.coefs) + 1
    with pytest.raises(z.ZkbError) as e:
        load(B=bad_ci)
    assert e.value.code == z.ZKB_E_ARG
    bad_rp = (r.C[0].copy(), r.C[1], r.C[2])
    bad_rp[0][10] = bad_rp[0][9] - 1 if bad_rp[0][9] > 0 else 10 ** 6
    with pytest.raises(z.ZkbError) as e:
        load(C=bad_rp)
    assert e.value.code == z.ZKB_E_ARG
    short = (r.C[0][:-1], r.C[1][:-1], r.C[2][:-1])                   # one row fewer than A and B
    with pytest.raises(z.ZkbError) as e:
        load(C=short)
    assert e.value.code == z.ZKB_E_ARG
    with pytest.raises(z.ZkbError) as e:
        load(n_vars=0)
    assert e.value.code == z.ZKB_E_ARG
    b = z.GpuBackend(-1)
    b.set_field(2)
    with pytest.raises(z.ZkbError) as e:
        b.r1cs_load(r.A, r.B, r.C, r.coef_table, r.n_vars)
    assert e.value.code == z.ZKB_E_UNSUPPORTED
