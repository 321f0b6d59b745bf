"""Random STRUCTURED relations (nested For / Switch / Call / AnonCall, Free, wire re-use, instance and witness
consumption inside functions and branches) for differential tests: oracle vs the C++ flattener (op for op,
no GPU) and vs the CUDA path (values, verdicts).  Programs are valid by construction; whether their
assertions hold depends on the data."""
from __future__ import annotations

import numpy as np

from oracle import ir


class Gen:
    def __init__(self, seed: int, p: int, boolean: bool = False, max_depth: int = 3):
        self.rng = np.random.default_rng(seed)
        self.p = p
        self.boolean = boolean
        self.max_depth = max_depth
        self.functions = []          # ir.Function, in definition order
        self.fn_sig = {}             # name -> (n_out, n_in, inst, wit)

    # ---- small helpers ------------------------------------------------------------------
    def ri(self, lo, hi):
        return int(self.rng.integers(lo, hi + 1))

    def chance(self, pr):
        return self.rng.random() < pr

    def pick(self, xs):
        return xs[int(self.rng.integers(0, len(xs)))]

    def const_bytes(self):
        if self.boolean:
            return bytes([self.ri(0, 1)])
        v = self.pick([0, 1, 2, self.p - 1, self.ri(0, min(self.p - 1, 1 << 30))])
        return ir.le_bytes(v)

    @staticmethod
    def wl(ids):
        """wire list: consecutive runs of >= 2 become WireRange"""
        out = []
        i = 0
        while i < len(ids):
            j = i
            while j + 1 < len(ids) and ids[j + 1] == ids[j] + 1:
                j += 1
            if j > i:
                out.append(ir.WireRange(ids[i], ids[j]))
            else:
                out.append(ir.Wire(ids[i]))
            i = j + 1
        return out

    # ---- a scope being generated -----------------------------------------------------------
    class Scope:
        def __init__(self, n_out, n_in):
            self.n_out = n_out
            self.avail = list(range(n_out, n_out + n_in))
            self.next = n_out + n_in
            self.freed = []
            self.gates = []
            self.inst = 0
            self.wit = 0

    def new_wire(self, sc, n=1):
        """n fresh CONSECUTIVE ids (re-using a freed id only for n == 1)"""
        if n == 1 and sc.freed and self.chance(0.6):
            return [sc.freed.pop(self.ri(0, len(sc.freed) - 1))]
        base = sc.next
        sc.next += n
        return list(range(base, base + n))

    def seed_scope(self, sc):
        if not sc.avail:
            w = self.new_wire(sc)[0]
            sc.gates.append(("Constant", w, self.const_bytes()))
            sc.avail.append(w)

    def simple_gate(self, sc, out=None):
        self.seed_scope(sc)
        a, b = self.pick(sc.avail), self.pick(sc.avail)
        w = out if out is not None else self.new_wire(sc)[0]
        if self.boolean:
            k = self.pick(["Xor", "And", "Not", "Copy", "Constant", "Instance", "Witness"])
        else:
            k = self.pick(["Add", "Mul", "AddConstant", "MulConstant", "Copy", "Constant", "Instance", "Witness", "Add", "Mul"])
        if k in ("Add", "Mul", "Xor", "And"):
            sc.gates.append((k, w, a, b))
        elif k in ("AddConstant", "MulConstant"):
            sc.gates.append((k, w, a, self.const_bytes()))
        elif k in ("Copy", "Not"):
            sc.gates.append((k, w, a))
        elif k == "Constant":
            sc.gates.append((k, w, self.const_bytes()))
        elif k == "Instance":
            sc.gates.append((k, w))
            sc.inst += 1
        else:
            sc.gates.append((k, w))
            sc.wit += 1
        if out is None:
            sc.avail.append(w)

    def assertion(self, sc):
        self.seed_scope(sc)
        t = self.pick(sc.avail)
        if self.chance(0.8):   # holds for every value
            if self.boolean:
                s = self.new_wire(sc)[0]
                sc.gates.append(("Xor", s, t, t))
            else:
                n = self.new_wire(sc)[0]
                sc.gates.append(("MulConstant", n, t, ir.le_bytes(self.p - 1)))
                s = self.new_wire(sc)[0]
                sc.gates.append(("Add", s, t, n))
                sc.avail.append(n)
            sc.avail.append(s)
            sc.gates.append(("AssertZero", s))
        else:                  # data dependent
            sc.gates.append(("AssertZero", t))

    def free_some(self, sc):
        cands = [w for w in sc.avail if w >= sc.n_out]
        if len(cands) < 4:
            return
        w = self.pick(cands)
        run = [w]
        while run[-1] + 1 in cands and len(run) < 3 and self.chance(0.5):
            run.append(run[-1] + 1)
        sc.gates.append(("Free", run[0], run[-1] if (len(run) > 1 or self.chance(0.3)) else None))
        for x in run:
            sc.avail.remove(x)
            sc.freed.append(x)

    # ---- bodies -------------------------------------------------------------------------------
    def body(self, n_out, n_in, depth, n_gates):
        """gate list of a sub-scope: inputs n_out..n_out+n_in-1 defined, outputs 0..n_out-1 assigned last"""
        sc = self.Scope(n_out, n_in)
        for _ in range(n_gates):
            self.step(sc, depth)
        for o in range(n_out):
            self.simple_gate(sc, out=o)
        return sc

    def define_function(self, depth):
        name = f"f{len(self.functions)}"
        n_out, n_in = self.ri(0, 2), self.ri(0, 3)
        sc = self.body(n_out, n_in, depth, self.ri(1, 4))
        self.functions.append(ir.Function(name, n_out, n_in, sc.inst, sc.wit, sc.gates))
        self.fn_sig[name] = (n_out, n_in, sc.inst, sc.wit)
        return name

    def inputs_for(self, sc, n):
        self.seed_scope(sc)
        return [self.pick(sc.avail) for _ in range(n)]

    def step(self, sc, depth):
        r = self.rng.random()
        if depth >= self.max_depth or r < 0.55:
            self.simple_gate(sc)
        elif r < 0.63:
            self.assertion(sc)
        elif r < 0.68:
            self.free_some(sc)
        elif r < 0.76 and self.fn_sig:
            name = self.pick(list(self.fn_sig))
            n_out, n_in, fi, fw = self.fn_sig[name]
            ins = self.inputs_for(sc, n_in)
            outs = self.new_wire(sc, n_out) if n_out else []
            sc.gates.append(("Call", name, self.wl(outs), self.wl(ins) if self.chance(0.5) else [ir.Wire(i) for i in ins]))
            sc.inst += fi
            sc.wit += fw
            sc.avail += outs
        elif r < 0.84:
            n_out, n_in = self.ri(0, 2), self.ri(0, 2)
            ins = self.inputs_for(sc, n_in)
            inner = self.body(n_out, n_in, depth + 1, self.ri(1, 3))
            outs = self.new_wire(sc, n_out) if n_out else []
            sc.gates.append(("AnonCall", self.wl(outs), [ir.Wire(i) for i in ins], inner.inst, inner.wit, inner.gates))
            sc.inst += inner.inst
            sc.wit += inner.wit
            sc.avail += outs
        elif r < 0.92:
            self.for_loop(sc, depth)
        else:
            self.switch(sc, depth)

    def for_loop(self, sc, depth):
        n_iter = self.ri(1, 3)
        it = self.pick(["i", "j", "k"])
        per = self.ri(1, 2)                     # outputs per iteration
        outs = self.new_wire(sc, per * n_iter)
        first = self.ri(0, 2)
        base = outs[0] - first * per            # output k of iteration i (first..): base + i*per + k
        I, C = ("Name", it), (lambda v: ("Const", v))
        if base >= 0:
            o_expr = lambda k: ("Add", ("Mul", I, C(per)), C(base + k))
        else:                                   # avoid u64 underflow: (i - first)*per + outs[0] + k
            o_expr = lambda k: ("Add", ("Mul", ("Sub", I, C(first)), C(per)), C(outs[0] + k))
        if per == 2 and self.chance(0.5):
            out_list = [("Range", o_expr(0), o_expr(1))]
        else:
            out_list = [("Single", o_expr(k)) for k in range(per)]
        named = [n for n, s in self.fn_sig.items() if s[0] == per]
        if named and self.chance(0.5):
            name = self.pick(named)
            _, n_in, fi, fw = self.fn_sig[name]
            ins = self.inputs_for(sc, n_in)
            in_list = [("Single", ("DivConst", C(w * 2), 2)) if self.chance(0.3) else ("Single", C(w)) for w in ins]
            body = ("IterExprCall", name, out_list, in_list)
            sc.inst += fi * n_iter
            sc.wit += fw * n_iter
        else:
            n_in = self.ri(0, 2)
            ins = self.inputs_for(sc, n_in)
            in_list = [("Single", C(w)) for w in ins]
            inner = self.body(per, n_in, depth + 1, self.ri(0, 2))
            body = ("IterExprAnonCall", out_list, in_list, inner.inst, inner.wit, inner.gates)
            sc.inst += inner.inst * n_iter
            sc.wit += inner.wit * n_iter
        sc.gates.append(("For", it, first, first + n_iter - 1, self.wl(outs), body))
        sc.avail += outs

    def switch(self, sc, depth):
        self.seed_scope(sc)
        cond = self.pick(sc.avail)
        n_out = self.ri(0, 2)
        n_br = self.ri(1, 3)
        cases, branches = [], []
        mi = mw = 0
        used = set()
        for _ in range(n_br):
            while True:
                cv = self.ri(0, 1) if self.boolean else self.ri(0, 4)
                if cv not in used or len(used) >= (2 if self.boolean else 5):
                    break
            used.add(cv)
            cases.append(bytes([cv]))
            named = [n for n, s in self.fn_sig.items() if s[0] == n_out]
            if named and self.chance(0.4):
                name = self.pick(named)
                _, n_in, fi, fw = self.fn_sig[name]
                ins = self.inputs_for(sc, n_in)
                branches.append(("AbstractGateCall", name, [ir.Wire(i) for i in ins]))
                mi, mw = max(mi, fi), max(mw, fw)
            else:
                n_in = self.ri(0, 2)
                ins = self.inputs_for(sc, n_in)
                inner = self.body(n_out, n_in, depth + 1, self.ri(0, 3))
                branches.append(("AbstractAnonCall", [ir.Wire(i) for i in ins], inner.inst, inner.wit, inner.gates))
                mi, mw = max(mi, inner.inst), max(mw, inner.wit)
        outs = self.new_wire(sc, n_out) if n_out else []
        sc.gates.append(("Switch", cond, self.wl(outs), cases, branches))
        sc.inst += mi
        sc.wit += mw
        sc.avail += outs

    # ---- whole statement --------------------------------------------------------------------------
    def statement(self, n_top=12, n_functions=3):
        """returns [Instance, Witness, Relation] with exactly as many values as the relation consumes"""
        for _ in range(n_functions):
            self.define_function(1)
        top = self.Scope(0, 0)
        for _ in range(n_top):
            self.step(top, 0)
        fc = ir.le_bytes(self.p)
        h = ir.Header(fc)
        gate_mask = ir.BOOL if self.boolean else ir.ARITH
        rel = ir.Relation(h, gate_mask, ir.FOR_FUNCTION_SWITCH, list(self.functions), top.gates)

        def val():
            if self.boolean:
                return bytes([self.ri(0, 1)])
            return ir.le_bytes(self.pick([0, 1, 2, 3, 4, self.ri(0, min(self.p - 1, 1 << 40))]) % self.p)

        inst = ir.Instance(h, [val() for _ in range(top.inst)])
        wit = ir.Witness(h, [val() for _ in range(top.wit)])
        return [inst, wit, rel]
