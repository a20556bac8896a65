"""Instrumented FP64 operation count of the reference's own assembly, per element.

SURVEY.md 8(d) fixes the roofline denominator by a HAND count of the reference source (52 kflop per 4-point tet
with lhs=1) "until replaced by an instrumented count".  This script produces that count: it executes the unmodified
reference Fortran (ElmGMRe's block loop: AsIq, AsIGMR -> e3 -> ..., bc3LHS) through the f77np interpreter with every
REAL*8 array replaced by an ndarray subclass that counts the element-wise add / subtract / multiply / divide / sqrt /
power operations it takes part in (one flop each, the SURVEY's convention; comparisons, min/max, abs, sign flips and
integer work are not counted; both sides of a WHERE are evaluated over the whole block, as the interpreter does).
The counts are exclusive per routine and divided by the number of elements.

    python tests/golden/count_flops_f77.py            # prints the table, writes tests/golden/flops_f77.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import f77np  # noqa: E402

COUNT = [0]
_FLOP = {np.add, np.subtract, np.multiply, np.divide, np.true_divide, np.sqrt, np.power, np.square, np.reciprocal}


class CountArr(np.ndarray):
    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        ins = tuple(np.asarray(i) if isinstance(i, CountArr) else i for i in inputs)
        if out is not None:
            kw["out"] = tuple(np.asarray(o) if isinstance(o, CountArr) else o for o in out)
        res = getattr(ufunc, method)(*ins, **kw)
        if ufunc in _FLOP and method == "__call__":
            r = res if isinstance(res, np.ndarray) else np.asarray(res)
            if r.dtype == np.float64:
                COUNT[0] += int(r.size)
        if isinstance(res, np.ndarray) and res.dtype == np.float64:
            return res.view(CountArr)
        return res


class _NP:
    """numpy with float64 allocations returned as CountArr (installed as f77np.np)"""

    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def full(shape, v, dtype=None, order="C"):
        a = np.full(shape, v, dtype=dtype, order=order)
        return a.view(CountArr) if a.dtype == np.float64 else a

    @staticmethod
    def zeros(shape, dtype=float, order="C"):
        a = np.zeros(shape, dtype=dtype, order=order)
        return a.view(CountArr) if a.dtype == np.float64 else a


def instrument():
    f77np.np = _NP()
    per, stack = {}, []
    orig = f77np.Program.call

    def call(self, name, *a):
        c0 = COUNT[0]
        stack.append(0)
        try:
            return orig(self, name, *a)
        finally:
            delta = COUNT[0] - c0
            kids = stack.pop()
            per[name] = per.get(name, 0) + delta - kids
            if stack:
                stack[-1] += delta
    f77np.Program.call = call
    return per


def main():
    per = instrument()
    import make_golden_f77 as mg
    from common import make_case
    out = {}
    for label, kw in (("tet_4pt_lhs1", dict(rule=2)), ("tet_1pt_lhs1", dict(rule=1)),
                      ("tet_4pt_lhs0", dict(rule=2, lhs0=True)), ("hex_8pt_lhs1", dict(rule=2, topo="hex"))):
        kw = dict(kw)
        lhs0 = kw.pop("lhs0", False)
        case = make_case(4, 4, 4, bc="none", periodic_z=False, ibksiz=64, **kw)
        numel = case[2][0].numel
        per.clear()
        COUNT[0] = 0
        prog = mg.make_program()
        wrap = lambda a: a.view(CountArr) if isinstance(a, np.ndarray) and a.dtype == np.float64 else a  # noqa: E731
        F0 = mg.F
        mg.F = lambda a, dtype=np.float64: wrap(F0(a, dtype))
        try:
            mg.run_elmgmre(prog, case, lhs=0 if lhs0 else 1, with_boundary=False)
        finally:
            mg.F = F0
        tot = sum(per.values())
        rows = {k: v / numel for k, v in sorted(per.items(), key=lambda kv: -kv[1]) if v}
        out[label] = {"numel": numel, "flop_per_element": tot / numel, "by_routine": rows}
        print("%s: %.0f flop/element over %d elements" % (label, tot / numel, numel))
        for k, v in rows.items():
            print("   %-12s %9.1f" % (k, v))
    json.dump(out, open(os.path.join(HERE, "flops_f77.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
