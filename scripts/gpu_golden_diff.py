"""Developer aid: field-by-field text comparison of the scan/eigen/heat files with the goldens."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alps_b200 import tables
from alps_b200.solver import Solver
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
pl = tables.config_kpar_fast(); sol = Solver(pl, emulate_nproc=4); sol.set_k(1e-2, 1e-2)
opts = sol.opts(D_threshold=1.0e-15)
d = tempfile.mkdtemp(); prefix = os.path.join(d, "t")
w, D = sol.refine_guess([complex(9.9e-3, -5.5e-6)], opts)
rows, w = sol.om_scan(w, opts, 4, 1e-3, 1e-1, True, 32, 1, True, True, prefix, 1)
for kind in ("scan", "eigen", "heat"):
    a = [l.split() for l in open(prefix + ".%s_kpara_1.root_1" % kind) if l.strip()]
    b = [l.split() for l in open(os.path.join(GOLD, "test_kpar_fast.%s_kpara_1.root_1" % kind)) if l.strip()]
    tot = bad = 0
    for i, (ra, rb) in enumerate(zip(a, b)):
        for j, (x, y) in enumerate(zip(ra, rb)):
            tot += 1
            if x != y:
                bad += 1
                if bad <= 12: print(kind, "row", i, "col", j, x, y)
    print(kind, "fields", tot, "different", bad)
