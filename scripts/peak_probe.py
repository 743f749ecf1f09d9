import sys; sys.path.insert(0,'/root/repo')
from alps_b200 import tables, _lib
from alps_b200.solver import Solver
sol = Solver(tables.config_small(16,32))
print("dfma peak (reuse)", sol.dfma_peak(), "no-reuse", sol.info(_lib.INFO_DFMA_NOREUSE))
