"""Writes compact copies of the reference's remaining test configurations into tests/inputs/suite/
(this container only: reads /root/reference/tests/*.in and distribution/*_dist.in).  Comments are dropped,
the scans are shortened (ns <= 4 with the reference's step size, numiter <= 40) so that the GPU suite runs each in about a second, and the
namelists are re-emitted in a canonical form.  The physics parameters are untouched.
With --full the scans, maps and iteration limits are left at the reference's values and the files go to
tests/inputs/full/ (timed end to end by scripts/full_configs.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alps_b200.namelist import read_namelists

REF = "/root/reference"
FULL = "--full" in sys.argv
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "inputs",
                   "full" if FULL else "suite")
NAMES = ["test_ICW", "test_electron_mode", "test_analytical", "test_chebyshev", "test_cold_plasma", "test_bimax",
         "test_kperp", "test_kperp_alpha", "test_double_scan", "test_map"] + (["test_kpar_fast", "test_relativistic"] if FULL else [])
# tests/test_kperp_alpha.in names arrayName='test_kperp_alpha', for which the reference ships no _dist.in: the three
# Maxwellians its &ffit blocks describe (p, e, alphas with T_j = T_ref: fit_2 = 1/(m_j tau_j) = 1, 1836, 0.25) as a
# generate_distribution input written here
EXTRA_DIST = {"test_kperp_alpha": {
    "system": dict(nspec=3, beta=1.0, va=1.0e-4, nperp=120, npar=240, maxp=6.0, writename="test_kperp_alpha"),
    **{"spec_%d" % (i + 1): dict(ms_read=m, taus=1.0, alphs=1.0, ps=0.0, kappas=8.0, distributions=1, autoscales=True,
                                 maxpperps=1.0, maxppars=1.0) for i, m in enumerate((1.0, 5.44662e-4, 4.0))}}}
os.makedirs(OUT, exist_ok=True)


def fmt(v):
    if isinstance(v, bool):
        return "T" if v else "F"
    if isinstance(v, float):
        return repr(v).replace("e", "d") if "e" in repr(v) else repr(v) + "d0"
    if isinstance(v, str):
        return "'%s'" % v
    return str(v)


def emit(nl, path):
    with open(path, "w") as fh:
        for group, d in nl.items():
            fh.write("&%s\n" % group)
            for k, v in d.items():
                fh.write("%s=%s\n" % (k, fmt(v)))
            fh.write("/\n")


for name in NAMES:
    nl = read_namelists(os.path.join(REF, "tests", name + ".in"))
    s = nl["system"]
    if not FULL:
        s["numiter"] = min(int(s.get("numiter", 50)), 40)
    for k, d in (nl.items() if not FULL else []):
        if k.startswith("scan_input_"):
            ns_old = int(d["ns"]) * int(d.get("nres", 1))
            ns_new = min(ns_old, 4 if int(s.get("scan_option", 1)) == 1 else 2)
            # same step size as the reference run (ns * nres steps from the current k to swf): move the end
            # point, not the step
            st = int(d["scan_type"])
            if ns_new < ns_old and st in (3, 4):
                k0, swf = float(s["kperp"] if st == 3 else s["kpar"]), float(d["swf"])
                if bool(d.get("swlog", False)):
                    d["swf"] = float(k0 * (swf / k0) ** (ns_new / ns_old))
                else:
                    d["swf"] = float(k0 + (swf - k0) * ns_new / ns_old)
            d["ns"] = ns_new
            d["nres"] = 1
        if k.startswith("maps_"):
            d["nr"], d["ni"] = min(int(d["nr"]), 24), min(int(d["ni"]), 24)
    emit(nl, os.path.join(OUT, name.replace("test_", "cfg_") + ".in"))
    arr = s.get("arrayname", "")
    dist = os.path.join(REF, "distribution", arr + "_dist.in")
    if os.path.exists(dist):
        emit(read_namelists(dist), os.path.join(OUT, name.replace("test_", "cfg_") + "_dist.in"))
    elif name in EXTRA_DIST:
        emit(EXTRA_DIST[name], os.path.join(OUT, name.replace("test_", "cfg_") + "_dist.in"))
    print(name, "->", arr, os.path.exists(dist))
