"""Host logic: the .in / _dist.in reader against an input written like the reference's test inputs."""
import numpy as np

from alps_b200.namelist import read_namelists

SAMPLE = """
&system
kperp=1.d-2        ! wave vector
kpar=1.d-2
nspec=2
use_map=.false.
arrayName='test_kpar_fast'
Bessel_zero = 1.d-50
D_threshold = 1.d-15
fit_check=T
/
&guess_1
g_om=9.9d-3    !real frequency
g_gam=-5.5d-6 !imaginary frequency
/
&spec_2
nn=1.d0
qq=-1.d0
mm=5.44662d-4
relat=F
use_bM=F
/
&ffit_2_1
fit_type_in=1
fit_1=1.4128D+4
perpcorr=1.836D+3 !renormalization factor
/
&scan_input_1
scan_type=4
swlog=.true.    !log (T) or linear (F) steps
ns=32
/
"""


def test_read_namelists(tmp_path):
    p = tmp_path / "t.in"
    p.write_text(SAMPLE)
    nl = read_namelists(str(p))
    assert nl["system"]["kperp"] == 1.0e-2 and nl["system"]["nspec"] == 2
    assert nl["system"]["use_map"] is False and nl["system"]["fit_check"] is True
    assert nl["system"]["arrayname"] == "test_kpar_fast"
    assert nl["system"]["bessel_zero"] == 1.0e-50
    assert nl["guess_1"] == {"g_om": 9.9e-3, "g_gam": -5.5e-6}
    assert nl["spec_2"]["mm"] == 5.44662e-4 and nl["spec_2"]["relat"] is False
    assert nl["ffit_2_1"]["fit_1"] == 1.4128e4 and nl["ffit_2_1"]["perpcorr"] == 1836.0
    assert nl["scan_input_1"] == {"scan_type": 4, "swlog": True, "ns": 32}


def test_plasma_from_inputs_matches_named_config(tmp_path):
    """the twin of the reference's start-up sequence builds the same tables as tables.config_kpar_fast"""
    from alps_b200 import tables
    from alps_b200.run import plasma_from_inputs
    inp = tmp_path / "t.in"
    inp.write_text("""&system
kperp=1.d-2
kpar=1.d-2
nspec=2
nperp=120
npar=240
vA=1.d-4
arrayName='x'
Bessel_zero=1.d-50
/
&spec_1
nn=1.d0
qq=1.d0
mm=1.d0
ff=1
/
&ffit_1_1
fit_type_in=1
/
&spec_2
nn=1.d0
qq=-1.d0
mm=5.44662d-4
ff=1
/
&ffit_2_1
fit_type_in=1
/
""")
    dist = tmp_path / "t_dist.in"
    dist.write_text("""&system
nspec=2
beta=1.d0
vA=1.d-4
nperp=120
npar=240
maxP=6.d0
/
&spec_1
ms_read=1.d0
taus=1.d0
alphs=1.d0
ps=0.d0
kappas=8.d0
distributions=1
autoscaleS=T
maxPperpS=1.d0
maxPparS=1.d0
/
&spec_2
ms_read=5.44662d-4
taus=1.d0
alphs=1.d0
ps=0.d0
kappas=8.d0
distributions=1
autoscaleS=T
maxPperpS=1.d0
maxPparS=1.d0
/
""")
    pl = plasma_from_inputs(read_namelists(str(inp)), read_namelists(str(dist)), base_dir=str(tmp_path))
    ref = tables.config_kpar_fast()
    assert np.array_equal(pl.pp, ref.pp) and np.array_equal(pl.f0, ref.f0)
    assert np.array_equal(pl.param_fit, ref.param_fit)
    assert [s.perp_correction for s in pl.species] == [s.perp_correction for s in ref.species]
