"""TEST INFRASTRUCTURE. The runs of the reference's own executables that pin the oracle: shared by
oracle/make_reference_goldens.py (which produces tests/golden/ref_<case>.npz) and the tests that read them."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stress-particle-sph_b200"))
from spsph import decks  # noqa: E402

# case -> (binary, deck spec, steps to run, steps whose frames are kept | plot cadence).
# A tuple of steps: the binary plots every step (plot_step = 1: its fp32 output clock then fires exactly once per
# step, 1_SPH_2018.f90:86-89) and the listed frames are kept. That only works for short runs: the reference opens
# its frame files on unit numbers nnode + step etc. (mat:2937-2939), which collide with its own input unit 997 at
# step 136 of the Bui problem. Long runs therefore plot every `cadence` steps and keep whatever frames appear
# (the output clock is fp32, so a frame can land one step late).
CASES = {
    # the three top-level example inputs (BASELINE configs 0-2), 100 steps as north_star asks
    "bui": ("bui", lambda n: decks.bui_spec(maxtimestep=n), 100, (1, 50, 100)),
    "vs": ("vs", lambda n: decks.vertical_slope_spec(maxtimestep=n), 100, (1, 50, 100)),
    "sl": ("sl", lambda n: decks.strain_localisation_spec(maxtimestep=n), 100, (100,)),
    # a long Bui run: the pair list grows and the split traversal order of SURVEY App. B is exercised
    "bui_long": ("bui", lambda n: decks.bui_spec(maxtimestep=n), 610, 300),
    # the other input sets the reference ships
    "bui_outside": ("bui", lambda n: decks.bui_spec(maxtimestep=n, mode="outside"), 60, (60,)),
    "bui_inside_sp1": ("bui", lambda n: decks.bui_spec(maxtimestep=n, mode="inside", npoints=1), 60, (60,)),
    "bui_inside_sp2": ("bui", lambda n: decks.bui_spec(maxtimestep=n, mode="inside", npoints=2), 60, (60,)),
    "bui_inside_sp3": ("bui", lambda n: decks.bui_spec(maxtimestep=n, mode="inside", npoints=3), 60, (60,)),
    "bui_standard": ("bui", lambda n: decks.bui_spec(maxtimestep=n, mode="standard"), 60, (60,)),
    "vs_sp2": ("vs", lambda n: decks.vertical_slope_spec(maxtimestep=n, npoints=2), 60, (60,)),
    "vs_sp3": ("vs", lambda n: decks.vertical_slope_spec(maxtimestep=n, npoints=3), 60, (60,)),
    "vs_standard": ("vs", lambda n: decks.vertical_slope_spec(maxtimestep=n, standard=True), 60, (60,)),
    "sl_sp2": ("sl", lambda n: decks.strain_localisation_spec(maxtimestep=n, npoints=2), 40, (40,)),
    "sl_sp3": ("sl", lambda n: decks.strain_localisation_spec(maxtimestep=n, npoints=3), 40, (40,)),
    "sl_standard": ("sl", lambda n: decks.strain_localisation_spec(maxtimestep=n, standard=True), 40, (40,)),
    # the synthetic refinements bench.py runs (BASELINE configs 3 and 4) at a size the reference can do
    "bui_refined": ("bui", lambda n: decks.refined_bui_spec(ncol=68, maxtimestep=n), 30, (30,)),
    "vs_wide": ("vs", lambda n: decks.wide_slope_spec(ncol=60, nslab=2, maxtimestep=n), 30, (30,)),
    # smoothing kernels other than the cubic spline (main:1494-1536)
    "bui_gauss": ("bui", lambda n: dict(decks.bui_spec(maxtimestep=n), skf=2), 30, (30,)),
    "bui_quintic": ("bui", lambda n: dict(decks.bui_spec(maxtimestep=n), skf=3), 30, (30,)),
    "vs_gauss": ("vs", lambda n: dict(decks.vertical_slope_spec(maxtimestep=n), skf=2), 30, (30,)),
    "sl_quintic": ("sl", lambda n: dict(decks.strain_localisation_spec(maxtimestep=n), skf=3), 20, (20,)),
    # options no shipped input switches on: artificial stress (main:908-1016), continuity density with and without
    # the smoothing-length update (main:706-713, 772-797, 807-821)
    "bui_art_stress": ("bui", lambda n: dict(decks.bui_spec(maxtimestep=n), art_stress=True), 40, (40,)),
    "sl_art_stress": ("sl", lambda n: dict(decks.strain_localisation_spec(maxtimestep=n), art_stress=True), 30, (30,)),
    "bui_cont_density": ("bui", lambda n: dict(decks.bui_spec(maxtimestep=n), cont_density=True), 40, (40,)),
    "vs_cont_density_sle2": ("vs", lambda n: dict(decks.vertical_slope_spec(maxtimestep=n), cont_density=True), 40, (40,)),
    "sl_cont_density_sle2": ("sl", lambda n: dict(decks.strain_localisation_spec(maxtimestep=n), cont_density=True), 30, (30,)),
    # Perzyna viscoplasticity with the other yield criteria of invar09 / yieldf09 (strain_localisation copy
    # mat:2277-2296, 2423-2461): Tresca, Mohr-Coulomb, Drucker-Prager; yield stress lowered so that the sample
    # yields over a wide zone within the run; exponential flow rule (nflow /= 1) and fnorm**delta with delta /= 1
    "sl_tresca": ("sl", lambda n: _sl_perzyna(n, ncrit=1), 60, (60,)),
    "sl_mohr_coulomb": ("sl", lambda n: _sl_perzyna(n, ncrit=3, frict=20.), 60, (60,)),
    "sl_dp_perzyna": ("sl", lambda n: _sl_perzyna(n, ncrit=4, frict=20.), 60, (60,)),
    "sl_vm_expflow": ("sl", lambda n: _sl_perzyna(n, ncrit=2, nflow=2, delta=1.5), 60, (60,)),
    "sl_vm_powflow": ("sl", lambda n: _sl_perzyna(n, ncrit=2, delta=1.5), 60, (60,)),
    # prescribed-traction free surface (ifsigman = 1: apply_stress_free, mat:1756-1839, on the nodes that
    # get_nodes_on_free_surface marked at the end of the previous step) and XSPH together with boundary conditions
    # (main:224-230 strips the BCs of the neighbours of free-surface nodes)
    "sl_sigman": ("sl", lambda n: dict(_sl_perzyna(n, ncrit=2, free_right=True), ifsigman=1), 60, (60,)),
    "vs_sigman": ("vs", lambda n: dict(_vs_free_right(n), ifsigman=1), 60, (60,)),
    "sl_xsph": ("sl", lambda n: dict(_sl_perzyna(n, ncrit=2, free_right=True), xsph=True), 60, (60,)),
    "sl_sigman_xsph": ("sl", lambda n: dict(_sl_perzyna(n, ncrit=2, free_right=True), ifsigman=1, xsph=True), 60, (60,)),
    # Check_Out_Domain (main:1170-1194): the domain box ends 1e-5 beyond the free face of the column,
    # so the wall particles right of it are out from the start (171) and velocity / stress particles leave one by one
    # as the face moves (195 flagged after 200 steps); a flagged particle drops out of the grid for good
    "bui_out_domain": ("bui", lambda n: dict(decks.bui_spec(maxtimestep=n), domain=[-10, -10, 4.00001, 41]), 210, 100),
    # corners of the input space no shipped deck visits: a sinusoidal boundary value with start-up factor (Normal_BCs
    # mat:1696-1705: tvar = 0), plane stress (ntype_solid = 1, Bui copy), the outside approach with 1 and 3 stress
    # particles per velocity particle (shift_stress_points main:296-316), re-seating every 5th step
    "sl_sine_bc": ("sl", lambda n: _sl_sine_bc(n), 60, (60,)),
    "bui_plane_stress": ("bui", lambda n: dict(decks.bui_spec(maxtimestep=n), ntype_solid=1), 60, (60,)),
    "bui_outside_sp1": ("bui", lambda n: decks.bui_spec(maxtimestep=n, mode="outside", npoints=1), 60, (60,)),
    "bui_outside_sp3": ("bui", lambda n: decks.bui_spec(maxtimestep=n, mode="outside", npoints=3), 60, (60,)),
    # a wider kernel support (sml = 1.5 instead of 1.2: up to 88 partners per particle instead of 56)
    "bui_sml15": ("bui", lambda n: dict(decks.bui_spec(maxtimestep=n), sml=1.5), 40, (40,)),
    "bui_shift5": ("bui", lambda n: dict(decks.bui_spec(maxtimestep=n), shift_update=5), 60, (60,)),
    # the three example problems run to the END of their shipped time span (north_star: "matching failure-surface
    # geometry at the end of the run"): Bui t_end = 2.5 s (16 667 steps; the shipped maxtimestep = 10 lifted), vertical
    # slope t_end = 2 s (2000 steps), strain localisation t_end = 0.021 s (2100 steps); the last frame the plot
    # cadence allows before the run ends is the final state
    "bui_full": ("bui", lambda n: decks.bui_spec(maxtimestep=n), 17000, 5550),
    "vs_full": ("vs", lambda n: decks.vertical_slope_spec(maxtimestep=n), 2100, 995),
    "sl_full": ("sl", lambda n: decks.strain_localisation_spec(maxtimestep=n), 2200, 1045),
    # the inside approach pressed against its walls long enough for boundary_forces to act
    "bui_inside_sp1_long": ("bui", lambda n: decks.bui_spec(maxtimestep=n, mode="inside", npoints=1), 1510, 1500),
}


def _sl_perzyna(n, ncrit, frict=0., nflow=1, delta=1., yield0=1.5e5, free_right=False):
    s = decks.strain_localisation_spec(maxtimestep=n)
    s["props"] = [2, ncrit, 8.e07, 0.25, 1., 2.e3, yield0, -8.e06, frict, 50., delta, nflow]
    if free_right:  # no boundary conditions on x = 0.5: get_nodes_on_free_surface marks that side (bc_or_not = 2)
        s["segments"] = [g for g in s["segments"] if not (g[0] == 0.5 and g[2] == 0.5)]
    return s


def _sl_sine_bc(n):
    s = _sl_perzyna(n, ncrit=2)
    s["bcs"] = list(s["bcs"])
    s["bcs"][4] = (5, 6, 0, 1.0, 0.5, 3000., 0.3, 2.e-4)  # top v_y = (a0 + a1 sin(w t - fi)) (1 - exp(-t/Tf))
    return s


def _vs_free_right(n):
    """vertical slope with moving particles and without the zero-traction BCs of its cut face x = 10"""
    s = decks.vertical_slope_spec(maxtimestep=n)
    s["update_x"] = True
    s["segments"] = [g for g in s["segments"] if not (g[0] == 10. and g[2] == 10.)]
    return s


def golden_path(case):
    return os.path.join(ROOT, "tests", "golden", f"ref_{case}.npz")


def spec_of(case):
    """deck specification of a case, with the step count the golden run used"""
    which, spec_fn, nsteps, _ = CASES[case]
    return which, spec_fn(nsteps)


# cases whose options the CUDA engine does not implement: spsph_create must refuse them
DEVICE_UNSUPPORTED = set()
# cases where the engine evaluates libm functions (atan, asin, sin, cos, tan, exp, pow) with CUDA's implementations
# instead of glibc's: agreement to the north star's 1e-9 relative L-inf instead of bit for bit
DEVICE_TOLERANCE = {"bui_art_stress": 1e-9, "sl_art_stress": 1e-9, "sl_vm_expflow": 1e-9, "sl_vm_powflow": 1e-9,
                    "sl_tresca": 1e-9, "sl_mohr_coulomb": 1e-9}
# (sl_sigman, vs_sigman, sl_sigman_xsph -- apply_stress_free with k_fs_normals -- carried 1e-9 until their first
# hardware run; round 2 confirmed them bit-exact on the B200, tools/strict_cases.py, and they are now asserted bitwise)
# device paths written after this round's GPU budget was spent: their first run on hardware
# is tests/test_zz_gpu_new_paths.py, the last file of the GPU suite, so that a surprise there cannot mask the
# verified cases of tests/test_gpu_reference.py (the driver runs pytest with -x)
DEVICE_UNVERIFIED = {"bui_sml15", "sl_sine_bc", "bui_plane_stress", "bui_outside_sp1", "bui_outside_sp3", "bui_shift5", "bui_out_domain", "bui_full", "vs_full", "sl_full", "sl_tresca", "sl_mohr_coulomb", "sl_dp_perzyna", "sl_vm_expflow", "sl_vm_powflow", "sl_sigman",
                     "vs_sigman", "sl_xsph", "sl_sigman_xsph"}
