"""Input-deck writer: produces input.txt / <name>.dat / <name>.pts triples in the reference's positional
free-format layout (documentation/input_files.docx; readers at 3_SPH_material_2018.f90:62-336, 383-467,
787-802 and 1_SPH_2018.f90:143-152).

Used for (a) the three shipped example problems, regenerated from their parameter values so that the GPU box
(which has no /root/reference) can run them, and (b) the synthetic refined problems of BASELINE.json
(configs 4 and 5). tests/test_decks.py checks, where the reference checkout is present, that the regenerated
decks parse to exactly the same particles and parameters as the shipped files.
"""
import os


def _fmt(v):
    if isinstance(v, bool):
        return "T" if v else "F"
    if isinstance(v, float):
        return repr(v)
    return str(v)


def _row(*vals):
    return " ".join(_fmt(v) for v in vals)


def write_deck(directory, spec):
    """spec: dict, see bui_spec() for the keys."""
    os.makedirs(directory, exist_ok=True)
    name = spec["name"]
    inp = ["problem_name", name, "SP_SPH art_stress particle_shift",
           _row(spec.get("sp_sph", True), spec.get("art_stress", False), False)]
    if spec.get("sp_sph", True):
        inp += ["inside_approach", _row(spec["inside_approach"]), "number_of_stress_points", _row(spec["npoints"])]
        if not spec["inside_approach"]:
            inp += ["SPH_shifting vel_vector shift_update rx_factor ry_factor disp_tol",
                    _row(spec["sph_shift"], spec["vel_vector"], spec["shift_update"], spec["rx_factor"],
                         spec["ry_factor"], spec["disp_tol"])]
    inp += ["smoothing_length_factor", _row(spec["sml"])]
    for blk in spec["blocks"]:
        inp += ["dt time_end maxtimestep", _row(blk["dt"], blk["time_end"], blk["maxtimestep"]),
                "print save plot", _row(blk.get("print_step", 100), blk.get("save_step", 100), blk.get("plot_step", 100))]
    inp += ["end_of_simulation", "-1,1,1", ""]
    with open(os.path.join(directory, "input.txt"), "w") as f:
        f.write("\n".join(inp))

    dat = ["1", spec.get("title", "generated_deck"), "ndimn", "2"]
    if spec["variant"] == "bui":
        dat += ["ntype_solid nstre nmats", _row(spec.get("ntype_solid", 2), 4, 1)]
    else:
        dat += ["nstre nmats", _row(4, 1)]
    props = spec["props"]  # 12 values: eco law E nu b rho yield H fi gamma delta n
    dat += ["material_row", _row(1, *props)]
    if int(props[1]) in (5, 12):
        dat += ["extra_model_parameters", _row(*spec["props_extra"])]  # 8 values (props 13..20)
    dat += ["ic_unks", "0"]
    bcs = spec.get("bcs", [])
    dat += ["BCs_nprer_sigman", _row(len(bcs), spec.get("ifsigman", 0))]
    if bcs:
        dat += ["bc_table"] + [_row(*b) for b in bcs]  # 8 values each: id var tvar a1 a0 w fi Tf
    segs = spec.get("segments", [])
    dat += ["number_of_segments_with_BCs", _row(len(segs))]
    if segs:
        dat += ["segment_table"] + [_row(*s) for s in segs]  # x1 y1 x2 y2 bc
    dat += ["number_of_nodes_with_BCs", "0"]
    curves = spec.get("curves", [])
    dat += ["ic_tcurve", "1", "ntcurves maxpts", _row(len(curves), 100)]
    for k, (tt, ff) in enumerate(curves):
        dat += [f"pts_in_curve_{k + 1}", _row(len(tt)), "times_and_factors", _row(*tt), _row(*ff)]
    dat += ["xmin_domain xmax_domain", _row(*spec["domain"]), "initial_conditions_index", "2"]
    for lab in ("s11", "s22", "s12", "s33", "v1", "v2"):
        dat += [f"ICtype_{lab}", "0"]
    dat += ["pa_sph nnps sle skf cspm update_x XSPH",
            _row(2, 2, spec.get("sle", 1), spec.get("skf", 1), spec["cspm"], spec["update_x"], spec["xsph"]),
            "summ_dens cont_dens", _row(False, spec.get("cont_density", False)), "damping", _row(spec["damping"]),
            "alpha beta", _row(spec["alpha"], spec["beta"])]
    if spec.get("gravity") is not None:
        gx, gy, curve, fac = spec["gravity"]
        dat += ["ic_grav", "1", "gx gy timecurve factor", _row(gx, gy, curve, fac)]
    else:
        dat += ["ic_grav", "0"]
    dat += ["output_values", "sxx syy sxy szz ux uy strain rho sml disp_10", _row(*spec.get("out", [1] * 7 + [0] * 3)), ""]
    with open(os.path.join(directory, name + ".dat"), "w") as f:
        f.write("\n".join(dat))

    g = spec["geom"]  # x1 x4 y1 y4 dx dy
    pts = ["geom_type", "1", "x1 x2 x3 x4 dx", _row(g["x1"], g["x4"], g["x1"], g["x4"], g["dx"]),
           "y1 y2 y3 y4 dy", _row(g["y1"], g["y1"], g["y4"], g["y4"], g["dy"]), "dummy_nodes",
           _row(bool(spec.get("walls")))]
    if spec.get("walls"):
        pts += ["number_of_walls", _row(len(spec["walls"]))]
        for k, wl in enumerate(spec["walls"]):  # (id, position, start, end)
            pts += [f"wall_{k + 1}", "id position x1 x2", _row(*wl)]
    pts += [""]
    with open(os.path.join(directory, name + ".pts"), "w") as f:
        f.write("\n".join(pts))
    return directory


# ---- the three shipped example problems (values from example_problems/*/{input.txt,*.dat,*.pts}) ----------

def bui_spec(dx=0.1, dt=1.5e-4, maxtimestep=10, walls=None, mode="vel_vector", npoints=2):
    """soil_failure_bui_et_al_2008. mode: "vel_vector" = outside_approach/velocity_vector_update (== the top-level
    input.txt), "outside" = outside_approach/, "inside" = inside_approach/SP<npoints>, "standard" = standard_sph/."""
    if walls is None:
        walls = [(1, -0.1, -0.1, 2.1), (2, -0.1, -0.3, 9)]
    return dict(
        name="co_soil", variant="bui", title="cohesive_soil_failure_Bui_2008", ntype_solid=2,
        sp_sph=(mode != "standard"), inside_approach=(mode == "inside"), npoints=npoints, sph_shift=True,
        vel_vector=(mode == "vel_vector"), shift_update=1,
        rx_factor=0.3333333333333, ry_factor=0.3333333333333, disp_tol=0.125, sml=1.2,
        blocks=[dict(dt=dt, time_end=2.5, maxtimestep=maxtimestep, print_step=1, save_step=1, plot_step=1)],
        props=[2, 12, 1.8e06, 0.3, 1., 1850, 2000., 0, 0., 5., 1., 1],
        props_extra=[0.466308, 5000, 0, 0, 0, 0, 0, 0],
        curves=[([0, 1, 500], [1, 1, 1])], domain=[-10, -10, 41, 41],
        cspm=False, update_x=True, xsph=True, sle=1, damping=0, alpha=0.1, beta=0.1,
        gravity=(0, -9.81, 1, 1.), geom=dict(x1=0.0, x4=4, y1=0, y4=2, dx=dx, dy=dx), walls=walls)


_VS_BCS = [(1, 5, 1, 0, 0, 0, 0, 0), (2, 6, 1, 0, 0, 0, 0, 0), (3, 1, 1, 0, 0, 0, 0, 0), (4, 3, 1, 0, 0, 0, 0, 0),
           (5, 2, 1, 0, 0, 0, 0, 0)]


def vertical_slope_spec(dx=0.5, dt=0.001, width=10., maxtimestep=10000000, npoints=1, standard=False):
    """vertical_slope (elastic_cut): elastic block, CSPM, damping, segment BCs; `width` stretches it in x.
    npoints = 1,2,3 -> SP1/SP2/SP3 (the top-level deck is SP1); standard -> standard/ (SP_SPH = F)."""
    w = width
    segs = [(0., 0., w, 0., 1), (0., 0., w, 0., 2), (0, 0.5, 0., 10., 1), (0, 0.5, 0., 10., 4),
            (0., 10., w, 10., 5), (0., 10., w, 10., 4), (w, 0.5, w, 10., 3), (w, 0.5, w, 10., 4)]
    return dict(
        name="elastic_cut", variant="vs", title="square_elastic_vertical_cut",
        sp_sph=not standard, inside_approach=True, npoints=npoints, sml=0.8,
        blocks=[dict(dt=dt, time_end=2, maxtimestep=maxtimestep)],
        props=[1, 2, 8.e07, 0.3, 1., 2.e3, 200000., 0, 0., 2, 1., 1],
        bcs=_VS_BCS, segments=segs, curves=[([0, 1, 100], [0, 1, 1])], domain=[-1, -1, max(41, w + 31), 41],
        cspm=True, update_x=False, xsph=False, sle=2, damping=50, alpha=0, beta=0,
        gravity=(0., -9.81, 1, 1.), geom=dict(x1=0, x4=w, y1=0, y4=10, dx=dx, dy=dx),
        out=[1] * 7 + [0, 0, 0])


def strain_localisation_spec(dx=0.0125, dt=0.00001, maxtimestep=10000000, npoints=1, standard=False):
    """strain_localisation_in_soil_sample (localisation): von-Mises Perzyna with softening, CSPM, BC curves.
    npoints = 1,2,3 -> SP1/SP2/SP3; standard -> standard/ (SP_SPH = F, alpha = beta = 0.5)."""
    bcs = [(1, 5, 1, 0, 0, 0, 0, 0), (2, 6, 1, 0, 0, 0, 0, 0), (3, 1, 2, 0, 0, 0, 0, 0), (4, 3, 1, 0, 0, 0, 0, 0),
           (5, 6, 1, 1, 0, 0, 0, 0)]
    segs = [(0., 0., 0., 1., 1), (0., 0., 0., 1., 4), (0., 1., 0.5, 1., 1), (0., 1., 0.5, 1., 5),
            (0.5, 0., 0.5, 1., 3), (0.5, 0., 0.5, 1., 4), (0., 0., 0.5, 0., 1), (0., 0., 0.5, 0., 2)]
    return dict(
        name="localisation", variant="sl", title="strain_localisation_test",
        sp_sph=not standard, inside_approach=True, npoints=npoints, sml=1.2,
        blocks=[dict(dt=dt, time_end=0.021, maxtimestep=maxtimestep)],
        props=[2, 2, 8.e07, 0.25, 1., 2.e3, 5.e5, -8.e06, 0., 50., 1., 1],
        bcs=bcs, segments=segs,
        curves=[([0., 0.0005, 0.0005, 0.2, 1], [0, 1, 1, 1, 1]), ([0., 0.005, 0.0051, 1], [1.0, 1.0, 1.0, 1.0])],
        domain=[-1, -1, 41, 41], cspm=True, update_x=True, xsph=False, sle=2, damping=0,
        alpha=0.5 if standard else 0.0, beta=0.5 if standard else 0.0,
        gravity=None, out=[0, 0, 1, 0, 0, 0, 1, 0, 0, 0] if standard else [0, 0, 1, 0, 1, 1, 1, 0, 0, 0],
        geom=dict(x1=0, x4=0.5, y1=0, y4=1, dx=dx, dy=dx))


# ---- synthetic refined problems (BASELINE.json configs 4 and 5) ---------------------

def refined_bui_spec(ncol=1632, maxtimestep=100):
    """Bui column refined so that the 4 m x 2 m block holds (ncol+1) x (ncol/2+1) velocity particles;
    ncol = 1632 gives 1633 x 817 = 1 334 161 nodes + 2 668 322 stress particles = 4 002 483 particles.
    dt is scaled with dx (constant Courant number); walls follow the shipped layout shifted to -dx."""
    dx = 4.0 / ncol
    s = bui_spec(dx=dx, dt=1.5e-4 * dx / 0.1, maxtimestep=maxtimestep,
                 walls=[(1, -dx, -dx, 2 + dx), (2, -dx, -3 * dx, 9)])
    s["title"] = f"refined_Bui_column_ncol_{ncol}"
    return s


def wide_slope_spec(ncol=1414, nslab=1, maxtimestep=100):
    """Vertical slope refined to dx = 10/ncol and stretched to nslab x 10 m in x (one 10 m slab per GPU for the
    weak-scaling run); ncol = 1414 gives ~2.0 M + 2.0 M particles per slab."""
    dx = 10.0 / ncol
    s = vertical_slope_spec(dx=dx, dt=0.001 * dx / 0.5, width=10.0 * nslab, maxtimestep=maxtimestep)
    s["title"] = f"wide_vertical_slope_ncol_{ncol}_nslab_{nslab}"
    return s


SHIPPED = {"bui": bui_spec, "vs": vertical_slope_spec, "sl": strain_localisation_spec}
