"""Checkpoint / resume of a run (the reference has none: its .chk file is an input echo and save_step is read and
ignored, 1_SPH_2018.f90:150). A checkpoint is the downloaded state (every spsph_state array) plus
the three scalars of the driver's clock and the length the pair list has grown to, which decides the traversal order
of the following steps; restoring all of it continues the run bit for bit.

    save(path, engine, itimestep, time_sph)              after step `itimestep`, time_sph = the clock after that step
    itimestep, time_sph = resume(path, engine, problem)  on a freshly created engine for the same problem

Only the arrays the time step changes are stored (DYNAMIC); the set-up arrays (mass, rho, hsml, itype, wall geometry,
boundary-condition tables, x00) come from the problem the deck reader built, exactly as in a fresh run.

Not carried (the reference keeps them in module arrays that are not part of the state): the velocity gradient of the
last sweep (read by the first density_update of a step when cont_density = T) and the free-surface normals
(read by apply_stress_free when ifsigman = 1); save() refuses such runs.
"""
import numpy as np

DYNAMIC = ("x", "vel", "stress", "internal_vars", "f_drucker", "displ", "x_10", "disp_10", "n_int", "bc_int",
           "if_out_domain", "bc_or_not")


def save(path, engine, itimestep, time_sph):
    if engine.p.cont_density or (engine.p.ifsigman == 1 and engine.p.no_bcs > 0 and engine.p.update_x):
        raise ValueError("checkpoint: cont_density = T and ifsigman = 1 keep state outside spsph_state")
    if getattr(engine, "in_dist_mode", False):
        # a slab rank only holds its own particles (the rest of its download is stale): merge the ranks' downloads with
        # spsph.dist.merge_owned and checkpoint the merged state from one single-GPU engine instead
        raise ValueError("checkpoint: this engine is one slab of a multi-GPU run; save the merged state (dist.merge_owned)")
    arrays = engine.download()
    np.savez(path, itimestep=np.int64(itimestep), time_sph=np.float64(time_sph),
             list_capacity=np.int64(engine.list_capacity()), ntotal2=np.int64(engine.p.ntotal2),
             **{k: arrays[k] for k in DYNAMIC})


def resume(path, engine, problem):
    """uploads the problem's set-up arrays with the checkpointed state on top into `engine`; returns
    (itimestep, time_sph): continue with engine.run(itimestep + 1, time_sph, dt, n)"""
    with np.load(path) as z:
        if int(z["ntotal2"]) != engine.p.ntotal2 or problem.params.ntotal2 != engine.p.ntotal2:
            raise ValueError("checkpoint: written for a different problem (particle count differs)")
        arrays = dict(problem.arrays)
        for k in DYNAMIC:
            arrays[k] = np.ascontiguousarray(z[k])
        engine.upload(arrays)
        engine.set_list_capacity(int(z["list_capacity"]))
        return int(z["itimestep"]), float(z["time_sph"])
