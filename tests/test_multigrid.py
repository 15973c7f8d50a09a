"""Device residual / divergence norms and the multigrid pressure solve (BASELINE.json config 3).

The reference only sketches these (compute_residual, restrict_to_*, prolongate are declared in src/Cuda/Kernels.cuh:38-49 without a
definition, v_cycle is commented out in src/Cuda/HNanoSolver.cu:399-507), so there is no reference output: PARITY UNPINNED for this
block. The product is checked against the CPU restatement of the same algorithm (oracle.OracleMultigrid), against fp64 host norms, and
against the acceptance gates of SURVEY.md Appendix A-9 (iii): relative Poisson residual <= 1e-4 and a projected velocity whose
divergence is no larger than what the reference's red-black solve reaches at I = 100."""
import numpy as np
import pytest

from hnanosolver_b200 import synth

NORM_TOL = 1e-6      # device fp64 reduction vs host fp64
MG_FIELD_TOL = 1e-5  # product V-cycle vs its CPU restatement, relative to max |p|


# a full box, a scattered leaf set that straddles the origin (negative coordinates, several root tiles of coarse parents), a plume
CASES = {"sphere40": lambda: synth.smoke_sphere(40, 2), "soup": lambda: synth.random_leaves(n_leaves=60, extent=7, seed=11, offset=(-24, -8, -40)),
         "c2": lambda: synth.WORKLOADS["c2"]()}


def _div_of_advected(O, w):
    ix = O.OracleIndex(w.coords)
    adv = ix.advect_vector(w.velocity, w.dt, w.voxel_size)
    return ix, adv, ix.divergence(adv, w.voxel_size)


# ---- CPU: the oracle's own properties ---------------------------------------------------------------------------------------
def test_oracle_multigrid_converges_and_beats_the_fixed_count_solve(oracle_mod):
    O = oracle_mod
    w = synth.smoke_sphere(40, 2)
    ix, _, div = _div_of_advected(O, w)
    mg = O.OracleMultigrid(w.coords, w.voxel_size)
    assert len(mg.levels) >= 3 and np.unique(mg.levels[-1][0].coords >> 3, axis=0).shape[0] == 1
    p, cycles, rel = mg.solve(div, 12, rel_tol=1e-4, omega=1.15)
    assert rel <= 1e-4 and cycles <= 6
    rbgs = ix.rbgs(div, np.zeros_like(div), w.voxel_size, 40, O.omega_compute(w.voxel_size))
    a, b = mg.residual_sums(rbgs, div)
    assert rel < np.sqrt(a / b)                                  # 40 red-black iterations are still far from converged
    # coarse diagonals: 6 inside, larger next to the domain boundary, by the documented amount
    d1 = mg.diag[1]
    assert d1.min() == np.float32(6.0) and np.isclose(d1.max(), 6.0 + 3 * (1 / 0.75 - 1), rtol=1e-6)


def test_oracle_residual_is_zero_for_an_exact_solution(oracle_mod):
    O = oracle_mod
    w = synth.smoke_sphere(32, 5)
    ix = O.OracleIndex(w.coords)
    rng = np.random.default_rng(3)
    p = rng.standard_normal(ix.n).astype(np.float32)
    mg = O.OracleMultigrid(w.coords, w.voxel_size, max_levels=1)
    # rhs := L p  =>  residual 0 up to fp32 rounding of L p itself
    r0 = mg.residual(0, p, np.zeros(ix.n, np.float32))             # = -L p
    a, b = mg.residual_sums(p, -r0)
    assert np.sqrt(a / b) < 1e-6


# ---- GPU ----------------------------------------------------------------------------------------------------------------------
def _gpu_state(w, div=None, p=None):
    import hnanosolver_b200 as H

    grid = H.create_index_grid_from_origins(w.origins, w.voxel_size)
    sim = H.Simulation(grid, 0)
    sim.upload(w.velocity)
    return grid, sim


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere40", "soup", "c2"])
def test_device_norms_match_fp64_host_norms(oracle_mod, name):
    O = oracle_mod
    w = CASES[name]()
    ix, adv, div = _div_of_advected(O, w)
    grid, sim = _gpu_state(w)
    sim.advect_velocity(w.dt)
    sim.divergence(True)
    sim.pressure_solve(7, O.omega_compute(w.voxel_size))
    p = sim.aux(1)
    assert np.array_equal(sim.aux(0), div)
    a, b = sim.residual_sums()
    mg = O.OracleMultigrid(w.coords, w.voxel_size, max_levels=1)
    ha, hb = mg.residual_sums(p, div)
    assert abs(a - ha) <= NORM_TOL * ha and abs(b - hb) <= NORM_TOL * hb
    # ||div(u)||^2 of the advected velocity through the divergence kernel + the reduction
    got = sim.divergence_sum_squares(of_advected=True)
    want = O.sum_squares(div)
    assert abs(got - want) <= NORM_TOL * want


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere40", "soup", "c2"])
def test_v_cycles_match_the_cpu_restatement(oracle_mod, name):
    import hnanosolver_b200 as H

    O = oracle_mod
    w = CASES[name]()
    ix, adv, div = _div_of_advected(O, w)
    grid, sim = _gpu_state(w)
    sim.advect_velocity(w.dt)
    sim.divergence(True)
    mg = H.Multigrid(grid)
    omg = O.OracleMultigrid(w.coords, w.voxel_size)
    assert mg.num_levels == len(omg.levels)
    for k in range(mg.num_levels):
        assert mg.level_cells(k) == omg.levels[k][0].n                 # same hierarchy, cell for cell
    cycles, rel = sim.pressure_solve_mg(mg, 2, 0.0, 2, 2, 1.15)
    assert cycles == 2 and rel == -1.0
    want, _, _ = omg.solve(div, 2, 2, 2, 1.15)
    got = sim.aux(1)
    err = np.abs(got.astype(np.float64) - want).max() / np.abs(want).max()
    assert err <= MG_FIELD_TOL, err
    assert abs(sim.relative_residual() - np.sqrt(np.divide(*omg.residual_sums(want, div)))) <= 1e-3 * sim.relative_residual() + 1e-7


@pytest.mark.gpu
def test_config3_multigrid_reaches_1e_4_and_the_converged_divergence(oracle_mod):
    """BASELINE.json config 3: 256^3-bounded sparse plume, V-cycles to a relative Poisson residual of 1e-4."""
    import hnanosolver_b200 as H

    O = oracle_mod
    w = synth.WORKLOADS["c3"](with_coords=False)
    grid = H.create_index_grid_from_origins(w.origins, w.voxel_size)
    sim = H.Simulation(grid, 0)
    sim.upload(w.velocity)
    sim.advect_velocity(w.dt)
    sim.divergence(True)
    div0 = sim.divergence_sum_squares(of_advected=True)
    mg = H.Multigrid(grid)
    cycles, rel = sim.pressure_solve_mg(mg, 20, 1e-4, 2, 2, 1.15)
    assert rel <= 1e-4 and cycles <= 8, (cycles, rel)
    sim.subtract_gradient(True)
    div_mg = sim.divergence_sum_squares(of_advected=False)

    def rbgs(iterations):
        sim.divergence(True)
        sim.pressure_solve(iterations, O.omega_compute(w.voxel_size))
        r = sim.relative_residual()
        sim.subtract_gradient(True)
        return r, sim.divergence_sum_squares(of_advected=False)

    # the reference's solve at the top of the SOP's iteration range (I = 100) on the same input, and run far beyond it
    rel_100, div_100 = rbgs(100)
    rel_inf, div_inf = rbgs(3000)
    assert rel < rel_100 and div_mg < div0
    # The divergence and the gradient are 2h central differences while the operator is the compact 7-point Laplacian,
    # so ||div(u_new)|| does not go to zero: it goes to the value the CONVERGED pressure gives, which the reference's own sweep
    # approaches from below as I grows (an under-converged pressure leaves a slightly smaller central-difference divergence).
    # Gate: the multigrid result sits at that converged value (0.3 %), and within 2 % of the reference's I = 100 result.
    assert rel_inf < 5e-3
    assert abs(np.sqrt(div_mg) - np.sqrt(div_inf)) <= 3e-3 * np.sqrt(div_inf), (div_mg, div_inf)
    assert np.sqrt(div_mg) <= 1.02 * np.sqrt(div_100), (div0, div_mg, div_100)


@pytest.mark.gpu
def test_frame_with_the_multigrid_solver_is_a_frame(oracle_mod):
    """hns_state_set_pressure_solver: the frame's other stages are untouched, the pressure stage is the V-cycle solve."""
    import hnanosolver_b200 as H

    w = synth.smoke_sphere(40, 2)
    grid = H.create_index_grid_from_origins(w.origins, w.voxel_size)
    a, b = H.Simulation(grid, len(w.scalars)), H.Simulation(grid, len(w.scalars))
    mg = H.Multigrid(grid)
    for s in (a, b):
        s.upload(w.velocity, w.scalars)
    a.set_pressure_solver(mg, 2, 2, 2, 1.15)
    a.step(40, w.dt)
    b.advect_velocity(w.dt)
    b.divergence(True)
    b.pressure_solve_mg(mg, 2, 0.0, 2, 2, 1.15)
    b.subtract_gradient(True)
    b.advect_scalars(w.dt, 0)
    a.sync(), b.sync()
    assert np.array_equal(a.velocity(), b.velocity()) and np.array_equal(a.aux(1), b.aux(1))
    assert all(np.array_equal(a.scalar(i), b.scalar(i)) for i in range(len(w.scalars)))
    a.set_pressure_solver(None)
    a.upload(w.velocity, w.scalars)
    b.upload(w.velocity, w.scalars)
    a.step(5, w.dt), b.step(5, w.dt)
    assert np.array_equal(a.velocity(), b.velocity())
