"""Host-side setup of the product (mesh tables, FEM blocks, ordering, multifrontal plan)
cross-checked against the independently written oracle and against SciPy's SuperLU."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from flowcontrol_b200.actuator import ActuatorBCParabolicV, ActuatorBCUniformU, ActuatorForceGaussianV
from flowcontrol_b200.fem import RADON7_ETA, RADON7_W, RADON7_XI, ScalarBlocks, p2_shape
from flowcontrol_b200.mesh import TaylorHoodTables
from flowcontrol_b200.multifrontal import BlockFactor, SymbolicFactor, apply_plan_host, build_plan
from flowcontrol_b200.ordering import dissect, postorder
from flowcontrol_b200.problem import DirichletBC, DirichletSet, FlowProblem
from flowcontrol_b200.sensor import SENSOR_TYPE, SensorHorizontalWallShear, SensorPoint
from oracle import flow_oracle as fo
from util import near, unit_square_mesh


@pytest.fixture(scope="module")
def small():
    xy, tri = unit_square_mesh(12, jitter=0.2, seed=3)
    tab = TaylorHoodTables.from_arrays(xy, tri)
    return xy, tri, tab, ScalarBlocks(tab), fo.TaylorHoodMesh(xy, tri)


def test_integer_maps_bit_exact_vs_oracle(small):
    xy, tri, tab, blocks, om = small
    assert np.array_equal(tab.cell_nodes, om.cell_nodes)
    assert np.array_equal(tab.edges, om.edges)
    assert np.array_equal(tab.bnd_edges, om.bnd_edges)
    assert np.array_equal(tab.bnd_cells, om.bnd_edge_cell)
    assert tab.N == om.N and tab.nN == om.nN
    assert np.array_equal(tab.node_xy, om.node_xy)


def test_radon_rule_exact_to_degree_5():
    from math import factorial

    for a in range(6):
        for b in range(6 - a):
            exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            assert np.isclose(np.sum(RADON7_W * RADON7_XI**a * RADON7_ETA**b), exact, rtol=1e-14)


def test_p2_shape_partition_of_unity_and_nodal():
    phi, dphi = p2_shape(np.array([0.2, 0.7]), np.array([0.3, 0.1]))
    assert np.allclose(phi.sum(axis=1), 1.0) and np.allclose(dphi.sum(axis=1), 0.0)
    nodes = np.array([[0, 0], [1, 0], [0, 1], [0.5, 0.5], [0, 0.5], [0.5, 0]], dtype=float)
    phi, _ = p2_shape(nodes[:, 0], nodes[:, 1])
    assert np.allclose(phi, np.eye(6), atol=1e-15)


def test_operators_match_oracle(small):
    xy, tri, tab, blocks, om = small
    ops = fo.Operators(om)
    rng = np.random.default_rng(0)
    U0 = rng.standard_normal(tab.Nv)
    A_p = blocks.saddle_point(200.0, 100.0, U0, shift=0.3)
    A_o = ops.lhs(200.0, 100.0, U0, shift=0.3)
    assert abs(A_p - A_o).max() < 1e-12 * abs(A_o).max()
    A_p = blocks.saddle_point(0.0, 50.0, U0, linearised=False)
    A_o = ops.lhs(0.0, 50.0, U0, newton_terms=False)
    assert abs(A_p - A_o).max() < 1e-12 * abs(A_o).max()
    assert np.allclose(blocks.convection(U0), ops.convection(U0), rtol=0, atol=1e-13 * abs(ops.convection(U0)).max())


def test_colouring_is_conflict_free(small):
    _, _, tab, _, _ = small
    cptr, cells = tab.element_colouring()
    assert sorted(cells.tolist()) == list(range(tab.nT))
    for c in range(len(cptr) - 1):
        nodes = tab.cell_nodes[cells[cptr[c] : cptr[c + 1]]].ravel()
        assert len(np.unique(nodes)) == len(nodes)


def test_dissection_tree_invariants(small):
    _, _, tab, _, _ = small
    tree = dissect(tab, leaf_cells=6)
    owned = np.concatenate([t.own_nodes for t in tree])
    assert sorted(owned.tolist()) == list(range(tab.nN))  # every node eliminated exactly once
    po = postorder(tree)
    seen = set()
    for t in po:
        assert all(c in seen for c in tree[t].children)
        seen.add(t)
    # boundary nodes of a subdomain are owned by strict ancestors
    owner = np.empty(tab.nN, dtype=int)
    for i, t in enumerate(tree):
        owner[t.own_nodes] = i
    for i, t in enumerate(tree):
        anc = set()
        p = t.parent
        while p >= 0:
            anc.add(p)
            p = tree[p].parent
        assert set(owner[t.bnd_nodes].tolist()) <= anc


def _bcs(acts):
    return [
        DirichletBC(lambda x, y: near(y, 1.0), (0, 1), acts[0]),
        DirichletBC(lambda x, y: near(x, 0.0), (0, 1), (0.0, 0.0)),
        DirichletBC(lambda x, y: near(y, 0.0), (1,), (0.0,)),
    ]


def test_dirichlet_override_semantics(small):
    _, _, tab, _, _ = small
    acts = [ActuatorBCUniformU()]
    d = DirichletSet(tab, _bcs(acts), acts)
    corner = int(np.flatnonzero(near(tab.node_xy[:, 0], 0.0) & near(tab.node_xy[:, 1], 1.0))[0])
    j = int(np.searchsorted(d.dofs, corner))
    assert d.dofs[j] == corner and d.shape[0, j] == 0.0  # left wall (later BC) overrides the lid
    lid_mid = int(np.flatnonzero(near(tab.node_xy[:, 1], 1.0) & (tab.node_xy[:, 0] > 0.3) & (tab.node_xy[:, 0] < 0.7))[0])
    assert d.shape[0, np.searchsorted(d.dofs, lid_mid)] == 1.0
    # bottom: only u_y constrained
    bot = int(np.flatnonzero(near(tab.node_xy[:, 1], 0.0) & (tab.node_xy[:, 0] > 0.3))[0])
    assert bot not in set(d.dofs.tolist()) and (bot + tab.nN) in set(d.dofs.tolist())
    assert np.allclose(d.values([2.5]), d.const + 2.5 * d.shape[0])


def test_multifrontal_matches_superlu_and_plan_matches_factor(small):
    _, _, tab, blocks, _ = small
    rng = np.random.default_rng(1)
    U0 = 0.3 * rng.standard_normal(tab.Nv)
    acts = [ActuatorBCUniformU()]
    d = DirichletSet(tab, _bcs(acts), acts)
    A = blocks.saddle_point(300.0, 100.0, U0)
    sym = SymbolicFactor(tab, d.free, leaf_cells=4)
    assert sorted(sym.perm.tolist()) == np.flatnonzero(d.free).tolist()
    fac = BlockFactor(sym, A)
    b = rng.standard_normal((sym.n, 3))
    x = fac.solve(b)
    Aff = A[sym.perm][:, sym.perm].tocsc()
    xref = spla.splu(Aff).solve(b)
    assert np.linalg.norm(x - xref) / np.linalg.norm(xref) < 1e-11
    plan0 = build_plan(fac, top_levels=0, cluster_rows=0)
    assert np.abs(apply_plan_host(plan0, b) - x).max() < 1e-12 * np.abs(x).max()
    assert plan0.i0.max() <= plan0.zrow and plan0.nnz == sym.factor_entries() and len(plan0.asm_dst) == 0
    for tl in (1, 2, 4, 99):  # merged top of the tree (explicit inverse of the top Schur complement)
        assert np.abs(apply_plan_host(build_plan(fac, top_levels=tl, cluster_rows=0), b) - x).max() < 1e-12 * np.abs(x).max()
    # shared-memory subtree clusters below the pull-form launches: every cut (rows resident per cluster, height limit)
    # gives the same solution, produces every x and y row exactly once, and keeps the factor entry count
    seen_tiers = set()
    for rows, hmax in ((40, 1), (60, 99), (120, 2), (120, 99), (400, 3), (10000, 99)):  # (min_tier_clusters keeps small upper tiers out)
        pc = build_plan(fac, top_levels=2, cluster_rows=rows, cluster_height=hmax, min_tier_clusters=1 if rows != 120 else 6)
        assert np.abs(apply_plan_host(pc, b) - x).max() < 1e-12 * np.abs(x).max(), (rows, hmax)
        ncl = len(pc.cl_fptr) - 1
        assert ncl > 0 and pc.tier_ptr[-1] == ncl
        seen_tiers.add(len(pc.tier_ptr) - 1)
        own = np.concatenate([np.arange(c, c + w) for c, w in zip(pc.fr_c0, pc.fr_w)])
        bwp = np.arange(len(pc.blk_K)) >= pc.launch_ptr[pc.n_forward_launches]
        pull = [np.arange(o, o + r) for o, r in zip(pc.blk_out0[bwp], pc.blk_M[bwp])]
        assert sorted(np.concatenate([own] + pull).tolist()) == list(range(sym.n))
        for q in range(ncl):  # the resident vector of a cluster respects the requested size
            f0, f1 = pc.cl_fptr[q], pc.cl_fptr[q + 1]
            assert pc.fr_w[f0:f1].sum() + pc.fr_m[f1 - 1] <= rows
        assert pc.nU <= plan0.nU  # update vectors inside clusters never reach global memory
    assert max(seen_tiers) >= 2  # some cut produced clusters that import from lower clusters
    # amalgamated upper tree: fronts with four children (virtual update vectors summed right before their launch)
    for above in (1, 2):
        syma = SymbolicFactor(tab, d.free, leaf_cells=4, amalgamate_above=above)
        assert max(len(c) for c in syma.children) == 4 and max(s_.height for s_ in syma.supernodes) < max(s_.height for s_ in sym.supernodes)
        faca = BlockFactor(syma, A)
        ba = rng.standard_normal((syma.n, 2))
        xa = spla.splu(A[syma.perm][:, syma.perm].tocsc()).solve(ba)
        assert np.linalg.norm(faca.solve(ba) - xa) / np.linalg.norm(xa) < 1e-11
        for tl, rows in ((1, 0), (0, 0), (1, 80)):
            pa = build_plan(faca, top_levels=tl, cluster_rows=rows, min_tier_clusters=1)
            assert np.abs(apply_plan_host(pa, ba) - xa).max() < 1e-11 * np.abs(xa).max(), (above, tl, rows)
            if rows == 0:
                assert pa.asm_lptr[-1] == len(pa.asm_dst) and len(pa.asm_lptr) == len(pa.launch_ptr)
                assert len(pa.asm_dst) > (len(pa.launch_ptr) > 1 and tl) * 0  # gather-sum rows exist for the four-child fronts
    # depth-bounded dissection: no deeper than an even dissection, leaves within the requested size, same solution
    symb = SymbolicFactor(tab, d.free, leaf_cells=4, balanced=True)
    assert max(s_.depth for s_ in symb.supernodes) == int(np.ceil(np.log2(tab.nT / 4)))
    assert max(s_.depth for s_ in symb.supernodes) <= max(s_.depth for s_ in sym.supernodes)
    assert sorted(symb.perm.tolist()) == sorted(sym.perm.tolist())
    facb = BlockFactor(symb, A)
    bb = rng.standard_normal((symb.n, 2))
    xb = spla.splu(A[symb.perm][:, symb.perm].tocsc()).solve(bb)
    for rows in (0, 80):
        pb = build_plan(facb, top_levels=2, cluster_rows=rows, min_tier_clusters=1)
        assert np.abs(apply_plan_host(pb, bb) - xb).max() < 1e-11 * np.abs(xb).max(), rows
    # pre-summed gathers: above the given height every forward block reads ONE plane (y_t, written by a gather-sum)
    for hp in (1, 3):
        pp = build_plan(facb, top_levels=2, presum_height=hp)
        fwd = np.arange(len(pp.blk_K)) < pp.launch_ptr[pp.n_forward_launches]
        assert (pp.blk_nsrc[fwd] == 1).sum() > (pb.blk_nsrc[: pb.launch_ptr[pb.n_forward_launches]] == 1).sum() or hp > 1
        assert len(pp.asm_dst) > len(pb.asm_dst) and pp.asm_lptr[-1] == len(pp.asm_dst)
        assert np.abs(apply_plan_host(pp, bb) - xb).max() < 1e-11 * np.abs(xb).max(), hp
    # in-place leaves: no y rows stored for them, their backward blocks read b where it lies and are marked ystore = -2
    pl = build_plan(fac, top_levels=2, leaf_inplace=True)
    assert np.abs(apply_plan_host(pl, b) - x).max() < 1e-12 * np.abs(x).max()
    for rows in (80, 400):  # with subtree clusters below (the plan of ensembles of up to 32 trajectories): leaves outside clusters only
        plc = build_plan(fac, top_levels=2, cluster_rows=rows, cluster_height=3, min_tier_clusters=1, leaf_inplace=True)
        assert np.abs(apply_plan_host(plc, b) - x).max() < 1e-12 * np.abs(x).max(), rows
    marked = pl.blk_ystore == -2
    assert marked.sum() == sum(1 for i, c in enumerate(sym.children) if not c and len(sym.supernodes[i].struct) and sym.supernodes[i].c1 > sym.supernodes[i].c0)
    assert (np.flatnonzero(marked) >= pl.launch_ptr[pl.n_forward_launches]).all()
    first_bw = int(pl.launch_ptr[pl.n_forward_launches])
    for q in np.flatnonzero(marked):  # in the backward sweep nobody but the block itself reads the rows it leaves as b
        rows = np.arange(pl.blk_out0[q], pl.blk_out0[q] + pl.blk_M[q])
        assert np.isin(rows, pl.i0[pl.blk_iptr[q] : pl.blk_iptr[q] + pl.blk_K[q]]).all()
        for q2 in range(first_bw, len(pl.blk_K)):
            if q2 != q:
                assert not np.isin(pl.i0[pl.blk_iptr[q2] : pl.blk_iptr[q2] + pl.blk_K[q2]], rows).any()
    ysl = pl.blk_ystore >= 0  # y rows are stored for every front but the in-place leaves
    rows_yl = np.concatenate([np.arange(o, o + k) for o, k in zip(pl.blk_ystore[ysl], pl.blk_K[ysl])] + [pl.asm_dst])
    leaf_rows = np.concatenate([np.arange(pl.blk_out0[q], pl.blk_out0[q] + pl.blk_M[q]) for q in np.flatnonzero(marked)])
    assert sorted(rows_yl.tolist()) == sorted(set(range(sym.n, 2 * sym.n)) - set((leaf_rows + sym.n).tolist()))
    plan = build_plan(fac, top_levels=2, cluster_rows=0)
    # every x row and every y row is produced exactly once; blocks of one launch never read rows
    # that the same launch writes
    n = sym.n
    bw = np.arange(len(plan.blk_K)) >= plan.launch_ptr[plan.n_forward_launches]
    rows_x = np.concatenate([np.arange(o, o + r) for o, r in zip(plan.blk_out0[bw], plan.blk_M[bw])])
    assert sorted(rows_x.tolist()) == list(range(n))
    ys = plan.blk_ystore >= 0
    rows_y = np.concatenate([np.arange(o, o + k) for o, k in zip(plan.blk_ystore[ys], plan.blk_K[ys])] + [plan.asm_dst])
    assert sorted(rows_y.tolist()) == list(range(n, 2 * n))  # y rows: stored by a forward block or assembled for the top
    for l in range(len(plan.launch_ptr) - 1):
        written, read = set(), set()
        for q in range(plan.launch_ptr[l], plan.launch_ptr[l + 1]):
            K, M = int(plan.blk_K[q]), int(plan.blk_M[q])
            written |= set(range(plan.blk_out0[q], plan.blk_out0[q] + M))
            if plan.blk_ystore[q] >= 0:
                written |= set(range(plan.blk_ystore[q], plan.blk_ystore[q] + K))
            sl = slice(plan.blk_iptr[q], plan.blk_iptr[q] + K)
            read |= set(plan.i0[sl].tolist())
            if plan.blk_nsrc[q] == 3:
                read |= set(plan.i1[sl].tolist()) | set(plan.i2[sl].tolist())
            if plan.blk_eptr[q] >= 0:
                es = slice(plan.blk_eptr[q], plan.blk_eptr[q] + M)
                read |= set(plan.e0[es].tolist()) | set(plan.e1[es].tolist())
        read.discard(-1)
        assert plan.zrow not in written
        assert not (written & read)


def test_problem_rhs_and_lifting_match_oracle():
    """FlowProblem.host_step_rhs + block solve == oracle step (BC actuator with u_ctrl != 0, force actuator)."""
    xy, tri = unit_square_mesh(10, jitter=0.15, seed=5)
    tab = TaylorHoodTables.from_arrays(xy, tri)
    blocks = ScalarBlocks(tab)
    rng = np.random.default_rng(2)
    acts = [ActuatorBCUniformU(), ActuatorForceGaussianV(sigma=0.15, position=np.array([0.4, 0.5]))]
    bcs = _bcs(acts)
    sensors = [SensorPoint(sensor_type=SENSOR_TYPE.V, position=np.array([0.31, 0.52])),
               SensorPoint(sensor_type=SENSOR_TYPE.P, position=np.array([0.6, 0.4])),
               SensorHorizontalWallShear(sensor_type=SENSOR_TYPE.OTHER, x_sensor_left=0.2, x_sensor_right=0.8, y_sensor=0.0)]
    UP0 = np.concatenate([0.5 * rng.standard_normal(tab.Nv), rng.standard_normal(tab.nV)])
    prob = FlowProblem(tab, blocks, 80.0, 0.01, bcs, acts, sensors, UP0, leaf_cells=4)
    case = fo.CaseSpec(
        name="t", mesh_file="", Re=80.0, dt=0.01, uinf=1.0,
        bcs_pert=[fo.DirichletSpec(lambda x, y: fo.near(y, 1.0), (0, 1), ("actuator", 0)),
                  fo.DirichletSpec(lambda x, y: fo.near(x, 0.0), (0, 1), (0.0, 0.0)),
                  fo.DirichletSpec(lambda x, y: fo.near(y, 0.0), (1,), (0.0,))],
        bcs_full=[], actuators=[fo.ActuatorSpec("bc", fo.uniform_u()), fo.ActuatorSpec("force", fo.gaussian_v(0.15, (0.4, 0.5)))],
        sensors=[fo.SensorSpec("point", comp=1, position=(0.31, 0.52)), fo.SensorSpec("point", comp=2, position=(0.6, 0.4)),
                 fo.SensorSpec("wall_shear", x_left=0.2, x_right=0.8, y=0.0)],
        initial_guess=None,
    )
    orc = fo.FlowOracle(case, xy, tri)
    orc.set_base_flow(UP0)
    ic = np.concatenate([0.1 * rng.standard_normal(tab.Nv), np.zeros(tab.nV)])
    orc.init_time_stepping(ic=ic)
    assert np.array_equal(prob.dirichlet.dofs, orc.bc_pert.dofs)
    u_n, u_nn, order = ic[: tab.Nv].copy(), ic[: tab.Nv].copy(), 1
    S = np.zeros((prob.ns, tab.N))
    for s in range(prob.ns):
        sl = slice(prob.sensor_ptr[s], prob.sensor_ptr[s + 1])
        S[s, prob.sensor_idx[sl]] = prob.sensor_val[sl]
    for step, uc in enumerate(([0.3, -0.7], [0.1, 0.4], [-0.2, 0.9])):
        orc.step(uc)
        b = prob.host_step_rhs(order, u_n, u_nn, uc)
        x = np.zeros(tab.N)
        x[prob.sym.perm] = prob.factors[order].solve(b)
        x[prob.dirichlet.dofs] = prob.dirichlet.values(uc)
        assert np.linalg.norm(x[: tab.Nv] - orc.up[: tab.Nv]) / np.linalg.norm(orc.up[: tab.Nv]) < 1e-10
        assert np.linalg.norm(x[tab.Nv :] - orc.up[tab.Nv :]) / np.linalg.norm(orc.up[tab.Nv :]) < 1e-9
        assert np.allclose(S @ x, orc.y_meas, rtol=1e-9, atol=1e-12)
        u_nn, u_n, order = u_n, x[: tab.Nv].copy(), 2


def test_force_actuator_unit_norm(small):
    _, _, tab, blocks, _ = small
    a = ActuatorForceGaussianV(sigma=0.1, position=np.array([0.5, 0.5]))
    a.normalise(tab.node_xy, blocks.Mv)
    sx, sy = a.shape(tab.node_xy[:, 0], tab.node_xy[:, 1])
    s = np.concatenate([sx, sy])
    assert np.isclose(s @ (blocks.Mv @ s), 1.0, rtol=1e-13)  # test_actuator.py:155-161 in the reference


def test_parabolic_width_helper():
    assert np.isclose(ActuatorBCParabolicV.angular_size_deg_to_width(10, 0.5), 0.5 * np.sin(np.deg2rad(5)))


def test_force_coefficient_rows_match_oracle_and_closed_surface_identity():
    """SensorForceCoefficient rows (product) vs the oracle's quadrature evaluation on random fields, and the exact
    identity  oint p n ds = -Area * grad p  for a linear pressure on a closed polygonal hole (u = 0)."""
    from flowcontrol_b200.sensor import SensorForceCoefficient

    # square [0,1]^2 with a square hole [0.25,0.75]^2 removed (structured 8x8 grid)
    xy, tri = unit_square_mesh(8)
    cent = xy[tri].mean(axis=1)
    keep = ~((cent[:, 0] > 0.25) & (cent[:, 0] < 0.75) & (cent[:, 1] > 0.25) & (cent[:, 1] < 0.75))
    tri = tri[keep]
    used = np.unique(tri)
    remap = -np.ones(len(xy), dtype=np.int64)
    remap[used] = np.arange(len(used))
    xy, tri = xy[used], remap[tri].astype(np.int32)
    tab = TaylorHoodTables.from_arrays(xy, tri)
    om = fo.TaylorHoodMesh(xy, tri)

    def hole(x, y):
        return (x > 0.2) & (x < 0.8) & (y > 0.2) & (y < 0.8)

    nu, uinf, D = 0.013, 1.3, 0.5
    rows = [SensorForceCoefficient(sensor_type=SENSOR_TYPE.OTHER, inside=hole, component=c, nu=nu, uinf=uinf, D=D).row(tab)
            for c in (1, 0)]  # lift, drag
    rng = np.random.default_rng(4)
    up = rng.standard_normal(tab.N)
    cl, cd = fo.force_coefficients(om, hole, up, nu, uinf, D)
    assert np.isclose(rows[0][1] @ up[rows[0][0]], cl, rtol=1e-12)
    assert np.isclose(rows[1][1] @ up[rows[1][0]], cd, rtol=1e-12)
    # u = 0, p = 2x - 3y: force on the hole = -Area * grad p (n points into the hole)
    up = np.zeros(tab.N)
    up[tab.Nv :] = 2.0 * tab.xy[:, 0] - 3.0 * tab.xy[:, 1]
    q = 0.5 * uinf**2 * D
    assert np.isclose(rows[1][1] @ up[rows[1][0]], -0.25 * 2.0 / q, rtol=1e-12)   # drag
    assert np.isclose(rows[0][1] @ up[rows[0][0]], -0.25 * -3.0 / q, rtol=1e-12)  # lift


def test_crank_nicolson_host_step_matches_oracle():
    """time_scheme='cn' (nsforms.py:191-236): product operators (A_cn factor, explicit operator E_cn, averaged force,
    BC lifting) vs the oracle's Crank-Nicolson step, BC + force actuators with time-varying u_ctrl."""
    xy, tri = unit_square_mesh(10, jitter=0.15, seed=5)
    tab = TaylorHoodTables.from_arrays(xy, tri)
    blocks = ScalarBlocks(tab)
    rng = np.random.default_rng(2)
    acts = [ActuatorBCUniformU(), ActuatorForceGaussianV(sigma=0.15, position=np.array([0.4, 0.5]))]
    sensors = [SensorPoint(sensor_type=SENSOR_TYPE.V, position=np.array([0.31, 0.52]))]
    UP0 = np.concatenate([0.5 * rng.standard_normal(tab.Nv), rng.standard_normal(tab.nV)])
    prob = FlowProblem(tab, blocks, 80.0, 0.01, _bcs(acts), acts, sensors, UP0, leaf_cells=4, time_scheme="cn")
    case = fo.CaseSpec(
        name="t", mesh_file="", Re=80.0, dt=0.01, uinf=1.0,
        bcs_pert=[fo.DirichletSpec(lambda x, y: fo.near(y, 1.0), (0, 1), ("actuator", 0)),
                  fo.DirichletSpec(lambda x, y: fo.near(x, 0.0), (0, 1), (0.0, 0.0)),
                  fo.DirichletSpec(lambda x, y: fo.near(y, 0.0), (1,), (0.0,))],
        bcs_full=[], actuators=[fo.ActuatorSpec("bc", fo.uniform_u()), fo.ActuatorSpec("force", fo.gaussian_v(0.15, (0.4, 0.5)))],
        sensors=[fo.SensorSpec("point", comp=1, position=(0.31, 0.52))], initial_guess=None,
    )
    orc = fo.FlowOracle(case, xy, tri, time_scheme="cn")
    orc.set_base_flow(UP0)
    ic = np.concatenate([0.1 * rng.standard_normal(tab.Nv), np.zeros(tab.nV)])
    orc.init_time_stepping(ic=ic)
    u_n, prev = ic[: tab.Nv].copy(), np.zeros(2)
    for uc in ([0.3, -0.7], [0.1, 0.4], [-0.2, 0.9], [0.0, 0.0]):
        orc.step(uc)
        b = prob.host_step_rhs(2, u_n, None, uc, prev)
        x = np.zeros(tab.N)
        x[prob.sym.perm] = prob.factors[2].solve(b)
        x[prob.dirichlet.dofs] = prob.dirichlet.values(uc)
        assert np.linalg.norm(x[: tab.Nv] - orc.up[: tab.Nv]) / np.linalg.norm(orc.up[: tab.Nv]) < 1e-10
        assert np.linalg.norm(x[tab.Nv :] - orc.up[tab.Nv :]) / np.linalg.norm(orc.up[tab.Nv :]) < 1e-9
        u_n, prev = x[: tab.Nv].copy(), np.asarray(uc, dtype=float)


def test_assembly_position_map_is_bit_exact(small):
    """Integer maps of the device matrix assembly (assembly.py): every element-matrix entry (cell, a, b) lands on the CSR
    entry (node_a, node_b) of the scalar P2 pattern, and cells of one colour share no node (atomic-free scatter)."""
    import scipy.sparse as sp

    from flowcontrol_b200.assembly import DeviceAdvectionAssembler

    _, _, tab, blocks, _ = small
    asm = DeviceAdvectionAssembler(tab, blocks)
    cn = np.asarray(tab.cell_nodes)
    rows, cols = np.repeat(cn, 6, axis=1).ravel(), np.tile(cn, (1, 6)).ravel()
    row_of = np.repeat(np.arange(tab.nN), np.diff(asm.indptr))
    assert np.array_equal(row_of[asm.pos], rows) and np.array_equal(asm.indices[asm.pos], cols)
    counts = np.bincount(asm.pos, minlength=len(asm.indices))
    ref = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(tab.nN, tab.nN)).tocsr()
    ref.sort_indices()
    assert np.array_equal(ref.indices, asm.indices) and np.array_equal(ref.data.astype(np.int64), counts)
    cptr, ccells = tab.element_colouring()
    assert sorted(ccells.tolist()) == list(range(tab.nT))
    for c in range(len(cptr) - 1):
        nodes = cn[ccells[cptr[c] : cptr[c + 1]]].ravel()
        assert len(np.unique(nodes)) == len(nodes)


def test_front_maps_of_the_device_factorisation_reproduce_the_host_factor(small):
    """devfactor.FrontMaps (integer maps of the GPU numeric factorisation: matrix entry -> front position, child boundary ->
    parent front position, levels) drive a numpy emulation of the device algorithm to exactly the host factor's blocks."""
    import scipy.sparse as sp

    from flowcontrol_b200.devfactor import FrontMaps

    _, _, tab, blocks, _ = small
    rng = np.random.default_rng(3)
    U0 = 0.3 * rng.standard_normal(tab.Nv)
    acts = [ActuatorBCUniformU()]
    d = DirichletSet(tab, _bcs(acts), acts)
    A = blocks.saddle_point(300.0, 100.0, U0)
    for above in (0, 2):
        sym = SymbolicFactor(tab, d.free, leaf_cells=4, amalgamate_above=above)
        fac = BlockFactor(sym, A)
        Ap = sp.csr_matrix(A)[sym.perm][:, sym.perm].tocsr()
        Ap.sort_indices()
        maps = FrontMaps(sym, Ap)
        assert maps.a_ptr[-1] == Ap.nnz and sorted(maps.a_src.tolist()) == list(range(Ap.nnz))  # every entry lands in exactly one front
        for lv in range(len(maps.level_ptr) - 1):  # children sit in earlier levels
            lvl = set(maps.level_fronts[maps.level_ptr[lv] : maps.level_ptr[lv + 1]].tolist())
            for f in lvl:
                assert not (set(sym.children[f]) & lvl)
        for (E, Fi, G), (E2, Fi2, G2) in zip(fac.blocks, maps.emulate(Ap.data)):
            assert np.array_equal(E, E2) and np.array_equal(Fi, Fi2) and np.array_equal(G, G2)
