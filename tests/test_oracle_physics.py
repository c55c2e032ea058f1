"""Known answers and invariants for the fp64 oracle: these need no MuJoCo.
PARITY UNPINNED against MuJoCo itself -- see oracle/oracle_internal.h."""
import numpy as np
import pytest

from gym_so100_c_b200 import model as M
from gym_so100_c_b200.mjcf import quat_to_mat
from oracle import so100_oracle as O

START = np.array(M.SO100_START_ARM_POSE)


def _orc(blob, n=1):
    return O.Oracle(blob, n)


def test_loader_structure(model_rec):
    m = model_rec
    assert (int(m["nq"]), int(m["nv"]), int(m["nu"]), int(m["nbody"])) == (13, 12, 6, 13)
    assert int(m["ngeom_all"]) == 38 and int(m["ngeom"]) == 25 and int(m["nsite"]) == 5
    assert int(m["npair"]) == 191
    vnum = [int(v) for v in m["geom_vnum"][:25] if v > 0]
    assert vnum == [8, 150, 400, 421, 516, 525, 16, 221, 12, 8, 187]        # hull sizes, SURVEY 8a-M
    assert int(m["geom_mjid"][int(m["cg_cube"])]) == 32 and int(m["geom_mjid"][int(m["cg_table"])]) == 0
    pads = [int(m["geom_mjid"][g]) for g in range(25) if (int(m["pad_mask"]) >> g) & 1]
    assert pads == [20, 21, 22, 23, 28, 29, 30, 31]
    assert int(m["nsubstep"]) == 10 and float(m["timestep"]) == 0.002 and float(m["impratio"]) == 10.0


def test_setconst_known_answers(model_rec):
    m = model_rec
    np.testing.assert_allclose(m["dof_M0"][:6], [0.13149, 0.125088, 0.108872, 0.10116, 0.100043, 0.100028], atol=2e-6)
    np.testing.assert_allclose(m["act_kv"], [5.1282, 5.0018, 4.6663, 4.4980, 4.4731, 4.4728], atol=1e-4)
    np.testing.assert_allclose(m["dof_invweight0"][6:], [20, 20, 20, 500, 500, 500], rtol=1e-12)
    np.testing.assert_allclose(m["body_invweight0"][11], [20, 500], rtol=1e-12)
    # cube contacts: mixed parameters (SURVEY 8a-M "Pair mixing")
    for p in range(191):
        g1, g2 = int(m["geom_mjid"][m["pair_g1"][p]]), int(m["geom_mjid"][m["pair_g2"][p]])
        if 32 in (g1, g2):
            assert int(m["pair_condim"][p]) == 4
            other = g2 if g1 == 32 else g1
            if other in (20, 21, 22, 23, 28, 29, 30, 31):
                np.testing.assert_allclose(m["pair_solref"][p], [0.01, 1]); np.testing.assert_allclose(m["pair_solimp"][p][:3], [2, 1, 0.01])
                assert g1 == other                      # lower id first for box-box
            else:
                np.testing.assert_allclose(m["pair_solref"][p], [0.015, 1]); np.testing.assert_allclose(m["pair_solimp"][p], [1.45, 0.975, 0.0055, 0.5, 2])
                assert g1 == 32                         # box type sorts before mesh: ("red_box","table")
        else:
            assert int(m["pair_condim"][p]) == 3


def test_fk_known_answers(model_blob, model_rec):
    o = _orc(model_blob)
    q = np.zeros((1, 13)); q[0, 6:] = model_rec["qpos0"][6:]
    for arm, ee in ((np.zeros(6), (-0.025629, 0.500002, 0.096198)), (START, (-0.054752, 0.500002, 0.149592)),
                    (np.array([0, -1.57, 1.57, 1.57, -1.57, 0.0]), (-0.230116, 0.500001, 0.116647)),
                    (np.array([0, -3.32, 3.11, 1.18, 0, -0.174]), (-0.310208, 0.500001, 0.060715))):
        q[0, :6] = arm
        o.set_state(q, np.zeros((1, 12)), np.zeros((1, 6)), np.zeros((1, 12)))
        o.forward()
        d = o.dyn(0)
        np.testing.assert_allclose(d["sites"][2], ee, atol=2e-6)
        np.testing.assert_allclose(d["sites"][4], (-0.2, 0.7, 0.021), atol=1e-12)
        np.testing.assert_allclose(d["M"], M.mass_matrix(model_rec, q[0]), atol=1e-12)   # C vs numpy definition
    np.testing.assert_allclose(np.diag(d["M"])[:6] * 0 + 1, 1)
    q[0, :6] = START
    np.testing.assert_allclose(np.diag(M.mass_matrix(model_rec, q[0]))[:6],
                               [0.125589, 0.12127, 0.108873, 0.10116, 0.100044, 0.100028], atol=2e-6)


def test_rne_bias_against_lagrangian_finite_differences(model_blob, model_rec):
    """bias = Mdot qd - 1/2 d(qd^T M qd)/dq + dV/dq, with M(q) from the numpy definition (independent of the C RNE)."""
    rng = np.random.default_rng(0)
    o = _orc(model_blob)
    g = 9.81
    masses = model_rec["body_mass"]

    def potential(q):
        xpos, xquat = M.fk(model_rec, q)
        v = 0.0
        for b in range(1, 13):
            if masses[b] > 0 and model_rec["body_weldid"][b] != 0:
                v += masses[b] * g * (xpos[b] + quat_to_mat(xquat[b]) @ model_rec["body_ipos"][b])[2]
        return v

    for _ in range(3):
        q = np.zeros(13); q[:6] = START + rng.uniform(-0.5, 0.5, 6); q[6:] = model_rec["qpos0"][6:]
        qd = np.zeros(12); qd[:6] = rng.uniform(-1, 1, 6)
        o.set_state(q[None], qd[None], q[None, :6], np.zeros((1, 12)))
        o.forward()
        bias = o.dyn(0)["bias"]
        eps = 1e-6
        dM = np.zeros((6, 12, 12)); dV = np.zeros(6)
        for k in range(6):
            qp, qm = q.copy(), q.copy(); qp[k] += eps; qm[k] -= eps
            dM[k] = (M.mass_matrix(model_rec, qp) - M.mass_matrix(model_rec, qm)) / (2 * eps)
            dV[k] = (potential(qp) - potential(qm)) / (2 * eps)
        Mdot = np.einsum("kij,k->ij", dM, qd[:6])
        want = (Mdot @ qd)[:6] - 0.5 * np.einsum("kij,i,j->k", dM, qd, qd) + dV
        np.testing.assert_allclose(bias[:6], want, atol=2e-7)
        np.testing.assert_allclose(bias[6:], [0, 0, 0.05 * g, 0, 0, 0], atol=1e-12)


def test_free_fall_and_touchdown(model_blob):
    """Appendix B (i): frictionloss saturates, qacc_z = -9.81 + 0.01/0.05 = -9.61 until the cube reaches the table."""
    o = _orc(model_blob)
    o.reset(box_pose=np.array([[-0.2, 0.45, 0.05, 1, 0, 0, 0]]))
    o.forward()
    assert o.dyn(0)["qacc"][8] == pytest.approx(-9.61, abs=1e-9)
    k_touch = None
    for k in range(1, 60):
        o.substeps(1)
        qp, qv, _, _ = o.get_state()
        if o.contacts(0):
            k_touch = k
            break
        assert qv[0, 8] == pytest.approx(-9.61 * 0.002 * k, rel=1e-9)
    # z(k) = 0.05 - 9.61 h^2 k(k+1)/2 first drops below 0.02 at k = 40 (40*41 = 1640 > 1560.9 > 39*40);
    # the contact list seen after call k is the one built at the START of substep k (state after k-1 steps)
    assert k_touch == 41
    c = o.contacts(0)[0]
    assert (c["geom1"], c["geom2"]) == (32, 0) and c["normal"][2] == pytest.approx(-1.0)


def test_cube_comes_to_rest_and_arm_holds(model_blob):
    """Cube lands, rests with normal force m g; Appendix B (ii): the arm holds its start pose within actuator error."""
    o = _orc(model_blob)
    o.reset(box_pose=np.array([[-0.2, 0.45, 0.05, 1, 0, 0, 0]]))
    o.substeps(600)
    qp, qv, _, _ = o.get_state()
    o.forward()
    cs = o.contacts(0)
    assert len(cs) == 1 and cs[0]["force"][0] == pytest.approx(0.05 * 9.81, rel=2e-3)
    assert qp[0, 8] == pytest.approx(0.02, abs=2e-5) and np.abs(qv[0, 6:]).max() < 1e-3
    assert np.abs(qp[0, :6] - START).max() < 0.02 and np.abs(qv[0, :6]).max() < 1e-3
    s = o.solver(0)
    assert s["grad"] < 1e-10 and s["iters"] < 20                                   # Appendix B (iv): KKT residual


def test_solver_kkt_and_cone_feasibility(model_blob):
    """Forces lie in the friction cone and the gradient vanishes on contact-rich states."""
    import scenarios
    n = 16
    qpos, qvel, ctrl = scenarios.cube_in_bin(n)
    o = _orc(model_blob, n)
    o.set_state(qpos, qvel, ctrl, np.zeros((n, 12)))
    o.forward()
    for i in range(n):
        assert o.solver(i)["grad"] < 1e-9
        for c in o.contacts(i):
            f = c["force"]
            assert f[0] >= -1e-12
            # elliptic cone: (f1/mu1)^2 + (f2/mu2)^2 + (f3/mu3)^2 <= f0^2
            assert (f[1] / 1.0) ** 2 + (f[2] / 1.0) ** 2 + (f[3] / 0.005) ** 2 <= f[0] ** 2 * (1 + 1e-6) + 1e-12


def test_sat_equals_epa_on_boxes():
    """Separating-axis result == GJK/EPA minimum translation on random overlapping boxes (both narrow phases)."""
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(400):
        def rand_rot():
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            return quat_to_mat(q)
        hA, hB = rng.uniform(0.01, 0.05, 3), rng.uniform(0.01, 0.05, 3)
        RA, RB = rand_rot(), rand_rot()
        cA = np.zeros(3); cB = rng.normal(size=3); cB *= rng.uniform(0.02, 0.07) / np.linalg.norm(cB)
        sn, sd, en, ed, npts, hit = O.test_box_pair(cA, RA.ravel(), hA, cB, RB.ravel(), hB)
        assert (npts > 0) == bool(hit)
        if not hit:
            continue
        # EPA is the exact minimum translation; SAT may keep a face axis that is up to 5% (+ the edge
        # penalty 2e-6/sin) deeper than the best edge axis.  Where the depths agree the axes must agree.
        assert ed - 1e-9 <= sd <= ed * 1.05 + 1e-4
        if abs(sd - ed) < 1e-9:
            # depth is stationary in the normal, so an EPA depth tolerance of 1e-11 pins the normal to ~1e-5
            np.testing.assert_allclose(sn, en, atol=2e-4)
            checked += 1
    assert checked > 50


def test_hull_contacts_match_qhull_minkowski_depth(model_blob, model_rec):
    """Convex (GJK/EPA) contacts against an independent computation: the penetration depth of two convex hulls is the distance from
    the origin to the nearest facet of their Minkowski difference, computed here with qhull (scipy) on all pairwise vertex
    differences -- no code shared with the oracle's GJK/EPA.  Depth must agree to 2e-6 m and, where the nearest facet is unique,
    the contact normal to 1e-5 (the oracle snaps normals within 1e-3 rad of a box face, cos = 1 - 5e-7)."""
    from scipy.spatial import ConvexHull
    import scenarios
    m = model_rec
    ng = int(m["ngeom"])
    by_mjid = {int(m["geom_mjid"][g]): g for g in range(ng)}

    def world_vertices(g, xpos, xquat):
        b = int(m["geom_body"][g])
        Rb = quat_to_mat(xquat[b])
        if int(m["geom_vnum"][g]) > 0:
            a = int(m["geom_vadr"][g])           # hull vertices are stored in the BODY frame (model.compile_model)
            return xpos[b] + np.asarray(m["vert"][a:a + int(m["geom_vnum"][g])]) @ Rb.T
        Rg = Rb @ quat_to_mat(m["geom_quat"][g])
        pg = xpos[b] + Rb @ m["geom_pos"][g]
        h = np.asarray(m["geom_size"][g])
        local = np.array([[sx * h[0], sy * h[1], sz * h[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
        return pg + local @ Rg.T

    n, checked, normals = 24, 0, 0
    for name in ("arm_hull_contacts", "grasp_hull_contacts"):
        qpos, qvel, ctrl = scenarios.ALL[name](n)
        o = _orc(model_blob, n)
        o.set_state(qpos, qvel, ctrl, np.zeros((n, 12)))
        o.forward()
        for i in range(n):
            d = o.dyn(i)
            for c in o.contacts(i):
                g1, g2 = by_mjid[c["geom1"]], by_mjid[c["geom2"]]
                if max(int(m["geom_vnum"][g1]), int(m["geom_vnum"][g2])) <= 8:
                    continue                                  # box-like pair: separating-axis path, covered by test_sat_equals_epa_on_boxes
                A, B = world_vertices(g1, d["xpos"], d["xquat"]), world_vertices(g2, d["xpos"], d["xquat"])
                diff = (A[:, None, :] - B[None, :, :]).reshape(-1, 3)
                hull = ConvexHull(diff)
                off = hull.equations[:, 3]                    # facet: normal . x + off <= 0 inside
                assert (off < 1e-9).all(), "the oracle reports a contact but the hulls do not overlap"
                order = np.argsort(-off)
                depth = -off[order[0]]
                assert abs(depth - (-c["dist"])) < 2e-6, (name, i, c["geom1"], c["geom2"], depth, -c["dist"])
                checked += 1
                # unique nearest facet (coplanar qhull triangles share their normal): compare the direction
                nf = hull.equations[order[0], :3]
                others = [k for k in order[1:40] if np.dot(hull.equations[k, :3], nf) < 1 - 1e-9]
                if not others or -off[others[0]] > depth + 1e-5:
                    assert abs(abs(np.dot(nf, c["normal"])) - 1.0) < 1e-5, (name, i, nf, c["normal"])
                    normals += 1
        o.close()
    print(f"qhull check: {checked} hull contacts, {normals} with a unique nearest facet")
    assert checked >= 30 and normals >= 15


def test_box_manifold_matches_independent_clipping():
    """The box-box contact MANIFOLD (SAT + Sutherland-Hodgman in the oracle, SAT + 24 candidate points in the CUDA path, which the
    GPU parity tests compare with the oracle point by point) against a computation that shares nothing with either:
      * normal and depth = nearest facet of the Minkowski difference of the two boxes' corners, from qhull;
      * contact polygon = reference-face rectangle intersected with the projected incident face, from
        scipy.spatial.HalfspaceIntersection (qhull's dual) -- its vertices that lie below the reference face, lifted half their
        depth along the normal, must be exactly the oracle's contact points, with the same per-point depths.
    Face-face manifolds only (the axis preference of 5 % / 2 um documented in DESIGN.md can pick a face where qhull's nearest
    facet is an edge-edge one; those cases are skipped, edge contacts are covered by test_sat_equals_epa_on_boxes)."""
    from scipy.spatial import ConvexHull, HalfspaceIntersection
    rng = np.random.default_rng(11)
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
    checked = multi = 0
    for trial in range(600):
        hA, hB = rng.uniform(0.01, 0.05, 3), rng.uniform(0.01, 0.05, 3)
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        RA = quat_to_mat(q)
        # B nearly face-aligned with A (small relative tilt): face-face manifolds with 3..8 points
        axes = np.eye(3)[rng.permutation(3)] * rng.choice([-1, 1], size=(3, 1))
        tilt = rng.normal(size=4) * (0.04 if trial % 2 else 0.002); tilt[0] = 1; tilt /= np.linalg.norm(tilt)
        RB = RA @ axes.T @ quat_to_mat(tilt)
        if np.linalg.det(RB) < 0:
            RB[:, 2] *= -1
        k = rng.integers(3)
        sgn = rng.choice([-1, 1])
        extB = np.abs(RB.T @ RA[:, k]) @ hB                  # half extent of B along A's axis k
        off = rng.uniform(-0.6, 0.6, 3) * hA
        off[k] = sgn * (hA[k] + extB - rng.uniform(2e-4, 3e-3))
        cA = rng.uniform(-0.1, 0.1, 3)
        cB = cA + RA @ off
        nrm, pos, dist = O.test_box_manifold(cA, RA.ravel(), hA, cB, RB.ravel(), hB)
        if len(pos) == 0:
            continue
        # ---- independent normal / depth: nearest facet of A (-) B
        VA, VB = cA + (corners * hA) @ RA.T, cB + (corners * hB) @ RB.T
        hull = ConvexHull((VA[:, None, :] - VB[None, :, :]).reshape(-1, 3))
        offs = hull.equations[:, 3]
        assert (offs < 1e-12).all()
        best = int(np.argmax(offs))
        depth, nq = -offs[best], hull.equations[best, :3]
        # face axis of A or B?  (otherwise qhull's nearest facet is an edge-edge one: skipped)
        face = [(0, j) for j in range(3) if abs(abs(nq @ RA[:, j]) - 1) < 1e-9] + [(1, j) for j in range(3) if abs(abs(nq @ RB[:, j]) - 1) < 1e-9]
        if not face or abs(abs(nq @ nrm) - 1) > 1e-9:
            continue
        assert abs(depth - (-dist.min())) < 1e-10, (trial, depth, dist)
        assert nrm @ (cB - cA) > 0                                        # geom1 -> geom2
        # ---- independent polygon: reference face rectangle (of the box that owns the normal) x projected incident face
        ref_is_A = face[0][0] == 0
        cR, RR, hR = (cA, RA, hA) if ref_is_A else (cB, RB, hB)
        cI, RI, hI = (cB, RB, hB) if ref_is_A else (cA, RA, hA)
        nref = nrm if ref_is_A else -nrm                                  # outward normal of the reference face
        ax = int(np.argmax(np.abs(RR.T @ nref)))
        u, v = RR[:, (ax + 1) % 3], RR[:, (ax + 2) % 3]
        hu, hv = hR[(ax + 1) % 3], hR[(ax + 2) % 3]
        iax = int(np.argmax(np.abs(RI.T @ nref)))
        isg = -np.sign(RI[:, iax] @ nref)
        fc = cI + isg * hI[iax] * RI[:, iax]
        eu, ev = RI[:, (iax + 1) % 3] * hI[(iax + 1) % 3], RI[:, (iax + 2) % 3] * hI[(iax + 2) % 3]
        quad3 = np.array([fc + eu + ev, fc - eu + ev, fc - eu - ev, fc + eu - ev])
        quad = np.stack([(quad3 - cR) @ u, (quad3 - cR) @ v], axis=1)
        hs = [[1, 0, -hu], [-1, 0, -hu], [0, 1, -hv], [0, -1, -hv]]
        ctr = quad.mean(axis=0)
        for a, b in zip(quad, np.roll(quad, -1, axis=0)):
            e = b - a
            nn = np.array([e[1], -e[0]]); nn /= np.linalg.norm(nn)
            if nn @ (ctr - a) > 0:
                nn = -nn
            hs.append([nn[0], nn[1], -(nn @ a)])
        hs = np.array(hs)
        # interior point: Chebyshev centre by a tiny LP
        from scipy.optimize import linprog
        res = linprog([0, 0, -1], A_ub=np.hstack([hs[:, :2], np.ones((len(hs), 1))]), b_ub=-hs[:, 2], bounds=[(None, None)] * 2 + [(0, None)])
        if res.status != 0 or res.x[2] < 1e-6:
            continue                                                      # sliver overlap: vertex set ill-conditioned
        poly = HalfspaceIntersection(hs, res.x[:2]).intersections
        # incident plane height above each polygon vertex, depth below the reference face
        ni = np.cross(quad3[1] - quad3[0], quad3[3] - quad3[0])
        pts, deps = [], []
        for pu, pv in poly:
            base = cR + pu * u + pv * v
            tpar = ((quad3[0] - base) @ ni) / (nref @ ni)                 # base + t nref on the incident plane
            d = hR[ax] - tpar
            if d > 1e-9:
                pts.append(base + (tpar + 0.5 * d) * nref); deps.append(d)
        pts, deps = np.array(pts), np.array(deps)
        # de-duplicate qhull's repeated vertices
        keep = []
        for i, p in enumerate(pts):
            if all(np.linalg.norm(p - pts[j]) > 1e-9 for j in keep):
                keep.append(i)
        pts, deps = pts[keep], deps[keep]
        near_zero = np.sum(np.abs(deps) < 1e-7) + np.sum(np.abs(dist) < 1e-7)
        if near_zero:
            continue                                                      # a vertex within 1e-7 m of the reference plane: either answer is right
        assert len(pts) == len(pos), (trial, len(pts), len(pos))
        for p, d in zip(pos, dist):
            j = int(np.argmin(np.linalg.norm(pts - p, axis=1)))
            assert np.linalg.norm(pts[j] - p) < 1e-9 and abs(deps[j] + d) < 1e-9, (trial, p, pts[j], d, deps[j])
        checked += 1
        multi += len(pos) >= 4
    print(f"box manifold check: {checked} face-face manifolds, {multi} with >= 4 points")
    assert checked >= 150 and multi >= 80


def test_oracle_converges_quickly_on_the_float32_two_cycle_states(model_blob):
    """The ten states of tests/golden/solver_two_cycle_states.npz made the float32 solver alternate between two points until the iteration
    cap (tests/test_gpu_parity.py checks that it no longer does).  In fp64 they are ordinary solves: a few iterations at MuJoCo's
    tolerances, and the tight checker tolerance reaches the same accelerations."""
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "solver_two_cycle_states.npz"))
    k = d["qpos"].shape[0]
    args = [d[n].astype(np.float64) for n in ("qpos", "qvel", "ctrl", "warm")]
    qacc = {}
    try:
        for mode in (1, 0):
            O.lib().so100o_set_solver_mode(mode)
            orc = O.Oracle(model_blob, k)
            orc.set_state(*args)
            orc.forward()
            its = np.array([orc.solver(i)["iters"] for i in range(k)])
            qacc[mode] = np.array([orc.dyn(i)["qacc"] for i in range(k)])
            if mode == 1:
                assert its.max() <= 15, its
            orc.close()
    finally:
        O.lib().so100o_set_solver_mode(0)
    assert np.abs(qacc[0] - qacc[1]).max() / (1 + np.abs(qacc[0]).max()) < 1e-5
