"""Independent anchor of the articulated dynamics: Isaac Gym / PhysX cannot be had, and both the CUDA kernel and
oracle/physics_ref.c are articulated-body algorithms written by the same hand.  This file checks them against a formulation that
shares NOTHING with ABA: the joint-space equations of motion assembled from body Jacobians,

    M(q) = sum_b J_b^T diag(I_b^world, m_b 1) J_b,      c(q, u) = sum_b J_b^T [ I_b alpha_b0 + w_b x I_b w_b ;  m_b (a_b0 - g) ],

(alpha_b0, a_b0: accelerations of body b with du/dt = 0, from the velocity recursion differentiated by hand) and the implicit
PD / armature terms on the diagonal exactly as DESIGN.md 3 states them:

    (M + diag(armature + h kd + h^2 kp)) du/dt = kp (target - q - h qd) - kd qd - c.

One substep of the integrator (contact-free, no damping, no clamp) must return velocities u+ = u + h du/dt of that linear system.
Generalised velocity u = [root angular velocity (world), root linear velocity of the body origin (world), relative angular velocity
of every spherical joint in the CHILD frame]; q of a joint = rotation vector of child-in-parent (humanoid_smpl.py:619-622).
CPU only (the GPU twin of the first test lives in tests/test_gpu_parity.py)."""
import numpy as np
import pytest


def rodrigues(v):
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.eye(3)
    k = v / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def quat_to_mat(q):   # xyzw
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class Tree:
    """plain-numpy rigid-body tree read from the packed model struct (the integrator's own inputs)"""

    def __init__(self, m):
        self.nb, self.nd = m.nb, m.nd
        self.parent = [m.parent[b] for b in range(m.nb)]
        self.offset = np.array([[m.offset[b][k] for k in range(3)] for b in range(m.nb)])
        self.mass = np.array([m.mass[b] for b in range(m.nb)])
        self.com = np.array([[m.com[b][k] for k in range(3)] for b in range(m.nb)])
        I = np.array([[m.inertia[b][k] for k in range(6)] for b in range(m.nb)])
        self.inertia = np.array([[[i[0], i[3], i[4]], [i[3], i[1], i[5]], [i[4], i[5], i[2]]] for i in I])
        self.dof = [m.dof_of_body[b] for b in range(m.nb)]
        self.kp = np.array([m.kp[k] for k in range(m.nd)])
        self.kd = np.array([m.kd[k] for k in range(m.nd)])
        self.arm = np.array([m.armature[k] for k in range(m.nd)])

    def kinematics(self, root_pos, root_quat, q):
        R, p = [None] * self.nb, [None] * self.nb
        R[0], p[0] = quat_to_mat(root_quat), np.array(root_pos, float)
        for b in range(1, self.nb):
            pa = self.parent[b]
            R[b] = R[pa] @ rodrigues(q[self.dof[b]:self.dof[b] + 3])
            p[b] = p[pa] + R[pa] @ self.offset[b]
        return R, p

    def velocities(self, R, p, u):
        """u [75] or [75, k] (columns = independent velocity vectors) -> w_b, v_b (body origin), per body"""
        u = u.reshape(6 + self.nd, -1)
        w, v = [None] * self.nb, [None] * self.nb
        w[0], v[0] = u[0:3], u[3:6]
        for b in range(1, self.nb):
            pa = self.parent[b]
            r = (p[b] - p[pa])[:, None]
            w[b] = w[pa] + R[b] @ u[6 + self.dof[b]:6 + self.dof[b] + 3]
            v[b] = v[pa] + np.cross(w[pa], r, axis=0)
        return w, v

    def mass_matrix_and_bias(self, root_pos, root_quat, q, u, gz):
        n = 6 + self.nd
        R, p = self.kinematics(root_pos, root_quat, q)
        wJ, vJ = self.velocities(R, p, np.eye(n))                 # Jacobians of origin velocities, column by column
        w, v = self.velocities(R, p, u)
        w, v = [x[:, 0] for x in w], [x[:, 0] for x in v]
        # accelerations with du/dt = 0: alpha_b = alpha_p + w_p x (R_b wt_b);  a_b = a_p + alpha_p x r + w_p x (w_p x r)
        al, a = [None] * self.nb, [None] * self.nb
        al[0], a[0] = np.zeros(3), np.zeros(3)
        for b in range(1, self.nb):
            pa = self.parent[b]
            r = p[b] - p[pa]
            wj = R[b] @ u[6 + self.dof[b]:6 + self.dof[b] + 3]
            al[b] = al[pa] + np.cross(w[pa], wj)
            a[b] = a[pa] + np.cross(al[pa], r) + np.cross(w[pa], np.cross(w[pa], r))
        M, c = np.zeros((n, n)), np.zeros(n)
        ke = pe = 0.0
        for b in range(self.nb):
            if self.mass[b] <= 0:
                continue
            cw = R[b] @ self.com[b]
            Iw = R[b] @ self.inertia[b] @ R[b].T
            Jw = wJ[b]
            Jc = vJ[b] + np.cross(wJ[b], cw[:, None], axis=0)     # velocity of the centre of mass
            M += Jw.T @ Iw @ Jw + self.mass[b] * Jc.T @ Jc
            ac = a[b] + np.cross(al[b], cw) + np.cross(w[b], np.cross(w[b], cw))
            c += Jw.T @ (Iw @ al[b] + np.cross(w[b], Iw @ w[b])) + Jc.T @ (self.mass[b] * (ac - np.array([0, 0, gz])))
            vc = v[b] + np.cross(w[b], cw)
            ke += 0.5 * self.mass[b] * vc @ vc + 0.5 * w[b] @ Iw @ w[b]
            pe += -self.mass[b] * gz * (p[b] + cw)[2]
        return M, c, ke, pe


def random_state(rng, n, nd):
    root = np.zeros((n, 13))
    root[:, 0:2] = rng.uniform(-2, 2, (n, 2))
    root[:, 2] = rng.uniform(4, 6, n)                      # far above the ground: contact-free
    qq = rng.normal(size=(n, 4))
    root[:, 3:7] = qq / np.linalg.norm(qq, axis=1, keepdims=True)
    root[:, 7:10] = rng.normal(0, 1.0, (n, 3))
    root[:, 10:13] = rng.normal(0, 2.0, (n, 3))
    q = rng.uniform(-0.25, 0.25, (n, nd))                  # inside every joint limit: no limit springs
    qd = rng.normal(0, 2.0, (n, nd))
    tar = q + rng.normal(0, 0.3, (n, nd))
    return root, q, qd, tar


@pytest.fixture(scope="module")
def setup():
    from vid2player3d_b200 import abi, model_compiler
    mod = model_compiler.load_compiled("smpl_mesh_humanoid_amass_v1")
    m, verts = abi.pack_model(mod, float(mod["mass"].sum()) / 90.0)
    return mod, m, verts


def implicit_reference(tree, root, q, qd, tar, h, gz):
    u = np.concatenate([root[10:13], root[7:10], qd])
    M, c, _, _ = tree.mass_matrix_and_bias(root[0:3], root[3:7], q, u, gz)
    n = 6 + tree.nd
    E = np.zeros(n)
    E[6:] = tree.arm + h * tree.kd + h * h * tree.kp
    tau = np.zeros(n)
    tau[6:] = tree.kp * (tar - q - h * qd) - tree.kd * qd
    ud = np.linalg.solve(M + np.diag(E), tau - c)
    return u, ud, M


def test_aba_restatement_matches_jacobian_equations_of_motion(setup):
    """256 random contact-free states incl. PD, armature: du/dt of one substep of oracle/physics_ref.c == (M + E)^-1 (tau - c)"""
    from vid2player3d_b200 import abi
    from oracle import physics_ref as P
    mod, m, verts = setup
    tree = Tree(m)
    assert abs(tree.mass.sum() - float(mod["mass"].sum())) < 1e-4
    cfg = abi.make_cfg(mod, sim_dt=1.0 / 120.0, substeps=1, control_freq_inv=1, ang_damping=0.0, max_ang_vel=1e6)
    h = float(cfg.sim_dt)            # the struct holds float32: the integrator's h is that value, not 1/120 in double
    rng = np.random.default_rng(0)
    n = 256
    root, q, qd, tar = random_state(rng, n, 69)
    r1, q1, qd1 = root.copy(), q.copy(), qd.copy()
    P.control_step(m, verts, cfg, r1, q1, qd1, tar.copy())
    worst = 0.0
    for i in range(n):
        u, ud, M = implicit_reference(tree, root[i], q[i], qd[i], tar[i], h, cfg.gravity_z)
        u1 = np.concatenate([r1[i, 10:13], r1[i, 7:10], qd1[i]])
        ud_aba = (u1 - u) / h
        worst = max(worst, np.abs(ud_aba - ud).max() / (1.0 + np.abs(ud).max()))
        assert np.all(np.linalg.eigvalsh(M) > 0)
    assert worst < 1e-10, worst      # cond(M + E) ~ 1e3: two float64 formulations of the same linear system


def test_mass_matrix_consistency(setup):
    """the Jacobian mass matrix reproduces the kinetic energy and the total momentum the restatement reports (diagnostics)"""
    from vid2player3d_b200 import abi
    from oracle import physics_ref as P
    mod, m, verts = setup
    tree = Tree(m)
    cfg = abi.make_cfg(mod)
    rng = np.random.default_rng(1)
    root, q, qd, _ = random_state(rng, 8, 69)
    for i in range(8):
        u = np.concatenate([root[i, 10:13], root[i, 7:10], qd[i]])
        M, _, ke, pe = tree.mass_matrix_and_bias(root[i, 0:3], root[i, 3:7], q[i], u, cfg.gravity_z)
        d = P.diagnostics(m, cfg, root[i], q[i], qd[i])
        assert abs(0.5 * u @ M @ u - ke) < 1e-9 * (1 + ke)
        assert abs(d["ke"] - ke) < 1e-8 * (1 + ke) and abs(d["pe"] - pe) < 1e-8 * (1 + abs(pe))


def test_energy_drift_is_first_order(setup):
    """gravity only (no PD, no armature, no damping), free flight: the total energy of the semi-implicit Euler integrator drifts
    O(h) - halving the substep roughly halves the drift over the same simulated time"""
    from vid2player3d_b200 import abi
    from oracle import physics_ref as P
    mod, _, _ = setup
    md = dict(mod)
    md["armature"] = np.zeros(69)
    m0, verts = abi.pack_model(md, 0.0)
    tree = Tree(m0)
    drifts = []
    for substeps in (4, 8, 16):
        cfg = abi.make_cfg(mod, ang_damping=0.0, max_ang_vel=1e6, substeps=substeps, control_freq_inv=1)
        rng = np.random.default_rng(5)
        root, q, qd, tar = random_state(rng, 1, 69)
        qd *= 0.5
        u = np.concatenate([root[0, 10:13], root[0, 7:10], qd[0]])
        _, _, ke0, pe0 = tree.mass_matrix_and_bias(root[0, :3], root[0, 3:7], q[0], u, cfg.gravity_z)
        P.control_step(m0, verts, cfg, root, q, qd, tar, n_steps=12)          # 0.2 s of flight
        u = np.concatenate([root[0, 10:13], root[0, 7:10], qd[0]])
        _, _, ke1, pe1 = tree.mass_matrix_and_bias(root[0, :3], root[0, 3:7], q[0], u, cfg.gravity_z)
        drifts.append(abs((ke1 + pe1) - (ke0 + pe0)) / (ke0 + abs(pe0)))
    assert drifts[0] < 5e-2
    assert 0.3 < drifts[1] / drifts[0] < 0.7 and 0.3 < drifts[2] / drifts[1] < 0.7, drifts


def test_ground_contact_against_analytic_coulomb_friction(setup):
    """the compliant-implicit ground contact against textbook physics: a humanoid held rigid by stiff PD (a ragdoll locked in its pose),
    lying on the ground and pushed along +x at 2 m/s, (a) carries exactly its weight, (b) decelerates at mu g = 9.81 m/s^2 while it
    slides (regularised Coulomb friction, mu = 1 from the yaml), (c) comes to rest and stays there (stick)."""
    from vid2player3d_b200 import abi
    from oracle import physics_ref as P
    mod, _, _ = setup
    m, verts = abi.pack_model(mod, 50.0)                      # very stiff joints: the body behaves as one rigid piece
    cfg = abi.make_cfg(mod, substeps=8)
    g = 9.81
    root = np.zeros((1, 13))
    root[0, 2], root[0, 3:7] = 0.3, [0.0, 0.0, 0.0, 1.0]      # lying (the asset is y-up in its body frame): drop from 30 cm and settle
    q, qd, tar = np.zeros((1, 69)), np.zeros((1, 69)), np.zeros((1, 69))
    P.control_step(m, verts, cfg, root, q, qd, tar, n_steps=90)
    assert abs(root[0, 9]) < 2e-2 and np.abs(root[0, 7:9]).max() < 2e-2, "did not settle"
    rb, cf = P.control_step(m, verts, cfg, root, q, qd, tar, n_steps=1)
    weight = float(mod["mass"].sum()) * g
    assert abs(cf[0, :, 2].sum() - weight) < 0.02 * weight
    root[0, 7] = 2.0                                          # push
    vs = [2.0]
    for _ in range(12):                                       # 12 control steps = 0.4 s; mu g dt = 0.327 m/s per control step
        P.control_step(m, verts, cfg, root, q, qd, tar, n_steps=1)
        vs.append(root[0, 7])
    vs = np.array(vs)
    sliding = vs[1:] > 0.4                                    # well above the regularisation speed v_s = 0.05 m/s
    dec = -(vs[1:] - vs[:-1]) / (2 * (1 / 60.0))
    assert sliding.sum() >= 3
    assert np.all(np.abs(dec[sliding] - g) < 0.15 * g), dec
    assert abs(vs[-1]) < 0.05, vs                             # stuck


def test_ball_bounce_and_drag_against_closed_forms():
    """the ball model against closed forms: (a) a vertical drop rebounds with the restitution the scene's materials give (PhysX 'average'
    of the ball 0.9 and the plane 0.5 = 0.7, humanoid_smpl_im_mvae.py:414-438), (b) with no spin the horizontal deceleration in flight is
    the quadratic drag k_f C_d |v| v / m of apply_external_force_to_ball (:711-739; C_d 0.55, rho 1.204 kg/m^3, r 0.032 m, m 0.057 kg)"""
    from vid2player3d_b200 import abi, model_compiler
    from oracle import physics_ref as P
    mod = model_compiler.canonical_racket_last(model_compiler.load_compiled("smpl_mesh_humanoid_federer"))
    m, verts = abi.pack_model(mod, float(mod["mass"].sum()) / 90.0)
    cfg = abi.make_cfg(mod, substeps=6, task_mode=1, pd_mode=1, contact_bodies=(), key_bodies=(), enable_early_termination=False,
                       ball=dict(spin_scale=5.0, ball_e_ground=0.7, ball_mu_ground=0.6, ball_e_racket=0.9, ball_mu_racket=0.5))
    root = np.zeros((1, 13)); root[0, :3] = [50.0, 50.0, 0.95]; root[0, 3:7] = [0.5, 0.5, 0.5, 0.5]     # the player far away from the ball
    q, qd, tar = np.zeros((1, 69)), np.zeros((1, 69)), np.zeros((1, 69))
    ball = np.zeros((1, 13)); ball[0, :3] = [0.0, 0.0, 1.0]; ball[0, 6] = 1.0
    vz = []
    for _ in range(40):
        P.control_step(m, verts, cfg, root, q, qd, tar, ball=ball, hits=np.zeros(1, np.int32))
        vz.append(ball[0, 9])
    vz = np.array(vz)
    i = int(np.argmax(vz))                       # first step after the bounce: largest upward speed
    v_in = -vz[i - 1]
    assert v_in > 3.0 and abs(vz[i] / v_in - 0.7) < 0.06, (v_in, vz[i])
    ball[:] = 0; ball[0, :3] = [0.0, 0.0, 30.0]; ball[0, 6] = 1.0; ball[0, 7] = 25.0      # fast, high, no spin
    v0 = ball[0, 7]
    P.control_step(m, verts, cfg, root, q, qd, tar, ball=ball, hits=np.zeros(1, np.int32))
    dt = 1.0 / 30.0
    kf = 1.204 * np.pi * 0.032 ** 2 / 2
    a_drag = kf * 0.55 * v0 * v0 / 0.057
    assert abs((v0 - ball[0, 7]) / dt - a_drag) < 0.1 * a_drag, ((v0 - ball[0, 7]) / dt, a_drag)


def test_exact_sphere_hull_query_against_surface_sampling():
    """oracle/physics_ref.c::hull_sphere_ref (planes for separation / containment, closest point over the hull triangles otherwise)
    against an independent estimate: the distance to 400 random points on every triangle of the hull (an upper bound that converges
    to the true distance) and scipy's point-in-hull test."""
    from scipy.spatial import Delaunay
    from vid2player3d_b200 import abi, model_compiler
    from oracle import physics_ref as P
    mod = model_compiler.load_compiled("smpl_mesh_humanoid_amass_v1")
    m, verts = abi.pack_model(mod, 1.0)
    planes, tris, ntris, tmax = abi.pack_faces(mod, verts)
    assert ntris[:24].min() >= 50 and tmax == 128
    P.set_hull_faces(planes, tris, ntris, tmax)
    rng = np.random.default_rng(0)
    try:
        for b in (0, 3, 4, 13, 16, 22):
            nv, k = int(mod["nverts"][b]), int(ntris[b])
            V = verts[b, :nv].astype(np.float64)
            T = V[tris[b, :k, :3].astype(int)]                              # [k, 3, 3]
            w = rng.dirichlet([1, 1, 1], size=(k, 400))                     # barycentric samples
            S = np.einsum("ksj,kjd->ksd", w, T).reshape(-1, 3)
            S = np.concatenate([S, V])
            tri = Delaunay(V)
            ctr, ext = V.mean(0), np.ptp(V, axis=0).max()
            R = 0.032
            hits = 0
            for _ in range(300):
                c = ctr + rng.normal(size=3) * ext * 0.45
                hit, pen, nl = P.hull_sphere(m, verts, b, c, R)
                d_s = np.linalg.norm(S - c, axis=1).min()
                inside = tri.find_simplex(c) >= 0
                if inside:
                    assert hit == 1 and pen >= R - 1e-6 and abs(np.linalg.norm(nl) - 1) < 1e-5
                    hits += 1
                    continue
                if hit == 1:
                    d = R - pen
                    assert -1e-6 <= d < R and d <= d_s + 1e-6 and d_s - d < 6e-3, (d, d_s)
                    assert abs(np.linalg.norm(nl) - 1) < 1e-6
                    hits += 1
                else:
                    assert d_s >= R - 1e-6, d_s                             # no contact -> no surface sample inside the sphere
            assert hits > 10
    finally:
        P.clear_hull_faces()
