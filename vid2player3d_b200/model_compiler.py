"""Model compiler: MJCF (+ binary STL meshes, + primitive cylinders) -> flat constant block.

Replaces what `gym.load_asset` / `create_actor` / `set_actor_dof_properties` do for the
reference (embodied_pose/env/tasks/humanoid_smpl_im.py:273-300,356-389 and
vid2player/env/tasks/humanoid_smpl_im_mvae.py:367-442): it turns the humanoid asset into
the per-asset constants the CUDA step kernel stages into shared memory:

  parent / depth / children tables, joint offsets, mass, COM, inertia (from the convex
  hull of each body mesh at the geom density), armature, PD gains (joint `stiffness` /
  `damping`, later scaled by mass/90 * kp_scale like humanoid_smpl_im.py:376-383), joint
  ranges, and the hull vertices used for ground contact.

Every body with three co-located hinges becomes ONE spherical joint whose coordinates are
the exp-map of the child-in-parent rotation (SURVEY.md §7 "hard parts").  Bodies without a
joint (the `Racket` body of the vid2player assets) are welded to their parent: their mass
is folded into the parent link, they keep their own rigid-body state row.

The compiled block is committed under vid2player3d_b200/assets/compiled/*.npz (derived
data, like the golden fixtures) because /root/reference does not exist on the GPU box.
"""
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

VMAX = 64  # hull vertices per body are padded to this many


def _read_stl(path):
    with open(path, "rb") as f:
        data = f.read()
    n = struct.unpack_from("<I", data, 80)[0]
    if 84 + 50 * n == len(data):
        rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                            count=n, offset=84)
        return rec["v"].reshape(-1, 3).astype(np.float64)
    # ascii fallback
    verts = []
    for line in data.decode("ascii", "ignore").splitlines():
        t = line.split()
        if len(t) == 4 and t[0] == "vertex":
            verts.append([float(x) for x in t[1:]])
    return np.asarray(verts, dtype=np.float64)


def _hull_mass_props(points, density):
    """Mass, COM and inertia about the COM of the convex hull of `points` (uniform density)."""
    from scipy.spatial import ConvexHull
    hull = ConvexHull(points)
    hv = points[hull.vertices]
    centre = hv.mean(axis=0)
    vol = 0.0
    first = np.zeros(3)
    second = np.zeros((3, 3))  # integral of x x^T dV
    canon = (np.ones((3, 3)) + np.eye(3)) / 120.0
    for simplex, eq in zip(hull.simplices, hull.equations):
        a, b, c = points[simplex]
        if np.dot(np.cross(b - a, c - a), eq[:3]) < 0:
            b, c = c, b
        # tetrahedron (centre, a, b, c)
        A = np.stack([a - centre, b - centre, c - centre], axis=1)
        det = np.linalg.det(A)
        v = det / 6.0
        vol += v
        first += v * (a + b + c + centre) / 4.0
        # second moment about `centre`, then shifted to origin
        C = det * A @ canon @ A.T
        tc = (a + b + c - 3 * centre) / 4.0 * v  # first moment about centre
        second += C + np.outer(centre, tc) + np.outer(tc, centre) + v * np.outer(centre, centre)
    mass = density * vol
    com = first / vol
    S = density * second
    S_c = S - mass * np.outer(com, com)
    inertia = np.trace(S_c) * np.eye(3) - S_c
    return mass, com, inertia, hv


def _cylinder_props(p0, p1, radius, density):
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    axis = p1 - p0
    L = np.linalg.norm(axis)
    e = axis / L
    mass = density * np.pi * radius ** 2 * L
    com = 0.5 * (p0 + p1)
    i_ax = 0.5 * mass * radius ** 2
    i_tr = mass * (3 * radius ** 2 + L ** 2) / 12.0
    inertia = i_tr * np.eye(3) + (i_ax - i_tr) * np.outer(e, e)
    return mass, com, inertia


def _combine(parts):
    """parts: list of (mass, com, inertia_about_com) in one frame -> combined."""
    m = sum(p[0] for p in parts)
    com = sum(p[0] * p[1] for p in parts) / m
    I = np.zeros((3, 3))
    for pm, pc, pI in parts:
        d = pc - com
        I += pI + pm * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    return m, com, I


def compile_mjcf(xml_path):
    tree = ET.parse(xml_path)
    root = tree.getroot()
    xml_dir = os.path.dirname(os.path.abspath(xml_path))
    comp = root.find("compiler")
    angle_deg = (comp is None) or comp.get("angle", "degree") == "degree"
    dj = root.find("default/joint")
    d_arm = float(dj.get("armature", 0.0)) if dj is not None else 0.0
    meshes = {m.get("name"): os.path.join(xml_dir, m.get("file")) for m in root.findall("asset/mesh")}

    bodies = []

    def visit(node, parent):
        idx = len(bodies)
        pos = np.array([float(x) for x in node.get("pos", "0 0 0").split()])
        quat = np.array([float(x) for x in node.get("quat", "1 0 0 0").split()])
        assert np.allclose(quat, [1, 0, 0, 0]), "non-identity body quat not supported"
        hinges = [j for j in node.findall("joint") if j.get("type", "hinge") == "hinge"]
        free = node.find("freejoint") is not None
        b = dict(name=node.get("name"), parent=parent, pos=pos, free=free, fixed=(not free and len(hinges) == 0))
        if hinges:
            assert len(hinges) == 3, "expected 3 co-located hinges (-> one spherical joint)"
            axes = np.array([[float(x) for x in h.get("axis").split()] for h in hinges])
            assert np.allclose(axes, np.eye(3)), "hinge axes must be x,y,z"
            b["kp"] = [float(h.get("stiffness", 0.0)) for h in hinges]
            b["kd"] = [float(h.get("damping", 0.0)) for h in hinges]
            b["arm"] = [float(h.get("armature", d_arm)) for h in hinges]
            rng = np.array([[float(x) for x in h.get("range", "-180 180").split()] for h in hinges])
            b["range"] = np.deg2rad(rng) if angle_deg else rng
            b["dof_names"] = [h.get("name") for h in hinges]
        parts, verts, prims = [], [], []
        for g in node.findall("geom"):
            dens = float(g.get("density", 1000.0))
            gtype = g.get("type", "sphere")
            if gtype == "mesh":
                pts = _read_stl(meshes[g.get("mesh")])
                m, c, I, hv = _hull_mass_props(np.unique(pts, axis=0), dens)
                parts.append((m, c, I))
                verts.append(hv)
            elif gtype == "cylinder":
                ft = [float(x) for x in g.get("fromto").split()]
                r = float(g.get("size").split()[0])
                parts.append(_cylinder_props(ft[:3], ft[3:], r, dens))
                prims.append(ft + [r])
            else:
                raise NotImplementedError(gtype)
        b["mass"], b["com"], b["inertia"] = _combine(parts)
        b["verts"] = np.concatenate(verts, 0) if verts else np.zeros((0, 3))
        b["prims"] = np.array(prims).reshape(-1, 7)
        bodies.append(b)
        for ch in node.findall("body"):
            visit(ch, idx)

    wb = root.find("worldbody")
    tops = wb.findall("body")
    assert len(tops) == 1
    visit(tops[0], -1)
    return _pack(bodies, os.path.basename(xml_path))


TMAX = 128  # hull triangles per body are padded to this many (2 V - 4 <= 124 for V <= 64)


def hull_faces(verts, nverts):
    """Triangulated faces of every body's convex hull (the shapes PhysX collides the ball with): tris [nb, TMAX, 3] uint8 vertex indices,
    counter-clockwise seen from outside (outward normal = (v1 - v0) x (v2 - v0)), ntris [nb], and tri_lmax [nb] = the longest triangle edge
    (a point within distance r of the hull is within r + tri_lmax of one of its vertices: the cheap vertex pass filters the exact one)."""
    from scipy.spatial import ConvexHull
    nb = len(nverts)
    tris = np.zeros((nb, TMAX, 3), np.uint8)
    ntris = np.zeros(nb, np.int32)
    lmax = np.zeros(nb)
    for b in range(nb):
        nv = int(nverts[b])
        if nv < 4:
            continue
        V = np.asarray(verts[b, :nv], np.float64)
        hull = ConvexHull(V)                       # Qhull with triangulated output: 2 V - 4 facets for points in general position
        assert len(hull.simplices) <= TMAX
        for k, (simp, eq) in enumerate(zip(hull.simplices, hull.equations)):
            i, j, l = (int(x) for x in simp)
            if np.cross(V[j] - V[i], V[l] - V[i]) @ eq[:3] < 0:
                j, l = l, j
            tris[b, k] = (i, j, l)
            lmax[b] = max(lmax[b], np.linalg.norm(V[j] - V[i]), np.linalg.norm(V[l] - V[j]), np.linalg.norm(V[i] - V[l]))
        ntris[b] = len(hull.simplices)
    return tris, ntris, lmax


def _pack(bodies, name):
    nb = len(bodies)
    parent = np.array([b["parent"] for b in bodies], np.int32)
    fixed = np.array([b["fixed"] for b in bodies], np.int32)
    assert bodies[0]["free"] and not fixed[0]
    # the reference relies on body order == depth-first MJCF order and dof order body-major (x,y,z)
    jointed = [i for i, b in enumerate(bodies) if "kp" in b]
    dof_body_ids = np.array(jointed, np.int32)
    nd = 3 * len(jointed)
    dof_of_body = -np.ones(nb, np.int32)
    for j, i in enumerate(jointed):
        dof_of_body[i] = 3 * j
    depth = np.zeros(nb, np.int32)
    for i in range(1, nb):
        depth[i] = depth[parent[i]] + 1
    offset = np.stack([b["pos"] for b in bodies]).astype(np.float64)
    mass = np.array([b["mass"] for b in bodies])
    com = np.stack([b["com"] for b in bodies])
    inertia = np.stack([b["inertia"] for b in bodies])
    # dynamics view: welded bodies are folded into their parent link
    dmass, dcom, dinertia = mass.copy(), com.copy(), inertia.copy()
    for i in range(nb - 1, 0, -1):
        if fixed[i]:
            p = parent[i]
            m, c, I = _combine([(dmass[p], dcom[p], dinertia[p]), (dmass[i], dcom[i] + offset[i], dinertia[i])])
            dmass[p], dcom[p], dinertia[p] = m, c, I
            dmass[i] = 0.0
    kp = np.zeros(nd)
    kd = np.zeros(nd)
    arm = np.zeros(nd)
    lim = np.zeros((nd, 2))
    dof_names = []
    for j, i in enumerate(jointed):
        kp[3 * j:3 * j + 3] = bodies[i]["kp"]
        kd[3 * j:3 * j + 3] = bodies[i]["kd"]
        arm[3 * j:3 * j + 3] = bodies[i]["arm"]
        lim[3 * j:3 * j + 3] = bodies[i]["range"]
        dof_names += bodies[i]["dof_names"]
    nverts = np.array([len(b["verts"]) for b in bodies], np.int32)
    assert nverts.max() <= VMAX
    verts = np.zeros((nb, VMAX, 3))
    radius = np.zeros(nb)
    for i, b in enumerate(bodies):
        if nverts[i]:
            verts[i, :nverts[i]] = b["verts"]
            verts[i, nverts[i]:] = b["verts"][0]
            radius[i] = np.linalg.norm(b["verts"], axis=1).max()
    prims = [np.concatenate([np.full((len(b["prims"]), 1), i), b["prims"]], 1) for i, b in enumerate(bodies) if len(b["prims"])]
    prims = np.concatenate(prims, 0) if prims else np.zeros((0, 8))
    tris, ntris, tri_lmax = hull_faces(verts, nverts)
    return dict(
        tris=tris, ntris=ntris, tri_lmax=tri_lmax,
        name=name, body_names=np.array([b["name"] for b in bodies]), dof_names=np.array(dof_names),
        parent=parent, depth=depth, fixed=fixed, dof_of_body=dof_of_body, dof_body_ids=dof_body_ids,
        offset=offset, mass=mass, com=com, inertia=inertia,
        dyn_mass=dmass, dyn_com=dcom, dyn_inertia=dinertia,
        kp=kp, kd=kd, armature=arm, limits=lim,
        nverts=nverts, verts=verts, radius=radius, prims=prims,
    )


def save(model, path):
    np.savez_compressed(path, **model)


def load(path):
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


COMPILED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "compiled")


def load_compiled(name):
    """name: e.g. 'smpl_mesh_humanoid_amass_v1'"""
    return load(os.path.join(COMPILED_DIR, name + ".npz"))


def canonical_racket_last(model):
    """Left-handed assets (nadal: Racket welded to L_Wrist, body 19) in the right-handed body order [humanoid 0..23, Racket 24].
    This is the order the reference's consumers see through `_humanoid_body_ids_lefthand`
    (vid2player/env/tasks/humanoid_smpl_im_mvae.py:67,197-206); here the simulator itself runs in it, so no permuted copy of the
    rigid-body tensors is needed.  DOF order is untouched (the Racket has no DOF).  No-op when the Racket is already last."""
    names = [str(x) for x in model["body_names"]]
    if "Racket" not in names or names[-1] == "Racket":
        return model
    nb, r = len(names), names.index("Racket")
    perm = [i for i in range(nb) if i != r] + [r]          # new index -> old index
    inv = {old: new for new, old in enumerate(perm)}
    out = dict(model)
    for k, v in model.items():
        v = np.asarray(v)
        if k in ("parent", "dof_body_ids"):
            continue
        if v.ndim >= 1 and v.shape[0] == nb and k not in ("kp", "kd", "armature", "limits", "dof_names", "prims"):
            out[k] = v[perm]
    out["parent"] = np.array([inv[int(p)] if p >= 0 else -1 for p in np.asarray(model["parent"])[perm]], np.int32)
    out["dof_body_ids"] = np.array([inv[int(b)] for b in model["dof_body_ids"]], np.int32)
    if "prims" in model and len(model["prims"]):
        pr = np.array(model["prims"], np.float64)
        pr[:, 0] = [inv[int(b)] for b in pr[:, 0]]
        out["prims"] = pr
    assert all(out["parent"][i] < i for i in range(nb))     # parents still precede their children
    return out


if __name__ == "__main__":
    import sys
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    srcs = [
        "embodied_pose/data/assets/mjcf/smpl_mesh_humanoid_amass_v1.xml",
        "vid2player/data/assets/smpl_mesh_humanoid_federer.xml",
        "vid2player/data/assets/smpl_mesh_humanoid_djokovic.xml",
        "vid2player/data/assets/smpl_mesh_humanoid_nadal.xml",
    ]
    os.makedirs(COMPILED_DIR, exist_ok=True)
    for s in srcs:
        m = compile_mjcf(os.path.join(ref, s))
        out = os.path.join(COMPILED_DIR, os.path.splitext(os.path.basename(s))[0] + ".npz")
        save(m, out)
        print(f"{s}: bodies={len(m['parent'])} dof={len(m['kp'])} mass={m['mass'].sum():.3f} "
              f"max verts={m['nverts'].max()} -> {out}")
