"""Reference MoCap buffer ("MotionLib") as flat SoA arrays.

Same per-frame tensors the reference keeps after loading
(embodied_pose/utils/motion_lib.py:68-93): gts[F,B,3] grs[F,B,4] lrs[F,B,4] grvs[F,3]
gravs[F,3] dvs[F,D] concatenated over motions, plus the per-motion scalars
(_motion_lengths, _motion_num_frames, _motion_dt, length_starts, _motion_min_verts_h,
_motion_bodies).  The CUDA sampler (csrc/b200env.cu: motion_state) gathers two frames per
env from these arrays and blends them exactly like MotionLib.get_motion_state (:164-266).

The real AMASS / tennis motion files are not shipped with the reference (git-ignored,
SURVEY.md §0.5), so `synthetic()` fabricates a smooth random-walk library on the shipped
skeleton for tests and benchmarks.  `from_reference()` accepts a loaded reference MotionLib
object when one is available.
"""
import os

import numpy as np

BASE_ROT = np.array([0.5, 0.5, 0.5, 0.5])  # SMPL y-up -> z-up base rotation (humanoid_smpl_im.py:766-770)


def _qmul(a, b):
    x1, y1, z1, w1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    x2, y2, z2, w2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                     w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                     w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], axis=-1)


def _qconj(a):
    return np.concatenate([-a[..., :3], a[..., 3:]], -1)


def _qrot(q, v):
    qv, qw = q[..., :3], q[..., 3:4]
    t = 2.0 * np.cross(qv, v)
    return v + qw * t + np.cross(qv, t)


def _qexp(v):
    ang = np.linalg.norm(v, axis=-1, keepdims=True)
    half = 0.5 * ang
    s = np.where(ang > 1e-12, np.sin(half) / np.maximum(ang, 1e-12), 0.5)
    return np.concatenate([v * s, np.cos(half)], -1)


def _qlog(q):
    q = np.where(q[..., 3:4] < 0, -q, q)
    sn = np.linalg.norm(q[..., :3], axis=-1, keepdims=True)
    ang = 2.0 * np.arctan2(sn, q[..., 3:4])
    return q[..., :3] * np.where(sn > 1e-12, ang / np.maximum(sn, 1e-12), 2.0)


class FlatMotionLib:
    FIELDS = ("gts", "grs", "lrs", "grvs", "gravs", "dvs", "motion_lengths", "num_frames", "motion_dt",
              "length_starts", "min_verts_h", "motion_bodies")

    def __init__(self, **kw):
        for k in self.FIELDS:
            setattr(self, k, kw[k])
        self.gts = np.ascontiguousarray(self.gts, np.float32)
        self.grs = np.ascontiguousarray(self.grs, np.float32)
        self.lrs = np.ascontiguousarray(self.lrs, np.float32)
        self.grvs = np.ascontiguousarray(self.grvs, np.float32)
        self.gravs = np.ascontiguousarray(self.gravs, np.float32)
        self.dvs = np.ascontiguousarray(self.dvs, np.float32)
        self.motion_lengths = np.ascontiguousarray(self.motion_lengths, np.float32)
        self.num_frames = np.ascontiguousarray(self.num_frames, np.int64)
        self.motion_dt = np.ascontiguousarray(self.motion_dt, np.float32)
        self.length_starts = np.ascontiguousarray(self.length_starts, np.int64)
        self.min_verts_h = np.ascontiguousarray(self.min_verts_h, np.float32)
        self.motion_bodies = np.ascontiguousarray(self.motion_bodies, np.float32)

    def num_motions(self):
        return len(self.motion_lengths)

    def as_dict(self, key_body_ids, dof_body_ids):
        d = {k: getattr(self, k) for k in self.FIELDS}
        d["key_body_ids"] = np.asarray(key_body_ids, np.int64)
        d["dof_body_ids"] = np.asarray(dof_body_ids, np.int64)
        return d

    def save(self, path):
        np.savez_compressed(path, **{k: getattr(self, k) for k in self.FIELDS})

    @classmethod
    def load(cls, path):
        z = np.load(path)
        return cls(**{k: z[k] for k in cls.FIELDS})

    # ---- flat on-disk format: one file, mmap-able, arrays 64-byte aligned in the order the sampler reads them.
    #   bytes 0-7   magic "B200ML01"      bytes 8-15  little-endian uint64 = length of the JSON header that follows
    #   header      {"fields": {name: {"dtype": "<f4"|"<i8", "shape": [...], "offset": bytes from file start}}, "num_motions": M}
    #   payload     the arrays of FIELDS, C order.  Replaces `torch.save(motion_lib)` pickles of Python objects
    #   (embodied_pose/utils/motion_lib.py:68-93 keeps the same tensors after `_load_motions`).
    MAGIC = b"B200ML01"

    def save_flat(self, path):
        import json
        arrays = {k: np.ascontiguousarray(getattr(self, k)) for k in self.FIELDS}
        base = 4096                                   # header region: magic + length + JSON, space padded; payload starts here
        while True:
            fields, off = {}, base
            for k, a in arrays.items():
                fields[k] = {"dtype": a.dtype.str, "shape": list(a.shape), "offset": off}
                off += (a.nbytes + 63) // 64 * 64
            hdr = json.dumps({"fields": fields, "num_motions": int(self.num_motions())}).encode()
            if 16 + len(hdr) <= base:
                break
            base *= 2
        with open(path, "wb") as fh:
            fh.write(self.MAGIC)
            fh.write(np.uint64(len(hdr)).tobytes())
            fh.write(hdr)
            for k, a in arrays.items():
                fh.seek(fields[k]["offset"])
                fh.write(a.tobytes())
            fh.truncate(off)

    @classmethod
    def load_flat(cls, path, mmap=True):
        """mmap=True: the arrays are read-only views of the file (no copy until the upload to the device)."""
        import json
        with open(path, "rb") as fh:
            if fh.read(8) != cls.MAGIC:
                raise ValueError(f"{path}: not a B200ML01 motion library")
            n = int(np.frombuffer(fh.read(8), np.uint64)[0])
            hdr = json.loads(fh.read(n).decode())
        missing = [k for k in cls.FIELDS if k not in hdr["fields"]]
        if missing:
            raise ValueError(f"{path}: missing fields {missing}")
        kw = {}
        for k in cls.FIELDS:
            f = hdr["fields"][k]
            if mmap:
                kw[k] = np.memmap(path, dtype=np.dtype(f["dtype"]), mode="r", offset=f["offset"], shape=tuple(f["shape"]))
            else:
                kw[k] = np.fromfile(path, dtype=np.dtype(f["dtype"]), count=int(np.prod(f["shape"])), offset=f["offset"]).reshape(f["shape"])
        return cls(**kw)

    @classmethod
    def load_any(cls, path, motion_file_range=None):
        """What the reference's `motion_file` may name (embodied_pose/env/tasks/humanoid_smpl_im.py:420-440) plus our own formats:
        a `.b200ml` flat file, an `.npz` archive (save()), a reference `torch.save(motion_lib)` pickle (`.pth` / `.pt`; needs the
        reference's `utils.motion_lib` and `poselib` importable, like the reference's own `torch.load` of it), or a DIRECTORY of such
        files: sorted, sliced by `motion_file_range` = [first, last) and merged like `merge_multiple_motion_libs` (:101-118)."""
        path = os.fspath(path)
        if os.path.isdir(path):
            files = []
            for pat in ("*.pth", "*.pt", "*.b200ml", "*.npz"):
                files = sorted(f for f in (os.path.join(path, n) for n in os.listdir(path)) if f.endswith(pat[1:]))
                if files:
                    break                         # the reference globs *.pth; a directory of our own formats works the same way
            if motion_file_range is not None:
                files = files[motion_file_range[0]:motion_file_range[1]]
            if not files:
                raise FileNotFoundError(f"{path}: no motion library files (*.pth, *.pt, *.b200ml, *.npz) in the directory / range")
            return cls.merge([cls.load_any(f) for f in files])
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        ext = os.path.splitext(path)[1].lower()
        if ext == ".npz":
            return cls.load(path)
        if ext in (".pth", ".pt"):
            import torch
            return cls.from_reference(torch.load(path, map_location="cpu", weights_only=False))
        with open(path, "rb") as fh:
            if fh.read(8) == cls.MAGIC:
                return cls.load_flat(path)
        raise ValueError(f"{path}: unknown motion library format (expected .b200ml, .npz, or a reference .pth / .pt pickle)")

    @classmethod
    def merge(cls, libs):
        """Concatenation of motion libraries (the reference's `merge_multiple_motion_libs`, motion_lib.py:101-118): per-frame arrays and
        per-motion tables are concatenated, the frame offsets of every motion are rebuilt (`generate_length_starts`)."""
        libs = list(libs)
        if len(libs) == 1:
            return libs[0]
        if len({(l.gts.shape[1], l.dvs.shape[1], l.motion_bodies.shape[1]) for l in libs}) != 1:
            raise ValueError("motion libraries with different skeletons / shape widths cannot be merged")
        kw = {k: np.concatenate([np.asarray(getattr(l, k)) for l in libs], 0) for k in cls.FIELDS if k != "length_starts"}
        nf = kw["num_frames"].astype(np.int64)
        kw["length_starts"] = np.concatenate([[0], np.cumsum(nf)[:-1]]).astype(np.int64)
        return cls(**kw)

    @classmethod
    def from_reference(cls, ml):
        """ml: a loaded reference MotionLib (embodied_pose/utils/motion_lib.py)."""
        g = lambda t: t.detach().cpu().numpy()  # noqa: E731
        return cls(gts=g(ml.gts), grs=g(ml.grs), lrs=g(ml.lrs), grvs=g(ml.grvs), gravs=g(ml.gravs), dvs=g(ml.dvs),
                   motion_lengths=g(ml._motion_lengths), num_frames=g(ml._motion_num_frames),
                   motion_dt=g(ml._motion_dt), length_starts=g(ml.length_starts),
                   min_verts_h=g(ml._motion_min_verts_h), motion_bodies=g(ml._motion_bodies))


def synthetic(model, num_motions=64, num_frames=300, fps=30.0, seed=7, sigma=0.05, root_height=0.9,
              shape_dim=11, ragged=False):
    """Smooth random-walk motions on the model's skeleton.

    local rotations: per-joint random walk of the rotation vector (sigma rad / frame, mean
    reverting), root: heading random walk + slow planar drift at `root_height`.
    `ragged=True` draws a different frame count per motion (tests the length tables).
    """
    rng = np.random.default_rng(seed)
    parent = model["parent"]
    B = len(parent)
    jointed = model["dof_body_ids"]
    offset = model["offset"]
    dt = 1.0 / fps
    chunks = {k: [] for k in ("gts", "grs", "lrs", "grvs", "gravs", "dvs")}
    nfs = []
    for m in range(num_motions):
        F = int(rng.integers(num_frames // 2, num_frames + 1)) if ragged else num_frames
        nfs.append(F)
        rv = np.zeros((F, B, 3))
        steps = rng.normal(0.0, sigma, size=(F, B, 3))
        for f in range(1, F):
            rv[f] = 0.98 * rv[f - 1] + steps[f]
        lrs = _qexp(rv)
        heading = np.cumsum(rng.normal(0.0, 0.02, size=F)) + rng.uniform(-np.pi, np.pi)
        hq = _qexp(np.stack([np.zeros(F), np.zeros(F), heading], -1))
        tilt = _qexp(rv[:, 0] * 0.3)
        lrs[:, 0] = _qmul(hq, _qmul(np.broadcast_to(BASE_ROT, (F, 4)), tilt))
        for b in range(B):
            if model["fixed"][b]:
                lrs[:, b] = np.array([0.0, 0.0, 0.0, 1.0])
        vel = np.cumsum(rng.normal(0.0, 0.05, size=(F, 2)), 0) * 0.2
        root_xy = np.cumsum(vel * dt, 0) + rng.uniform(-2, 2, size=2)
        root_p = np.concatenate([root_xy, np.full((F, 1), root_height) + 0.03 * np.sin(np.arange(F)[:, None] * 0.2)], -1)
        grs = np.zeros((F, B, 4))
        gts = np.zeros((F, B, 3))
        grs[:, 0] = lrs[:, 0]
        gts[:, 0] = root_p
        for b in range(1, B):
            p = parent[b]
            grs[:, b] = _qmul(grs[:, p], lrs[:, b])
            gts[:, b] = gts[:, p] + _qrot(grs[:, p], np.broadcast_to(offset[b], (F, 3)))
        grs /= np.linalg.norm(grs, axis=-1, keepdims=True)
        grvs = np.zeros((F, 3))
        grvs[:-1] = (root_p[1:] - root_p[:-1]) / dt
        grvs[-1] = grvs[-2]
        gravs = np.zeros((F, 3))
        dq = _qmul(grs[1:, 0], _qconj(grs[:-1, 0]))
        gravs[:-1] = _qlog(dq) / dt
        gravs[-1] = gravs[-2]
        # dof velocities: axis*angle of q0^-1 q1 per joint / dt  (motion_lib.py:490-518)
        dl = _qlog(_qmul(_qconj(lrs[:-1]), lrs[1:])) / dt
        dvs = np.zeros((F, len(jointed) * 3))
        dvs[:-1] = dl[:, jointed].reshape(F - 1, -1)
        dvs[-1] = dvs[-2]
        for k, v in (("gts", gts), ("grs", grs), ("lrs", lrs), ("grvs", grvs), ("gravs", gravs), ("dvs", dvs)):
            chunks[k].append(v)
    nfs = np.array(nfs, np.int64)
    starts = np.concatenate([[0], np.cumsum(nfs)[:-1]])
    bodies = np.zeros((num_motions, shape_dim), np.float32)
    return FlatMotionLib(
        **{k: np.concatenate(v, 0) for k, v in chunks.items()},
        motion_lengths=(nfs - 1) * np.float32(dt), num_frames=nfs, motion_dt=np.full(num_motions, dt, np.float32),
        length_starts=starts, min_verts_h=np.zeros(num_motions, np.float32), motion_bodies=bodies)
