"""Import harness for the reference's own PyTorch functions.

Only usable where /root/reference exists (the build container).  It is used by
tests/golden/make_golden*.py to GENERATE the committed fixtures; nothing under
`-m gpu`, smoke() or bench.py imports this file.
"""
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

REF = os.environ.get("V2P_REFERENCE", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# absent third-party packages of the reference: any `import pkg.sub.mod` resolves to a MagicMock
MOCK_TOP = {"imageio", "gym", "mujoco_py", "lxml", "stl", "smpl_visualizer", "pyvista", "vtk", "vtkmodules", "cv2",
            "tensorboardX", "rl_games", "scenepic", "horovod", "wandb", "glfw", "OpenGL", "mujoco", "chumpy", "smplx", "pyrender", "trimesh", "open3d", "ffmpeg", "mediapy"}


class _MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in MOCK_TOP:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__path__ = []
        m.__spec__ = spec
        m.__name__ = spec.name
        return m

    def exec_module(self, module):
        return


def available():
    return os.path.isdir(os.path.join(REF, "embodied_pose"))


def setup(which="embodied_pose"):
    """Put the isaacgym shim + one reference sub-project on sys.path; mock absent deps."""
    shim = os.path.join(REPO, "vid2player3d_b200", "shim")
    for p in (os.path.join(REF, which), REF, os.path.join(REF, "poselib"), shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    real = set()
    for top in list(MOCK_TOP):
        try:
            __import__(top)
            real.add(top)
        except Exception:
            pass
    MOCK_TOP.difference_update(real)
    if not any(isinstance(f, _MockFinder) for f in sys.meta_path):
        sys.meta_path.append(_MockFinder())
    if "uhc.smpllib.smpl_local_robot" not in sys.modules:
        sys.modules["uhc.smpllib.smpl_local_robot"] = MagicMock()
