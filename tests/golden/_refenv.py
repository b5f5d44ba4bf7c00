"""Import harness for the reference's own PyTorch functions.

Only usable where /root/reference exists (the build container).  It is used by
tests/golden/make_golden.py to GENERATE the committed fixtures; nothing under
`-m gpu`, smoke() or bench.py imports this file.
"""
import os
import sys
from unittest.mock import MagicMock

REF = os.environ.get("V2P_REFERENCE", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def available():
    return os.path.isdir(os.path.join(REF, "embodied_pose"))


def setup(which="embodied_pose"):
    """Put the isaacgym shim + one reference sub-project on sys.path; mock absent deps."""
    shim = os.path.join(REPO, "vid2player3d_b200", "shim")
    for p in (os.path.join(REF, which), REF, os.path.join(REF, "poselib"), shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    for m in ("imageio", "gym", "gym.spaces", "mujoco_py", "lxml", "lxml.etree", "stl", "stl.mesh",
              "uhc.smpllib.smpl_local_robot", "smpl_visualizer", "smpl_visualizer.smpl", "pyvista",
              "vtk", "cv2", "tensorboardX", "rl_games"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
