"""mc3d-b200: the B200-native hot path of sashapersonxyz/Multi-camera_3D_Pose_Estimation.

Python mirror of the reference's call surface for that path (``utils``, ``pose_estimation``,
``mmpose_pose_estimation``, ``pose_refinement``) over libmc3d.so (include/mc3d.h, csrc/*.cu,
hand-written sm_100a CUDA).  The directory name is the project's; import it as ``mc3d_b200``
(the shim ``mc3d_b200.py`` at the repository root maps that name onto this directory).

No CPU fallback exists: every compute entry point raises ``Mc3dError`` when the library is not built
or no B200 is visible.
"""
from ._lib import Mc3dError, launch_count, lib  # noqa: F401

__version__ = '0.1.0'
