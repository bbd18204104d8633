"""Shared helpers of the pose tests: the seeded synthetic 3-D <-> 2-D problems live in alvaar_b200.synth."""
from alvaar_b200.synth import make_pose_problem, quat_to_R  # noqa: F401
