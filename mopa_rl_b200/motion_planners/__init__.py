"""Drop-in counterparts of the reference's ``motion_planners`` package
(``planner.PyKinematicPlanner``, ``sampling_based_planner.SamplingBasedPlanner``) backed by
libmopa_b200.so."""
from .planner import PyKinematicPlanner  # noqa: F401
from .sampling_based_planner import SamplingBasedPlanner, joint_convert  # noqa: F401
