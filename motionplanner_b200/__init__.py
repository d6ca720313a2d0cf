"""motionplanner_b200 -- B200-native batched motion-primitive collision check.

A drop-in for the one data-parallel hot path of KochdumperNiklas/MotionPlanner (reference files
``auxiliary/collisionChecker.py``, ``collision_check`` in ``src/lowLevelPlannerManeuverAutomaton.py`` and
``src/maneuverAutomatonPlannerStandalone.py``).  Host Python here is a thin mirror of the reference's call surface;
all geometry runs in hand-written sm_100a kernels behind the C ABI of ``include/mpc.h`` (``csrc/libmpcheck.so``).
There is no CPU fallback: importing :mod:`motionplanner_b200.engine` without the built library raises.
"""
__version__ = "0.1.0"
