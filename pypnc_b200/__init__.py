"""pypnc_b200: batched whole-body control (robot_system update -> IHWBC QP) on B200."""
__version__ = "0.1.0"
