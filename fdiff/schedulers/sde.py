"""``fdiff.schedulers.sde`` (reference: src/fdiff/schedulers/sde.py:14-246).  A reference checkpoint pickles its
scheduler under this module path; the attribute names are the reference's, so it unpickles into these classes."""

from fastfourierdiffusion_b200.sde import SDE, SamplingOutput, VEScheduler, VPScheduler

__all__ = ["SDE", "SamplingOutput", "VEScheduler", "VPScheduler"]
