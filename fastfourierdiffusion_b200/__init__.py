"""fastfourierdiffusion_b200 -- B200-native reverse-SDE sampling path of FastFourierDiffusion.

Host-side mirror of the reference's ``fdiff`` surface for the sampling path
(sampler, score model, SDE schedulers, E2-CRF cache, dft/idft) over a C-ABI
CUDA library (``include/ffd_b200.h``, ``csrc/``).  See DESIGN.md / INTEGRATION.md.
"""

from .batch import DiffusableBatch
from .caching import E2CRFCache
from .fourier import dft, idft, spectral_density
from .fresca import apply_fresca_to_score, frequency_scale
from .metrics import MarginalWasserstein, MetricCollection, SlicedWasserstein, WassersteinDistances
from .sampler import DiffusionSampler
from .score_model import ScoreModule
from .sde import SDE, SamplingOutput, VEScheduler, VPScheduler

__all__ = ["DiffusableBatch", "E2CRFCache", "dft", "idft", "apply_fresca_to_score", "frequency_scale", "DiffusionSampler", "ScoreModule", "SDE",
           "SamplingOutput", "VEScheduler", "VPScheduler", "SlicedWasserstein", "MarginalWasserstein", "WassersteinDistances", "MetricCollection", "spectral_density"]
