"""``fdiff`` import surface of the reference, served by ``fastfourierdiffusion_b200``.

The reference's entry points and Hydra configs name their classes by module path
(``cmd/conf/sampler/default.yaml:1`` ``_target_: fdiff.sampling.sampler.DiffusionSampler``,
``cmd/conf/score_model/default.yaml:1``, ``noise_scheduler/{vpsde,vesde}.yaml:1``,
``cmd/conf/metrics/default.yaml:1,6,11``; ``cmd/{sample,benchmark_cache,ablation_cache}.py`` import
``fdiff.models.score_models``, ``fdiff.sampling.sampler``, ``fdiff.utils.caching``, ``fdiff.schedulers.sde``,
``fdiff.utils.fourier``).  Every module here re-exports the B200 product class of the same name, so those
strings resolve unchanged when this repository's root is on ``sys.path`` instead of the reference's ``src/``.
Only the sampling path is served: training losses, Lightning callbacks, plotting and the MLP / LSTM
score models raise ``NotImplementedError`` / are absent.
"""

from fastfourierdiffusion_b200 import __all__ as _all  # noqa: F401
