"""``fdiff.models.score_models`` (reference: src/fdiff/models/score_models.py:24-288).

``ScoreModule`` is the transformer score model of the sampling path.  ``MLPScoreModule`` / ``LSTMScoreModule``
(score_models.py:363-520) are other model families, outside the hot path: naming them fails loudly."""

from fastfourierdiffusion_b200.score_model import ScoreModule


def _out_of_scope(name):
    class _Absent:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"fdiff.models.score_models.{name} is outside the B200 sampling path "
                                      "(only the transformer ScoreModule is built)")

        @classmethod
        def load_from_checkpoint(cls, *a, **k):
            cls()

    _Absent.__name__ = name
    return _Absent


MLPScoreModule = _out_of_scope("MLPScoreModule")
LSTMScoreModule = _out_of_scope("LSTMScoreModule")

__all__ = ["ScoreModule", "MLPScoreModule", "LSTMScoreModule"]
