"""``fdiff.dataloaders.datamodules`` -- the datamodule surface the SAMPLING path touches.

``cmd/sample.py:66-70,107-109`` instantiates the training run's datamodule only to read three things: the tensor
shape ``(max_len, n_channels)`` (reference: src/fdiff/dataloaders/datamodules.py:194-203 ECG 187x1, :349-353
MIMIC-III 24x40, :406-410 NASDAQ 252x5, :469-476 NASA 251x4 / 134x5, :530-538 USDroughts 365x13), the
``fourier_transform`` / ``standardize`` flags and ``feature_mean_and_std`` (:153-161).  Downloading, cleaning and
batching the datasets (Kaggle / PhysioNet access) is training-side data engineering and is out of scope
(SURVEY 8): ``setup`` reads the tensors the reference's own preprocessing cached under ``data_dir`` when they
exist (``X_train.pt`` / ``mitbih_train.csv``) and otherwise creates SYNTHETIC N(0, 1) data of the named shape
(``self.synthetic = True``), which is what the benchmarks of this repository run on.
"""

from __future__ import annotations

from pathlib import Path
from typing import Any, Optional, Tuple

import torch

from fastfourierdiffusion_b200.cli import DATASET_SHAPES


class Datamodule:
    dataset_name = "synthetic"
    shape_key = "synthetic"
    n_synthetic = 1024

    def __init__(self, data_dir: Path | str = Path.cwd() / "data", random_seed: int = 42, batch_size: int = 32,
                 fourier_transform: bool = False, standardize: bool = False, **_ignored: Any) -> None:
        self.data_dir = Path(data_dir) / self.dataset_name
        self.random_seed = random_seed
        self.batch_size = batch_size
        self.fourier_transform = fourier_transform
        self.standardize = standardize
        self.X_train = torch.Tensor()
        self.y_train: Optional[torch.Tensor] = None
        self.X_test = torch.Tensor()
        self.y_test: Optional[torch.Tensor] = None
        self.synthetic = False

    def prepare_data(self) -> None:  # datamodules.py:91-95 downloads; nothing is fetched here
        return None

    def _cached_dir(self) -> Path:
        return self.data_dir

    def setup(self, stage: str = "fit") -> None:
        d = self._cached_dir()
        if (d / "X_train.pt").exists():
            self.X_train = torch.load(d / "X_train.pt", map_location="cpu").float()
            if (d / "X_test.pt").exists():
                self.X_test = torch.load(d / "X_test.pt", map_location="cpu").float()
        else:
            max_len, n_channels = DATASET_SHAPES[self.shape_key]
            g = torch.Generator().manual_seed(self.random_seed)
            self.X_train = torch.randn(self.n_synthetic, max_len, n_channels, generator=g)
            self.X_test = torch.randn(self.n_synthetic // 4, max_len, n_channels, generator=g)
            self.synthetic = True
        assert tuple(self.X_train.shape[1:]) == DATASET_SHAPES[self.shape_key], self.X_train.shape

    @property
    def dataset_parameters(self) -> dict:  # datamodules.py:145-151
        return {"n_channels": self.X_train.size(2), "max_len": self.X_train.size(1),
                "num_training_steps": (len(self.X_train) + self.batch_size - 1) // self.batch_size}

    @property
    def feature_mean_and_std(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """datamodules.py:40-53,153-161: per-(position, channel) statistics of the training set, in the frequency
        domain when ``fourier_transform``."""
        X = self.X_train
        if self.fourier_transform:
            from fastfourierdiffusion_b200.fourier import dft

            X = dft(X).detach().cpu()
        return X.mean(dim=0), X.std(dim=0)


class ECGDatamodule(Datamodule):
    dataset_name = "ecg"
    shape_key = "ecg"

    def setup(self, stage: str = "fit") -> None:
        csv = self.data_dir / "mitbih_train.csv"
        if csv.exists():  # datamodules.py:187-203
            import pandas as pd

            df = pd.read_csv(csv)
            self.X_train = torch.tensor(df.iloc[:, :187].values, dtype=torch.float32).unsqueeze(2)
            self.y_train = torch.tensor(df.iloc[:, 187].values, dtype=torch.long)
            return
        super().setup(stage)


class SyntheticDatamodule(Datamodule):
    """datamodules.py:244-303: sine waves with random phase and Beta(2, 2) frequency -- the one dataset of the
    reference that is generated rather than downloaded, so it is generated here too (same process; the reference
    draws from the unseeded global numpy generator, here the generator is seeded with ``random_seed``)."""

    dataset_name = "synthetic"
    shape_key = "synthetic"

    def __init__(self, *args: Any, max_len: int = 100, num_samples: int = 1000, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.max_len = max_len
        self.num_samples = num_samples

    def setup(self, stage: str = "fit") -> None:
        import numpy as np

        rng = np.random.RandomState(self.random_seed)
        n = 2 * self.num_samples
        phase = rng.normal(size=n).reshape(-1, 1)
        frequency = rng.beta(a=2, b=2, size=n).reshape(-1, 1)
        X = np.sin(np.arange(self.max_len) * frequency + phase)
        self.X_train = torch.tensor(X[: self.num_samples], dtype=torch.float32).unsqueeze(2)
        self.X_test = torch.tensor(X[self.num_samples:], dtype=torch.float32).unsqueeze(2)
        self.synthetic = True


class MIMICIIIDatamodule(Datamodule):
    dataset_name = "mimiciii"
    shape_key = "mimiciii"


class NASDAQDatamodule(Datamodule):
    dataset_name = "nasdaq"
    shape_key = "nasdaq"


class NASADatamodule(Datamodule):
    dataset_name = "nasa"

    def __init__(self, *args: Any, subdataset: str = "charge", remove_outlier_feature: bool = True,
                 **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.subdataset = subdataset
        self.remove_outlier_feature = remove_outlier_feature
        self.shape_key = "nasa_" + subdataset

    def _cached_dir(self) -> Path:
        return self.data_dir / self.subdataset


class USDroughtsDatamodule(Datamodule):
    dataset_name = "usdroughts"
    shape_key = "droughts"
