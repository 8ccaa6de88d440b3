"""The reference's ``fdiff`` import surface and run-directory format, served by this repository.

The ``_target_`` strings and ``from fdiff...`` imports below are the ones the reference's configs and entry points
contain (cmd/conf/sampler/default.yaml:1, cmd/conf/score_model/default.yaml:1, noise_scheduler/{vpsde,vesde}.yaml:1,
cmd/conf/metrics/default.yaml:1,6,11, cmd/sample.py:10-15, cmd/benchmark_cache.py:20-25).  The run directory
``tests/golden/ref_run/tinyve01`` was written by ``oracle/make_golden_ckpt.py`` from the UNMODIFIED reference classes.
"""

import os

import numpy as np
import pytest
import torch

from ffd_testutil import GOLDEN, rel

RUN_ROOT = os.path.join(GOLDEN, "ref_run")


def test_reference_import_lines_resolve():
    from fdiff.dataloaders.datamodules import Datamodule, ECGDatamodule, NASADatamodule  # noqa: F401
    from fdiff.models.score_models import ScoreModule
    from fdiff.sampling.metrics import MetricCollection  # noqa: F401
    from fdiff.sampling.sampler import DiffusionSampler
    from fdiff.schedulers.sde import VEScheduler, VPScheduler
    from fdiff.utils.caching import E2CRFCache
    from fdiff.utils.dataclasses import DiffusableBatch  # noqa: F401
    from fdiff.utils.extraction import dict_to_str, get_best_checkpoint, get_model_type
    from fdiff.utils.fourier import dft, idft  # noqa: F401
    from fdiff.utils.fresca import apply_fresca_to_score  # noqa: F401

    import fastfourierdiffusion_b200 as F

    assert ScoreModule is F.ScoreModule and DiffusionSampler is F.DiffusionSampler and E2CRFCache is F.E2CRFCache
    assert VPScheduler is F.VPScheduler and VEScheduler is F.VEScheduler
    assert get_model_type({"score_model": {"_target_": "fdiff.models.score_models.ScoreModule"}}) is ScoreModule
    with pytest.raises(NotImplementedError):
        get_model_type({"score_model": {"_target_": "fdiff.models.score_models.Nope"}})
    best = get_best_checkpoint(os.path.join(RUN_ROOT, "tinyve01", "checkpoints"))
    assert str(best).endswith("epoch=1-val_loss=0.50.ckpt")
    assert dict_to_str({"a": 1}) == "a: 1\n"
    dm = NASADatamodule(fourier_transform=True, standardize=True, subdataset="charge")
    dm.setup()
    assert dm.synthetic and dm.dataset_parameters["max_len"] == 251 and dm.dataset_parameters["n_channels"] == 4
    assert ECGDatamodule().shape_key == "ecg"


def test_hydra_targets_instantiate():
    """The reference's YAML nodes, verbatim, through the ``_target_`` strings."""
    from fastfourierdiffusion_b200.cli import instantiate, resolve_config

    import fastfourierdiffusion_b200 as F

    cfg = resolve_config({
        "fourier_transform": True,
        "sampler": {"_target_": "fdiff.sampling.sampler.DiffusionSampler", "_partial_": True, "sample_batch_size": 50},
        "score_model": {"_target_": "fdiff.models.score_models.ScoreModule", "_partial_": True, "d_model": 72,
                        "num_layers": 10, "n_head": 12, "lr_max": 1.0e-3,
                        "fourier_noise_scaling": "${fourier_transform}", "likelihood_weighting": False,
                        "noise_scheduler": {"_target_": "fdiff.schedulers.sde.VPScheduler", "eps": 1e-5,
                                            "beta_min": 0.1, "beta_max": 20,
                                            "fourier_noise_scaling": "${score_model.fourier_noise_scaling}"}},
        "metrics": {"_target_": "fdiff.sampling.metrics.MetricCollection", "_partial_": True,
                    "metrics": [{"_target_": "fdiff.sampling.metrics.SlicedWasserstein", "_partial_": True,
                                 "random_seed": 42, "num_directions": 10, "save_all_distances": False}]},
    })
    assert cfg["score_model"]["noise_scheduler"]["fourier_noise_scaling"] is True
    model_partial = instantiate(cfg["score_model"])
    torch.manual_seed(0)
    model = model_partial(n_channels=1, max_len=187, num_training_steps=10)
    assert isinstance(model, F.ScoreModule) and isinstance(model.noise_scheduler, F.VPScheduler)
    assert model.scale_noise is True and model.noise_scheduler.noise_scaling is True and model.dims.d_ff == 2048
    sampler = instantiate(cfg["sampler"])(score_model=model)
    assert isinstance(sampler, F.DiffusionSampler) and sampler.sample_batch_size == 50
    assert callable(instantiate(cfg["metrics"]))


def test_run_directory_and_checkpoint_loading(tmp_path):
    from fastfourierdiffusion_b200.cli import SAMPLE_DEFAULTS, find_best_checkpoint, load_run, parse_overrides

    import fastfourierdiffusion_b200 as F

    cfg = parse_overrides(["num_samples=7", "use_cache=true", "cache_kwargs={'K': 3, 'R': 20}"], SAMPLE_DEFAULTS)
    assert cfg["num_samples"] == 7 and cfg["cache_kwargs"] == {"K": 3, "R": 20}
    run = load_run(RUN_ROOT, "tinyve01")
    assert run["train_cfg"]["datamodule"]["fourier_transform"] is True
    assert run["train_cfg"]["score_model"]["noise_scheduler"]["fourier_noise_scaling"] is True
    assert run["checkpoint"].endswith("epoch=1-val_loss=0.50.ckpt")
    # hyper-parameters win over any default: VE scheduler with sigma_max = 2, d_model 12, 2 layers, 2 heads
    model = F.ScoreModule.load_from_checkpoint(run["checkpoint"])
    assert isinstance(model.noise_scheduler, F.VEScheduler) and model.noise_scheduler.sigma_max == 2.0
    d = model.dims
    assert (d.max_len, d.n_channels, d.d_model, d.n_head, d.n_layers) == (16, 1, 12, 2, 2)
    assert model.scale_noise is True and model.noise_scheduler.G is not None
    # a model id that has no run directory / no config / no checkpoint raises like the reference (cmd/sample.py:38-59)
    with pytest.raises(FileNotFoundError):
        load_run(RUN_ROOT, "missing")
    (tmp_path / "abc" / "checkpoints").mkdir(parents=True)
    with pytest.raises(FileNotFoundError):
        load_run(str(tmp_path), "abc")
    (tmp_path / "abc" / "train_config.yaml").write_text("fourier_transform: false\n")
    with pytest.raises(FileNotFoundError):
        load_run(str(tmp_path), "abc")
    assert find_best_checkpoint(str(tmp_path / "nope")) is None


@pytest.mark.gpu
def test_checkpoint_scores_match_reference():
    """Scores of the model loaded from the reference-format checkpoint equal what the reference computed from it."""
    import fastfourierdiffusion_b200 as F

    run = os.path.join(RUN_ROOT, "tinyve01")
    e = np.load(os.path.join(run, "expected.npz"))
    model = F.ScoreModule.load_from_checkpoint(os.path.join(run, "checkpoints", "epoch=1-val_loss=0.50.ckpt"),
                                               precision="fp32").to("cuda").eval()
    b = F.DiffusableBatch(X=torch.from_numpy(e["X"]).cuda(), timesteps=torch.from_numpy(e["t"]).cuda())
    score = model(b).cpu().numpy()
    err = rel(score, e["score"])
    print(f"checkpoint scores vs reference: rel {err:.2e}")
    assert err < 2e-5


@pytest.mark.gpu
def test_sample_entry_point_with_trained_run(tmp_path):
    """cmd/sample.py with a model id: architecture / SDE / domain / standardisation from the run directory."""
    import importlib.util

    from ffd_testutil import ROOT

    spec = importlib.util.spec_from_file_location("cmd_sample", os.path.join(ROOT, "cmd", "sample.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = tmp_path / "out"
    mod.main([f"model_path={RUN_ROOT}", "model_id=tinyve01", "num_samples=4", "num_diffusion_steps=6",
              "sampler.sample_batch_size=4", "precision=fp32", f"output_dir={out}"])
    X = torch.load(out / "samples.pt")
    assert X.shape == (4, 16, 1) and torch.isfinite(X).all()
    with pytest.raises(FileNotFoundError):
        mod.main([f"model_path={RUN_ROOT}", "model_id=absent", f"output_dir={out}"])
