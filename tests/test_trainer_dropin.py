"""Row f1: the UNMODIFIED reference trainer `medimgen/train_ldm.py` driving the B200 modules.

The trainer source is imported from /root/reference (build container) or from the byte-for-byte copies staged into the
git-ignored baseline/_ref/ (GPU box; oracle/stage_reference.py). Apart from scaffolding for packages that are absent
from this image and irrelevant to the step (matplotlib, torchinfo, the zarr/batchgenerators data pipeline), the only
thing taken from this repository is `compat.install()`."""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    for cand in (os.environ.get("MEDIMGEN_REFERENCE", "/root/reference"), os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(cand, "medimgen", "train_ldm.py")):
            return cand
    return None


@pytest.fixture()
def trainer_env():
    """sys.path / sys.modules for importing the unmodified trainer; fully restored afterwards."""
    ref = _reference_root()
    if ref is None:
        pytest.skip("reference trainer sources not present (run python -m oracle.stage_reference in the build container)")
    import importlib.util
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    extra = [ref, os.path.join(ROOT, "oracle", "shim")]               # medimgen + the 4-symbol MONAI shim
    for name in ("matplotlib", "torchinfo"):                           # absent from this image, never exercised
        if importlib.util.find_spec(name) is None:
            extra.append(os.path.join(ROOT, "tests", "stubs"))
            break
    sys.path[:0] = [p for p in extra if p not in sys.path]
    # the data pipeline (zarr / blosc2 / batchgenerators) is replaced by a synthetic loader in the test
    dp = types.ModuleType("medimgen.data_processing")
    dp.get_data_loaders = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("synthetic loaders only in this test"))
    sys.modules["medimgen.data_processing"] = dp
    import medical_image_generation_b200.compat as compat
    patched = compat.install()
    try:
        yield patched
    finally:
        compat.uninstall()
        sys.path[:] = saved_path
        # drop what this fixture made importable (and only that: torch lazily imports sub-packages that register
        # operator libraries once per process -- deleting and re-importing those fails)
        scoped = ("medimgen", "generative", "monai", "matplotlib", "torchinfo")
        for k in list(sys.modules):
            if k not in saved_mods and k.split(".")[0] in scoped:
                del sys.modules[k]
        for k, v in saved_mods.items():
            if k.split(".")[0] in scoped:
                sys.modules[k] = v


def test_unmodified_trainer_imports_and_binds_b200_classes(trainer_env):
    """CPU: `from medimgen.train_ldm import LDM` works and the names it bound at import time are the B200 classes."""
    import medical_image_generation_b200 as mig
    tl = __import__("medimgen.train_ldm", fromlist=["LDM"])
    assert tl.__file__.endswith(os.path.join("medimgen", "train_ldm.py")) and "medical_image_generation_b200" not in tl.__file__
    assert tl.DiffusionModelUNet is mig.DiffusionModelUNet and tl.AutoencoderKL is mig.AutoencoderKL
    assert tl.DDPMScheduler is mig.DDPMScheduler and tl.LatentDiffusionInferer is mig.LatentDiffusionInferer
    with pytest.raises(NotImplementedError):
        tl.VQVAE()                                 # placeholder of the shipped minimal `generative` (or the real class)


def _config(tmp_path, ae_ckpt, steps_T=12):
    down = [[[1, 1, 1], [3, 3, 3], [1, 1, 1]], [[2, 2, 2], [3, 3, 3], [1, 1, 1]]]
    vae = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=3, num_res_blocks=1,
               with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False, use_flash_attention=False,
               use_checkpointing=False, use_convtranspose=False, num_channels=[32, 64], attention_levels=[False, False],
               norm_num_groups=16, downsample_parameters=down, upsample_parameters=list(reversed(down))[:-1])
    lat = [[[1, 1, 1], [3, 3, 3], [1, 1, 1]], [[2, 2, 2], [3, 3, 3], [1, 1, 1]], [[2, 2, 2], [3, 3, 3], [1, 1, 1]]]
    ddpm = dict(spatial_dims=3, in_channels=3, out_channels=3, num_res_blocks=1, use_flash_attention=False,
                num_channels=[64, 128, 128], attention_levels=[False, True, True], num_head_channels=[0, 64, 128],
                strides=[q[0] for q in lat], kernel_sizes=[q[1] for q in lat], paddings=[q[2] for q in lat])
    return dict(vae_params=vae, ddpm_params=ddpm, load_autoencoder_path=str(ae_ckpt), load_model_path=None,
                time_scheduler_params=dict(num_train_timesteps=steps_T, schedule="scaled_linear_beta", beta_start=0.0015,
                                           beta_end=0.0205, prediction_type="epsilon"),
                ddpm_learning_rate=1e-4, lr_scheduler=None, lr_scheduler_params=None, grad_accumulate_step=2,
                grad_clip_max_norm=1, output_mode="log", progress_bar=False, results_path=str(tmp_path / "ldm"),
                n_epochs=1, val_plot_interval=10, ddpm_batch_size=2, ae_transformations=dict(patch_size=[16, 16, 16]))


@pytest.mark.gpu
def test_unmodified_train_ldm_runs_on_b200_modules(trainer_env, tmp_path):
    """LDM.__init__ (checkpoint load), scale-factor probe, train_one_epoch under fp16 autocast + GradScaler with gradient
    accumulation + clip + torch AdamW, validate_epoch, save_model / load_model (resume order: optimiser first) and
    sample_images -- all through the reference's own code (train_ldm.py:41-556)."""
    from torch.cuda.amp import GradScaler
    tl = __import__("medimgen.train_ldm", fromlist=["LDM"])
    torch.manual_seed(0)
    cfg = _config(tmp_path, tmp_path / "ae.pth")
    ae = tl.AutoencoderKL(**cfg["vae_params"])
    torch.save({"epoch": 3, "network_state_dict": ae.state_dict()}, cfg["load_autoencoder_path"])
    ldm = tl.LDM(cfg, latent_space_type="vae")
    # zero-initialised convs (unet:649-659,1934-1944) make the fresh U-Net output exactly 0: re-randomise so the loss moves
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for _, p in sorted(ldm.ddpm.named_parameters()):
            if float(p.abs().max()) == 0.0:
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))
    gen = torch.Generator().manual_seed(2)
    loader = [{"id": i, "image": torch.rand(2, 1, 16, 16, 16, generator=gen)} for i in range(5)]
    inferer, z_shape = ldm.get_inferer_and_latent_shape(loader)
    assert z_shape == (2, 3, 8, 8, 8) and float(inferer.scale_factor) > 0
    optimizer, lr_sched = ldm.get_optimizer_and_lr_schedule()
    assert isinstance(optimizer, torch.optim.AdamW) and lr_sched is None
    before = {k: v.detach().clone() for k, v in ldm.ddpm.state_dict().items()}
    scaler = GradScaler()
    for epoch in (1, 2, 3):
        ldm.train_one_epoch(epoch, loader, optimizer, scaler, inferer)      # 5 batches, accumulate 2 -> 3 optimiser steps
    ldm.validate_epoch(loader[:2], inferer)
    losses = ldm.loss_dict["rec_loss"]
    assert len(losses) == 3 and all(l == l and 0 < l < 5 for l in losses), losses
    assert losses[-1] < losses[0], losses
    after = ldm.ddpm.state_dict()
    moved = [k for k in before if "proj_attn" not in k and not torch.equal(before[k], after[k])]
    assert len(moved) >= 0.9 * len([k for k in before if "proj_attn" not in k])
    assert all(torch.equal(before[k], after[k]) for k in before if "proj_attn" in k)    # never used, never updated
    assert float(scaler.get_scale()) > 0
    # checkpoint round trip in the reference's format and resume order (train_ldm.py:466-505, 522-525)
    ldm.save_model(3, ldm.loss_dict["val_rec_loss"][-1], optimizer)
    ckpt_path = os.path.join(cfg["results_path"], "checkpoints", "last_model.pth")
    ckpt = torch.load(ckpt_path)
    assert set(ckpt) == {"epoch", "network_state_dict", "optimizer_state_dict", "validation_loss"}
    ldm2 = tl.LDM(cfg, latent_space_type="vae")
    opt2, _ = ldm2.get_optimizer_and_lr_schedule()
    assert ldm2.load_model(ckpt_path, optimizer=opt2, for_training=True) == 4
    x = torch.randn(1, 3, 8, 8, 8, device="cuda")
    t = torch.tensor([5], device="cuda")
    with torch.no_grad():
        a, b = ldm.ddpm.eval()(x, t), ldm2.ddpm.eval()(x, t)
    # same weights -> same output up to the bf16 noise of atomically reduced GroupNorm statistics (random weights: ~1)
    assert float((a - b).norm() / b.norm()) < 2e-2
    # sampling through the reference's sample_images (train_ldm.py:332-366): full reverse process of T = 12 steps + decode
    imgs = ldm.sample_images(z_shape, inferer, verbose=False, seed=42)
    assert tuple(imgs.shape) == (2, 1, 16, 16, 16) and torch.isfinite(imgs).all()
    # and the whole orchestration LDM.train (train_ldm.py:507-556): GradScaler, scale-factor probe, torchinfo dry run,
    # one epoch of train + validate, loss plots (no-op pyplot), checkpoint + loss_dict.pkl
    cfg3 = dict(cfg, results_path=str(tmp_path / "ldm_full"))
    ldm3 = tl.LDM(cfg3, latent_space_type="vae")
    ldm3.train(train_loader=loader, val_loader=loader[:2])
    assert os.path.isfile(os.path.join(cfg3["results_path"], "checkpoints", "best_model.pth"))
    assert os.path.isfile(os.path.join(cfg3["results_path"], "loss_dict.pkl"))
    assert len(ldm3.loss_dict["rec_loss"]) == 1 and len(ldm3.loss_dict["val_rec_loss"]) == 1


# ---- train_autoencoder.AutoEncoder (SURVEY.md 8f-1 second trainer, 8f-3 adversarial loss) --------------------------------
def test_unmodified_autoencoder_trainer_imports_and_binds_b200_classes(trainer_env):
    """CPU: `from medimgen.train_autoencoder import AutoEncoder` works; AutoencoderKL is the B200 class; the adversarial
    pieces it imports from `generative` are real (shipped restatements), the perceptual loss is the shipped LPIPS-VGG restatement and needs a local weight file."""
    import medical_image_generation_b200 as mig
    ta = __import__("medimgen.train_autoencoder", fromlist=["AutoEncoder"])
    assert ta.__file__.endswith(os.path.join("medimgen", "train_autoencoder.py"))
    assert ta.AutoencoderKL is mig.AutoencoderKL
    disc = ta.PatchDiscriminator(spatial_dims=3, in_channels=1, out_channels=1, num_channels=8, num_layers_d=3)
    feats = disc(torch.rand(2, 1, 32, 32, 32))
    assert [tuple(f.shape[1:]) for f in feats] == [(8, 16, 16, 16), (16, 8, 8, 8), (32, 4, 4, 4), (64, 3, 3, 3), (1, 2, 2, 2)]
    keys = list(disc.state_dict())
    assert keys[:3] == ["initial_conv.conv.weight", "initial_conv.conv.bias", "0.conv.weight"] and "0.conv.bias" not in keys
    assert "2.adn.N.running_var" in keys and keys[-2:] == ["final_conv.conv.weight", "final_conv.conv.bias"]
    adv = ta.PatchAdversarialLoss(criterion="least_squares")
    logits = torch.tensor([[2.0, -2.0]])
    # least squares on LeakyReLU(0.05)(logits): ((2-1)^2 + (-0.1-1)^2) / 2 and (2^2 + 0.1^2) / 2
    assert abs(float(adv(logits, target_is_real=True, for_discriminator=True)) - (1.0 + 1.21) / 2) < 1e-6
    assert abs(float(adv(logits, target_is_real=False, for_discriminator=True)) - (4.0 + 0.01) / 2) < 1e-6
    with pytest.warns(UserWarning):
        g = adv(logits, target_is_real=False, for_discriminator=False)     # a generator target is always "real"
    assert abs(float(g) - (1.0 + 1.21) / 2) < 1e-6
    os.environ.pop("MEDIMGEN_LPIPS_WEIGHTS", None)
    with pytest.raises(RuntimeError, match="local file"):      # published LPIPS weights cannot be downloaded here
        ta.PerceptualLoss(spatial_dims=3, network_type="vgg")
    assert ta.PerceptualLoss(spatial_dims=3, network_type="vgg", is_fake_3d=True, fake_3d_ratio=0.2, pretrained=False)


def _ae_config(tmp_path):
    down = [[[1, 1, 1], [3, 3, 3], [1, 1, 1]], [[2, 2, 2], [3, 3, 3], [1, 1, 1]]]
    vae = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=3, num_res_blocks=1,
               with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False, use_flash_attention=False,
               use_checkpointing=False, use_convtranspose=False, num_channels=[32, 64], attention_levels=[False, False],
               norm_num_groups=16, downsample_parameters=down, upsample_parameters=list(reversed(down))[:-1])
    return dict(vae_params=vae, load_model_path=None, ae_learning_rate=1e-3, d_learning_rate=1e-3, lr_scheduler=None,
                lr_scheduler_params=None, grad_accumulate_step=2, grad_clip_max_norm=1, kl_weight=1e-6, adv_weight=0.05,
                perc_weight=0.125, autoencoder_warm_up_epochs=2, output_mode="log", progress_bar=False,
                results_path=str(tmp_path / "ae"), n_epochs=3, val_plot_interval=10, ae_batch_size=2,
                ae_transformations=dict(patch_size=[16, 16, 16]),
                discriminator_params=dict(spatial_dims=3, in_channels=1, out_channels=1, num_channels=8, num_layers_d=2))


@pytest.mark.gpu
def test_unmodified_train_autoencoder_runs_on_b200_modules(trainer_env, tmp_path):
    """AutoEncoder.train_one_epoch through the reference's own code (train_autoencoder.py:331-436): generator step (B200
    AutoencoderKL forward, L1 + KL + perceptual + -- after the warm-up epochs -- adversarial loss), discriminator step,
    the per-step requires_grad toggling (:374-377,401-404), fp16 autocast + two GradScalers, accumulation, clipping, Adam;
    then validate_one_epoch and save_model / load_model. The perceptual loss is the shipped LPIPS-VGG
    `PerceptualLoss` with fake-3D slice sampling as configuration.py:964 configures it, randomly initialised (its published
    weights cannot be downloaded here)."""
    from torch.cuda.amp import GradScaler
    ta = __import__("medimgen.train_autoencoder", fromlist=["AutoEncoder"])
    torch.manual_seed(0)
    cfg = _ae_config(tmp_path)
    ae = ta.AutoEncoder(cfg, latent_space_type="vae", print_summary=False)
    disc = ta.PatchDiscriminator(**cfg["discriminator_params"]).to(ae.device)
    opt_g, opt_d, sched_g, sched_d = ae.get_optimizers_and_lr_schedules(disc)
    assert isinstance(opt_g, torch.optim.Adam) and sched_g is None and sched_d is None

    # the reference builds PerceptualLoss(**config['perceptual_params']) (train_autoencoder.py:601) with downloaded LPIPS
    # weights; there is no network here, so the same class is built with pretrained=False (random VGG-16 trunk)
    perceptual = ta.PerceptualLoss(spatial_dims=3, network_type="vgg", is_fake_3d=True, fake_3d_ratio=0.2,
                                   pretrained=False).to(ae.device)

    gen = torch.Generator().manual_seed(3)
    base = torch.rand(2, 1, 16, 16, 16, generator=gen)
    loader = [{"id": i, "image": (base + 0.05 * torch.rand(2, 1, 16, 16, 16, generator=gen)).clamp(0, 1)} for i in range(5)]
    w_before = {k: v.detach().clone() for k, v in ae.autoencoder.state_dict().items()}
    d_before = {k: v.detach().clone() for k, v in disc.state_dict().items()}
    sg, sd = GradScaler(), GradScaler()
    for epoch in (1, 2, 3, 4):     # epochs 1 (< warm-up 2): no adversarial terms; 2-4: generator + discriminator steps
        ae.train_one_epoch(epoch, loader, disc, perceptual, opt_g, opt_d, sg, sd)
    ld = ae.loss_dict
    assert len(ld["rec_loss"]) == 4 and all(v == v and v > 0 for v in ld["rec_loss"]), ld
    assert ld["rec_loss"][-1] < ld["rec_loss"][0], ld["rec_loss"]
    assert ld["gen_loss"][0] == 0 and ld["disc_loss"][0] == 0                   # warm-up epoch
    assert all(v > 0 for v in ld["gen_loss"][1:]) and all(v > 0 for v in ld["disc_loss"][1:]), ld
    assert all(v == v and v >= 0 for v in ld["reg_loss"] + ld["perc_loss"])
    w_after, d_after = ae.autoencoder.state_dict(), disc.state_dict()
    assert sum(not torch.equal(w_before[k], w_after[k]) for k in w_before) >= 0.9 * len(w_before)
    assert any(not torch.equal(d_before[k], d_after[k]) for k in d_before if k.endswith("conv.weight"))
    assert all(p.requires_grad is False for p in ae.autoencoder.parameters())   # state the discriminator step leaves behind
    ae.validate_one_epoch(loader[:2])
    assert len(ld["val_rec_loss"]) == 1 and 0 < ld["val_rec_loss"][0] < 1
    ae.save_model(4, ld["val_rec_loss"][-1], opt_g, disc, opt_d)
    ckpt_path = os.path.join(cfg["results_path"], "checkpoints", "last_model.pth")
    ckpt = torch.load(ckpt_path)
    assert set(ckpt) == {"epoch", "network_state_dict", "optimizer_state_dict", "validation_loss", "discriminator_state_dict",
                         "disc_optimizer_state_dict"}
    ae2 = ta.AutoEncoder(cfg, latent_space_type="vae", print_summary=False)
    disc2 = ta.PatchDiscriminator(**cfg["discriminator_params"]).to(ae2.device)
    og2, od2, _, _ = ae2.get_optimizers_and_lr_schedules(disc2)
    assert ae2.load_model(ckpt_path, optimizer=og2, discriminator=disc2, disc_optimizer=od2, for_training=True) == 5
    x = loader[0]["image"].cuda()
    with torch.no_grad():
        mu1, _ = ae.autoencoder.eval().encode(x)
        mu2, _ = ae2.autoencoder.eval().encode(x)
    assert float((mu1.float() - mu2.float()).norm() / mu2.float().norm()) < 2e-2
