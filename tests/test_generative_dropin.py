"""The reference's import path (`generative.*`) served by this repository: the sampling cells
of the reference tutorials run with ONLY ``sys.path`` changed.  The cell texts below are the tutorials' own API usage
(cited per cell); what is shortened is the number of inference steps, and trained weights are replaced by a
``load_state_dict`` of the seeded recipe, as a user would load a checkpoint.  Each script runs in a fresh interpreter
(the parent process may have the reference itself imported as ``generative`` for the oracle tests) and its result is
compared with the CPU oracle."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]

_IMPORTS = '''
import sys
sys.path.insert(0, {root!r})                 # <- the only change: this repository instead of the reference checkout
import torch
from torch.cuda.amp import autocast
from generative.inferers import DiffusionInferer
from generative.networks.nets import DiffusionModelUNet
from generative.networks.schedulers import DDPMScheduler, DDIMScheduler
'''

# tutorials/generative/3d_ddpm/3d_ddpm_tutorial.py:157-167 (network), 175 (scheduler), 186 (inferer),
# 256-260 (sampling cell inside the training loop), 318-330 (DDIM sampling cell)
_CELL_3D_DDPM = _IMPORTS + '''
device = torch.device("cuda")

model = DiffusionModelUNet(
    spatial_dims=3,
    in_channels=1,
    out_channels=1,
    num_channels=[256, 256, 512],
    attention_levels=[False, False, True],
    num_head_channels=[0, 0, 512],
    num_res_blocks=2,
)
model.load_state_dict(torch.load({weights!r}))
model.to(device)

scheduler = DDPMScheduler(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0005, beta_end=0.0195)

inferer = DiffusionInferer(scheduler)

model.eval()
torch.manual_seed(11)
image = torch.randn((1, 1, 32, 40, 32))
image = image.to(device)
scheduler.set_timesteps(num_inference_steps={steps})
with autocast(enabled=True):
    image = inferer.sample(input_noise=image, diffusion_model=model, scheduler=scheduler)

scheduler_ddim = DDIMScheduler(
    num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0005, beta_end=0.0195, clip_sample=False
)

scheduler_ddim.set_timesteps(num_inference_steps={steps})

model.eval()
torch.manual_seed(12)
noise = torch.randn((1, 1, 32, 40, 32))
noise = noise.to(device)

image_ddim = inferer.sample(input_noise=noise, diffusion_model=model, scheduler=scheduler_ddim)
torch.save(dict(ddpm=image.cpu(), ddim=image_ddim.cpu()), {out!r})
'''

# tutorials/generative/classifier_free_guidance/2d_ddpm_classifier_free_guidance_tutorial.py:193-203 (network),
# 206 (scheduler), 296-312 (sampling with classifier-free guidance)
_CELL_CFG = _IMPORTS + '''
device = torch.device("cuda")

model = DiffusionModelUNet(
    spatial_dims=2,
    in_channels=1,
    out_channels=1,
    num_channels=(64, 64, 64),
    attention_levels=(False, False, True),
    num_res_blocks=1,
    num_head_channels=(0, 0, 64),
    with_conditioning=True,
    cross_attention_dim=1,
)
model.load_state_dict(torch.load({weights!r}))
model.to(device)

scheduler = DDPMScheduler(num_train_timesteps=1000)

model.eval()
guidance_scale = 7.0
conditioning = torch.cat([-1 * torch.ones(1, 1, 1).float(), torch.ones(1, 1, 1).float()], dim=0).to(device)

torch.manual_seed(21)
noise = torch.randn((1, 1, 64, 64))
noise = noise.to(device)
scheduler.set_timesteps(num_inference_steps={steps})
progress_bar = iter(scheduler.timesteps)
for t in progress_bar:
    with autocast(enabled=True):
        with torch.no_grad():
            noise_input = torch.cat([noise] * 2)
            model_output = model(noise_input, timesteps=torch.Tensor((t,)).to(noise.device), context=conditioning)
            noise_pred_uncond, noise_pred_text = model_output.chunk(2)
            noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)

    noise, _ = scheduler.step(noise_pred, t, noise)
torch.save(dict(sample=noise.cpu()), {out!r})
'''


def _run(script: str):
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", script], capture_output=True, text=True, env=env,
                       cwd="/tmp", timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_generative_import_path_cpu():
    """Every import form the reference's tutorials / tests / bundles use for the sampling path resolves to this
    implementation (no GPU needed), and the out-of-scope subpackages fail loudly."""
    _run(f'''
import sys
sys.path.insert(0, {str(ROOT)!r})
from generative.inferers import (ControlNetDiffusionInferer, ControlNetLatentDiffusionInferer, DiffusionInferer,
                                 LatentDiffusionInferer, VQVAETransformerInferer)
from generative.networks.nets import (VQVAE, AutoencoderKL, ControlNet, DecoderOnlyTransformer, DiffusionModelUNet,
                                      SPADEAutoencoderKL, SPADEDiffusionModelUNet)
from generative.networks.nets.controlnet import ControlNet as C2
from generative.networks.nets.diffusion_model_unet import DiffusionModelUNet as D2
from generative.networks.nets.vqvae import VQVAE as V2
from generative.networks.layers import EMAQuantizer, VectorQuantizer
from generative.networks.schedulers import DDIMScheduler, DDPMScheduler, PNDMScheduler
from generative.networks.schedulers.ddim import DDIMScheduler as S2
from generative.networks.schedulers.ddpm import DDPMScheduler as S3
from generative.networks.blocks.selfattention import SABlock
from generative.utils import ComponentStore, unsqueeze_left, unsqueeze_right
from generative.utils.enums import OrderingTransformations, OrderingType
from generative.utils.ordering import Ordering
import generativemodels_b200.networks.nets as N
import generativemodels_b200.inferers as I
assert DiffusionModelUNet is N.DiffusionModelUNet is D2 and C2 is ControlNet and V2 is VQVAE
assert S2 is DDIMScheduler and S3 is DDPMScheduler and DiffusionInferer is I.DiffusionInferer
assert DiffusionModelUNet.__module__ == "generativemodels_b200.networks.nets.diffusion_model_unet"
for name in ("generative.losses", "generative.metrics", "generative.engines"):
    try:
        __import__(name)
    except ModuleNotFoundError as e:
        assert "out of scope" in str(e)
    else:
        raise AssertionError(name + " must not resolve")
m = DiffusionModelUNet(spatial_dims=2, in_channels=1, out_channels=1, num_channels=(32, 64), num_res_blocks=1,
                       attention_levels=(False, True), num_head_channels=(0, 32), norm_num_groups=8)
assert "down_blocks.1.attentions.0.to_q.weight" in m.state_dict()
''')


@pytest.mark.gpu
def test_tutorial_3d_ddpm_sampling_cells(cuda_device, tmp_path):
    from generativemodels_b200.networks.nets import DiffusionModelUNet
    from oracle import torch_oracle as O
    from tests.fixture_checks import TOL_TRAJ, close
    from tests.golden import configs as G
    steps = 3
    sd = G.recipe_state_dict(DiffusionModelUNet(**G.C3_UNET), 13)
    w, out = tmp_path / "w.pt", tmp_path / "out.pt"
    torch.save(sd, w)
    _run(_CELL_3D_DDPM.format(root=str(ROOT), weights=str(w), out=str(out), steps=steps))
    got = torch.load(out)
    cfg = G.unet_oracle_cfg(G.C3_UNET)
    fn = lambda x, t, c: O.unet_forward(sd, cfg, x, t, context=c)
    kw = dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0005, beta_end=0.0195)
    s = O.DDPMOracle(**kw)
    s.set_timesteps(steps)
    torch.manual_seed(11)
    want = O.diffusion_sample(fn, s, torch.randn((1, 1, 32, 40, 32)))
    close(got["ddpm"], want, "3d_ddpm tutorial, DDPM sampling cell", TOL_TRAJ, 2 * TOL_TRAJ)
    s = O.DDIMOracle(clip_sample=False, **kw)
    s.set_timesteps(steps)
    torch.manual_seed(12)
    want = O.diffusion_sample(fn, s, torch.randn((1, 1, 32, 40, 32)))
    close(got["ddim"], want, "3d_ddpm tutorial, DDIM sampling cell", TOL_TRAJ, 2 * TOL_TRAJ)


@pytest.mark.gpu
def test_tutorial_classifier_free_guidance_loop(cuda_device, tmp_path):
    from generativemodels_b200.networks.nets import DiffusionModelUNet
    from oracle import torch_oracle as O
    from tests.fixture_checks import TOL_TRAJ, close
    from tests.golden import configs as G
    steps = 4
    kw = dict(spatial_dims=2, in_channels=1, out_channels=1, num_channels=(64, 64, 64),
              attention_levels=(False, False, True), num_res_blocks=1, num_head_channels=(0, 0, 64),
              with_conditioning=True, cross_attention_dim=1)
    sd = G.recipe_state_dict(DiffusionModelUNet(**kw), 17)
    w, out = tmp_path / "w.pt", tmp_path / "out.pt"
    torch.save(sd, w)
    _run(_CELL_CFG.format(root=str(ROOT), weights=str(w), out=str(out), steps=steps))
    got = torch.load(out)["sample"]
    cfg = G.unet_oracle_cfg(kw)
    s = O.DDPMOracle(num_train_timesteps=1000)
    s.set_timesteps(steps)
    ctx = torch.cat([-1 * torch.ones(1, 1, 1), torch.ones(1, 1, 1)], dim=0)
    torch.manual_seed(21)
    x = torch.randn((1, 1, 64, 64))
    for t in s.timesteps:
        o = O.unet_forward(sd, cfg, torch.cat([x] * 2), torch.Tensor((t,)), context=ctx)
        eu, et = o.chunk(2)
        x, _ = s.step(eu + 7.0 * (et - eu), int(t), x)
    close(got, x, "classifier-free-guidance tutorial loop", TOL_TRAJ, 2 * TOL_TRAJ)
