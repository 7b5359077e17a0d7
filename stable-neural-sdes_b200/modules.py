"""Parameter containers with the reference's state_dict layout.

The engine never calls Python ``f``/``g``: it needs ``state_dict()``, ``coeffs`` and ``times``
only.  In production that object is the reference's own ``Diffusion_model``
(/root/reference/benchmark_classification/models_sde/neuralsde.py:123-184); these classes are
stand-ins with identical parameter names, shapes, declaration order and default init
(neuralsde.py:146-179; tutorial notebook cell 7) for benchmarks, serving and tests where the
reference package is not importable.
"""
import math

import torch
from torch import nn

from .packing import CONTROL_EMB_OPTS, TIME_INPUT_OPTS


class _ControlMixin:
    sde_type = "ito"
    noise_type = "diagonal"

    def set_X(self, coeffs, times):
        if isinstance(coeffs, (tuple, list)):
            coeffs = coeffs[0] if len(coeffs) == 1 else torch.cat(list(coeffs), dim=-1)
        self.coeffs, self.times = coeffs, times

    def f(self, t, y):
        raise NotImplementedError("parameter container: the drift is evaluated inside the CUDA engine")

    g = f


class DiffusionModelParams(_ControlMixin, nn.Module):
    def __init__(self, input_channels, hidden_channels, hidden_hidden_channels, num_hidden_layers,
                 theta=1.0, sigma=1.0, input_option=0, noise_option=0):
        super().__init__()
        H, HH = hidden_channels, hidden_hidden_channels
        self.input_option, self.noise_option = input_option, noise_option
        self.input_channels, self.hidden_channels = input_channels, H
        self.initial_network = nn.Linear(input_channels, H)
        self.linear_in = nn.Linear(H + (2 if input_option in TIME_INPUT_OPTS else 0), HH)
        if input_option in CONTROL_EMB_OPTS:
            self.emb = nn.Linear(2 * H, H)
        self.linears = nn.ModuleList(nn.Linear(HH, HH) for _ in range(num_hidden_layers - 1))
        self.linear_out = nn.Linear(HH, H)
        self.theta = nn.Parameter(torch.tensor([[theta]]))
        if noise_option in (1, 2, 3):
            self.sigma = nn.Parameter(torch.tensor([sigma]))
        if noise_option in (4, 5, 6):
            self.sigma_diag = nn.Parameter(torch.tensor([sigma] * H))
        if noise_option in (12, 13):
            self.noise_t = nn.Linear(2, H)
        if noise_option in (14, 15):
            self.noise_y = nn.Linear(H + 2, H)
        if noise_option in (16, 17):
            self.noise_t = nn.Sequential(nn.Linear(2, H), nn.ReLU(), nn.Linear(H, H))
        if noise_option in (18, 19):
            self.noise_y = nn.Sequential(nn.Linear(H + 2, H), nn.ReLU(), nn.Linear(H, H))


class _MLPParams(nn.Module):
    def __init__(self, in_size, out_size, hidden_dim, num_layers):
        super().__init__()
        mods = [nn.Linear(in_size, hidden_dim), nn.Identity()]
        for _ in range(num_layers - 1):
            mods += [nn.Linear(hidden_dim, hidden_dim), nn.Identity()]
        mods.append(nn.Linear(hidden_dim, out_size))
        self._model = nn.Sequential(*mods)


class TutorialLSDEParams(_ControlMixin, nn.Module):
    def __init__(self, input_dim, hidden_dim, hidden_hidden_dim, num_layers):
        super().__init__()
        self.linear_X = nn.Linear(input_dim, hidden_dim)
        self.emb = nn.Linear(hidden_dim * 2, hidden_dim)
        self.f_net = _MLPParams(hidden_dim, hidden_dim, hidden_hidden_dim, num_layers)
        self.linear_out = nn.Linear(hidden_dim, hidden_dim)
        self.noise_in = nn.Linear(1, hidden_dim)
        self.g_net = _MLPParams(hidden_dim, hidden_dim, hidden_hidden_dim, num_layers)


class LatentSDEParams(nn.Module):
    """Stand-in for the reference ``LatentSDE`` (torch-ists/torch_ists/diff_module/NSDE/latent_sde.py:29-55): the
    same parameters / buffers under the same names; ``hidden_channels`` counts the KL accumulator channel."""
    sde_type = "ito"
    noise_type = "diagonal"

    def __init__(self, input_channels, hidden_channels, hidden_hidden_channels, num_hidden_layers,
                 theta=1.0, mu=0.0, sigma=0.5):
        super().__init__()
        H, HH = hidden_channels, hidden_hidden_channels
        logvar = math.log(sigma ** 2 / (2. * theta))
        self.register_buffer("theta", torch.tensor([[theta]]))
        self.register_buffer("mu", torch.tensor([[mu]]))
        self.register_buffer("sigma", torch.tensor([[sigma]]))
        self.register_buffer("py0_mean", torch.tensor([[mu]]))
        self.register_buffer("py0_logvar", torch.tensor([[logvar]]))
        self.initial_network = nn.Sequential(nn.Linear(input_channels, H - 1))
        self.linear_in = nn.Linear(H + 2 - 1, HH)
        self.linears = nn.ModuleList(nn.Linear(HH, HH) for _ in range(num_hidden_layers - 1))
        self.linear_out = nn.Linear(HH, H - 1)
        self.embedding = nn.Linear(H - 1, H)
        self.qy0_mean = nn.Parameter(torch.tensor([[mu]]))
        self.qy0_logvar = nn.Parameter(torch.tensor([[logvar]]))

    @property
    def py0_std(self):
        return torch.exp(.5 * self.py0_logvar)

    @property
    def qy0_std(self):
        return torch.exp(.5 * self.qy0_logvar)

    def f_aug(self, t, y):
        raise NotImplementedError("parameter container: the augmented drift is evaluated inside the CUDA engine")

    g_aug = f_aug
