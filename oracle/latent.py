"""Oracle (test infrastructure): the LatentSDE augmented system and its forward, restated.

Follows /root/reference/torch-ists/torch_ists/diff_module/NSDE/latent_sde.py:
``_stable_division`` :24-26, the posterior drift ``f`` :57-66, shared diffusion ``g`` :68-69, prior drift ``h`` :71-72,
``f_aug`` :74-82 (drift of the extra channel = 0.5 |u|^2, u = (f - h) / g), ``g_aug`` :84-90 (no noise on it) and
``forward`` :92-147 (controlled initial state, KL(t=0) + KL(path), default method 'srk', dt = max(min dt, 1e-3),
``torchsde.sdeint_adjoint(..., names={'drift': 'f_aug', 'diffusion': 'g_aug'})``).  PINNED: tests/golden/make_golden.py
imports the reference's own unmodified class (torchsde.SDEIto / torchcde shimmed) and freezes f_aug / g_aug and the
forward outputs into tests/golden/latent_golden.pt; the solver underneath is the oracle's (parity unpinned, see
oracle/__init__.py).
"""
import math

import torch
from torch import distributions, nn

from . import solver, spline

NAMES = {"drift": "f_aug", "diffusion": "g_aug"}


def stable_division(a, b, epsilon=1e-7):
    b = torch.where(b.abs().detach() > epsilon, b, torch.full_like(b, fill_value=epsilon) * b.sign())
    return a / b


class LatentSDE(nn.Module):
    sde_type = "ito"
    noise_type = "diagonal"

    def __init__(self, input_channels, hidden_channels, hidden_hidden_channels, num_hidden_layers,
                 theta=1.0, mu=0.0, sigma=0.5):
        super().__init__()
        lat = hidden_channels - 1                      # the last state channel accumulates the KL path term
        logvar = math.log(sigma ** 2 / (2. * theta))
        for name, v in (("theta", theta), ("mu", mu), ("sigma", sigma), ("py0_mean", mu), ("py0_logvar", logvar)):
            self.register_buffer(name, torch.tensor([[v]]))
        self.initial_network = nn.Sequential(nn.Linear(input_channels, lat))
        self.linear_in = nn.Linear(lat + 2, hidden_hidden_channels)
        self.linears = nn.ModuleList(nn.Linear(hidden_hidden_channels, hidden_hidden_channels)
                                     for _ in range(num_hidden_layers - 1))
        self.linear_out = nn.Linear(hidden_hidden_channels, lat)
        self.embedding = nn.Linear(lat, hidden_channels)
        self.qy0_mean = nn.Parameter(torch.tensor([[mu]]))
        self.qy0_logvar = nn.Parameter(torch.tensor([[logvar]]))

    def f(self, t, y):
        if t.dim() == 0:
            t = torch.full_like(y[:, 0], fill_value=t).unsqueeze(-1)
        z = self.linear_in(torch.cat((torch.sin(t), torch.cos(t), y), dim=-1)).relu()
        for lin in self.linears:
            z = lin(z).relu()
        return self.linear_out(z)

    def g(self, t, y):
        return self.sigma.expand(y.size(0), y.size(1))

    def h(self, t, y):
        return self.theta * (self.mu - y)

    def f_aug(self, t, y):
        y = y[:, :-1]
        f, g, h = self.f(t, y), self.g(t, y), self.h(t, y)
        u = stable_division(f - h, g)
        return torch.cat([f, .5 * (u ** 2).sum(dim=1, keepdim=True)], dim=1)

    def g_aug(self, t, y):
        y = y[:, :-1]
        return torch.cat([self.g(t, y), torch.zeros(y.shape[0], 1).to(y.device)], dim=1)

    @property
    def py0_std(self):
        return torch.exp(.5 * self.py0_logvar)

    @property
    def qy0_std(self):
        return torch.exp(.5 * self.qy0_logvar)

    def forward(self, coeffs, times, bm=None, method=None, with_grad=False):
        y0 = spline.CubicSpline(coeffs, times).evaluate(times[0])
        logqp0 = distributions.kl_divergence(distributions.Normal(self.qy0_mean, self.qy0_std),
                                             distributions.Normal(self.py0_mean, self.py0_std)).sum(dim=1)
        lat0 = self.initial_network(y0)
        aug_y0 = torch.cat([lat0, torch.zeros(coeffs.shape[0], 1).to(lat0)], dim=1)
        integ = solver.sdeint_with_grad if with_grad else solver.sdeint
        aug_ys = integ(self, aug_y0, times, solver.solver_dt(times), bm, method=method or "srk", names=NAMES)
        aug_ys = aug_ys.permute(1, 0, 2)
        latent = aug_ys[:, :, :-1]
        logqp = (logqp0 + aug_ys[:, -1, -1]).mean(dim=0)
        return self.embedding(latent), latent, logqp
