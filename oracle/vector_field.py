"""Oracle (test infrastructure): the SDE vector fields of the hot path.

``DiffusionModel`` restates reference ``Diffusion_model``
(benchmark_classification/models_sde/neuralsde.py:123-307; byte-identical copies at
benchmark_forecasting/models_sde/neuralsde.py:189-373 and
torch-ists/torch_ists/diff_module/NSDE/nsde_model.py:147-331) with the SAME
state_dict key names and shapes, so a reference state_dict loads unchanged.

``TutorialLSDEFunc`` restates ``NeuralLSDEFunc`` of
tutorial/"simple OU process - Neural LSDE.ipynb" cell 7 (BASELINE config c1).

f/g here evaluate one op per ATen call in the reference's order - including the
B-fold redundant ``noise_t(time_features)`` and the dead ``X.evaluate`` for input
options 1/3/5 - because this module doubles as the timed CPU baseline.
"""
import torch
from torch import nn

from .spline import CubicSpline

TIME_INPUT_OPTS = (3, 4, 5, 6)      # neuralsde.py:148  sin/cos(t) prepended to y
CONTROL_EMB_OPTS = (2, 4, 6)        # neuralsde.py:153  emb(cat(yy, Xt))
LATENT_ONLY_OPTS = (1, 3, 5)        # neuralsde.py:208  z = yy
GEOMETRIC_OPTS = (5, 6)             # neuralsde.py:220  z * tanh(y)

NAMED_MODELS = {                    # benchmark_classification/common_sde.py:303-342, README.md:32
    "staticsde": (1, 0), "naivesde": (1, 18), "neurallsde": (2, 16),
    "neurallnsde": (4, 17), "neuralgsde": (6, 17), "neuralsde_3_18": (3, 18),
}


class DiffusionModel(nn.Module):
    sde_type = "ito"
    noise_type = "diagonal"

    def __init__(self, input_channels, hidden_channels, hidden_hidden_channels, num_hidden_layers,
                 theta=1.0, sigma=1.0, input_option=0, noise_option=0):
        super().__init__()
        H, HH = hidden_channels, hidden_hidden_channels
        self.input_option, self.noise_option = input_option, noise_option
        self.input_channels, self.hidden_channels = input_channels, H
        self.hidden_hidden_channels, self.num_hidden_layers = HH, num_hidden_layers
        self.initial_network = nn.Linear(input_channels, H)
        self.linear_in = nn.Linear(H + 2 if input_option in TIME_INPUT_OPTS else H, HH)
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

    # -- control ----------------------------------------------------------------
    def set_X(self, coeffs, times):
        self.coeffs, self.times = coeffs, times
        self.X = CubicSpline(coeffs, times)

    # -- helpers ----------------------------------------------------------------
    @staticmethod
    def _time_column(t, y):
        if t.dim() == 0:
            t = torch.full_like(y[:, 0], fill_value=t).unsqueeze(-1)     # host sync in the reference too
        return t

    def _time_features(self, t, y):
        t = self._time_column(t, y)
        return t, torch.cat((torch.sin(t), torch.cos(t)), dim=-1)

    # -- drift (neuralsde.py:295-302) ------------------------------------------
    def f(self, t, y):
        Xt = self.initial_network(self.X.evaluate(t))
        if self.input_option in TIME_INPUT_OPTS:
            _, tf = self._time_features(t, y)
            yy = self.linear_in(torch.cat((tf, y), dim=-1))
        else:
            yy = self.linear_in(y)
        if self.input_option == 0:
            z = Xt
        elif self.input_option in LATENT_ONLY_OPTS:
            z = yy
        else:
            z = self.emb(torch.cat([yy, Xt], dim=-1))
        z = z.relu()
        for lin in self.linears:
            z = lin(z).relu()
        z = self.linear_out(z)
        if self.input_option in GEOMETRIC_OPTS:
            z = z * y.tanh()
        return z.tanh()

    # -- diffusion (neuralsde.py:233-293, 304-307) ------------------------------
    def raw_diffusion(self, t, y):
        t, tf = self._time_features(t, y)
        n = self.noise_option
        B, H = y.size(0), y.size(1)
        if n == 0:
            return torch.zeros(B, H).to(y.device)
        if n in (1, 2, 3):
            s = self.sigma.exp().expand(B, H)
            return s if n == 1 else (s * t if n == 2 else s * y)
        if n in (4, 5, 6):
            s = self.sigma_diag.exp().repeat(B, 1)
            return s if n == 4 else (s * t if n == 5 else s * y)
        if n == 7:
            return torch.sqrt(y)
        if n == 8:
            return y ** 3
        if n == 9:
            return y.sigmoid()
        if n == 10:
            return y.relu()
        if n == 11:
            return t * y
        if n == 12:
            return self.noise_t(tf)
        if n == 13:
            return self.noise_t(tf) * y
        if n == 14:
            return self.noise_y(torch.cat([tf, y], dim=-1))
        if n == 15:
            return self.noise_y(torch.cat([tf, y], dim=-1)) * y
        if n == 16:
            return self.noise_t(tf).relu()
        if n == 17:
            return self.noise_t(tf).relu() * y
        if n == 18:
            return self.noise_y(torch.cat([tf, y], dim=-1)).relu()
        if n == 19:
            return self.noise_y(torch.cat([tf, y], dim=-1)).relu() * y
        raise ValueError(f"Unknown noise_option {n}.")

    def g(self, t, y):
        noise = self.theta.sigmoid() * torch.nan_to_num(self.raw_diffusion(t, y))
        return noise.tanh()


class _LipSwish(nn.Module):
    def forward(self, x):
        return 0.909 * nn.functional.silu(x)


class _MLP(nn.Module):
    """Linear -> act -> (Linear -> act)*(L-1) -> Linear; keys ``_model.{i}`` as in the notebook."""

    def __init__(self, in_size, out_size, hidden_dim, num_layers, activation="lipswish"):
        super().__init__()
        act = _LipSwish() if activation == "lipswish" else nn.ReLU()
        mods = [nn.Linear(in_size, hidden_dim), act]
        for _ in range(num_layers - 1):
            mods += [nn.Linear(hidden_dim, hidden_dim), act]
        mods.append(nn.Linear(hidden_dim, out_size))
        self._model = nn.Sequential(*mods)

    def forward(self, x):
        return self._model(x)


class TutorialLSDEFunc(nn.Module):
    """f = linear_out(f_net(emb(cat(y, linear_X(X(t)))))), g = g_net(noise_in(t)) (row-independent)."""
    sde_type = "ito"
    noise_type = "diagonal"

    def __init__(self, input_dim, hidden_dim, hidden_hidden_dim, num_layers, activation="lipswish"):
        super().__init__()
        self.input_channels, self.hidden_channels = input_dim, hidden_dim
        self.hidden_hidden_channels, self.num_hidden_layers = hidden_hidden_dim, num_layers
        self.linear_X = nn.Linear(input_dim, hidden_dim)
        self.emb = nn.Linear(hidden_dim * 2, hidden_dim)
        self.f_net = _MLP(hidden_dim, hidden_dim, hidden_hidden_dim, num_layers, activation)
        self.linear_out = nn.Linear(hidden_dim, hidden_dim)
        self.noise_in = nn.Linear(1, hidden_dim)
        self.g_net = _MLP(hidden_dim, hidden_dim, hidden_hidden_dim, num_layers, activation)

    def set_X(self, coeffs, times):
        self.coeffs, self.times = coeffs, times
        self.X = CubicSpline(coeffs, times)

    def f(self, t, y):
        Xt = self.linear_X(self.X.evaluate(t))
        z = self.emb(torch.cat([y, Xt], dim=-1))
        return self.linear_out(self.f_net(z))

    def g(self, t, y):
        if t.dim() == 0:
            t = torch.full_like(y[:, 0], fill_value=t).unsqueeze(-1)
        return self.g_net(self.noise_in(t))
