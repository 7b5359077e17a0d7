"""state_dict -> flat fp32 weight blob in the order include/snsde.h documents.

Parameter names/shapes are the reference's (pinned by its own test,
/root/reference/tests/test_neuralsde_core_alignment.py:102-104;
declaration order neuralsde.py:146-179).
"""
import torch

from . import _lib

TIME_INPUT_OPTS = (3, 4, 5, 6)
CONTROL_EMB_OPTS = (2, 4, 6)


def describe(sde):
    """Infer the model descriptor fields from a duck-typed SDE module."""
    sd = sde.state_dict()
    if "linear_X.weight" in sd and "f_net._model.0.weight" in sd:
        H, C = sd["linear_X.weight"].shape
        HH = sd["f_net._model.0.weight"].shape[0]
        L = sum(1 for k in sd if k.startswith("f_net._model.") and k.endswith(".weight")) - 1
        return dict(family=_lib.FAMILY_TUTORIAL_LSDE, input_option=0, noise_option=0,
                    input_channels=C, hidden=H, hidden_hidden=HH, num_hidden_layers=L)
    if "qy0_mean" in sd and "py0_mean" in sd and "embedding.weight" in sd and "linear_in.weight" in sd:
        # LatentSDE (torch-ists/torch_ists/diff_module/NSDE/latent_sde.py:29-55): hidden = latent width + 1
        H = sd["embedding.weight"].shape[0]
        HH = sd["linear_in.weight"].shape[0]
        L = sum(1 for k in sd if k.startswith("linears.") and k.endswith(".weight")) + 1
        C = sd["initial_network.0.weight"].shape[1] if "initial_network.0.weight" in sd else 1
        if sd["linear_in.weight"].shape[1] != H + 1 or sd["linear_out.weight"].shape[0] != H - 1:
            raise ValueError("snsde: LatentSDE layer shapes do not match latent_sde.py:48-52")
        return dict(family=_lib.FAMILY_LATENT_SDE, input_option=0, noise_option=0, input_channels=C,
                    hidden=H, hidden_hidden=HH, num_hidden_layers=L)
    if "initial_network.weight" in sd and "linear_in.weight" in sd and hasattr(sde, "input_option"):
        H, C = sd["initial_network.weight"].shape
        HH = sd["linear_in.weight"].shape[0]
        L = sum(1 for k in sd if k.startswith("linears.") and k.endswith(".weight")) + 1
        return dict(family=_lib.FAMILY_BENCHMARK, input_option=int(sde.input_option),
                    noise_option=int(sde.noise_option), input_channels=C, hidden=H,
                    hidden_hidden=HH, num_hidden_layers=L)
    raise ValueError("snsde: SDE object is neither a Diffusion_model (neuralsde.py:123), the tutorial "
                     "NeuralLSDEFunc nor a LatentSDE (latent_sde.py:29); the engine does not call Python f/g")


def blob_keys(desc):
    """state_dict keys in blob order.  The trainable ones (see :func:`grad_keys`) always form a prefix."""
    L = desc["num_hidden_layers"]
    if desc["family"] == _lib.FAMILY_LATENT_SDE:
        mods = ["linear_in"] + [f"linears.{l}" for l in range(L - 1)] + ["linear_out"]
        return [f"{m}.{p}" for m in mods for p in ("weight", "bias")] + ["theta", "mu", "sigma"]
    if desc["family"] == _lib.FAMILY_TUTORIAL_LSDE:
        mlp = [f"_model.{2 * i}" for i in range(L + 1)]
        mods = (["linear_X", "emb"] + [f"f_net.{m}" for m in mlp] + ["linear_out", "noise_in"]
                + [f"g_net.{m}" for m in mlp])
        return [f"{m}.{p}" for m in mods for p in ("weight", "bias")]
    io, no = desc["input_option"], desc["noise_option"]
    mods = ["initial_network", "linear_in"] + (["emb"] if io in CONTROL_EMB_OPTS else [])
    mods += [f"linears.{l}" for l in range(L - 1)] + ["linear_out"]
    keys = [f"{m}.{p}" for m in mods for p in ("weight", "bias")] + ["theta"]
    if no in (1, 2, 3):
        keys.append("sigma")
    if no in (4, 5, 6):
        keys.append("sigma_diag")
    noise = {12: ["noise_t"], 13: ["noise_t"], 14: ["noise_y"], 15: ["noise_y"],
             16: ["noise_t.0", "noise_t.2"], 17: ["noise_t.0", "noise_t.2"],
             18: ["noise_y.0", "noise_y.2"], 19: ["noise_y.0", "noise_y.2"]}.get(no, [])
    keys += [f"{m}.{p}" for m in noise for p in ("weight", "bias")]
    return keys


def grad_keys(desc):
    """The blob entries that are nn.Parameters (the LatentSDE prior's theta / mu / sigma are buffers, latent_sde.py:35-37)."""
    keys = blob_keys(desc)
    return keys[:-3] if desc["family"] == _lib.FAMILY_LATENT_SDE else keys


def pack(sde, desc):
    """Returns a contiguous fp32 CPU tensor (one D2H copy when the module lives on a GPU)."""
    sd = sde.state_dict()
    missing = [k for k in blob_keys(desc) if k not in sd]
    if missing:
        raise ValueError(f"snsde: state_dict lacks {missing}")
    flat = torch.cat([sd[k].detach().reshape(-1).to(torch.float32) for k in blob_keys(desc)])
    return flat.cpu().contiguous()


def weights_version(sde):
    """Changes whenever a tensor of the weight blob is updated in place or replaced (buffers included: the LatentSDE
    prior's theta / mu / sigma are buffers)."""
    return tuple((p.data_ptr(), p._version) for p in list(sde.parameters()) + list(sde.buffers()))
