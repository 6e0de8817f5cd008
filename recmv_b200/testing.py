"""Deterministic construction helpers shared by the golden generator (which runs them on the
REFERENCE's classes) and the tests (which run them on this package's classes): identical seeds and
identical constructor call order give bit-identical parameters, so fixtures only need to store
inputs, outputs and parameter checksums instead of megabytes of weights."""
import torch


def perturb_module(mod, seed, scale=0.02):
    """'Trained-like' weights: seeded noise on every parameter, proportional to its RMS (keeps
    softplus pre-activations out of the saturated regime of a fresh geometric init)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    with torch.no_grad():
        for name, p in sorted(mod.named_parameters()):
            rms = p.detach().float().pow(2).mean().sqrt().clamp_min(1e-3)
            noise = torch.randn(p.shape, generator=g, dtype=torch.float32) * (scale * rms.cpu())
            p.add_(noise.to(p.device, p.dtype))
    return mod


def param_checksums(mod):
    """{name: (sum, abs-sum)} in float64 -- asserts two modules hold the same parameters."""
    out = {}
    for name, p in sorted(mod.named_parameters()):
        d = p.detach().double().cpu()
        out[name] = (float(d.sum()), float(d.abs().sum()))
    return out


def build_sdf(cls_or_factory, seed=0, perturb_seed=None, device="cpu"):
    torch.manual_seed(seed)
    net = cls_or_factory(device, 6, 0.6, 256)
    if perturb_seed is not None:
        perturb_module(net, perturb_seed)
    return net
