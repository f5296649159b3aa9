import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """CPU generator + any device => sample on CPU then move (seed-reproducible)."""
    rand_device = device
    if generator is not None:
        gdev = generator.device.type if not isinstance(generator, list) else generator[0].device.type
        if gdev == "cpu":
            rand_device = "cpu"
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)
