"""Label helpers with the reference's signatures (utils/utils.py:22-38)."""
import torch


def make_D_label(input, value, device, random=False):
    """GAN target for ``input``'s shape: constant ``value``, or -- with
    ``random`` -- a smoothed label drawn on the CPU from U(0, 0.305) for value 0
    and U(0.7, 1.05) for value 1, then moved to ``device`` (utils/utils.py:22-31).
    The draw uses torch's CPU generator exactly as the reference does, so a
    seeded run produces the same labels on both sides."""
    shape = input.data.size()
    if random:
        bounds = {0: (0.0, 0.305), 1: (0.7, 1.05)}[value]
        label = torch.FloatTensor(shape).uniform_(*bounds)
    else:
        label = torch.FloatTensor(shape).fill_(value)
    return label.to(device)


def make_shape_label(input, npts):
    """argmax of the class one-hot repeated for every point (utils/utils.py:33-38)."""
    cls = torch.argmax(input, dim=1, keepdim=True)
    return cls.repeat(1, npts).long().to(input.device)
