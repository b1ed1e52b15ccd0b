"""Model factory / initialisation with the reference's interface
(utils/model_utils.py:12-140), building the libpcadv-backed modules."""
import torch
from torch.nn import init

from ..models import (PointNetCls, PointNetSeg, PointNetSeg_regulization, DeepConvDiscNet,
                      PointwiseDiscNet, BaseDiscNet, ShapeDiscNet, PointDiscNet, StackDiscNet)


def init_weights(net, init_type, init_gain=1.0):
    """Weights of every module whose class name contains Conv or Linear get the
    chosen initialiser, biases zero -- the selection rule and draw order of
    utils/model_utils.py:36-58, so a seeded build matches the reference's."""
    fillers = {
        "normal": lambda w: init.normal_(w, 0.0, init_gain),
        "xavier": lambda w: init.xavier_normal_(w, gain=init_gain),
        "kaiming": lambda w: init.kaiming_normal_(w, a=0, mode="fan_in"),
        "orthogonal": lambda w: init.orthogonal_(w, gain=init_gain),
    }
    if init_type not in fillers:
        raise NotImplementedError("initialization method [%s] is not implemented" % init_type)

    def visit(m):
        name = m.__class__.__name__
        if hasattr(m, "weight") and ("Conv" in name or "Linear" in name):
            fillers[init_type](m.weight.data)
            if getattr(m, "bias", None) is not None:
                init.constant_(m.bias.data, 0.0)

    net.apply(visit)


def init_net(net, device, init_type, init_gain=1.0):
    net.to(device)
    if init_type is not None:
        init_weights(net, init_type, init_gain=init_gain)
    return net


def load_models(mode, device, args):
    """Modes of utils/model_utils.py:60-140: cls, seg, seg_regu, disc, disc_seg,
    disc_dual, disc_stack."""
    def maybe_load(model):
        ckpt = getattr(args, "checkpoint", None)
        if ckpt:
            model.load_state_dict(torch.load(ckpt, map_location=device))
        return model

    if mode == "cls":
        return maybe_load(PointNetCls(k=40, feature_transform=False).to(device))
    if mode == "seg":
        return maybe_load(init_net(PointNetSeg(NUM_SEG_CLASSES=50), device, args.init_disc))
    if mode == "seg_regu":
        return maybe_load(init_net(PointNetSeg_regulization(NUM_SEG_CLASSES=50), device, args.init_disc))
    if mode == "disc":
        return init_net(DeepConvDiscNet(input_dim=args.disc_indim, output_dim=1), device, args.init_disc)
    if mode == "disc_seg":
        return init_net(PointwiseDiscNet(input_pts=args.input_pts, input_dim=args.disc_indim), device,
                        args.init_disc)
    if mode == "disc_dual":
        nets = (BaseDiscNet(input_pts=args.input_pts, input_dim=args.disc_indim, output_dim=256),
                ShapeDiscNet(shared_output_dim=256, num_shapes=16),
                PointDiscNet(shared_output_dim=256, input_pts=args.input_pts))
        return tuple(init_net(n, device, args.init_disc) for n in nets)
    if mode == "disc_stack":
        return init_net(StackDiscNet(input_pts=args.input_pts, input_dim=args.disc_indim, num_shapes=16),
                        device, args.init_disc)
    raise ValueError("Invalid mode {}!".format(mode))
