from .pointnet import (STN3d, STNkd, PointNetfeat, PointNetCls, PointNetSeg,          # noqa: F401
                       PointNetSeg_regulization, PointNetDenseCls,
                       feature_transform_regularizer)
from .discriminator import (ConvDiscNet, DeepConvDiscNet, PointwiseDiscNet, BaseDiscNet,  # noqa: F401
                            ShapeDiscNet, PointDiscNet, StackDiscNet)
