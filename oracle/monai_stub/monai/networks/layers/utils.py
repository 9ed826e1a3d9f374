"""get_norm_layer / get_act_layer as MONAI 0.7.0 resolves the arguments the reference passes."""
import torch.nn as nn


def get_norm_layer(name, spatial_dims=1, channels=1):
    if isinstance(name, (tuple, list)):
        name, kwargs = name[0], dict(name[1])
    else:
        kwargs = {}
    key = str(name).upper()
    if key == "INSTANCE":
        return nn.InstanceNorm3d(channels, **kwargs)
    if key == "BATCH":
        return nn.BatchNorm3d(channels, **kwargs)
    raise NotImplementedError(name)


def get_act_layer(name):
    if isinstance(name, (tuple, list)):
        name, kwargs = name[0], dict(name[1])
    else:
        kwargs = {}
    key = str(name).upper()
    if key == "LEAKYRELU":
        return nn.LeakyReLU(**kwargs)
    if key == "PRELU":
        return nn.PReLU(**kwargs)
    raise NotImplementedError(name)
