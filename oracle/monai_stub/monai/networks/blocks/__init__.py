"""vit.py:7 imports these three names and never uses them."""
import torch.nn as nn


class UnetrBasicBlock(nn.Module):
    pass


class UnetrPrUpBlock(nn.Module):
    pass


class UnetrUpBlock(nn.Module):
    pass
