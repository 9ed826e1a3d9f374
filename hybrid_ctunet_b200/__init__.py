"""B200-native (sm_100a) CTUNet hot path: drop-in modules over hand-written CUDA kernels.

Public surface mirrors the reference's modules:
    hybrid_ctunet_b200.networks.hybrid_CTUNet : CTUNet, CUNet, TUNet        (networks/hybrid_CTUNet.py)
    hybrid_ctunet_b200.networks.resnet        : ResNet, generate_model      (networks/resnet.py)
    hybrid_ctunet_b200.networks.vit           : ViT                         (networks/vit.py)
    hybrid_ctunet_b200.trainer_CTUNet         : sliding_window_inference    (trainer_CTUNet.py:417)
    hybrid_ctunet_b200.trainer_CUNet          : sliding_window_inference    (trainer_CUNet.py:268)
"""
__version__ = "0.1.0"
