"""Import shim: put `<site>/hybrid_ctunet_b200/dropin` FIRST on sys.path and the reference's own scripts
(`from networks.hybrid_CTUNet import CTUNet`, main_CTUNet.py:28; `from networks.resnet import generate_model`,
hybrid_CTUNet.py:21; `from trainer_CTUNet import sliding_window_inference`, test_CTUNet.py:25) resolve to the sm_100a
drop-in modules without editing a line of them."""
