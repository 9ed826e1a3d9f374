"""Drop-in replacements for the reference's `networks` package (resnet, vit, hybrid_CTUNet)."""
