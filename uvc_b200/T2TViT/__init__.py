"""T2T-ViT backbone behind the sm_100a engine (reference: UVC/T2TViT/)."""
