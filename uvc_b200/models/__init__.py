from .model_distilled import (DistilledVisionTransformer, VisionTransformer, Block, Attention, Mlp, PatchEmbed,  # noqa: F401
                              deit_tiny_patch16_224, deit_small_patch16_224, deit_base_patch16_224)
from .configs import CONFIGS  # noqa: F401
