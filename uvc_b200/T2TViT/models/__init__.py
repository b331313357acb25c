from .t2t_vit import T2T_ViT, T2T_module, t2t_vit_14  # noqa: F401
from .token_performer import Token_performer  # noqa: F401
