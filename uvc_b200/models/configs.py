"""Model hyper-parameter table read by joint_train.py:124,883-885 / post_train.py (`CONFIGS[args.model_type]`).

Re-statement of the entries of the reference's `models/configs.py:34-53,112-165` + `models/modeling.py:435-452`
that the UVC loops can reach, without the ml_collections dependency: attribute access (`config.embed_dim`)
and item access both work.
"""


class _Cfg(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _deit(hidden, heads, depth=12, mlp_ratio=4):
    tr = _Cfg(mlp_dim=hidden * mlp_ratio, num_heads=heads, num_layers=depth, attention_dropout_rate=0.0, dropout_rate=0.1)
    return _Cfg(patches=_Cfg(size=(16, 16)), hidden_size=hidden, transformer=tr, classifier='token', representation_size=None,
                patch_size=16, embed_dim=hidden, depth=depth, num_heads=heads, mlp_ratio=mlp_ratio)


def get_deit_tiny_config():
    return _deit(192, 3)


def get_deit_small_config():
    return _deit(384, 6)


def get_b16_config():
    return _deit(768, 12)


def get_t2t_vit_14_config():
    c = get_deit_small_config()
    c.depth, c.num_heads, c.mlp_ratio = 14, 6, 3
    return c


CONFIGS = {
    't2t_vit_14': get_t2t_vit_14_config(),
    'deit_base_patch16_224': get_b16_config(),
    'deit_small_patch16_224': get_deit_small_config(),
    'deit_tiny_patch16_224': get_deit_tiny_config(),
    'ViT-B_16': get_b16_config(),
}
