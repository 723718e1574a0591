def use_opt_einsum(*args, **kwargs):
    return None
