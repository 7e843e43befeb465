def _xy_to_ray_bundle(*a, **k):  # import-only stub
    raise NotImplementedError
