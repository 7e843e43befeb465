def padded_to_packed(*args, **kwargs):  # import-only stub
    raise NotImplementedError
