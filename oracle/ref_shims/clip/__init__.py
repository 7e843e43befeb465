def load(*a, **k):  # import-only stub
    raise NotImplementedError("CLIP weights are not available offline")
