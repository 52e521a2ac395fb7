def get(*args, **kwargs):
    return None
