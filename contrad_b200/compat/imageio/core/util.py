def _precision_warn(*args, **kwargs):
    pass
