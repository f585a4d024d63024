def run_random(*a, **k):
    raise NotImplementedError
