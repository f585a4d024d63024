registry = {}


def register(id, entry_point=None, kwargs=None, **_):
    registry[id] = {'entry_point': entry_point, 'kwargs': kwargs or {}}
