class Registry(dict):
    """name -> callable table with a decorator form (upstream utils/registry.py)."""

    def register(self, name, fn=None):
        if fn is not None:
            self[name] = fn
            return fn

        def deco(f):
            self[name] = f
            return f
        return deco
