"""TEST INFRASTRUCTURE ONLY -- inert stand-in for matplotlib (absent from the image); the reference
imports it at module level (tobias/tools/atacorrect_functions.py:19-22) but the hot path never plots."""
__version__ = "shim"


def use(*a, **k):
    return None


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
