from . import _Inert


def subplots(*a, **k):
    return _Inert(), _Inert()


def __getattr__(name):
    return _Inert()
