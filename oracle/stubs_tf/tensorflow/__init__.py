"""Import-only stand-in for TensorFlow, used ONLY by oracle/make_golden.py::gen_tf_half_numpy_known to import the
reference modules whose *numpy* functions (DiagonalGaussian.kl / log_likelihood / entropy, conjugate_gradients) are
executed for golden vectors.  Every attribute resolves to an inert object; nothing here computes anything, and none of
the executed reference functions touches `tf`."""
import sys
import types


class _Inert(object):
    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *args, **kwargs):
        return _Inert()

    def __iter__(self):
        return iter(())


class _Module(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Inert()


sys.modules[__name__].__class__ = _Module
