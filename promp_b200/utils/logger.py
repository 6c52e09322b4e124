"""Minimal key/value logger with the call surface the hot path uses
(ref: meta_policy_search/utils/logger.py logkv :204, dumpkvs :228, log, save_itr_params :376-396).
The reference's file writers / snapshotting are out of scope."""
import sys
from collections import OrderedDict

_kvs = OrderedDict()
_last_dump = OrderedDict()
_quiet = [False]


def set_quiet(q=True):
    _quiet[0] = bool(q)


def logkv(key, val):
    _kvs[key] = val


def logkvs(d):
    for k, v in d.items():
        logkv(k, v)


def getkvs():
    return _kvs


def last_dump():
    return _last_dump


def dumpkvs():
    _last_dump.clear()
    _last_dump.update(_kvs)
    if not _quiet[0] and _kvs:
        width = max(len(str(k)) for k in _kvs)
        lines = ['%s | %s' % (str(k).ljust(width), ('%-8.5g' % v) if isinstance(v, float) else str(v))
                 for k, v in sorted(_kvs.items())]
        bar = '-' * max(len(l) for l in lines)
        sys.stdout.write('\n'.join([bar] + lines + [bar]) + '\n')
    _kvs.clear()


def log(*args):
    if not _quiet[0]:
        print(*args)


def save_itr_params(itr, params):
    pass


def configure(*args, **kwargs):
    pass
