"""Key/value logger + snapshot writer with the reference's call surface and file formats
(ref: meta_policy_search/utils/logger.py: logkv :204, dumpkvs :228, log :246, configure :401-427, output formats :37-146,
save_itr_params :376-396).  Same keys, same files (`log.txt`, `progress.csv`, `progress.json`, `params.pkl` /
`itr_N.pkl`), so plotting scripts written for the reference read a promp_b200 run unchanged.  Own implementation;
TensorBoard output is not provided (TensorFlow is not a dependency).
"""
import json
import os
import os.path as osp
import pickle
import sys
from collections import OrderedDict

_kvs = OrderedDict()
_last_dump = OrderedDict()
_quiet = [False]


def _plain(v):
    """numpy / torch scalars -> python numbers (for json / csv)."""
    if hasattr(v, 'item') and getattr(v, 'ndim', 0) == 0:
        try:
            return v.item()
        except Exception:
            return v
    if hasattr(v, 'tolist') and getattr(v, 'size', 2) == 1:
        return float(v.reshape(-1)[0])
    return v


class _TableWriter(object):
    """Sorted 'key | value' table (stdout or log.txt); values printed with %-8.3g like the reference."""

    def __init__(self, target):
        self.own = isinstance(target, str)
        self.f = open(target, 'wt') if self.own else target

    @staticmethod
    def _cut(s):
        return s[:20] + '...' if len(s) > 23 else s

    def writekvs(self, kvs):
        if not kvs:
            return
        rows = [(self._cut(str(k)), self._cut(('%-8.3g' % v) if isinstance(v, float) else str(v)))
                for k, v in sorted(kvs.items())]
        kw, vw = max(len(k) for k, _ in rows), max(len(v) for _, v in rows)
        bar = '-' * (kw + vw + 7)
        self.f.write('\n'.join([bar] + ['| %s | %s |' % (k.ljust(kw), v.ljust(vw)) for k, v in rows] + [bar]) + '\n')
        self.f.flush()

    def writeseq(self, args):
        self.f.write(''.join(str(a) for a in args) + '\n')
        self.f.flush()

    def close(self):
        if self.own:
            self.f.close()


class _JsonWriter(object):
    """One json object per dumpkvs() call."""

    def __init__(self, path):
        self.f = open(path, 'wt')

    def writekvs(self, kvs):
        self.f.write(json.dumps({k: _plain(v) for k, v in sorted(kvs.items())}) + '\n')
        self.f.flush()

    def close(self):
        self.f.close()


class _CsvWriter(object):
    """progress.csv: when a dump brings new keys the header grows and the earlier rows are padded with empty cells."""

    def __init__(self, path):
        self.path, self.keys, self.rows = path, [], []
        open(path, 'wt').close()

    def writekvs(self, kvs):
        new = [k for k in kvs if k not in self.keys]
        self.keys.extend(new)
        self.rows.append(['' if kvs.get(k) is None else str(_plain(kvs[k])) for k in self.keys])
        if new or len(self.rows) == 1:          # rewrite everything with the wider header
            with open(self.path, 'wt') as f:
                f.write(','.join(self.keys) + '\n')
                for r in self.rows:
                    f.write(','.join(r + [''] * (len(self.keys) - len(r))) + '\n')
        else:
            with open(self.path, 'at') as f:
                f.write(','.join(self.rows[-1]) + '\n')

    def close(self):
        pass


class _State(object):
    dir = None
    writers = None            # None -> default: table on stdout
    snapshot_mode = 'last'
    snapshot_gap = 1


_S = _State()


def _writers():
    if _S.writers is None:
        _S.writers = [_TableWriter(sys.stdout)]
    return _S.writers


def configure(dir=None, format_strs=None, snapshot_mode='last', snapshot_gap=1):
    """ref logger.py:401-427.  dir defaults to $OPENAI_LOGDIR; formats: stdout, log, csv, json."""
    if dir is None:
        dir = os.getenv('OPENAI_LOGDIR')
    if dir is None:
        raise ValueError("logger.configure: no directory given and OPENAI_LOGDIR is not set")
    os.makedirs(dir, exist_ok=True)
    if format_strs is None:
        env = os.getenv('OPENAI_LOG_FORMAT')
        format_strs = env.split(',') if env else ['stdout', 'log', 'csv']
    reset()
    ws = []
    for fmt in format_strs:
        if fmt == 'stdout':
            ws.append(_TableWriter(sys.stdout))
        elif fmt == 'log':
            ws.append(_TableWriter(osp.join(dir, 'log.txt')))
        elif fmt == 'json':
            ws.append(_JsonWriter(osp.join(dir, 'progress.json')))
        elif fmt == 'csv':
            ws.append(_CsvWriter(osp.join(dir, 'progress.csv')))
        else:
            raise ValueError('Unknown format specified: %s' % (fmt,))
    assert snapshot_mode in ('all', 'last', 'gap', 'last_gap', 'none'), snapshot_mode
    _S.dir, _S.writers, _S.snapshot_mode, _S.snapshot_gap = dir, ws, snapshot_mode, snapshot_gap
    log('Logging to %s' % dir)


def reset():
    if _S.writers:
        for w in _S.writers:
            w.close()
    _S.dir, _S.writers, _S.snapshot_mode, _S.snapshot_gap = None, None, 'last', 1


def get_dir():
    return _S.dir


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
    for w in _writers():
        if _quiet[0] and isinstance(w, _TableWriter) and not w.own:
            continue
        w.writekvs(_kvs)
    _kvs.clear()


record_tabular = logkv
dump_tabular = dumpkvs


def log(*args):
    for w in _writers():
        if isinstance(w, _TableWriter) and not (_quiet[0] and not w.own):
            w.writeseq(args)


def _dump(obj, path):
    try:
        import joblib
        joblib.dump(obj, path, compress=3)
    except ImportError:
        with open(path, 'wb') as f:
            pickle.dump(obj, f, protocol=pickle.HIGHEST_PROTOCOL)


def load_snapshot(path):
    """Read a snapshot written by save_itr_params (joblib, falling back to pickle)."""
    try:
        import joblib
        return joblib.load(path)
    except ImportError:
        with open(path, 'rb') as f:
            return pickle.load(f)


def snapshot_due(itr):
    """True when save_itr_params(itr, ...) would write a file (lets the Trainer skip building the snapshot - device->host
    copies of the parameters and optimizer slots - on the iterations that are not saved)."""
    if not _S.dir:
        return False
    mode, gap = _S.snapshot_mode, _S.snapshot_gap
    if mode in ('all', 'last'):
        return True
    if mode in ('gap', 'last_gap'):
        return itr % gap == 0
    return False


def save_itr_params(itr, params):
    """ref logger.py:376-396: snapshot_mode all / last / gap / last_gap / none.  `params` may be a zero-argument callable
    that builds the snapshot; it is only called when a file is written."""
    if not _S.dir:
        return None
    mode, gap = _S.snapshot_mode, _S.snapshot_gap
    name = None
    if mode == 'all':
        name = 'itr_%d.pkl' % itr
    elif mode == 'last':
        name = 'params.pkl'
    elif mode == 'gap':
        name = 'itr_%d.pkl' % itr if itr % gap == 0 else None
    elif mode == 'last_gap':
        name = 'params.pkl' if itr % gap == 0 else None
    elif mode == 'none':
        name = None
    else:
        raise NotImplementedError(mode)
    if name is None:
        return None
    path = osp.join(_S.dir, name)
    _dump(params() if callable(params) else params, path)
    return path
