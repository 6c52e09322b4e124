"""Device-resident trajectory storage and the lazy host views handed to reference-style callers.

One `PhaseData` holds everything a sampling phase produces, as SoA float32 tensors in HBM with the
reference's index contract (flat sample n = e*H + t inside task m):

    obs [M,N,Do]  act [M,N,Da]  mean [M,N,Da]  log_std [M,Da]  rew [M,N]  done [M,N] u8
    info [2,M,N] (env_infos, cheetah only)  returns [M,N]  adv [M,N]  coeffs [M,F] f64  stats [M,8] f64

`LazyPath` / `SamplesData` are dict-like views that copy to the host only when a caller actually
indexes them (the unchanged reference Trainer only needs them for diagnostics).
"""
from collections import OrderedDict

import numpy as np


class PhaseData(object):
    def __init__(self, M, E, H, obs_dim, act_dim, device):
        import torch
        self.M, self.E, self.H, self.N = M, E, H, E * H
        self.obs_dim, self.act_dim = obs_dim, act_dim
        f32 = dict(dtype=torch.float32, device=device)
        N = self.N
        self.obs = torch.empty(M, N, obs_dim, **f32)
        self.act = torch.empty(M, N, act_dim, **f32)
        self.mean = torch.empty(M, N, act_dim, **f32)
        self.log_std = torch.empty(M, act_dim, **f32)
        self.rew = torch.empty(M, N, **f32)
        self.done = torch.empty(M, N, dtype=torch.uint8, device=device)
        self.info = None
        self.info_keys = ()
        self.returns = None
        self.adv = None
        self.coeffs = None
        self.stats = None
        self.adj_avg_rewards = None
        self.generation = 0          # bumped whenever a kernel rewrites this phase's tensors in place
        self._host = {}

    def host(self, name):
        """numpy copy of a device tensor, fetched once per phase."""
        if name not in self._host:
            t = getattr(self, name)
            self._host[name] = None if t is None else t.detach().cpu().numpy()
        return self._host[name]

    def invalidate_host(self):
        """Called after every kernel that rewrote device tensors of this phase (rollout, processing)."""
        self._host = {}
        self.generation += 1

    def bytes_per_step(self):
        return 4 * (self.obs_dim + 2 * self.act_dim + 1) + 1 + 8 * len(self.info_keys)


class RaggedPhaseData(PhaseData):
    """Variable-length paths (early termination, meta_sampler.py:116-125): task m owns n_paths[m] paths stored back to
    back in its row of the [M, Nmax] tensors; rows past n_valid[m] are padding.  `N` (= Nmax) is the row stride the
    policy kernels see; per-task means run over n_valid[m] (promp_policy_*_ragged)."""

    def __init__(self, path_lens, obs_dim, act_dim, device):
        import torch
        self.path_lens = [list(map(int, l)) for l in path_lens]
        M = len(self.path_lens)
        n_valid = [sum(l) for l in self.path_lens]
        n_paths = [len(l) for l in self.path_lens]
        Pmax = max(n_paths)
        Nmax = (max(n_valid) + 3) // 4 * 4
        PhaseData.__init__(self, M, 1, Nmax, obs_dim, act_dim, device)
        self.E, self.H = Pmax, None
        for t in (self.obs, self.act, self.mean, self.rew):
            t.zero_()
        self.done.zero_()
        off = np.zeros((M, Pmax + 1), dtype=np.int32)
        for m, lens in enumerate(self.path_lens):
            off[m, 1:len(lens) + 1] = np.cumsum(lens)
            off[m, len(lens) + 1:] = off[m, len(lens)]
        self.path_off_host = off
        self.n_valid_host = np.asarray(n_valid, dtype=np.int32)
        self.n_paths_host = np.asarray(n_paths, dtype=np.int32)
        self.path_off = torch.from_numpy(off).to(device)
        self.n_valid = torch.from_numpy(self.n_valid_host).to(device)
        self.n_paths = torch.from_numpy(self.n_paths_host).to(device)

    @property
    def total_paths(self):
        return int(self.n_paths_host.sum())


class DeviceRaggedPhaseData(RaggedPhaseData):
    """RaggedPhaseData whose path table was built ON THE DEVICE (promp_paths_finalize after a fused early-termination
    rollout): path_off / n_paths / n_valid are device tensors with worst-case shapes (max_paths = max_samples = E * T per
    task); their host copies are fetched lazily, only when a caller asks for individual paths or logs path statistics."""

    def __init__(self, M, max_paths, max_samples, obs_dim, act_dim, device):
        import torch
        PhaseData.__init__(self, M, 1, max_samples, obs_dim, act_dim, device)
        self.E, self.H = max_paths, None
        i32 = dict(dtype=torch.int32, device=device)
        self.path_off = torch.zeros(M, max_paths + 1, **i32)
        self.n_paths = torch.zeros(M, **i32)
        self.n_valid = torch.zeros(M, **i32)
        self.src_slot = torch.zeros(M, max_paths, **i32)
        self.src_start = torch.zeros(M, max_paths, **i32)
        self.cut = torch.zeros(2, **i32)
        self._host_tables = None

    def _tables(self):
        if self._host_tables is None:
            n_paths = self.n_paths.cpu().numpy()
            self._host_tables = dict(n_paths=n_paths, n_valid=self.n_valid.cpu().numpy(),
                                     path_off=self.path_off[:, :int(n_paths.max()) + 1].cpu().numpy())
        return self._host_tables

    def invalidate_host(self):
        PhaseData.invalidate_host(self)

    n_paths_host = property(lambda self: self._tables()['n_paths'])
    n_valid_host = property(lambda self: self._tables()['n_valid'])
    path_off_host = property(lambda self: self._tables()['path_off'])

    @property
    def path_lens(self):
        t = self._tables()
        return [list(np.diff(t['path_off'][m, :t['n_paths'][m] + 1])) for m in range(self.M)]


class RaggedLazyPathList(object):
    """Paths of one task of a DeviceRaggedPhaseData as the list of dicts the reference builds (meta_sampler.py:116-123);
    the host copy of the path table is fetched on first use."""

    def __init__(self, phase, m):
        self.phase, self.m = phase, m
        self._paths = None

    def _get(self):
        if self._paths is None:
            p, m = self.phase, self.m
            off = p.path_off_host[m]
            out = []
            for k in range(int(p.n_paths_host[m])):
                s = slice(int(off[k]), int(off[k + 1]))
                out.append(dict(observations=p.host('obs')[m, s], actions=p.host('act')[m, s], rewards=p.host('rew')[m, s],
                                env_infos={}, agent_infos=dict(mean=p.host('mean')[m, s],
                                                               log_std=np.broadcast_to(p.host('log_std')[m], (s.stop - s.start, p.act_dim)))))
            self._paths = out
        return self._paths

    def __len__(self):
        return len(self._get())

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())

    def __add__(self, other):
        return list(self._get()) + list(other)

    def __radd__(self, other):
        return list(other) + list(self._get())


class RaggedSamplesData(dict):
    """SamplesData for a RaggedPhaseData: the 8 keys trimmed to the task's n_valid samples."""

    def __init__(self, phase, m, processor=None):
        super(RaggedSamplesData, self).__init__()
        self.phase, self.m, self._processor = phase, m, processor

    def __missing__(self, key):
        p, m = self.phase, self.m
        n = int(p.n_valid_host[m])
        if key == 'observations':
            v = p.host('obs')[m, :n]
        elif key == 'actions':
            v = p.host('act')[m, :n]
        elif key == 'rewards':
            v = p.host('rew')[m, :n]
        elif key == 'returns':
            v = p.host('returns')[m, :n]
        elif key == 'advantages':
            v = p.host('adv')[m, :n]
        elif key == 'agent_infos':
            v = dict(mean=p.host('mean')[m, :n], log_std=np.broadcast_to(p.host('log_std')[m], (n, p.act_dim)))
        elif key == 'env_infos':
            v = {}
        elif key == 'adj_avg_rewards':
            if p.adj_avg_rewards is None and self._processor is not None:
                self._processor.compute_adj_avg_rewards(p)
            v = p.host('adj_avg_rewards')[m, :n]
        else:
            raise KeyError(key)
        self[key] = v
        return v

    def keys(self):
        return list(SamplesData.KEYS)

    def __len__(self):
        return len(SamplesData.KEYS)

    def __iter__(self):
        return iter(SamplesData.KEYS)


class LazyPath(dict):
    """One trajectory (task m, env e) as the dict the reference builds at meta_sampler.py:116-123."""

    def __init__(self, phase, m, e):
        super(LazyPath, self).__init__()
        self.phase, self.m, self.e = phase, m, e

    def _slice(self):
        H = self.phase.H
        return slice(self.e * H, (self.e + 1) * H)

    def __missing__(self, key):
        p, s = self.phase, self._slice()
        if key == 'observations':
            v = p.host('obs')[self.m, s]
        elif key == 'actions':
            v = p.host('act')[self.m, s]
        elif key == 'rewards':
            v = p.host('rew')[self.m, s]
        elif key == 'returns' and p.returns is not None:
            v = p.host('returns')[self.m, s]
        elif key == 'agent_infos':
            v = dict(mean=p.host('mean')[self.m, s],
                     log_std=np.broadcast_to(p.host('log_std')[self.m], (p.H, p.act_dim)))
        elif key == 'env_infos':
            v = {k: p.host('info')[i, self.m, s] for i, k in enumerate(p.info_keys)}
        else:
            raise KeyError(key)
        self[key] = v
        return v

    def keys(self):
        base = ['observations', 'actions', 'rewards', 'env_infos', 'agent_infos']
        if self.phase.returns is not None:
            base.append('returns')
        return set(base) | set(dict.keys(self))

    def __len__(self):
        return self.phase.H


class LazyPathList(object):
    """A list of LazyPath views created on first access.  Building 800 dict objects per sampling phase cost more host time
    than the rollout kernel takes; the reference Trainer only indexes / concatenates / iterates these lists
    (meta_trainer.py:113: `sum(list(paths.values()), [])`), which this sequence supports."""

    def __init__(self, phase, tasks, cache=None):
        self.phase = phase
        self._tasks = list(tasks)          # task ids, E paths each
        # one LazyPath object per (task, env), shared by every list cut from the same obtain_samples() result (callers may
        # annotate paths).  NOT stored on the phase: phase -> path -> phase would be a reference cycle that keeps the
        # device buffers of old phases alive until the cyclic GC runs (observed: 10x slower eager iterations).
        self._cache = {} if cache is None else cache

    def __len__(self):
        return len(self._tasks) * self.phase.E

    def _make(self, i):
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        m, e = divmod(i, self.phase.E)
        key = (self._tasks[m], e)
        p = self._cache.get(key)
        if p is None:
            p = self._cache[key] = LazyPath(self.phase, key[0], e)
        return p

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._make(j) for j in range(*i.indices(len(self)))]
        return self._make(i)

    def __iter__(self):
        return (self._make(i) for i in range(len(self)))

    def _concat(self, a, b):
        if isinstance(a, LazyPathList) and isinstance(b, LazyPathList) and a.phase is b.phase:
            return LazyPathList(a.phase, a._tasks + b._tasks, a._cache)
        return list(a) + list(b)

    def __add__(self, other):
        return self._concat(self, other)

    def __radd__(self, other):
        if isinstance(other, list) and not other:
            return self                     # `[] + lazy` (the start value of sum(..., []))
        return self._concat(other, self)

    def __eq__(self, other):
        return list(self) == list(other)


class PathsMetaBatch(OrderedDict):
    """OrderedDict{task -> [path]*E} (what MetaSampler.obtain_samples returns) + the device phase."""
    phase = None


class SamplesData(dict):
    """Processed samples of one task: the 8 keys of samplers/meta_sample_processor.py:39-47, lazily."""
    KEYS = ('observations', 'actions', 'rewards', 'returns', 'advantages', 'env_infos', 'agent_infos',
            'adj_avg_rewards')

    def __init__(self, phase, m, processor=None):
        super(SamplesData, self).__init__()
        self.phase, self.m, self._processor = phase, m, processor

    def __missing__(self, key):
        p, m = self.phase, self.m
        if key == 'observations':
            v = p.host('obs')[m]
        elif key == 'actions':
            v = p.host('act')[m]
        elif key == 'rewards':
            v = p.host('rew')[m]
        elif key == 'returns':
            v = p.host('returns')[m]
        elif key == 'advantages':
            v = p.host('adv')[m]
        elif key == 'agent_infos':
            v = dict(mean=p.host('mean')[m], log_std=np.broadcast_to(p.host('log_std')[m], (p.N, p.act_dim)))
        elif key == 'env_infos':
            v = {k: p.host('info')[i, m] for i, k in enumerate(p.info_keys)}
        elif key == 'adj_avg_rewards':
            if p.adj_avg_rewards is None and self._processor is not None:
                self._processor.compute_adj_avg_rewards(p)
            v = p.host('adj_avg_rewards')[m]
        else:
            raise KeyError(key)
        self[key] = v
        return v

    def keys(self):
        return list(self.KEYS)

    def __len__(self):
        return len(self.KEYS)

    def __iter__(self):
        return iter(self.KEYS)

    def __contains__(self, key):
        return key in self.KEYS

    def items(self):
        return [(k, self[k]) for k in self.KEYS]
