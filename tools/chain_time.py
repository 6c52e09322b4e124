"""Time promp_policy_chain against the stand-alone launches of the same stages on synthetic data (CUDA events, L2-warm).
usage: python tools/chain_time.py [point|cheetah] [M]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promp_b200 import _lib  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else 'point'
    Do, Da, M, N = (2, 2, 40, 2000) if wl == 'point' else (17, 6, 40, 4000)
    if len(sys.argv) > 2:
        M = int(sys.argv[2])
    P = _lib.load().promp_num_params(Do, Da, 64)
    dev = torch.device('cuda')
    g = torch.Generator(device='cuda').manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device=dev)
    theta = 0.1 * r(P)
    ph = [dict(obs=r(M, N, Do), act=r(M, N, Da), adv=r(M, N), mean=r(M, N, Da), ls=0.1 * r(M, Da)) for _ in range(2)]
    g0, th1, v = torch.empty(M, P, device=dev), torch.empty(M, P, device=dev), torch.empty(M, P, device=dev)
    st = torch.zeros(3, M, 4, device=dev)
    s = _lib.stream()

    def stage(kind, p, params, stride, obj, **kw):
        t = _lib.PolicyStage()
        t.kind, t.N, t.params, t.param_stride = kind, N, _lib.ptr(params), stride
        t.obs, t.act, t.adv, t.old_mean, t.old_log_std = (_lib.ptr(p['obs']), _lib.ptr(p['act']), _lib.ptr(p['adv']), _lib.ptr(p['mean']),
                                                          _lib.ptr(p['ls']))
        t.obj_kind, t.obj_scale, t.clip_eps, t.kl_coeff, t.clip_log_std = obj, 1.0, 0.3, kw.get('klc', 0.0), kw.get('clip', 0)
        t.grad, t.out_params, t.sgd_lr = _lib.ptr(kw.get('grad')), _lib.ptr(kw.get('out_params')), 0.1
        t.inner_lr, t.vec, t.out, t.stats = 0.1, _lib.ptr(kw.get('vec')), _lib.ptr(kw.get('out')), _lib.ptr(kw.get('stats'))
        return t
    S0 = stage(0, ph[0], theta, 0, 0, clip=1, grad=g0, out_params=th1, stats=st[0])
    S1 = stage(0, ph[1], th1, P, 2, grad=v, stats=st[1])
    S2 = stage(1, ph[0], theta, 0, 0, clip=1, klc=5e-4, vec=v, out=v)
    wsd = {}

    def chain(stages):
        arr = (_lib.PolicyStage * len(stages))(*stages)
        need = _lib.load().promp_policy_chain_workspace_bytes(Do, Da, 64, M, len(stages), ctypes.cast(arr, ctypes.c_void_p))
        if 'ws' not in wsd or wsd['ws'].numel() * 4 < need:
            wsd['ws'] = torch.zeros((need + 3) // 4, dtype=torch.int32, device=dev)
        ws = wsd['ws']
        _lib.call('promp_policy_chain', Do, Da, 64, M, -13.8, len(stages), ctypes.cast(arr, ctypes.c_void_p), None, None, _lib.ptr(ws),
                  ws.numel() * 4, s)

    def timeit(fn, n=30):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3

    print('%s M=%d N=%d lib=%s' % (wl, M, N, os.path.basename(_lib.LIB_PATH)))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    if hasattr(lib, 'promp_debug_chain_clocks'):      # -DPROMP_EXP_CLOCKS build: where the CTAs of the dataflow kernel spend their time
        names = ['pop', 'dep-wait', 'load', 'tiles', 'flush', 'reduce', 'items', 'lasts', 'kernel', 'ctas']
        buf = (ctypes.c_ulonglong * 16)()
        for q in (int(x) for x in os.environ.get('CHAIN_QS', '0,1,4').split(',')):
            _lib.set_option('chain_q', q)
            for name, stages in (('grad0', [S0]), ('grad1', [S1]), ('hvp', [S2]), ('all3', [S0, S1, S2])):
                chain(stages)
                lib.promp_debug_chain_clocks(buf, 1)
                n = 10
                for _ in range(n):
                    chain(stages)
                lib.promp_debug_chain_clocks(buf, 1)
                v = list(buf)
                ctas = max(v[9] / n, 1)
                us = lambda c: c / n / ctas / 1965.0
                print('  q=%d %-6s per CTA: %s | items %.1f lasts %.2f | kernel %.1f us | per item: pop %.2f flush %.2f load %.2f' % (
                    q, name, ' '.join('%s %.1f' % (names[i], us(v[i])) for i in range(6)), v[6] / n / ctas, v[7] / n / ctas, us(v[8]),
                    v[0] / max(v[6], 1) / 1965.0, v[4] / max(v[6], 1) / 1965.0, v[2] / max(v[6], 1) / 1965.0))
        return
    _lib.set_option('chain', 0)
    for name, stages in (('grad0', [S0]), ('grad1', [S1]), ('hvp', [S2]), ('all3', [S0, S1, S2])):
        print('  separate launches %-6s %.1f us' % (name, timeit(lambda: chain(stages))))
    _lib.set_option('chain', 1)
    for q in (0, 1, 2, 4):
        for taper in (1, 0):
            _lib.set_option('chain_q', q)
            _lib.set_option('chain_taper', taper)
            res = ['%s %.1f' % (name, timeit(lambda: chain(stages)))
                   for name, stages in (('grad0', [S0]), ('grad1', [S1]), ('hvp', [S2]), ('all3', [S0, S1, S2]))]
            print('  chain q=%-2d taper=%d: %s us' % (q, taper, '  '.join(res)))


if __name__ == '__main__':
    main()
