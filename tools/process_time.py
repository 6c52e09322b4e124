"""Time promp_process_samples in isolation on synthetic data (CUDA events, L2-warm like the real loop).
usage: [PROMP_B200_LIB=promp_b200/libpromp_b200_clk.so] python tools/process_time.py
With a -DPROMP_EXP_CLOCKS build (tools/build_clk_variant.sh) also prints the per-phase clocks of front-stage CTA (0,0)
and of the finishing CTA of task 0."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promp_b200 import _lib  # noqa: E402

NAMES = ['front: load rewards', 'front: return scans', 'front: stats reduce + returns write', 'front: Gram tiles',
         'front: group reduce + partial write', 'front: fence + ticket', '', '',
         'finish: reduce partials (+reward load)', 'finish: Cholesky + solves', 'finish: predict', 'finish: GAE scans',
         'finish: moments + advantages write']


def main():
    lib = _lib.load()
    dev = torch.device('cuda')
    for name, (M, E, H, Do) in (('point 40x20x100 Do=2', (40, 20, 100, 2)), ('cheetah 40x20x200 Do=17', (40, 20, 200, 17))):
        g = torch.Generator(device='cuda').manual_seed(0)
        obs = torch.randn(M, E * H, Do, generator=g, device=dev)
        rew = torch.randn(M, E * H, generator=g, device=dev)
        ret, adv = torch.empty(M, E * H, device=dev), torch.empty(M, E * H, device=dev)
        coeffs = torch.empty(M, 2 * Do + 4, dtype=torch.float64, device=dev)
        stats = torch.empty(M, 8, dtype=torch.float64, device=dev)
        nbytes = lib.promp_process_workspace_bytes(M, E, H, Do)
        ws = torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=dev)

        def call():
            _lib.call('promp_process_samples', M, E, H, Do, _lib.ptr(obs), _lib.ptr(rew), 0.99, 1.0, 1e-5, 1, 1, 0, _lib.ptr(ret),
                      _lib.ptr(adv), _lib.ptr(coeffs), _lib.ptr(stats), _lib.ptr(ws), ws.numel() * 8, _lib.stream())
        for _ in range(5):
            call()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        a.record()
        for _ in range(n):
            call()
        b.record()
        torch.cuda.synchronize()
        print('%s: %.1f us per launch (lib %s)' % (name, a.elapsed_time(b) / n * 1e3, os.path.basename(_lib.LIB_PATH)))
        if hasattr(lib, 'promp_debug_proc_clocks'):
            buf = (ctypes.c_ulonglong * 16)()
            lib.promp_debug_proc_clocks(buf, 1)
            for _ in range(n):
                call()
            lib.promp_debug_proc_clocks(buf, 1)
            for i, nm in enumerate(NAMES):
                if nm:
                    print('    %-40s %8.0f clk' % (nm, buf[i] / n))


if __name__ == '__main__':
    main()
