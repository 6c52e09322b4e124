"""Debug aid: report the first C-ABI call after which an active stream capture is found invalidated.
usage: python tools/debug_capture.py <pytest args...>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from promp_b200 import _lib  # noqa: E402

_orig = _lib.call
_state = {'reported': False, 'n': 0}


def call(name, *args):
    cap_before = None
    try:
        cap_before = torch.cuda.is_current_stream_capturing()
    except Exception as e:      # invalidated before this call (by a torch op in between)
        if not _state['reported']:
            _state['reported'] = True
            sys.stderr.write('CAPTURE ALREADY INVALID before %s: %r\n' % (name, e))
    try:
        _orig(name, *args)
    except Exception as e:
        if not _state['reported']:
            _state['reported'] = True
            sys.stderr.write('CALL %s RAISED (capturing before: %s): %r\n' % (name, cap_before, e))
        raise
    if cap_before:
        _state['n'] += 1
        try:
            torch.cuda.is_current_stream_capturing()
        except Exception as e:
            if not _state['reported']:
                _state['reported'] = True
                sys.stderr.write('CAPTURE INVALIDATED BY %s (captured call #%d): %r\n' % (name, _state['n'], e))


_lib.call = call
import pytest  # noqa: E402
sys.exit(pytest.main(sys.argv[1:]))
