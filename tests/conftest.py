import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


WORST = {}   # name -> the largest error a GPU parity test saw (written to gpurun_out/test_errors.json when the session ends)


def note_err(name, value):
    WORST[name] = max(WORST.get(name, 0.0), float(value))


def pytest_sessionfinish(session, exitstatus):
    if WORST:
        import json
        d = os.path.join(ROOT, 'gpurun_out')
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, 'test_errors.json'), 'w') as f:
            json.dump(WORST, f, indent=1, sort_keys=True)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
