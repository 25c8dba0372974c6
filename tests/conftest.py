import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for path in (ROOT, HERE, os.path.join(HERE, 'golden')):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'needs_reference: needs /root/reference (dev container only)')


def _load_golden(name='finder_golden.json'):
    with open(os.path.join(HERE, 'golden', name)) as fh:
        blob = json.load(fh)
    for case in blob['cases']:
        case['parts'] = blob['partsets'][case['partset']]
    return blob['cases']


GOLDEN_CASES = _load_golden()
# inputs outside the 2-bit domain (RNA, other symbols, windows longer than 32 bases): tests/golden/make_golden_general.py
GOLDEN_GENERAL = _load_golden('finder_golden_general.json')


@pytest.fixture(scope='session')
def golden_cases():
    return GOLDEN_CASES


@pytest.fixture()
def scratch_cwd(tmp_path, monkeypatch):
    """finder() creates ./<uuid>/ in the current directory (reference projector.py:7-19)."""
    monkeypatch.chdir(tmp_path)
    return tmp_path
